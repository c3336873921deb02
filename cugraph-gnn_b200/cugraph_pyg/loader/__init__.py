from .node_loader import NodeLoader
from .neighbor_loader import NeighborLoader

__all__ = ["NodeLoader", "NeighborLoader"]
