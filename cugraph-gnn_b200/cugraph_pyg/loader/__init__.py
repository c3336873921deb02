from .node_loader import NodeLoader
from .neighbor_loader import NeighborLoader
from .link_loader import LinkLoader
from .link_neighbor_loader import LinkNeighborLoader

__all__ = ["NodeLoader", "NeighborLoader", "LinkLoader", "LinkNeighborLoader"]
