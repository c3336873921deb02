from .graph_store import GraphStore
from .feature_store import FeatureStore

__all__ = ["GraphStore", "FeatureStore"]
