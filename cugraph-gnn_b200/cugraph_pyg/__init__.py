"""cugraph_pyg API surface (GraphStore / FeatureStore / DistTensor / NeighborLoader) over the B200 hot path.

Mirrors rapidsai/cugraph-gnn's python/cugraph-pyg package for homogeneous node-based neighbour sampling
(SURVEY.md §2.1 rows 16-19, §8 L1/L2).  Works without torch_geometric (duck-typed stand-ins in _pyg_compat).
"""
__version__ = "26.10.00+b200"

from . import data, tensor, sampler, loader  # noqa: F401,E402
