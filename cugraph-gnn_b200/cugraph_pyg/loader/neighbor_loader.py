"""NeighborLoader (role of the reference's cugraph_pyg/loader/neighbor_loader.py:15-236)."""
import warnings
from typing import Callable, Dict, List, Optional, Union

import cugraph_pyg
from cugraph_pyg.sampler import BaseSampler, DistributedNeighborSampler
from .node_loader import NodeLoader


class NeighborLoader(NodeLoader):
    """Duck-typed torch_geometric.loader.NeighborLoader: GraphSAGE-style neighbour sampling of input nodes."""

    def __init__(self, data, num_neighbors: Union[List[int], Dict], input_nodes=None, input_time=None, replace: bool = False,
                 subgraph_type: str = "directional", disjoint: bool = False, temporal_strategy: str = "uniform",
                 time_attr: Optional[str] = None, weight_attr: Optional[str] = None, transform: Optional[Callable] = None,
                 transform_sampler_output: Optional[Callable] = None, is_sorted: bool = False,
                 filter_per_worker: Optional[bool] = None, neighbor_sampler=None, directed: bool = True,
                 batch_size: int = 16, compression: Optional[str] = None, local_seeds_per_call: Optional[int] = None,
                 temporal_comparison: Optional[str] = None, **kwargs):
        if str(getattr(subgraph_type, "value", subgraph_type)) != "directional" or not directed:
            raise ValueError("Only directional subgraphs are currently supported")
        if neighbor_sampler is not None:
            raise ValueError("Passing a neighbor sampler is currently unsupported")
        if is_sorted:
            warnings.warn("The 'is_sorted' argument is ignored by cuGraph.")
        if time_attr is not None or input_time is not None:
            raise NotImplementedError("temporal sampling is outside the B200 hot path")
        if replace:
            raise NotImplementedError("sampling with replacement is outside the B200 hot path")
        if disjoint:
            raise NotImplementedError("disjoint sampling is outside the B200 hot path")
        if not isinstance(data, (list, tuple)) or not isinstance(data[1], cugraph_pyg.data.GraphStore):
            raise NotImplementedError("Currently can't accept non-cugraph graphs")
        feature_store, graph_store = data
        if compression is None:
            compression = "CSR" if graph_store.is_homogeneous else "COO"
        elif compression not in ("CSR", "COO"):
            raise ValueError("Invalid value for compression (expected 'CSR' or 'COO')")
        if isinstance(num_neighbors, dict) or not graph_store.is_homogeneous:
            raise NotImplementedError("heterogeneous sampling is not built yet (SURVEY.md §8e C5)")
        if weight_attr is not None:
            graph_store._set_weight_attr((feature_store, weight_attr))
        sampler = BaseSampler(
            DistributedNeighborSampler(
                graph_store._graph, retain_original_seeds=True, fanout=num_neighbors, prior_sources_behavior="exclude",
                deduplicate_sources=True, compression=compression, compress_per_hop=False, with_replacement=replace,
                disjoint=disjoint, local_seeds_per_call=local_seeds_per_call, biased=(weight_attr is not None),
                heterogeneous=False, temporal=False, num_edge_types=1),
            (feature_store, graph_store), batch_size=batch_size)
        super().__init__((feature_store, graph_store), sampler, input_nodes=input_nodes, input_time=input_time,
                         transform=transform, transform_sampler_output=transform_sampler_output,
                         filter_per_worker=filter_per_worker, batch_size=batch_size, **kwargs)
