"""NeighborLoader (role of the reference's cugraph_pyg/loader/neighbor_loader.py:15-236)."""
import warnings

import numpy as np
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
        if temporal_strategy != "uniform":
            warnings.warn("Only the uniform temporal strategy is currently supported")
        if temporal_comparison is None:
            temporal_comparison = "monotonically_decreasing"
        is_temporal = time_attr is not None
        if input_time is not None and not is_temporal:
            raise ValueError("input_time needs time_attr (the name of the edge time attribute)")
        if replace:
            raise NotImplementedError("sampling with replacement is outside the B200 hot path")
        if not isinstance(data, (list, tuple)) or not isinstance(data[1], cugraph_pyg.data.GraphStore):
            raise NotImplementedError("Currently can't accept non-cugraph graphs")
        feature_store, graph_store = data
        if is_temporal:
            graph_store._set_time_attr((feature_store, time_attr))
            if input_time is None:
                # the time attribute is assumed to exist for the input nodes as well (reference :178-187)
                from cugraph_pyg._pyg_compat import get_input_nodes

                in_type, in_nodes, _ = get_input_nodes(data, input_nodes, None)
                if in_type is None:
                    in_type = list(graph_store._vertex_offsets.keys())[0]
                input_time = feature_store[in_type, time_attr, None][in_nodes]
        if compression is None:
            compression = "CSR" if graph_store.is_homogeneous else "COO"
        elif compression not in ("CSR", "COO"):
            raise ValueError("Invalid value for compression (expected 'CSR' or 'COO')")
        if not graph_store.is_homogeneous and compression != "COO":
            raise ValueError("Only COO format is supported for heterogeneous graphs!")
        heterogeneous = not graph_store.is_homogeneous
        num_edge_types = len(graph_store.get_all_edge_attrs())
        if isinstance(num_neighbors, dict):
            # fan-out vector laid out [hop * T + edge type], edge types in sorted order (reference :192-201)
            sorted_keys, _, _ = graph_store._numeric_edge_types
            hops = len(next(iter(num_neighbors.values())))
            na = np.zeros(hops * len(sorted_keys), dtype="int32")
            for i, key in enumerate(sorted_keys):
                if key in num_neighbors:
                    for hop in range(hops):
                        na[hop * len(sorted_keys) + i] = num_neighbors[key][hop]
            num_neighbors = na
        elif heterogeneous or num_edge_types > 1:
            # a plain list on a typed graph means "the same fan-out for every edge type" (PyG's convention)
            num_neighbors = np.repeat(np.asarray(num_neighbors, dtype="int32"), num_edge_types)
        heterogeneous = heterogeneous or num_edge_types > 1
        if heterogeneous:
            compression = "COO"
        if disjoint:
            compression = "COO"  # cross-tree edges are removed from the COO result (pylibcugraph._disjoint_filter[_hetero])
        if weight_attr is not None:
            graph_store._set_weight_attr((feature_store, weight_attr))
        sampler = BaseSampler(
            DistributedNeighborSampler(
                graph_store._graph, retain_original_seeds=True, fanout=num_neighbors, prior_sources_behavior="exclude",
                deduplicate_sources=True, compression=compression, compress_per_hop=False, with_replacement=replace,
                disjoint=disjoint, local_seeds_per_call=local_seeds_per_call, biased=(weight_attr is not None),
                heterogeneous=heterogeneous, temporal=is_temporal, temporal_comparison=temporal_comparison,
                vertex_type_offsets=graph_store._vertex_offset_array,
                num_edge_types=num_edge_types),
            (feature_store, graph_store), batch_size=batch_size)
        super().__init__((feature_store, graph_store), sampler, input_nodes=input_nodes, input_time=input_time,
                         transform=transform, transform_sampler_output=transform_sampler_output,
                         filter_per_worker=filter_per_worker, batch_size=batch_size, **kwargs)
