"""LinkNeighborLoader (role of the reference's cugraph_pyg/loader/link_neighbor_loader.py:16-239)."""
import warnings
from typing import Callable, Dict, List, Optional, Union

import numpy as np

import cugraph_pyg
from cugraph_pyg.sampler import BaseSampler, DistributedNeighborSampler
from .link_loader import LinkLoader


class LinkNeighborLoader(LinkLoader):
    """Duck-typed torch_geometric.loader.LinkNeighborLoader: neighbour sampling around the endpoints of seed edges."""

    def __init__(self, data, num_neighbors: Union[List[int], Dict], edge_label_index=None, edge_label=None,
                 edge_label_time=None, replace: bool = False, subgraph_type: str = "directional", disjoint: bool = False,
                 temporal_strategy: str = "uniform", neg_sampling=None, neg_sampling_ratio=None, time_attr: Optional[str] = None,
                 weight_attr: Optional[str] = None, transform: Optional[Callable] = None,
                 transform_sampler_output: Optional[Callable] = None, is_sorted: bool = False,
                 filter_per_worker: Optional[bool] = None, neighbor_sampler=None, directed: bool = True, batch_size: int = 16,
                 compression: Optional[str] = None, local_seeds_per_call: Optional[int] = None,
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
        is_temporal = (edge_label_time is not None) and (time_attr is not None)
        if not is_temporal and (edge_label_time is not None or time_attr is not None):
            warnings.warn("Edge-based temporal sampling requires that both edge_label_time and time_attr are provided. "
                          "Defaulting to non-temporal sampling.")
            edge_label_time = None
        if is_temporal and neg_sampling is not None:
            raise NotImplementedError("temporal negative sampling is not implemented (DESIGN.md §10)")
        if replace:
            raise NotImplementedError("sampling with replacement is outside the B200 hot path")
        if not isinstance(data, (list, tuple)) or not isinstance(data[1], cugraph_pyg.data.GraphStore):
            raise NotImplementedError("Currently can't accept non-cugraph graphs")
        feature_store, graph_store = data
        if is_temporal:
            graph_store._set_time_attr((feature_store, time_attr))
        if compression is None:
            compression = "CSR" if graph_store.is_homogeneous else "COO"
        elif compression not in ("CSR", "COO"):
            raise ValueError("Invalid value for compression (expected 'CSR' or 'COO')")
        if not graph_store.is_homogeneous and compression != "COO":
            raise ValueError("Only COO format is supported for heterogeneous graphs!")
        heterogeneous = not graph_store.is_homogeneous
        num_edge_types = len(graph_store.get_all_edge_attrs())
        if isinstance(num_neighbors, dict):
            sorted_keys, _, _ = graph_store._numeric_edge_types
            hops = len(next(iter(num_neighbors.values())))
            na = np.zeros(hops * len(sorted_keys), dtype="int32")
            for i, key in enumerate(sorted_keys):
                if key in num_neighbors:
                    for hop in range(hops):
                        na[hop * len(sorted_keys) + i] = num_neighbors[key][hop]
            num_neighbors = na
        elif heterogeneous or num_edge_types > 1:
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
        super().__init__((feature_store, graph_store), sampler, edge_label_index=edge_label_index, edge_label=edge_label,
                         edge_label_time=edge_label_time, neg_sampling=neg_sampling, neg_sampling_ratio=neg_sampling_ratio,
                         transform=transform, transform_sampler_output=transform_sampler_output,
                         filter_per_worker=filter_per_worker, batch_size=batch_size, **kwargs)
