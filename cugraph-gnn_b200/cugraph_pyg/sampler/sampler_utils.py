"""Helpers of the sampler layer (role of the reference's cugraph_pyg/sampler/sampler_utils.py:19-336)."""
from math import ceil
from typing import Dict, Optional, Tuple, Union

import torch

from cugraph_pyg._pyg_compat import Data, HeteroData


def verify_metadata(metadata: Optional[Dict[str, Union[str, Tuple[str, str, str]]]]):
    if metadata is None:
        return
    for k, v in metadata.items():
        if not isinstance(k, str):
            raise ValueError("Metadata keys must be strings.")
        if isinstance(v, tuple):
            if not all(isinstance(x, str) for x in v):
                raise ValueError("Metadata tuples must contain only strings.")
        elif not isinstance(v, str):
            raise ValueError("Metadata values must be strings or tuples of strings.")


def filter_cugraph_pyg_store(feature_store, graph_store, node, row, col, edge, clx=None) -> Data:
    """Mini-batch Data object: edge_index + every stored feature, fetched for the sampled nodes (or edges when
    the feature's group is an edge type).  The node index stays on the device all the way into the gather."""
    data = Data()
    data.edge_index = torch.stack([row, col], dim=0)
    attrs = []
    for attr in feature_store.get_all_tensor_attrs():
        attr.index = edge if isinstance(attr.group_name, tuple) else node
        attrs.append(attr)
        data.num_nodes = int(node.numel())
    for attr, tensor in zip(attrs, feature_store.multi_get_tensor(attrs)):
        data[attr.attr_name] = tensor
    return data


def filter_cugraph_pyg_hetero_store(feature_store, graph_store, node_dict, row_dict, col_dict, edge_dict, clx=None) -> HeteroData:
    """Heterogeneous mini-batch (what torch_geometric.loader.utils.filter_custom_hetero_store builds for the
    reference, sampler.py:118-132): per edge type an edge_index, per node / edge type every stored feature."""
    data = HeteroData()
    for et in edge_dict.keys():
        data[et].edge_index = torch.stack([row_dict[et], col_dict[et]], dim=0)
    for nt, node in node_dict.items():
        data[nt].num_nodes = int(node.numel())
    attrs = []
    for attr in feature_store.get_all_tensor_attrs():
        if isinstance(attr.group_name, tuple):
            if attr.group_name not in edge_dict:
                continue
            attr.index = edge_dict[attr.group_name]
        else:
            if attr.group_name not in node_dict:
                continue
            attr.index = node_dict[attr.group_name]
        attrs.append(attr)
    for attr, tensor in zip(attrs, feature_store.multi_get_tensor(attrs)):
        data[attr.group_name][attr.attr_name] = tensor
    return data


def neg_sample(graph_store, seed_src, seed_dst, input_type, batch_size: int, neg_sampling, seed_time=None,
               node_time_func=None) -> Tuple[torch.Tensor, torch.Tensor]:
    """Negative (source, destination) pairs for a whole epoch of seed edges (reference: sampler_utils.py:95-315).
    At least one negative per batch; ids come back type-offset, ready for the sampler."""
    import pylibcugraph

    if node_time_func is not None:
        raise NotImplementedError("temporal negative sampling is not implemented (DESIGN.md §10)")
    src_weight = getattr(neg_sampling, "src_weight", getattr(neg_sampling, "weight", None))
    dst_weight = getattr(neg_sampling, "dst_weight", getattr(neg_sampling, "weight", None))
    num_neg = max(int(ceil(neg_sampling.amount * seed_src.numel())), int(ceil(seed_src.numel() / batch_size)))
    nv = graph_store._num_vertices()
    if graph_store.is_homogeneous:
        num_src = num_dst = list(nv.values())[0]
        off_src = off_dst = 0
    else:
        num_src, num_dst = nv[input_type[0]], nv[input_type[2]]
        off_src, off_dst = graph_store._vertex_offsets[input_type[0]], graph_store._vertex_offsets[input_type[2]]
    for name, w, n in (("src_weight", src_weight, num_src), ("dst_weight", dst_weight, num_dst)):
        if w is not None and w.numel() != n:
            raise ValueError(f"The '{name}' attribute needs to match the number of nodes {n} (got {w.numel()})")
    if src_weight is not None and dst_weight is not None and src_weight.dtype != dst_weight.dtype:
        raise ValueError(f"The 'src_weight' and 'dst_weight' attributes need to have the same dtype "
                         f"(got {src_weight.dtype} and {dst_weight.dtype})")
    # sources and destinations are drawn from their own vertex-type ranges
    res_s = pylibcugraph.negative_sampling(graph_store._resource_handle, graph_store._graph, num_neg,
                                           vertices=torch.arange(num_src, device="cuda") + off_src, src_bias=src_weight,
                                           dst_bias=None)
    res_d = pylibcugraph.negative_sampling(graph_store._resource_handle, graph_store._graph, num_neg,
                                           vertices=torch.arange(num_dst, device="cuda") + off_dst, src_bias=dst_weight,
                                           dst_bias=None)
    return res_s["sources"], res_d["sources"]


def neg_cat(seed_pos: torch.Tensor, seed_neg: torch.Tensor, pos_batch_size: int) -> Tuple[torch.Tensor, int]:
    """Interleaves positives and negatives batch by batch: [pos batch 0 | neg batch 0 | pos batch 1 | ...]
    (reference: sampler_utils.py:318-336); returns the joined tensor and the negatives per batch."""
    num_batches = int(ceil(seed_pos.numel() / pos_batch_size))
    neg_batch_size = int(ceil(seed_neg.numel() / num_batches))
    pos = torch.tensor_split(seed_pos, list(range(pos_batch_size, pos_batch_size * num_batches, pos_batch_size)))
    neg = torch.tensor_split(seed_neg, list(range(neg_batch_size, neg_batch_size * num_batches, neg_batch_size)))
    return torch.cat([torch.cat(pair) for pair in zip(pos, neg)]), neg_batch_size
