"""Helpers of the sampler layer (role of the reference's cugraph_pyg/sampler/sampler_utils.py:19-63; negative
sampling lives on the link-prediction path, SURVEY.md §8f row 2, and is not built)."""
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
