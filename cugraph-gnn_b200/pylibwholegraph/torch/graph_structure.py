"""CSR graph in WholeMemory + WholeGraph-style multi-layer sampling
(role of the reference's pylibwholegraph/torch/graph_structure.py:13-196)."""
from typing import List, Union

from pylibwholegraph.utils.imports import import_optional
from . import graph_ops
from . import wholegraph_ops
from .tensor import WholeMemoryTensor

torch = import_optional("torch")


class GraphStructure(object):
    r"""One relation in CSR form (int64 row pointers, int32/int64 column ids) plus named
    node / edge attribute tensors, all WholeMemory tensors."""

    def __init__(self):
        self.node_count = 0
        self.edge_count = 0
        self.csr_row_ptr = None
        self.csr_col_ind = None
        self.node_attributes = {}
        self.edge_attributes = {}

    def set_csr_graph(self, csr_row_ptr: WholeMemoryTensor, csr_col_ind: WholeMemoryTensor):
        assert csr_row_ptr.dim() == 1 and csr_col_ind.dim() == 1
        assert csr_row_ptr.dtype == torch.int64
        assert csr_row_ptr.shape[0] > 1
        assert csr_col_ind.dtype in (torch.int32, torch.int64)
        self.node_count = csr_row_ptr.shape[0] - 1
        self.edge_count = csr_col_ind.shape[0]
        self.csr_row_ptr = csr_row_ptr
        self.csr_col_ind = csr_col_ind

    def set_node_attribute(self, attr_name: str, attr_tensor: WholeMemoryTensor):
        assert attr_name not in self.node_attributes
        assert attr_tensor.shape[0] == self.node_count
        self.node_attributes[attr_name] = attr_tensor

    def set_edge_attribute(self, attr_name: str, attr_tensor: WholeMemoryTensor):
        assert attr_name not in self.edge_attributes
        assert attr_tensor.shape[0] == self.edge_count
        self.edge_attributes[attr_name] = attr_tensor

    def unweighted_sample_without_replacement_one_hop(self, center_nodes_tensor: "torch.Tensor", max_sample_count: int, *,
                                                      random_seed: Union[int, None] = None,
                                                      need_center_local_output: bool = False,
                                                      need_edge_output: bool = False):
        """-> (csr_row_ptr, sampled_nodes[, center_node_local_id][, edge_index])"""
        return wholegraph_ops.unweighted_sample_without_replacement(
            self.csr_row_ptr.wmb_tensor, self.csr_col_ind.wmb_tensor, center_nodes_tensor, max_sample_count,
            random_seed, need_center_local_output, need_edge_output,
        )

    def weighted_sample_without_replacement_one_hop(self, weight_name: str, center_nodes_tensor: "torch.Tensor",
                                                    max_sample_count: int, *, random_seed: Union[int, None] = None,
                                                    need_center_local_output: bool = False,
                                                    need_edge_output: bool = False):
        assert weight_name in self.edge_attributes
        return wholegraph_ops.weighted_sample_without_replacement(
            self.csr_row_ptr.wmb_tensor, self.csr_col_ind.wmb_tensor, self.edge_attributes[weight_name].wmb_tensor,
            center_nodes_tensor, max_sample_count, random_seed, need_center_local_output, need_edge_output,
        )

    def multilayer_sample_without_replacement(self, node_ids: "torch.Tensor", max_neighbors: List[int],
                                              weight_name: Union[str, None] = None,
                                              random_seed: Union[int, None] = None):
        """WholeGraph-style layered sampling: layer i samples from ALL targets of layer i+1.

        Returns (target_gids, edge_indice, csr_row_ptr, csr_col_ind), lists indexed by layer with
        layer `hops` = the seeds (same contract as the reference).  `random_seed` (an extension)
        makes the whole call reproducible: hop h uses random_seed + h."""
        hops = len(max_neighbors)
        target_gids = [None] * hops + [node_ids]
        edge_indice, csr_row_ptr, csr_col_ind = [None] * hops, [None] * hops, [None] * hops
        for depth, layer in enumerate(range(hops - 1, -1, -1)):
            seed = None if random_seed is None else random_seed + depth
            fanout = max_neighbors[depth]
            if weight_name is None:
                offsets, nbr_gids, src_lids = self.unweighted_sample_without_replacement_one_hop(
                    target_gids[layer + 1], fanout, random_seed=seed, need_center_local_output=True)
            else:
                offsets, nbr_gids, src_lids = self.weighted_sample_without_replacement_one_hop(
                    weight_name, target_gids[layer + 1], fanout, random_seed=seed, need_center_local_output=True)
            unique_gids, raw_to_unique = graph_ops.append_unique(
                target_gids[layer + 1], nbr_gids, need_neighbor_raw_to_unique=True)
            csr_row_ptr[layer] = offsets
            csr_col_ind[layer] = raw_to_unique
            edge_indice[layer] = torch.stack([raw_to_unique, src_lids])
            target_gids[layer] = unique_gids
        return target_gids, edge_indice, csr_row_ptr, csr_col_ind
