"""One-hop samplers over a CSR graph in WholeMemory (role of the reference's
pylibwholegraph/torch/wholegraph_ops.py; same function names, arguments and return tuples)."""
import random
from typing import Union

import pylibwholegraph.binding.wholememory_binding as wmb
from pylibwholegraph.utils.imports import import_optional
from .wholegraph_env import get_stream, TorchMemoryContext, get_wholegraph_env_fns, wrap_torch_tensor

torch = import_optional("torch")


def _one_hop(native_call, graph_tensors, center_nodes_tensor, max_sample_count, random_seed,
             need_center_local_output, need_edge_output):
    for t in graph_tensors:
        assert t.dim() == 1
    assert center_nodes_tensor.dim() == 1
    if random_seed is None:
        random_seed = random.getrandbits(64)
    offsets = torch.empty(center_nodes_tensor.shape[0] + 1, device="cuda", dtype=torch.int)
    dest_ctx = TorchMemoryContext()
    lid_ctx = TorchMemoryContext() if need_center_local_output else None
    gid_ctx = TorchMemoryContext() if need_edge_output else None
    native_call(
        *graph_tensors,
        wrap_torch_tensor(center_nodes_tensor),
        max_sample_count,
        wrap_torch_tensor(offsets),
        dest_ctx.get_c_context(),
        lid_ctx.get_c_context() if lid_ctx else 0,
        gid_ctx.get_c_context() if gid_ctx else 0,
        random_seed,
        get_wholegraph_env_fns(),
        get_stream(),
    )
    result = [offsets, dest_ctx.get_tensor()]
    if lid_ctx:
        result.append(lid_ctx.get_tensor())
    if gid_ctx:
        result.append(gid_ctx.get_tensor())
    return tuple(result)


def unweighted_sample_without_replacement(
    wm_csr_row_ptr_tensor: wmb.PyWholeMemoryTensor,
    wm_csr_col_ptr_tensor: wmb.PyWholeMemoryTensor,
    center_nodes_tensor: "torch.Tensor",
    max_sample_count: int,
    random_seed: Union[int, None] = None,
    need_center_local_output: bool = False,
    need_edge_output: bool = False,
):
    """Uniform neighbour sampling without replacement.

    Returns (sample_offset int32[n+1], dest[, center_localid int32][, edge_gid int64])."""
    return _one_hop(
        wmb.csr_unweighted_sample_without_replacement,
        (wm_csr_row_ptr_tensor, wm_csr_col_ptr_tensor),
        center_nodes_tensor, max_sample_count, random_seed, need_center_local_output, need_edge_output,
    )


def weighted_sample_without_replacement(
    wm_csr_row_ptr_tensor: wmb.PyWholeMemoryTensor,
    wm_csr_col_ptr_tensor: wmb.PyWholeMemoryTensor,
    wm_csr_weight_ptr_tensor: wmb.PyWholeMemoryTensor,
    center_nodes_tensor: "torch.Tensor",
    max_sample_count: int,
    random_seed: Union[int, None] = None,
    need_center_local_output: bool = False,
    need_edge_output: bool = False,
):
    """Weight-biased (A-Res) neighbour sampling without replacement; same returns as the uniform op."""
    assert wm_csr_weight_ptr_tensor.shape[0] == wm_csr_col_ptr_tensor.shape[0]
    return _one_hop(
        wmb.csr_weighted_sample_without_replacement,
        (wm_csr_row_ptr_tensor, wm_csr_col_ptr_tensor, wm_csr_weight_ptr_tensor),
        center_nodes_tensor, max_sample_count, random_seed, need_center_local_output, need_edge_output,
    )


def generate_random_positive_int_cpu(random_seed, sub_sequence, output_random_value_count):
    out = torch.empty((output_random_value_count,), dtype=torch.int)
    wmb.host_generate_random_positive_int(random_seed, sub_sequence, wrap_torch_tensor(out))
    return out


def generate_exponential_distribution_negative_float_cpu(random_seed: int, sub_sequence: int, output_random_value_count: int):
    out = torch.empty((output_random_value_count,), dtype=torch.float)
    wmb.host_generate_exponential_distribution_negative_float(random_seed, sub_sequence, wrap_torch_tensor(out))
    return out
