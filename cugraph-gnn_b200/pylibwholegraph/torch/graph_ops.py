"""Graph helper ops (role of the reference's pylibwholegraph/torch/graph_ops.py)."""
import pylibwholegraph.binding.wholememory_binding as wmb
from pylibwholegraph.utils.imports import import_optional
from .wholegraph_env import get_stream, TorchMemoryContext, get_wholegraph_env_fns, wrap_torch_tensor

torch = import_optional("torch")


def append_unique(target_node_tensor: "torch.Tensor", neighbor_node_tensor: "torch.Tensor",
                  need_neighbor_raw_to_unique: bool = False):
    """unique = targets ++ (new neighbours in first-occurrence order).

    e.g. targets [3, 11, 2, 10], neighbours [4, 5, 2, 11, 6, 9, 10, 5]
      -> unique [3, 11, 2, 10, 4, 5, 6, 9], raw_to_unique [4, 5, 2, 1, 6, 7, 3, 5]
    (the reference leaves the order of the new ids unspecified; this build fixes it)."""
    assert target_node_tensor.dim() == 1 and neighbor_node_tensor.dim() == 1
    assert target_node_tensor.is_cuda and neighbor_node_tensor.is_cuda
    unique_ctx = TorchMemoryContext()
    mapping = None
    if need_neighbor_raw_to_unique:
        mapping = torch.empty(neighbor_node_tensor.shape[0], device="cuda", dtype=torch.int)
    wmb.append_unique(
        wrap_torch_tensor(target_node_tensor),
        wrap_torch_tensor(neighbor_node_tensor),
        unique_ctx.get_c_context(),
        wrap_torch_tensor(mapping),
        get_wholegraph_env_fns(),
        get_stream(),
    )
    if need_neighbor_raw_to_unique:
        return unique_ctx.get_tensor(), mapping
    return unique_ctx.get_tensor()
