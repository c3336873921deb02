"""Functional gather / scatter (role of the reference's pylibwholegraph/torch/wholememory_ops.py)."""
import pylibwholegraph.binding.wholememory_binding as wmb
from pylibwholegraph.utils.imports import import_optional
from .wholegraph_env import get_stream, get_wholegraph_env_fns, wrap_torch_tensor
from .utils import wholememory_dtype_to_torch_dtype

torch = import_optional("torch")


def _check_indices(indices_tensor):
    assert indices_tensor.dim() == 1
    assert indices_tensor.dtype in (torch.int32, torch.int64)


def wholememory_gather_forward_functor(wholememory_tensor: wmb.PyWholeMemoryTensor, indices_tensor: "torch.Tensor",
                                       requires_grad=False, torch_output_dtype=None):
    """Gather rows of a PyWholeMemoryTensor into a fresh CUDA tensor."""
    _check_indices(indices_tensor)
    one_d = wholememory_tensor.dim() == 1
    out = torch.empty(
        (indices_tensor.shape[0], 1 if one_d else wholememory_tensor.shape[1]),
        device="cuda",
        dtype=torch_output_dtype or wholememory_dtype_to_torch_dtype(wholememory_tensor.dtype),
        requires_grad=requires_grad,
    )
    wmb.wholememory_gather_op(
        wholememory_tensor, wrap_torch_tensor(indices_tensor), wrap_torch_tensor(out), get_wholegraph_env_fns(), get_stream()
    )
    return out.view(-1) if one_d else out


def wholememory_scatter_functor(input_tensor: "torch.Tensor", indices_tensor: "torch.Tensor",
                                wholememory_tensor: wmb.PyWholeMemoryTensor):
    """Scatter rows of input_tensor into a PyWholeMemoryTensor."""
    _check_indices(indices_tensor)
    wmb.wholememory_scatter_op(
        wrap_torch_tensor(input_tensor), wrap_torch_tensor(indices_tensor), wholememory_tensor, get_wholegraph_env_fns(), get_stream()
    )
