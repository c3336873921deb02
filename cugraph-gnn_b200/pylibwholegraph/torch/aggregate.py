"""CSR neighbourhood aggregation of the sampled block (B200 extension, include/wholememory/b200_ops.h).

Consumes the sampler's CSR directly (major_offsets / minors) -- the reference expands it back to COO and
lets torch_geometric scatter-add (cugraph_pyg/sampler/sampler.py:55-65)."""
import ctypes

import pylibwholegraph.binding.wholememory_binding as wmb
from pylibwholegraph.utils.imports import import_optional
from .wholegraph_env import get_stream, wrap_torch_tensor

torch = import_optional("torch")

_vp = ctypes.c_void_p
_fwd = wmb.native_symbol("wholegraph_csr_aggregate")
_fwd.restype = ctypes.c_int
_fwd.argtypes = [_vp, _vp, _vp, _vp, ctypes.c_int, _vp, _vp]
_bwd = wmb.native_symbol("wholegraph_csr_aggregate_backward")
_bwd.restype = ctypes.c_int
_bwd.argtypes = [_vp, _vp, _vp, ctypes.c_int, _vp, _vp]

_REDUCE = {"sum": 0, "add": 0, "mean": 1}


def _h(t):
    if t is None:
        return None, None
    if hasattr(t, "wmb_tensor"):
        return t.wmb_tensor.get_c_handle(), t
    if hasattr(t, "get_c_handle"):
        return t.get_c_handle(), t
    w = wrap_torch_tensor(t)
    return w.get_c_handle(), w


def csr_aggregate_forward(indptr, indices, x, reduce="mean", gather_map=None):
    """out[i] = reduce over e in [indptr[i], indptr[i+1]) of x[map[indices[e]]] (fp32 out).

    x: CUDA tensor [n, F] (fp32/fp16/bf16) or a WholeMemoryTensor / WholeMemoryEmbedding table together with
    gather_map (int64 global ids, i.e. the renumber map) -- the feature gather is then fused in."""
    if hasattr(x, "get_embedding_tensor"):
        x = x.get_embedding_tensor()
    feat = x.shape[1]
    out = torch.empty((indptr.shape[0] - 1, feat), device=indptr.device, dtype=torch.float32)
    handles = [_h(indptr), _h(indices), _h(gather_map), _h(x)]
    ho = _h(out)
    err = _fwd(handles[0][0], handles[1][0], handles[2][0], handles[3][0], _REDUCE[reduce], ho[0], get_stream())
    wmb.check_wholememory_error_code(err)
    return out


class _CsrAggregate(torch.autograd.Function):
    @staticmethod
    def forward(ctx, indptr, indices, x, reduce):
        ctx.save_for_backward(indptr, indices)
        ctx.reduce = reduce
        ctx.x_shape = x.shape
        ctx.x_dtype = x.dtype
        return csr_aggregate_forward(indptr, indices, x, reduce)

    @staticmethod
    def backward(ctx, grad_out):
        indptr, indices = ctx.saved_tensors
        grad_out = grad_out.contiguous().float()
        grad_x = torch.zeros(ctx.x_shape, device=grad_out.device, dtype=torch.float32)
        hs = [_h(indptr), _h(indices), _h(grad_out), _h(grad_x)]
        err = _bwd(hs[0][0], hs[1][0], hs[2][0], _REDUCE[ctx.reduce], hs[3][0], get_stream())
        wmb.check_wholememory_error_code(err)
        return None, None, grad_x.to(ctx.x_dtype), None


def csr_aggregate(indptr, indices, x, reduce="mean"):
    """Differentiable (w.r.t. x) aggregation over a sampled CSR block."""
    return _CsrAggregate.apply(indptr, indices, x, reduce)


class SAGEConv(torch.nn.Module):
    """GraphSAGE layer on the sampler's CSR block: W_l * mean_{j in N(i)} x_j + W_r * x_i  (PyG SAGEConv maths,
    torch_geometric.nn.SAGEConv with aggr='mean', root_weight=True); destination rows are the first n_dst rows of x."""

    def __init__(self, in_channels, out_channels, aggr="mean", bias=True):
        super().__init__()
        self.aggr = aggr
        self.lin_l = torch.nn.Linear(in_channels, out_channels, bias=bias)
        self.lin_r = torch.nn.Linear(in_channels, out_channels, bias=False)

    def forward(self, x, indptr, indices):
        n_dst = indptr.shape[0] - 1
        agg = csr_aggregate(indptr, indices, x, self.aggr)
        return self.lin_l(agg.to(x.dtype)) + self.lin_r(x[:n_dst])
