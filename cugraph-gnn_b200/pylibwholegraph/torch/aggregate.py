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


def csr_transpose(indptr, indices, n_src):
    """(indptr_T [n_src + 1], indices_T) of the block with the edges reversed (rows = sources, entries = destination rows),
    int64 / int32: lets the backward pass of the aggregation run as a gather-reduce (the forward kernel, no atomics).
    A few torch ops per block (host glue, built once per mini-batch block and shared by every layer)."""
    n_dst = indptr.shape[0] - 1
    deg = (indptr[1:] - indptr[:-1]).to(torch.int64)
    dst = torch.repeat_interleave(torch.arange(n_dst, device=indptr.device), deg)
    src = indices.to(torch.int64)
    order = torch.argsort(src, stable=True)
    indptr_t = torch.zeros(n_src + 1, dtype=torch.int64, device=indptr.device)
    indptr_t[1:] = torch.bincount(src, minlength=n_src).cumsum(0)
    return indptr_t, dst[order].to(torch.int32).contiguous()


class _CsrAggregate(torch.autograd.Function):
    @staticmethod
    def forward(ctx, indptr, indices, x, reduce, transposed):
        ctx.save_for_backward(indptr, indices, *(transposed if transposed is not None else ()))
        ctx.has_t = transposed is not None
        ctx.reduce = reduce
        ctx.x_shape = x.shape
        ctx.x_dtype = x.dtype
        return csr_aggregate_forward(indptr, indices, x, reduce)

    @staticmethod
    def backward(ctx, grad_out):
        if not ctx.needs_input_grad[2]:  # first layer: the gathered features carry no gradient
            return None, None, None, None, None
        saved = ctx.saved_tensors
        indptr, indices = saved[0], saved[1]
        grad_out = grad_out.contiguous().float()
        if ctx.has_t:
            # gather-reduce over the reversed block: grad_x[s] = sum over edges (s -> d) of grad_out[d] (/ deg(d) for mean)
            indptr_t, indices_t = saved[2], saved[3]
            if _REDUCE[ctx.reduce] == 1:
                deg = (indptr[1:] - indptr[:-1]).clamp(min=1).to(torch.float32)
                grad_out = grad_out / deg[:, None]
            if indptr_t.shape[0] - 1 != ctx.x_shape[0]:
                raise ValueError("csr_transpose was built for %d source rows, x has %d" % (indptr_t.shape[0] - 1, ctx.x_shape[0]))
            grad_x = csr_aggregate_forward(indptr_t, indices_t, grad_out, "sum")
            return None, None, grad_x.to(ctx.x_dtype), None, None
        grad_x = torch.zeros(ctx.x_shape, device=grad_out.device, dtype=torch.float32)
        hs = [_h(indptr), _h(indices), _h(grad_out), _h(grad_x)]
        err = _bwd(hs[0][0], hs[1][0], hs[2][0], _REDUCE[ctx.reduce], hs[3][0], get_stream())
        wmb.check_wholememory_error_code(err)
        return None, None, grad_x.to(ctx.x_dtype), None, None


def csr_aggregate(indptr, indices, x, reduce="mean", transposed=None):
    """Differentiable (w.r.t. x) aggregation over a sampled CSR block.  transposed = csr_transpose(indptr, indices,
    x.shape[0]): the backward pass is then a second gather-reduce instead of an atomicAdd scatter."""
    return _CsrAggregate.apply(indptr, indices, x, reduce, transposed)


class SAGEConv(torch.nn.Module):
    """GraphSAGE layer on the sampler's CSR block: W_l * mean_{j in N(i)} x_j + W_r * x_i  (PyG SAGEConv maths,
    torch_geometric.nn.SAGEConv with aggr='mean', root_weight=True); destination rows are the first n_dst rows of x."""

    def __init__(self, in_channels, out_channels, aggr="mean", bias=True):
        super().__init__()
        self.aggr = aggr
        self.lin_l = torch.nn.Linear(in_channels, out_channels, bias=bias)
        self.lin_r = torch.nn.Linear(in_channels, out_channels, bias=False)

    def forward(self, x, indptr, indices, transposed=None):
        n_dst = indptr.shape[0] - 1
        agg = csr_aggregate(indptr, indices, x, self.aggr, transposed)
        return self.lin_l(agg.to(x.dtype)) + self.lin_r(x[:n_dst])

    def forward_fused(self, x, indptr, indices):
        """Inference form of forward() for bf16 rows of width 128 with mean aggregation: ONE kernel (csrc/sage_tile.cu) --
        warp-per-row gather-mean written straight into the shared-memory operand tile, [mean || self] . [W_l || W_r]^T on the
        tensor cores (tcgen05.mma, accumulator in tensor memory), bias added in the epilogue.  fp32 result [n_dst, out]."""
        if self.aggr != "mean":
            raise ValueError("the fused layer implements mean aggregation")
        w_cat = torch.cat([self.lin_l.weight, self.lin_r.weight], dim=1).detach().to(torch.bfloat16).contiguous()
        bias = self.lin_l.bias.detach().float() if self.lin_l.bias is not None else None
        return sage_layer_forward(indptr, indices, x, w_cat, bias)


_sage = wmb.native_symbol("wholegraph_sage_layer_forward")
_sage.restype = ctypes.c_int
_sage.argtypes = [_vp, _vp, _vp, _vp, _vp, _vp, _vp]


def sage_layer_forward(indptr, indices, x, w_cat, bias=None):
    """out[i] = [mean_{e in row i} x[indices[e]] || x[i]] . w_cat^T (+ bias), fp32 [n_dst, F_out]  (wholegraph_sage_layer_forward,
    include/wholememory/b200_ops.h).  x: bf16 [n_src, 128], destination rows first; w_cat: bf16 [F_out, 256] = [W_l || W_r];
    bias: fp32 [F_out] or None.  F_out is padded to a multiple of 16 here (the MMA's N granularity) and sliced back."""
    if x.dtype != torch.bfloat16 or w_cat.dtype != torch.bfloat16:
        raise TypeError("sage_layer_forward takes bf16 feature rows and bf16 weights (fp32 accumulation and output)")
    f_out = w_cat.shape[0]
    pad = (-f_out) % 16
    if pad:
        w_cat = torch.cat([w_cat, w_cat.new_zeros((pad, w_cat.shape[1]))], dim=0)
        if bias is not None:
            bias = torch.cat([bias, bias.new_zeros(pad)])
    w_cat = w_cat.contiguous()
    n_dst = indptr.shape[0] - 1
    out = torch.empty((n_dst, f_out + pad), device=x.device, dtype=torch.float32)
    hs = [_h(indptr), _h(indices), _h(x), _h(w_cat), _h(bias.contiguous() if bias is not None else None), _h(out)]
    err = _sage(hs[0][0], hs[1][0], hs[2][0], hs[3][0], hs[4][0], hs[5][0], get_stream())
    wmb.check_wholememory_error_code(err)
    return out[:, :f_out] if pad else out
