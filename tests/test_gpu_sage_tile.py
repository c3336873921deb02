"""A1, fused form (csrc/sage_tile.cu): one GraphSAGE layer with the dense part on the tensor cores (tcgen05.mma, W by TMA,
accumulator in tensor memory) against an fp64 restatement of the reference consumer's maths
(pylibwholegraph/torch/gnn_model.py:119-125 -> PyG SAGEConv, aggr="mean", root weight):
    out[i] = W_l . mean_{j in N(i)} x_j + W_r . x_i + b
Tolerance: BASELINE.json north_star, 1e-3 relative on the fp32 output (inputs are the same bf16-rounded values on both sides)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
RTOL = 1e-3


def _block(rng, n_dst, n_src, max_deg, empty_frac=0.1):
    deg = rng.integers(0, max_deg + 1, n_dst)
    deg[rng.random(n_dst) < empty_frac] = 0
    indptr = np.concatenate([[0], np.cumsum(deg)]).astype(np.int64)
    indices = rng.integers(0, n_src, indptr[-1]).astype(np.int64)
    return indptr, indices


def _expected(oracle, indptr, indices, x, w_cat, bias):
    """fp64: the oracle's aggregation (fp64 accumulate) then the dense part in fp64 numpy."""
    n_dst = indptr.shape[0] - 1
    mean = oracle.csr_aggregate(indptr, indices, x, mean=True).astype(np.float64)
    # the oracle returns fp32; redo the mean in fp64 where it matters (same values to ~1e-7, far inside the tolerance)
    w = w_cat.astype(np.float64)
    f_in = x.shape[1]
    out = mean @ w[:, :f_in].T + x[:n_dst].astype(np.float64) @ w[:, f_in:].T
    if bias is not None:
        out += bias.astype(np.float64)[None, :]
    return out


def _run(oracle, n_dst, n_src, max_deg, f_out, ptr_dtype=np.int64, idx_dtype=np.int64, bias=True, seed=0, scale=1.0):
    import torch
    from pylibwholegraph.torch import sage_layer_forward

    rng = np.random.default_rng(seed)
    indptr, indices = _block(rng, n_dst, n_src, max_deg)
    xt = torch.from_numpy((scale * rng.standard_normal((n_src, 128))).astype(np.float32)).to(torch.bfloat16)
    wt = torch.from_numpy((rng.standard_normal((f_out, 256)) / 16).astype(np.float32)).to(torch.bfloat16)
    b = rng.standard_normal(f_out).astype(np.float32) if bias else None
    got = sage_layer_forward(torch.from_numpy(indptr.astype(ptr_dtype)).cuda(), torch.from_numpy(indices.astype(idx_dtype)).cuda(),
                             xt.cuda(), wt.cuda(), torch.from_numpy(b).cuda() if bias else None)
    torch.cuda.synchronize()
    got = got.cpu().numpy().astype(np.float64)
    exp = _expected(oracle, indptr, indices, xt.float().numpy(), wt.float().numpy(), b)
    assert got.shape == exp.shape == (n_dst, f_out)
    err = np.max(np.abs(got - exp) / np.maximum(np.abs(exp), 1.0))
    assert err <= RTOL, "max relative error %.3e" % err
    return err


@pytest.mark.parametrize("n_dst", [1, 127, 128, 129, 3001])
@pytest.mark.parametrize("f_out", [16, 48, 128, 256])
def test_sage_tile_vs_fp64(oracle, n_dst, f_out):
    _run(oracle, n_dst, max(n_dst, 9000), 25, f_out, seed=n_dst + f_out)


@pytest.mark.parametrize("ptr_dtype,idx_dtype", [(np.int64, np.int32), (np.int32, np.int64), (np.int32, np.int32)])
def test_sage_tile_index_types(oracle, ptr_dtype, idx_dtype):
    _run(oracle, 2000, 6000, 15, 64, ptr_dtype, idx_dtype, seed=5)


def test_sage_tile_many_tiles_per_cta_no_bias_padded_width(oracle):
    # more 128-row tiles than SMs: the persistent CTAs loop (accumulator and operand tile reused, mbarrier phase flips);
    # F_out = 47 (the class count of the products / papers100M shapes) is padded to 48 inside the wrapper
    _run(oracle, 148 * 128 * 2 + 77, 60000, 10, 47, bias=False, seed=9)


def test_sage_tile_large_values_keep_fp32_mean(oracle):
    # feature values with a large common offset: a bf16-rounded mean alone would lose ~3 decimal digits here; the hi + lo
    # split of the aggregated operand keeps the result inside the tolerance
    import torch
    from pylibwholegraph.torch import sage_layer_forward

    rng = np.random.default_rng(11)
    indptr, indices = _block(rng, 1000, 4000, 25, empty_frac=0.0)
    x = (100.0 + rng.standard_normal((4000, 128))).astype(np.float32)
    xt = torch.from_numpy(x).to(torch.bfloat16)
    wt = torch.from_numpy((rng.standard_normal((32, 256)) / 16).astype(np.float32)).to(torch.bfloat16)
    got = sage_layer_forward(torch.from_numpy(indptr).cuda(), torch.from_numpy(indices).cuda(), xt.cuda(), wt.cuda()).cpu().numpy()
    exp = _expected(oracle, indptr, indices, xt.float().numpy(), wt.float().numpy(), None)
    err = np.max(np.abs(got - exp) / np.maximum(np.abs(exp), 1.0))
    assert err <= RTOL, "max relative error %.3e" % err


def test_sage_tile_matches_unfused_layer_and_rejects_other_shapes():
    import torch
    import pylibwholegraph.torch as wgth

    rng = np.random.default_rng(2)
    indptr, indices = _block(rng, 1500, 5000, 20)
    ip, ix = torch.from_numpy(indptr).cuda(), torch.from_numpy(indices).cuda()
    layer = wgth.SAGEConv(128, 64).cuda()
    x = torch.randn((5000, 128), device="cuda").to(torch.bfloat16)
    fused = layer.forward_fused(x, ip, ix)
    with torch.no_grad():
        # the unfused path (aggregation kernel + two GEMMs) on the same bf16-rounded operands, fp32 arithmetic
        layer.lin_l.weight.copy_(layer.lin_l.weight.to(torch.bfloat16).float())
        layer.lin_r.weight.copy_(layer.lin_r.weight.to(torch.bfloat16).float())
        plain = layer(x.float(), ip, ix)
    assert torch.allclose(fused, plain, rtol=2e-3, atol=2e-3)
    with pytest.raises(TypeError):
        wgth.sage_layer_forward(ip, ix, x.float(), torch.zeros((64, 256), device="cuda", dtype=torch.bfloat16))
    with pytest.raises(Exception):  # F_in = 64: NOT_IMPLEMENTED
        wgth.sage_layer_forward(ip, ix, x[:, :64].contiguous(), torch.zeros((64, 128), device="cuda", dtype=torch.bfloat16))
    # empty block
    out = wgth.sage_layer_forward(torch.zeros(1, dtype=torch.int64, device="cuda"), ix[:0], x, torch.zeros((16, 256), device="cuda", dtype=torch.bfloat16))
    assert out.shape == (0, 16)
