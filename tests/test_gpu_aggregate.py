"""A1: CSR aggregation against the fp64-accumulating oracle (tolerance from BASELINE.json: 1e-3 relative) and
against a plain torch fp32 reference of the same op; C3-shaped 3-layer GraphSAGE forward on sampled blocks."""
import numpy as np
import pytest

from graphs import random_csr

pytestmark = pytest.mark.gpu
RTOL = 1e-3  # north_star: "within 1e-3 rel for fp32 aggregation output"


def _block(rng, n_dst, n_src, max_deg):
    deg = rng.integers(0, max_deg + 1, n_dst)
    deg[rng.random(n_dst) < 0.1] = 0
    indptr = np.concatenate([[0], np.cumsum(deg)]).astype(np.int64)
    indices = rng.integers(0, n_src, indptr[-1]).astype(np.int64)
    return indptr, indices


def _close(got, exp):
    scale = np.maximum(np.abs(exp), 1.0)
    return np.max(np.abs(got - exp) / scale) <= RTOL


@pytest.mark.parametrize("dim", [128, 256, 32, 4, 100, 520])
@pytest.mark.parametrize("reduce", ["mean", "sum"])
@pytest.mark.parametrize("ptr_dtype,idx_dtype", [(np.int64, np.int32), (np.int32, np.int64), (np.int64, np.int64)])
def test_aggregate_fp32_vs_oracle(oracle, dim, reduce, ptr_dtype, idx_dtype):
    import torch
    from pylibwholegraph.torch import csr_aggregate_forward

    rng = np.random.default_rng(dim)
    n_dst, n_src = 3001, 9000
    indptr, indices = _block(rng, n_dst, n_src, 25)
    x = rng.standard_normal((n_src, dim)).astype(np.float32)
    got = csr_aggregate_forward(torch.from_numpy(indptr.astype(ptr_dtype)).cuda(), torch.from_numpy(indices.astype(idx_dtype)).cuda(),
                                torch.from_numpy(x).cuda(), reduce).cpu().numpy()
    exp = oracle.csr_aggregate(indptr, indices, x, mean=(reduce == "mean"))
    assert got.shape == exp.shape and _close(got, exp)
    # torch fp32 reference of the same op
    ref = torch.zeros((n_dst, dim))
    rows = torch.repeat_interleave(torch.arange(n_dst), torch.from_numpy(np.diff(indptr)))
    ref.index_add_(0, rows, torch.from_numpy(x)[torch.from_numpy(indices)])
    if reduce == "mean":
        ref /= torch.from_numpy(np.maximum(np.diff(indptr), 1)).float()[:, None]
    assert torch.allclose(torch.from_numpy(got), ref, rtol=1e-4, atol=1e-4)


@pytest.mark.parametrize("dtype", ["float16", "bfloat16"])
def test_aggregate_half_inputs(oracle, dtype):
    import torch
    from pylibwholegraph.torch import csr_aggregate_forward

    rng = np.random.default_rng(1)
    indptr, indices = _block(rng, 2000, 5000, 15)
    xt = torch.from_numpy(rng.standard_normal((5000, 128)).astype(np.float32)).to(getattr(torch, dtype))
    got = csr_aggregate_forward(torch.from_numpy(indptr).cuda(), torch.from_numpy(indices).cuda(), xt.cuda(), "mean").cpu().numpy()
    exp = oracle.csr_aggregate(indptr, indices, xt.float().numpy(), mean=True)  # same rounded inputs, fp64 accumulate
    assert _close(got, exp)


def test_aggregate_fused_gather_from_feature_table_and_errors(oracle):
    import torch
    import pylibwholegraph.torch as wgth

    torch.cuda.set_device(0)
    wgth.init(0, 1, 0, 1)
    comm = wgth.get_global_communicator()
    rng = np.random.default_rng(3)
    nodes, dim = 50000, 128
    emb = wgth.create_embedding(comm, "chunked", "cuda", torch.float32, [nodes, dim])
    table = rng.standard_normal((nodes, dim)).astype(np.float32)
    emb.get_embedding_tensor().get_local_tensor()[0].copy_(torch.from_numpy(table).cuda())
    renumber_map = rng.permutation(nodes)[:8000].astype(np.int64)
    indptr, indices = _block(rng, 1500, 8000, 20)
    got = wgth.csr_aggregate_forward(torch.from_numpy(indptr).cuda(), torch.from_numpy(indices).cuda(), emb, "mean",
                                     gather_map=torch.from_numpy(renumber_map).cuda()).cpu().numpy()
    exp = oracle.csr_aggregate(indptr, indices, table[renumber_map], mean=True)
    assert _close(got, exp)
    with pytest.raises(RuntimeError):  # rows must be 16-byte aligned
        wgth.csr_aggregate_forward(torch.from_numpy(indptr).cuda(), torch.from_numpy(indices).cuda(), torch.zeros((8000, 3)).cuda())
    empty = wgth.csr_aggregate_forward(torch.zeros(1, dtype=torch.int64).cuda(), torch.zeros(0, dtype=torch.int64).cuda(), torch.zeros((4, 8)).cuda())
    assert empty.shape == (0, 8)


def test_aggregate_backward_matches_autograd():
    import torch
    from pylibwholegraph.torch import csr_aggregate

    rng = np.random.default_rng(5)
    n_dst, n_src, dim = 700, 2000, 64
    indptr, indices = _block(rng, n_dst, n_src, 12)
    ip, ix = torch.from_numpy(indptr).cuda(), torch.from_numpy(indices).cuda()
    x = torch.randn((n_src, dim), device="cuda", requires_grad=True)
    w = torch.randn((n_dst, dim), device="cuda")
    for reduce in ("mean", "sum"):
        x.grad = None
        (csr_aggregate(ip, ix, x, reduce) * w).sum().backward()
        g = x.grad.clone()
        x.grad = None
        rows = torch.repeat_interleave(torch.arange(n_dst, device="cuda"), ip.diff())
        ref = torch.zeros((n_dst, dim), device="cuda").index_add(0, rows, x[ix])
        if reduce == "mean":
            ref = ref / ip.diff().clamp(min=1)[:, None]
        (ref * w).sum().backward()
        assert torch.allclose(g, x.grad, rtol=1e-4, atol=1e-4)


def test_aggregate_backward_through_the_reversed_block_matches_the_scatter_path():
    """csr_transpose + gather-reduce backward == atomicAdd backward == autograd, incl. empty rows and repeated sources"""
    import torch
    from pylibwholegraph.torch import csr_aggregate
    from pylibwholegraph.torch.aggregate import csr_transpose

    rng = np.random.default_rng(9)
    n_dst, n_src, dim = 900, 2500, 128
    indptr, indices = _block(rng, n_dst, n_src, 20)
    indices[: len(indices) // 4] = rng.integers(0, 30, len(indices) // 4)  # hub sources
    ip, ix = torch.from_numpy(indptr).cuda(), torch.from_numpy(indices).cuda()
    tr = csr_transpose(ip, ix, n_src)
    assert tr[0].shape[0] == n_src + 1 and int(tr[0][-1]) == len(indices)
    x = torch.randn((n_src, dim), device="cuda", requires_grad=True)
    w = torch.randn((n_dst, dim), device="cuda")
    for reduce in ("mean", "sum"):
        grads = []
        for t in (None, tr):
            x.grad = None
            (csr_aggregate(ip, ix, x, reduce, t) * w).sum().backward()
            grads.append(x.grad.clone())
        assert torch.allclose(grads[0], grads[1], rtol=1e-4, atol=1e-4)
    # no gradient wanted for x: backward returns without touching anything
    y = csr_aggregate(ip, ix, x.detach(), "mean", tr)
    assert not y.requires_grad


def test_c3_three_layer_sage_on_sampled_blocks(oracle):
    """BASELINE config C3 in miniature: products-shaped degree skew, fan-out [15, 10, 5], 3 SAGE layers whose
    aggregation runs on the sampler's CSR output; checked layer by layer against an fp64 numpy model."""
    import torch
    import pylibwholegraph.torch as wgth

    torch.cuda.set_device(0)
    wgth.init(0, 1, 0, 1)
    comm = wgth.get_global_communicator()
    nodes, edges, dim, hidden = 20000, 600000, 128, 256
    row_ptr, col = random_csr(nodes, edges, seed=13)
    feat = np.random.default_rng(0).standard_normal((nodes, dim)).astype(np.float32)

    def wm(arr):
        t = torch.from_numpy(arr)
        w = wgth.create_wholememory_tensor(comm, "chunked", "cuda", [arr.shape[0]], t.dtype, [1])
        w.get_local_tensor()[0].copy_(t.cuda())
        return w

    wm_rp, wm_col = wm(row_ptr), wm(col)
    seeds = np.random.default_rng(1).permutation(nodes)[:256].astype(np.int64)
    fanout = [15, 10, 5]
    L = 3
    res = wgth.MultiHopSampler().sample(wm_rp, wm_col, torch.from_numpy(seeds).cuda(), torch.tensor([0, 256]).cuda(), fanout, 62, compression="CSR")
    mo, minors, lho = res["major_offsets"], res["minors"], res["label_hop_offsets"].cpu().numpy()
    n_id = res["renumber_map"]
    x = torch.from_numpy(feat).cuda()[n_id]
    torch.manual_seed(0)
    layers = [wgth.SAGEConv(dim, hidden).cuda(), wgth.SAGEConv(hidden, hidden).cuda(), wgth.SAGEConv(hidden, 47).cuda()]
    xs_ref = x.double().cpu().numpy()
    h = x
    mo_np, mn_np = mo.cpu().numpy(), minors.cpu().numpy()
    for k, layer in enumerate(layers):
        # layer k aggregates the edges of hops < L - k: destination rows are the sources of those hops
        n_rows = int(lho[L - k])
        indptr = mo[: n_rows + 1]
        indices = minors[: int(mo_np[n_rows])]
        h = layer(h, indptr, indices)
        # fp64 reference
        agg = oracle.csr_aggregate(mo_np[: n_rows + 1], mn_np[: mo_np[n_rows]], xs_ref.astype(np.float32) if k == 0 else xs_ref, mean=True) if k == 0 else None
        if agg is None:
            agg = np.zeros((n_rows, xs_ref.shape[1]))
            for i in range(n_rows):
                s, e = mo_np[i], mo_np[i + 1]
                if e > s:
                    agg[i] = xs_ref[mn_np[s:e]].mean(0)
        wl, bl, wr = (p.detach().double().cpu().numpy() for p in (layer.lin_l.weight, layer.lin_l.bias, layer.lin_r.weight))
        xs_ref = agg @ wl.T + bl + xs_ref[:n_rows] @ wr.T
        if k < L - 1:
            h = torch.relu(h)
            xs_ref = np.maximum(xs_ref, 0)
        got = h.detach().cpu().numpy()
        assert got.shape == xs_ref.shape
        assert np.max(np.abs(got - xs_ref) / np.maximum(np.abs(xs_ref), 1.0)) < 5e-3  # dense layers run in fp32/TF32-free torch
    assert h.shape == (256, 47)
