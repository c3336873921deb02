"""GPU parity tests: the CUDA path, called through the C ABI (ctypes -> libwholegraph_b200.so),
against the CPU oracle on the same seeded inputs.  Integer / index results must be bit-exact.

Test matrices mirror the reference's own:
  gather   cpp/tests/wholememory_ops/wholememory_gather_tests.cu:277-450
  sampler  cpp/tests/wholegraph_ops/wholegraph_csr_unweighted_sample_without_replacement_tests.cu:382-403
  weighted cpp/tests/wholegraph_ops/wholegraph_csr_weighted_sample_without_replacement_tests.cu:429-455
  unique   cpp/tests/graph_ops/append_unique_tests.cu:219-231
"""
import numpy as np
import pytest

from graphs import karate_csr, random_csr

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def wg():
    import torch
    import pylibwholegraph.torch as wgth

    assert torch.cuda.is_available()
    torch.cuda.set_device(0)
    wgth.init(0, 1, 0, 1)
    comm = wgth.get_global_communicator()
    return wgth, comm


def _wm_from_numpy(wgth, comm, arr, memory_type="chunked", stride=None):
    """WholeMemory tensor holding `arr` (1-D or 2-D numpy), optional padded row stride."""
    import torch

    t = torch.from_numpy(np.ascontiguousarray(arr))
    if arr.ndim == 2:
        strides = [stride or arr.shape[1], 1]
    else:
        strides = [1]
    wm = wgth.create_wholememory_tensor(comm, memory_type, "cuda", list(arr.shape), t.dtype, strides)
    local, start = wm.get_local_tensor()
    assert start == 0 and local.shape[0] == arr.shape[0]
    local.copy_(t.cuda())
    return wm


def _torch_dtype(name):
    import torch

    return {"f32": torch.float32, "f16": torch.float16, "bf16": torch.bfloat16, "f64": torch.float64,
            "i32": torch.int32, "i64": torch.int64, "i16": torch.int16, "i8": torch.int8}[name]


# ------------------------------------------------------------------------------------------------
# G1: gather
# ------------------------------------------------------------------------------------------------
_DIMS = [(32, 32), (32, 33), (127, 127), (128, 128), (129, 129), (513, 513), (256, 256), (1, 1)]
# the full width / stride sweep runs on the chunked type; the other two memory types take three shapes
_GATHER_SHAPES = [("chunked", d, s) for d, s in _DIMS] + [(m, d, s) for m in ("continuous", "distributed") for d, s in ((32, 32), (128, 128), (32, 33))]


@pytest.mark.parametrize("memory_type,dim,stride", _GATHER_SHAPES)
@pytest.mark.parametrize("emb,out", [("f32", "f32"), ("f16", "f16"), ("f32", "f16"), ("f16", "f32")])
@pytest.mark.parametrize("idx", ["i32", "i64"])
def test_gather_matrix(wg, oracle, memory_type, dim, stride, emb, out, idx):
    import torch

    wgth, comm = wg
    rows, n = 20011, 5003
    rng = np.random.default_rng(dim * 7 + stride)
    # closed form à la embedding_test_utils.cu:186-226 / test_wholegraph_gather_scatter.py:16-31
    table = ((np.arange(rows)[:, None] * 3 + np.arange(dim)[None, :]) % 2048).astype(np.float32)
    emb_t, out_t = _torch_dtype(emb), _torch_dtype(out)
    wm = wgth.create_wholememory_tensor(comm, memory_type, "cuda", [rows, dim], emb_t, [stride, 1])
    local, _ = wm.get_local_tensor()
    local.copy_(torch.from_numpy(table).to(emb_t).cuda())
    indices = rng.integers(0, rows, n)
    indices[rng.random(n) < 0.05] = -1  # skipped rows
    idx_t = torch.from_numpy(indices.astype(np.int32 if idx == "i32" else np.int64)).cuda()
    got = wm.gather(idx_t, force_dtype=out_t)
    exp = oracle.gather(table.astype(np.float16 if emb == "f16" else np.float32), indices.astype(np.int64),
                        out_dtype=np.float16 if out == "f16" else np.float32)
    keep = indices >= 0
    got_np = got.cpu().numpy()
    assert got_np.shape == (n, dim)
    assert np.array_equal(got_np[keep], exp[keep])
    wgth.destroy_wholememory_tensor(wm)


def test_gather_output_stride_negative_rows_untouched_and_1d(wg, oracle):
    import torch
    import pylibwholegraph.binding.wholememory_binding as wmb
    from pylibwholegraph.torch.wholegraph_env import get_stream, get_wholegraph_env_fns, wrap_torch_tensor

    wgth, comm = wg
    rows, dim = 4099, 32
    table = np.random.default_rng(0).standard_normal((rows, dim)).astype(np.float32)
    wm = _wm_from_numpy(wgth, comm, table)
    idx = np.random.default_rng(1).integers(0, rows, 1000)
    idx[::7] = -1
    out = torch.full((1000, 33), 7.0, device="cuda")  # output_stride = 33
    view = out[:, :32]
    wmb.wholememory_gather_op(wm.wmb_tensor, wrap_torch_tensor(torch.from_numpy(idx).cuda()), wrap_torch_tensor(view),
                              get_wholegraph_env_fns(), get_stream())
    got = out.cpu().numpy()
    exp = oracle.gather(table, idx)
    keep = idx >= 0
    assert np.array_equal(got[keep, :32], exp[keep])
    assert (got[~keep] == 7.0).all() and (got[:, 32] == 7.0).all()
    # 1-D table (reference: gather_op.cpp:33-38, returns .view(-1))
    vec = np.arange(5000, dtype=np.int64) * 3
    wm1 = _wm_from_numpy(wgth, comm, vec)
    g1 = wm1.gather(torch.from_numpy(idx[keep]).cuda())
    assert g1.dim() == 1 and np.array_equal(g1.cpu().numpy(), vec[idx[keep]])
    # column sub-tensor of a 2-D table
    sub = wm.get_sub_tensor([0, 8], [-1, 24])
    gs = sub.gather(torch.from_numpy(idx[keep]).cuda())
    assert np.array_equal(gs.cpu().numpy(), table[idx[keep], 8:24])


@pytest.mark.parametrize("emb,out", [("i32", "i64"), ("i64", "i32"), ("i8", "i8"), ("i16", "i32"), ("f64", "f32"), ("bf16", "f32"), ("f32", "bf16"), ("bf16", "bf16")])
def test_gather_other_dtypes(wg, emb, out):
    import torch

    wgth, comm = wg
    rows, dim, n = 3001, 24, 777
    emb_t, out_t = _torch_dtype(emb), _torch_dtype(out)
    src = (torch.arange(rows * dim).reshape(rows, dim) % 97 - 40).to(emb_t)
    wm = wgth.create_wholememory_tensor(comm, "chunked", "cuda", [rows, dim], emb_t, [dim, 1])
    wm.get_local_tensor()[0].copy_(src.cuda())
    idx = torch.randint(0, rows, (n,), generator=torch.Generator().manual_seed(0))
    got = wm.gather(idx.cuda(), force_dtype=out_t).cpu()
    # conversion spec: half/bf16 through float, everything else static_cast (gather_scatter_func.cuh:150-197)
    exp = src[idx].to(torch.float32 if emb in ("bf16", "f16") else emb_t).to(out_t) if emb_t.is_floating_point else src[idx].to(out_t)
    assert torch.equal(got, exp)
    wgth.destroy_wholememory_tensor(wm)


def test_gather_rejects_mixed_number_classes(wg):
    import torch

    wgth, comm = wg
    wm = wgth.create_wholememory_tensor(comm, "chunked", "cuda", [16, 4], torch.float32, [4, 1])
    with pytest.raises(RuntimeError):  # LOGIC_ERROR, as gather_func.cu:62-66
        wm.gather(torch.zeros(3, dtype=torch.int64, device="cuda"), force_dtype=torch.int32)
    with pytest.raises(ValueError):  # indices must be 1-D int32/int64
        import pylibwholegraph.binding.wholememory_binding as wmb
        from pylibwholegraph.torch.wholegraph_env import get_stream, get_wholegraph_env_fns, wrap_torch_tensor

        wmb.wholememory_gather_op(wm.wmb_tensor, wrap_torch_tensor(torch.zeros(3, device="cuda")),
                                  wrap_torch_tensor(torch.zeros((3, 4), device="cuda")), get_wholegraph_env_fns(), get_stream())
    assert wm.gather(torch.zeros(0, dtype=torch.int64, device="cuda")).shape == (0, 4)


# ------------------------------------------------------------------------------------------------
# G2: scatter
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("dim,stride", [(32, 33), (128, 128), (129, 129), (1, 1)])
@pytest.mark.parametrize("src,emb", [("f32", "f32"), ("f32", "f16"), ("f16", "f32")])
@pytest.mark.parametrize("pinned_host_input", [False, True])
def test_scatter_matrix(wg, oracle, dim, stride, src, emb, pinned_host_input):
    import torch

    wgth, comm = wg
    rows, n = 10007, 3001
    wm = wgth.create_wholememory_tensor(comm, "chunked", "cuda", [rows, dim], _torch_dtype(emb), [stride, 1])
    local, _ = wm.get_local_tensor()
    local.zero_()
    rng = np.random.default_rng(dim + stride)
    idx = rng.permutation(rows)[:n]  # unique rows: the reference scatter has no defined order for duplicates
    rows_np = (rng.integers(-1000, 1000, (n, dim))).astype(np.float16 if src == "f16" else np.float32)
    inp = torch.from_numpy(rows_np)
    inp = inp.pin_memory() if pinned_host_input else inp.cuda()  # DistTensor.__setitem__ passes pinned host rows
    wm.scatter(inp, torch.from_numpy(idx).cuda())
    torch.cuda.synchronize()
    table = np.zeros((rows, dim), dtype=np.float16 if emb == "f16" else np.float32)
    oracle.scatter(rows_np, idx, table)
    assert np.array_equal(local.cpu().numpy(), table)
    wgth.destroy_wholememory_tensor(wm)


def test_scatter_gather_roundtrip_large(wg):
    """size-independent property at a table larger than L2: gather(scatter(x)) == x."""
    import torch

    wgth, comm = wg
    rows, dim = 1_000_003, 128
    wm = wgth.create_wholememory_tensor(comm, "chunked", "cuda", [rows, dim], torch.float32, [dim, 1])
    g = torch.Generator(device="cuda").manual_seed(0)
    idx = torch.randperm(rows, device="cuda", generator=g)[:300_000]
    x = torch.randn((300_000, dim), device="cuda", generator=g)
    wm.scatter(x, idx)
    assert torch.equal(wm.gather(idx), x)
    wgth.destroy_wholememory_tensor(wm)


# ------------------------------------------------------------------------------------------------
# S1: uniform sampling
# ------------------------------------------------------------------------------------------------
def _graph_tensors(wgth, comm, row_ptr, col, memory_type="chunked"):
    return _wm_from_numpy(wgth, comm, row_ptr, memory_type), _wm_from_numpy(wgth, comm, col, memory_type)


@pytest.mark.parametrize("nodes,edges,seeds,M", [
    (9703, 104323, 512, 50), (9703, 104323, 512, 10), (9703, 104323, 512, 25), (9703, 104323, 512, 5),
    (23289, 689403, 35, 10), (9703, 104323, 512, 1), (9703, 104323, 512, 32), (9703, 104323, 512, 33),
    (9703, 104323, 300, 100), (9703, 104323, 100, 300), (9703, 104323, 64, 1024), (9703, 104323, 512, -1), (9703, 104323, 0, 10),
])
@pytest.mark.parametrize("center_dtype,col_dtype", [(np.int32, np.int32), (np.int64, np.int64), (np.int64, np.int32)])
def test_uniform_sampler_bit_exact(wg, oracle, nodes, edges, seeds, M, center_dtype, col_dtype):
    import torch
    from pylibwholegraph.torch import wholegraph_ops

    wgth, comm = wg
    row_ptr, col = random_csr(nodes, edges, seed=nodes + M, col_dtype=col_dtype)
    centers = np.random.default_rng(M + 2).integers(0, nodes, seeds).astype(center_dtype)
    wm_rp, wm_col = _graph_tensors(wgth, comm, row_ptr, col)
    seed = 0x9E3779B97F4A7C15 ^ (M & 0xFFFF)
    off, dest, lid, gid = wholegraph_ops.unweighted_sample_without_replacement(
        wm_rp.wmb_tensor, wm_col.wmb_tensor, torch.from_numpy(centers).cuda(), M, seed,
        need_center_local_output=True, need_edge_output=True)
    eoff, edest, elid, egid = oracle.unweighted_sample(row_ptr, col, centers, M, seed)
    assert off.dtype == torch.int32 and lid.dtype == torch.int32 and gid.dtype == torch.int64
    assert np.array_equal(off.cpu().numpy(), eoff)
    assert np.array_equal(gid.cpu().numpy(), egid)  # exact, in order
    assert np.array_equal(dest.cpu().numpy(), edest)
    assert np.array_equal(lid.cpu().numpy(), elid)
    for t in (wm_rp, wm_col):
        wgth.destroy_wholememory_tensor(t)


def test_uniform_sampler_karate_c1(wg, oracle):
    """BASELINE config C1: karate, 1 hop, fanout [5], sampler seed 62 -- COO bit-exact."""
    import torch
    from pylibwholegraph.torch import wholegraph_ops

    wgth, comm = wg
    row_ptr, col = karate_csr(np.int64)
    wm_rp, wm_col = _graph_tensors(wgth, comm, row_ptr, col, "continuous")
    centers = np.arange(34, dtype=np.int64)
    off, dest, lid, gid = wholegraph_ops.unweighted_sample_without_replacement(
        wm_rp.wmb_tensor, wm_col.wmb_tensor, torch.from_numpy(centers).cuda(), 5, 62, True, True)
    eoff, edest, elid, egid = oracle.unweighted_sample(row_ptr, col, centers, 5, 62)
    # COO: majors = centers[lid], minors = dest, edge ids = gid
    assert np.array_equal(dest.cpu().numpy(), edest)
    assert np.array_equal(centers[lid.cpu().numpy()], centers[elid])
    assert np.array_equal(gid.cpu().numpy(), egid)
    assert np.array_equal(off.cpu().numpy(), eoff)


def test_uniform_sampler_optional_outputs_and_errors(wg):
    import torch
    from pylibwholegraph.torch import wholegraph_ops

    wgth, comm = wg
    row_ptr, col = random_csr(1000, 20000, seed=3)
    wm_rp, wm_col = _graph_tensors(wgth, comm, row_ptr, col)
    c = torch.arange(100, device="cuda")
    r = wholegraph_ops.unweighted_sample_without_replacement(wm_rp.wmb_tensor, wm_col.wmb_tensor, c, 7, 1)
    assert len(r) == 2
    r3 = wholegraph_ops.unweighted_sample_without_replacement(wm_rp.wmb_tensor, wm_col.wmb_tensor, c, 7, 1, need_edge_output=True)
    assert len(r3) == 3 and r3[2].dtype == torch.int64 and torch.equal(r3[1], r[1])
    with pytest.raises(NotImplementedError):
        wholegraph_ops.unweighted_sample_without_replacement(wm_rp.wmb_tensor, wm_col.wmb_tensor, c, 2000, 1)
    bad_rp = _wm_from_numpy(wgth, comm, row_ptr.astype(np.int32))
    with pytest.raises(ValueError):  # row_ptr must be int64
        wholegraph_ops.unweighted_sample_without_replacement(bad_rp.wmb_tensor, wm_col.wmb_tensor, c, 7, 1)


def test_uniform_sampler_large_properties(wg):
    """properties at a size the oracle would not finish quickly: offsets, membership, no repeats."""
    import torch
    from pylibwholegraph.torch import wholegraph_ops

    wgth, comm = wg
    nodes, edges = 2_000_000, 32_000_000
    g = torch.Generator(device="cuda").manual_seed(0)
    deg = torch.poisson(torch.full((nodes,), edges / nodes, device="cuda"), generator=g).long()
    row_ptr = torch.zeros(nodes + 1, dtype=torch.int64, device="cuda")
    row_ptr[1:] = deg.cumsum(0)
    E = int(row_ptr[-1])
    col = torch.randint(0, nodes, (E,), device="cuda", dtype=torch.int32, generator=g)
    wm_rp = wgth.create_wholememory_tensor(comm, "chunked", "cuda", [nodes + 1], torch.int64, [1])
    wm_rp.get_local_tensor()[0].copy_(row_ptr)
    wm_col = wgth.create_wholememory_tensor(comm, "chunked", "cuda", [E], torch.int32, [1])
    wm_col.get_local_tensor()[0].copy_(col)
    centers = torch.randint(0, nodes, (500_000,), device="cuda", generator=g)
    M = 10
    off, dest, lid, gid = wholegraph_ops.unweighted_sample_without_replacement(
        wm_rp.wmb_tensor, wm_col.wmb_tensor, centers, M, 62, True, True)
    cnt = torch.minimum(deg[centers], torch.tensor(M, device="cuda"))
    assert torch.equal(off.long(), torch.cat([torch.zeros(1, dtype=torch.long, device="cuda"), cnt.cumsum(0)]))
    assert torch.equal(col[gid], dest)
    src = centers[lid.long()]
    assert bool(((gid >= row_ptr[src]) & (gid < row_ptr[src + 1])).all())
    # without replacement: (seed row, edge id) pairs are unique
    key = lid.long() * (1 << 40) + gid
    assert key.unique().numel() == key.numel()
    for t in (wm_rp, wm_col):
        wgth.destroy_wholememory_tensor(t)


# ------------------------------------------------------------------------------------------------
# S2: weighted sampling  (compared as per-seed sets, as the reference does :341-393)
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("M", [10, 25, 50, 300, -1])
@pytest.mark.parametrize("wdtype", [np.float32, np.float64])
def test_weighted_sampler_sets(wg, oracle, M, wdtype):
    import torch
    from pylibwholegraph.torch import wholegraph_ops

    wgth, comm = wg
    nodes, edges, seeds = 9703, 204323, 400
    row_ptr, col = random_csr(nodes, edges, seed=17)
    w = np.random.default_rng(2).uniform(1.0, 20.0, edges).astype(wdtype)
    centers = np.random.default_rng(3).integers(0, nodes, seeds).astype(np.int64)
    wm_rp, wm_col = _graph_tensors(wgth, comm, row_ptr, col)
    wm_w = _wm_from_numpy(wgth, comm, w)
    off, dest, lid, gid = wholegraph_ops.weighted_sample_without_replacement(
        wm_rp.wmb_tensor, wm_col.wmb_tensor, wm_w.wmb_tensor, torch.from_numpy(centers).cuda(), M, 4242, True, True)
    eoff, edest, elid, egid = oracle.weighted_sample(row_ptr, col, w, centers, M, 4242)
    off, gid, dest, lid = off.cpu().numpy(), gid.cpu().numpy(), dest.cpu().numpy(), lid.cpu().numpy()
    assert np.array_equal(off, eoff)
    assert np.array_equal(lid, elid)
    assert np.array_equal(col[gid], dest)
    mismatched = 0
    for b in range(seeds):
        got = np.sort(gid[off[b]:off[b + 1]])
        exp = np.sort(egid[off[b]:off[b + 1]])
        if np.array_equal(got, exp):
            continue
        # device log1pf and glibc log1pf may differ by an ulp: a differing element must sit within a
        # few ulps of the M-th key
        keys = oracle.weighted_row_keys(row_ptr, w, int(centers[b]), b, M, 4242)
        start = row_ptr[centers[b]]
        thr = np.sort(keys)[-M]
        diff = np.setxor1d(got, exp)
        assert len(np.unique(got)) == len(got)
        assert np.all(np.abs(keys[diff - start] - thr) <= 4 * np.spacing(np.abs(thr)))
        mismatched += 1
    assert mismatched <= max(1, seeds // 100)


def test_weighted_sampler_zero_weight_edges_never_chosen(wg):
    """test_neighbor_loader.py:99-133: a zero-weight edge is excluded while positive ones remain."""
    import torch
    from pylibwholegraph.torch import wholegraph_ops

    wgth, comm = wg
    row_ptr, col = random_csr(2000, 80000, seed=23)
    rng = np.random.default_rng(5)
    w = rng.uniform(1, 20, 80000).astype(np.float32)
    w[rng.random(80000) < 0.4] = 0.0
    wm_rp, wm_col = _graph_tensors(wgth, comm, row_ptr, col)
    wm_w = _wm_from_numpy(wgth, comm, w)
    centers = np.arange(2000)
    off, dest, gid = wholegraph_ops.weighted_sample_without_replacement(
        wm_rp.wmb_tensor, wm_col.wmb_tensor, wm_w.wmb_tensor, torch.from_numpy(centers).cuda(), 5, 7, need_edge_output=True)
    off, gid = off.cpu().numpy(), gid.cpu().numpy()
    for b in range(2000):
        s, e = row_ptr[b], row_ptr[b + 1]
        if e - s > 5 and (w[s:e] > 0).sum() >= 5:
            assert (w[gid[off[b]:off[b + 1]]] > 0).all()


# ------------------------------------------------------------------------------------------------
# S3: append unique
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("T,N,dtype", [(3, 10, np.int32), (53, 123, np.int32), (57, 1235, np.int64), (0, 100, np.int64),
                                       (100, 0, np.int32), (20000, 300000, np.int64), (1, 1, np.int32)])
def test_append_unique_bit_exact(wg, oracle, T, N, dtype):
    import torch
    from pylibwholegraph.torch import graph_ops

    rng = np.random.default_rng(T + 31 * N)
    space = 3 * (T + N) + 10
    targets = rng.permutation(space)[:T].astype(dtype)
    neighbors = rng.integers(0, space, N).astype(dtype)
    uniq, r2u = graph_ops.append_unique(torch.from_numpy(targets).cuda(), torch.from_numpy(neighbors).cuda(), True)
    euniq, er2u = oracle.append_unique(targets, neighbors)
    assert np.array_equal(uniq.cpu().numpy(), euniq)  # first-occurrence order: exact
    assert np.array_equal(r2u.cpu().numpy(), er2u)
    only = graph_ops.append_unique(torch.from_numpy(targets).cuda(), torch.from_numpy(neighbors).cuda())
    assert torch.equal(only, uniq)


# ------------------------------------------------------------------------------------------------
# WholeGraph-style multi-layer sampling (graph_structure.py:136-196) == oracle composition
# ------------------------------------------------------------------------------------------------
def test_multilayer_sample_matches_oracle_composition(wg, oracle):
    import torch

    wgth, comm = wg
    row_ptr, col = random_csr(5000, 90000, seed=41)
    gs = wgth.GraphStructure()
    wm_rp, wm_col = _graph_tensors(wgth, comm, row_ptr, col)
    gs.set_csr_graph(wm_rp, wm_col)
    seeds = np.random.default_rng(0).permutation(5000)[:64].astype(np.int32)
    fanouts = [5, 3]
    tg, ei, rp, ci = gs.multilayer_sample_without_replacement(torch.from_numpy(seeds).cuda(), fanouts, random_seed=100)
    cur = seeds
    for depth, layer in enumerate([1, 0]):
        off, dest, lid, _ = oracle.unweighted_sample(row_ptr, col, cur, fanouts[depth], 100 + depth)
        uniq, r2u = oracle.append_unique(cur, dest)
        assert np.array_equal(rp[layer].cpu().numpy(), off)
        assert np.array_equal(ci[layer].cpu().numpy(), r2u)
        assert np.array_equal(ei[layer].cpu().numpy(), np.stack([r2u, lid]))
        assert np.array_equal(tg[layer].cpu().numpy(), uniq)
        cur = uniq


def test_embedding_gather_is_feature_fetch(wg):
    import torch

    wgth, comm = wg
    emb = wgth.create_embedding(comm, "chunked", "cuda", torch.float32, [10000, 128])
    local, start = emb.get_embedding_tensor().get_local_tensor()
    ref = (torch.arange(10000, device="cuda")[:, None] + torch.arange(128, device="cuda")[None, :]).float()
    local.copy_(ref)
    idx = torch.randint(0, 10000, (4096,), device="cuda")
    assert torch.equal(emb.gather(idx), ref[idx])
    assert torch.equal(wgth.WholeMemoryEmbeddingModule(emb)(idx), ref[idx])
    wgth.destroy_embedding(emb)
