"""S0, fused per-label path (csrc/multihop_fused.cuh) against the oracle AND against the per-hop kernel chain (WGB_MH_FUSED=0):
the two device paths must produce identical bytes for every output of the call, in every cluster-size regime (labels x CL
covers the SMs: B = 1 -> 8 CTAs per label ... B >= 75 -> 1), for ragged / empty / duplicate-laden labels, COO and CSR,
int32 and int64 ids, with edge ids, and for the local ids of the input seeds."""
import os

import numpy as np
import pytest

from graphs import random_csr

pytestmark = pytest.mark.gpu

OUT_KEYS = ("label_hop_offsets", "renumber_map_offsets", "renumber_map", "majors", "minors", "edge_id")


@pytest.fixture(scope="module")
def env():
    import torch
    import pylibwholegraph.torch as wgth

    torch.cuda.set_device(0)
    wgth.init(0, 1, 0, 1)
    return wgth, wgth.get_global_communicator(), wgth.MultiHopSampler()


def _wm(wgth, comm, arr):
    import torch

    t = torch.from_numpy(np.ascontiguousarray(arr))
    wm = wgth.create_wholememory_tensor(comm, "chunked", "cuda", [arr.shape[0]], t.dtype, [1])
    wm.get_local_tensor()[0].copy_(t.cuda())
    return wm


class _path(object):
    """with _path(fused): ... runs sampler calls on the fused path (default) or on the kernel chain"""

    def __init__(self, fused):
        self.fused = fused

    def __enter__(self):
        self.old = os.environ.get("WGB_MH_FUSED")
        os.environ["WGB_MH_FUSED"] = "1" if self.fused else "0"

    def __exit__(self, *a):
        if self.old is None:
            os.environ.pop("WGB_MH_FUSED", None)
        else:
            os.environ["WGB_MH_FUSED"] = self.old


def _both(sampler, *args, **kw):
    with _path(True):
        f = sampler.sample(*args, **kw)
        fl = sampler.seed_local_ids().cpu().numpy()
    with _path(False):
        c = sampler.sample(*args, **kw)
        cl = sampler.seed_local_ids().cpu().numpy()
    for k in c:
        assert k in f, k
        a, b = f[k].cpu().numpy(), c[k].cpu().numpy()
        assert a.dtype == b.dtype and a.shape == b.shape, (k, a.dtype, b.dtype, a.shape, b.shape)
        assert np.array_equal(a, b), "fused path and kernel chain differ in %s" % k
    assert np.array_equal(fl, cl), "seed_local_ids differ between the two paths"
    return f, fl


def _labels(rng, nodes, sizes, dup=False):
    parts = []
    for s in sizes:
        p = rng.permutation(nodes)[:s]
        if dup and s > 3:
            p[rng.integers(0, s, s // 3)] = p[rng.integers(0, s, s // 3)]  # repeated seeds inside the label
        parts.append(p)
    seeds = np.concatenate(parts).astype(np.int64) if parts else np.zeros(0, np.int64)
    lo = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)
    return seeds, lo


@pytest.mark.parametrize("sizes", [[1024], [300, 0, 7, 1024, 1], [64] * 9, [128] * 40, [32] * 100, [17] * 333, [0, 0, 5], [2000, 3000]],
                         ids=["B1", "ragged", "B9", "B40", "B100", "B333", "empty_labels", "big_labels"])
@pytest.mark.parametrize("fanout", [[25, 10], [15, 10, 5], [8], [32, 1], [3, 3, 3, 3]], ids=lambda f: "x".join(map(str, f)))
def test_fused_equals_chain_equals_oracle(env, oracle, sizes, fanout):
    import torch

    wgth, comm, sampler = env
    nodes, edges = 50_021, 900_000
    row_ptr, col = random_csr(nodes, edges, seed=11)
    wm_rp, wm_col = _wm(wgth, comm, row_ptr), _wm(wgth, comm, col)
    seeds, lo = _labels(np.random.default_rng(len(sizes) * 7 + fanout[0]), nodes, sizes, dup=True)
    got, _ = _both(sampler, wm_rp, wm_col, torch.from_numpy(seeds).cuda(), torch.from_numpy(lo).cuda(), fanout, 62)
    exp = oracle.multihop_sample(row_ptr, col, seeds, lo, fanout, 62)
    for k in OUT_KEYS:
        assert np.array_equal(got[k].cpu().numpy(), exp[k]), k


@pytest.mark.parametrize("col_dtype", [np.int32, np.int64])
@pytest.mark.parametrize("seed_dtype", [np.int32, np.int64])
def test_fused_dtypes_edge_ids_csr_int64_ids(env, oracle, col_dtype, seed_dtype):
    import torch

    wgth, comm, sampler = env
    nodes, edges = 9001, 150_000
    row_ptr, col = random_csr(nodes, edges, seed=5, col_dtype=col_dtype)
    eids = np.random.default_rng(2).permutation(edges).astype(np.int64)
    wm_rp, wm_col, wm_eid = _wm(wgth, comm, row_ptr), _wm(wgth, comm, col), _wm(wgth, comm, eids)
    seeds, lo = _labels(np.random.default_rng(3), nodes, [100, 31, 0, 64, 200, 1], dup=True)
    s, l = torch.from_numpy(seeds.astype(seed_dtype)).cuda(), torch.from_numpy(lo).cuda()
    exp = oracle.multihop_sample(row_ptr, col, seeds, lo, [7, 5, 3], 9, edge_ids=eids)
    got, local = _both(sampler, wm_rp, wm_col, s, l, [7, 5, 3], 9, csr_edge_id=wm_eid)
    for k in OUT_KEYS:
        assert np.array_equal(got[k].cpu().numpy(), exp[k]), k
    # local id of every input seed: the label's renumber map at that id is the seed
    rmo, m = exp["renumber_map_offsets"], exp["renumber_map"]
    for b in range(len(lo) - 1):
        assert np.array_equal(m[rmo[b]:rmo[b + 1]][local[lo[b]:lo[b + 1]]], seeds[lo[b]:lo[b + 1]])
    got64, _ = _both(sampler, wm_rp, wm_col, s, l, [7, 5, 3], 9, csr_edge_id=wm_eid, int64_ids=True)
    assert got64["majors"].dtype == torch.int64
    for k in OUT_KEYS:
        assert np.array_equal(got64[k].cpu().numpy(), exp[k]), k
    csr, _ = _both(sampler, wm_rp, wm_col, s, l, [7, 5, 3], 9, csr_edge_id=wm_eid, compression="CSR")
    assert "majors" not in csr and np.array_equal(csr["minors"].cpu().numpy(), exp["minors"])
    assert int(csr["major_offsets"][-1]) == len(exp["minors"])


def test_fused_repeated_calls_and_two_objects(env, oracle):
    """scratch reuse across calls of different shapes on one object, and two objects in flight (the loader's pipeline)"""
    import torch

    wgth, comm, sampler = env
    other = wgth.MultiHopSampler()
    row_ptr, col = random_csr(20_000, 300_000, seed=8)
    wm_rp, wm_col = _wm(wgth, comm, row_ptr), _wm(wgth, comm, col)
    rng = np.random.default_rng(1)
    shapes = [([256] * 8, [25, 10]), ([64] * 64, [10, 10]), ([1000], [5, 5, 5]), ([256] * 8, [25, 10])]
    calls = []
    for k, (sizes, fanout) in enumerate(shapes * 2):
        seeds, lo = _labels(rng, 20_000, sizes)
        calls.append((seeds, lo, fanout, 100 + k))
    objs = [sampler, other]
    with _path(True):
        pend = objs[0].sample_async(wm_rp, wm_col, torch.from_numpy(calls[0][0]).cuda(), torch.from_numpy(calls[0][1]).cuda(), calls[0][2], calls[0][3])
        for k in range(len(calls)):
            nxt = None
            if k + 1 < len(calls):
                s, l, f, seed = calls[k + 1]
                nxt = objs[(k + 1) & 1].sample_async(wm_rp, wm_col, torch.from_numpy(s).cuda(), torch.from_numpy(l).cuda(), f, seed)
            got = pend.result()
            exp = oracle.multihop_sample(row_ptr, col, calls[k][0], calls[k][1], calls[k][2], calls[k][3])
            for key in OUT_KEYS:
                assert np.array_equal(got[key].cpu().numpy(), exp[key]), (k, key)
            pend = nxt


def test_fused_call_group_of_bench_shape_equals_chain(env, oracle):
    """64 labels x 1024 seeds, fan-out [25, 10] on a 2 M-vertex power-law-ish graph: fused == chain on every byte, and the
    first four labels == oracle (labels are independent except for the stream geometry, which the chain shares)."""
    import torch

    wgth, comm, sampler = env
    nodes, edges = 2_000_003, 32_000_000
    rng = np.random.default_rng(0)
    row_ptr, col = random_csr(nodes, edges, seed=1)
    wm_rp, wm_col = _wm(wgth, comm, row_ptr), _wm(wgth, comm, col)
    seeds, lo = _labels(rng, nodes, [1024] * 64)
    s, l = torch.from_numpy(seeds).cuda(), torch.from_numpy(lo).cuda()
    got, _ = _both(sampler, wm_rp, wm_col, s, l, [25, 10], 62, int64_ids=True)
    exp = oracle.multihop_sample(row_ptr, col, seeds[:4096], lo[:5], [25, 10], 62)
    n_e, n_v = len(exp["minors"]), len(exp["renumber_map"])
    assert np.array_equal(got["minors"].cpu().numpy()[:n_e], exp["minors"])
    assert np.array_equal(got["majors"].cpu().numpy()[:n_e], exp["majors"])
    assert np.array_equal(got["edge_id"].cpu().numpy()[:n_e], exp["edge_id"])
    assert np.array_equal(got["renumber_map"].cpu().numpy()[:n_v], exp["renumber_map"])
