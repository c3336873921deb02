"""S0 heterogeneous: the fused multi-hop sampler over typed CSRs against the oracle (bit-exact), the reference's
deterministic pin, and the T = 1 / Vt = 1 degenerate case against the homogeneous entry point."""
import json
import os

import numpy as np
import pytest

from graphs import random_csr, random_typed_graph, typed_csrs

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
PINS = json.load(open(os.path.join(HERE, "golden", "reference_pins.json")))
KEYS = ("label_type_hop_offsets", "renumber_map_offsets", "renumber_map", "majors", "minors", "edge_id", "edge_type",
        "edge_renumber_map", "edge_renumber_map_offsets", "label_type_step_base")


@pytest.fixture(scope="module")
def env():
    import torch
    import pylibwholegraph.torch as wgth

    torch.cuda.set_device(0)
    wgth.init(0, 1, 0, 1)
    return wgth, wgth.get_global_communicator(), wgth.MultiHopSampler()


def _dev(arrs):
    import torch

    return [torch.from_numpy(np.ascontiguousarray(a)).cuda() for a in arrs]


def _assert_equal(got, exp):
    for k in KEYS:
        g = got[k].cpu().numpy()
        assert g.shape == exp[k].shape, (k, g.shape, exp[k].shape)
        assert np.array_equal(g, exp[k]), k


@pytest.mark.parametrize("fanout", [[3, 2, 4, 2, 2, 2], [-1, 3, 0, 2, -1, 1], [5, 5, 5], [0, 0, 0, 4, 4, 4], [10, 10, 10, 10, 10, 10],
                                    [2, 40, 1]])
@pytest.mark.parametrize("col_dtype", [np.int32, np.int64])
def test_hetero_bit_exact_vs_oracle(env, oracle, fanout, col_dtype):
    import torch

    wgth, comm, sampler = env
    edge_types = [(0, 1), (1, 0), (1, 1)]
    vto, row_ptrs, cols = random_typed_graph([700, 1500], edge_types, [9000, 14000, 21000], seed=len(fanout), col_dtype=col_dtype)
    rng = np.random.default_rng(4)
    seeds = np.concatenate([rng.integers(0, 2200, 60), rng.integers(700, 2200, 33), rng.integers(0, 700, 1)]).astype(np.int64)
    lo = np.array([0, 60, 60, 93, 94], dtype=np.int64)
    d_rp, d_col = _dev(row_ptrs), _dev(cols)
    for rep in range(2):
        got = sampler.sample_hetero(d_rp, d_col, vto.tolist(), torch.from_numpy(seeds).cuda(), torch.from_numpy(lo).cuda(), fanout, 31 + rep)
        exp = oracle.hetero_multihop_sample(row_ptrs, cols, vto, seeds, lo, fanout, 31 + rep)
        _assert_equal(got, exp)


def test_hetero_biased_and_edge_ids(env, oracle):
    """Weighted (A-Res) sampling per edge type + per-type edge ids.  One hop: the reference compares weighted samples per
    row as SETS, and a device/glibc log1pf ulp difference may flip a near-tie in <1% of rows (tests/test_gpu_multihop.py),
    so later hops are not comparable element-wise."""
    import torch

    wgth, comm, sampler = env
    edge_types = [(0, 0), (0, 1), (1, 0)]
    vto, row_ptrs, cols = random_typed_graph([900, 400], edge_types, [12000, 6000, 5000], seed=77)
    rng = np.random.default_rng(9)
    wts = [rng.random(c.shape[0]).astype(np.float32) + 0.01 for c in cols]
    eids = [rng.permutation(c.shape[0]).astype(np.int64) + 100000 * t for t, c in enumerate(cols)]
    eids[1] = None  # a type without edge ids reports CSR positions
    seeds = np.concatenate([rng.permutation(1300)[:50], rng.permutation(1300)[:30]]).astype(np.int64)
    lo = np.array([0, 50, 80], dtype=np.int64)
    fanout = [4, 3, 2]
    got = sampler.sample_hetero(_dev(row_ptrs), _dev(cols), vto.tolist(), torch.from_numpy(seeds).cuda(), torch.from_numpy(lo).cuda(),
                                fanout, 5, csr_weights=_dev(wts), csr_edge_ids=[None if e is None else _dev([e])[0] for e in eids])
    exp = oracle.hetero_multihop_sample(row_ptrs, cols, vto, seeds, lo, fanout, 5, weights=wts, edge_ids=eids)
    got = {k: v.cpu().numpy() for k, v in got.items()}
    for k in ("label_type_hop_offsets", "edge_renumber_map_offsets", "edge_type", "edge_id", "majors"):
        assert np.array_equal(got[k], exp[k]), k  # counts are min(deg, fanout): independent of the draws
    lho, rmo = exp["label_type_hop_offsets"], got["renumber_map_offsets"]
    rows = bad = 0
    for l in range(2):
        for t, (sv, dv) in enumerate(edge_types):
            a, b = lho[l * 3 + t], lho[l * 3 + t + 1]
            smap = got["renumber_map"][rmo[l * 2 + sv]:rmo[l * 2 + sv + 1]]
            dmap = got["renumber_map"][rmo[l * 2 + dv]:rmo[l * 2 + dv + 1]]
            src, dst = smap[got["majors"][a:b]], dmap[got["minors"][a:b]]
            ids = got["edge_renumber_map"][a:b]
            pos = ids if eids[t] is None else np.argsort(eids[t])[ids - 100000 * t]  # back to CSR positions
            assert np.array_equal(cols[t][pos].astype(np.int64), dst)
            assert ((row_ptrs[t][src] <= pos) & (pos < row_ptrs[t][src + 1])).all()
            starts = np.flatnonzero(np.r_[True, src[1:] != src[:-1]]) if len(src) else []
            for i, s0 in enumerate(starts):
                s1 = starts[i + 1] if i + 1 < len(starts) else len(src)
                rows += 1
                assert len(np.unique(pos[s0:s1])) == s1 - s0
                bad += not np.array_equal(np.sort(ids[s0:s1]), np.sort(exp["edge_renumber_map"][a:b][s0:s1]))
    assert rows > 100 and bad <= max(1, rows // 100)


def test_hetero_reference_pin(env, oracle):
    """python/cugraph-pyg/cugraph_pyg/tests/sampler/test_distributed_sampler.py:19-150 on the GPU path."""
    import torch

    wgth, comm, sampler = env
    pin = PINS["hetero_fanout_all"]
    srcs, dsts, eids, etps = (np.array(pin[k]) for k in ("srcs", "dsts", "eids", "etps"))
    row_ptrs, cols, pos = typed_csrs(srcs, dsts, etps, 2, 10)
    e = [eids[p].astype(np.int64) for p in pos]
    got = sampler.sample_hetero(_dev(row_ptrs), _dev(cols), [0, 4, 10], torch.tensor([4, 5]).cuda(), torch.tensor([0, 2]).cuda(),
                                [-1, -1, -1, -1], 62, csr_edge_ids=_dev(e))
    exp = oracle.hetero_multihop_sample(row_ptrs, cols, [0, 4, 10], np.array([4, 5]), np.array([0, 2]), [-1, -1, -1, -1], 62, edge_ids=e)
    _assert_equal(got, exp)
    out = {k: v.cpu().numpy() for k, v in got.items()}
    lho, rmo, ermo = out["label_type_hop_offsets"], out["renumber_map_offsets"], out["edge_renumber_map_offsets"]
    smap = out["renumber_map"][rmo[1]:rmo[2]]
    dmap0 = out["renumber_map"][rmo[0]:rmo[1]]
    expect = {(0, 0): ([0, 1], [4, 5], [0, 1]), (0, 1): ([4, 5], [8, 9], [0, 3]),
              (1, 0): ([5, 6, 7], [4, 5, 5], [8, 9, 9]), (1, 1): ([0, 1, 2], [8, 8, 9], [4, 5, 6])}
    for (t, h), (e_exp, s_exp, d_exp) in expect.items():
        a, b = lho[t * 2 + h], lho[t * 2 + h + 1]
        emap = out["edge_renumber_map"][ermo[t]:ermo[t + 1]]
        assert sorted(emap[out["edge_id"][a:b]].tolist()) == e_exp
        assert sorted(smap[out["majors"][a:b]].tolist()) == s_exp
        assert sorted((dmap0 if t == 0 else smap)[out["minors"][a:b]].tolist()) == d_exp


def test_hetero_one_type_equals_homogeneous_entry_point(env, oracle):
    import torch

    wgth, comm, sampler = env
    row_ptr, col = random_csr(5000, 60000, seed=12)
    rng = np.random.default_rng(1)
    seeds = rng.permutation(5000)[:300].astype(np.int64)
    lo = np.array([0, 100, 300], dtype=np.int64)
    d_rp, d_col = _dev([row_ptr])[0], _dev([col])[0]
    s, l = torch.from_numpy(seeds).cuda(), torch.from_numpy(lo).cuda()
    homo = sampler.sample(d_rp, d_col, s, l, [6, 4], 8)
    het = sampler.sample_hetero([d_rp], [d_col], [0, 5000], s, l, [6, 4], 8)
    assert torch.equal(het["majors"], homo["majors"]) and torch.equal(het["minors"], homo["minors"])
    assert torch.equal(het["renumber_map"], homo["renumber_map"])
    assert torch.equal(het["label_type_hop_offsets"], homo["label_hop_offsets"])
    assert torch.equal(het["edge_renumber_map"], homo["edge_id"])
    assert torch.equal(het["label_type_step_base"].reshape(3, -1), homo["label_step_base"])


def test_hetero_async_and_error_paths(env, oracle):
    import torch

    wgth, comm, sampler = env
    edge_types = [(0, 1), (1, 0)]
    vto, row_ptrs, cols = random_typed_graph([300, 300], edge_types, [3000, 3000], seed=5)
    seeds = np.arange(0, 600, 7, dtype=np.int64)
    lo = np.array([0, len(seeds)], dtype=np.int64)
    d_rp, d_col = _dev(row_ptrs), _dev(cols)
    s, l = torch.from_numpy(seeds).cuda(), torch.from_numpy(lo).cuda()
    pend = sampler.sample_hetero_async(d_rp, d_col, vto.tolist(), s, l, [3, 3, 2, 2], 1)
    _assert_equal(pend.result(), oracle.hetero_multihop_sample(row_ptrs, cols, vto, seeds, lo, [3, 3, 2, 2], 1))
    with pytest.raises(ValueError):  # vertex_type_offsets must start at 0
        sampler.sample_hetero(d_rp, d_col, [1, 300, 600], s, l, [3, 3], 1)
    with pytest.raises(ValueError):  # mixed col dtypes
        sampler.sample_hetero(d_rp, [d_col[0], d_col[1].long()], vto.tolist(), s, l, [3, 3], 1)
