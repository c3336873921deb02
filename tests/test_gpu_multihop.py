"""S0 on the GPU: the fused multi-hop sampler against the oracle's composition, bit-exact (COO), plus the
CSR form, the reference's deterministic pins and the decoder contract of cugraph-pyg's readers."""
import json
import os

import numpy as np
import pytest

from graphs import karate_csr, random_csr

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
PINS = json.load(open(os.path.join(HERE, "golden", "reference_pins.json")))


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


def _labels(rng, nodes, sizes):
    seeds = np.concatenate([rng.permutation(nodes)[:s] for s in sizes]).astype(np.int64)
    lo = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)
    return seeds, lo


def _assert_equal_coo(got, exp):
    for k in ("label_hop_offsets", "renumber_map_offsets", "renumber_map", "majors", "minors", "edge_id"):
        g = got[k].cpu().numpy()
        assert g.shape == exp[k].shape, (k, g.shape, exp[k].shape)
        assert np.array_equal(g, exp[k]), k


@pytest.mark.parametrize("fanout", [[25, 10], [5, 5, 5], [10], [3, 0, 4], [-1, 2], [2, -1], [-1, -1], [40, 3], [1, 1, 1, 1]])
@pytest.mark.parametrize("col_dtype", [np.int32, np.int64])
def test_multihop_bit_exact_vs_oracle(env, oracle, fanout, col_dtype):
    import torch

    wgth, comm, sampler = env
    nodes, edges = 6007, 70011
    row_ptr, col = random_csr(nodes, edges, seed=len(fanout) * 100 + abs(fanout[0]), col_dtype=col_dtype)
    seeds, lo = _labels(np.random.default_rng(7), nodes, [64, 1, 0, 200, 33])
    wm_rp, wm_col = _wm(wgth, comm, row_ptr), _wm(wgth, comm, col)
    for rep in range(2):  # second call exercises the epoch-tagged reuse of the hash table
        got = sampler.sample(wm_rp, wm_col, torch.from_numpy(seeds).cuda(), torch.from_numpy(lo).cuda(), fanout, 62 + rep)
        exp = oracle.multihop_sample(row_ptr, col, seeds, lo, fanout, 62 + rep)
        _assert_equal_coo(got, exp)


def test_multihop_many_calls_epoch_wrap(env, oracle):
    """> 254 calls force the epoch counter to wrap (table re-initialised), results stay exact."""
    import torch

    wgth, comm, sampler = env
    row_ptr, col = random_csr(500, 6000, seed=3)
    wm_rp, wm_col = _wm(wgth, comm, row_ptr), _wm(wgth, comm, col)
    seeds, lo = _labels(np.random.default_rng(1), 500, [16, 16])
    s, l = torch.from_numpy(seeds).cuda(), torch.from_numpy(lo).cuda()
    for k in range(300):
        got = sampler.sample(wm_rp, wm_col, s, l, [4, 3], 1000 + k)
        if k % 37 == 0 or k > 250:
            _assert_equal_coo(got, oracle.multihop_sample(row_ptr, col, seeds, lo, [4, 3], 1000 + k))


def test_multihop_duplicate_seeds_int32_seeds_and_edge_ids(env, oracle):
    import torch

    wgth, comm, sampler = env
    row_ptr, col = random_csr(3000, 40000, seed=9)
    eids = np.random.default_rng(2).permutation(40000).astype(np.int64)
    wm_rp, wm_col, wm_eid = _wm(wgth, comm, row_ptr), _wm(wgth, comm, col), _wm(wgth, comm, eids)
    seeds = np.array([5, 9, 5, 700, 9, 9, 12, 5, 40, 41, 40], dtype=np.int64)
    lo = np.array([0, 7, 11], dtype=np.int64)
    got = sampler.sample(wm_rp, wm_col, torch.from_numpy(seeds.astype(np.int32)).cuda(), torch.from_numpy(lo).cuda(), [6, 4], 5,
                         csr_edge_id=wm_eid)
    exp = oracle.multihop_sample(row_ptr, col, seeds, lo, [6, 4], 5, edge_ids=eids)
    _assert_equal_coo(got, exp)
    m = got["renumber_map"].cpu().numpy()
    assert m[:4].tolist() == [5, 9, 700, 12]  # first occurrence keeps the id


def test_multihop_int64_ids_and_csr_consistent_with_coo(env, oracle):
    import torch

    wgth, comm, sampler = env
    row_ptr, col = random_csr(8000, 120000, seed=21)
    wm_rp, wm_col = _wm(wgth, comm, row_ptr), _wm(wgth, comm, col)
    seeds, lo = _labels(np.random.default_rng(3), 8000, [128, 96, 1, 77])
    s, l = torch.from_numpy(seeds).cuda(), torch.from_numpy(lo).cuda()
    fanout = [7, 5, 3]
    B, L = 4, 3
    coo = sampler.sample(wm_rp, wm_col, s, l, fanout, 77, int64_ids=True)
    assert coo["majors"].dtype == torch.int64 and coo["minors"].dtype == torch.int64
    exp = oracle.multihop_sample(row_ptr, col, seeds, lo, fanout, 77)
    _assert_equal_coo({k: v for k, v in coo.items()}, {k: (v.astype(np.int64) if k in ("majors", "minors") else v) for k, v in exp.items()})
    csr = sampler.sample(wm_rp, wm_col, s, l, fanout, 77, compression="CSR")
    assert "majors" not in csr
    mo = csr["major_offsets"].cpu().numpy()
    lho = csr["label_hop_offsets"].cpu().numpy()
    assert np.array_equal(csr["minors"].cpu().numpy(), exp["minors"])
    assert np.array_equal(csr["edge_id"].cpu().numpy(), exp["edge_id"])
    assert np.array_equal(csr["renumber_map"].cpu().numpy(), exp["renumber_map"])
    assert mo[-1] == len(exp["minors"]) and (np.diff(mo) >= 0).all()
    # expanding major_offsets per label gives back the COO majors (the reference does this with ptr2index,
    # sampler/sampler.py:63-65)
    for b in range(B):
        seg = mo[lho[b * L]: lho[(b + 1) * L] + 1]
        majors = np.repeat(np.arange(len(seg) - 1), np.diff(seg))
        a, e = exp["label_hop_offsets"][b * L], exp["label_hop_offsets"][(b + 1) * L]
        assert seg[0] == a and seg[-1] == e
        assert np.array_equal(majors, exp["majors"][a:e])
        # hop boundaries inside the label: number of edges per hop as the decoder computes it (sampler.py:560)
        cur = lho[b * L: (b + 1) * L + 1] - lho[b * L]
        assert np.array_equal(np.diff(seg[cur] - seg[0]), np.diff(exp["label_hop_offsets"][b * L:(b + 1) * L + 1]))


def test_multihop_biased_matches_oracle_sets_and_zero_weight_pin(env, oracle):
    import torch

    wgth, comm, sampler = env
    nodes, edges = 4000, 90000
    row_ptr, col = random_csr(nodes, edges, seed=33)
    rng = np.random.default_rng(4)
    w = rng.uniform(1, 20, edges).astype(np.float32)
    w[rng.random(edges) < 0.3] = 0.0
    wm_rp, wm_col, wm_w = _wm(wgth, comm, row_ptr), _wm(wgth, comm, col), _wm(wgth, comm, w)
    seeds, lo = _labels(rng, nodes, [100, 50])
    got = sampler.sample(wm_rp, wm_col, torch.from_numpy(seeds).cuda(), torch.from_numpy(lo).cuda(), [5], 11, csr_weight=wm_w)
    exp = oracle.multihop_sample(row_ptr, col, seeds, lo, [5], 11, weights=w)
    assert np.array_equal(got["label_hop_offsets"].cpu().numpy(), exp["label_hop_offsets"])
    gm, em = got["renumber_map"].cpu().numpy(), exp["renumber_map"]
    eid = got["edge_id"].cpu().numpy()
    mj = got["majors"].cpu().numpy()
    lho = exp["label_hop_offsets"]
    for b in range(2):
        a, e = lho[b], lho[b + 1]
        # per (label, source) edge sets agree with the oracle up to near-threshold ties; zero weights never chosen
        rmo = got["renumber_map_offsets"].cpu().numpy()
        src = gm[rmo[b]:rmo[b + 1]][mj[a:e]]
        esrc = em[exp["renumber_map_offsets"][b]:exp["renumber_map_offsets"][b + 1]][exp["majors"][a:e]]
        assert np.array_equal(src, esrc)
        # per source row: identical sets when enough positive-weight edges exist (keys are continuous; a device /
        # glibc log1pf ulp difference may flip a near-tie in <1% of rows); otherwise all positive edges are taken
        # and the zero-weight remainder (all keys == -inf, order unspecified) only has to come from the row.
        deg = row_ptr[src + 1] - row_ptr[src]
        bad = 0
        starts = np.flatnonzero(np.r_[True, src[1:] != src[:-1]])
        for i, s0 in enumerate(starts):
            s1 = starts[i + 1] if i + 1 < len(starts) else len(src)
            v = src[s0]
            g, x = eid[a:e][s0:s1], exp["edge_id"][a:e][s0:s1]
            wrow = w[row_ptr[v]:row_ptr[v + 1]]
            pos = int((wrow > 0).sum())
            assert ((g >= row_ptr[v]) & (g < row_ptr[v + 1])).all() and len(np.unique(g)) == len(g)
            if deg[s0] <= 5:
                assert np.array_equal(np.sort(g), np.sort(x))
            elif pos >= 5:
                assert (w[g] > 0).all()
                bad += not np.array_equal(np.sort(g), np.sort(x))
            else:
                assert (w[g] > 0).sum() == pos
        assert bad <= max(1, len(starts) // 100)


def test_multihop_reference_pin_fanout_all(env):
    """test_distributed_sampler.py:19-150 projected per edge type (see tests/test_oracle_cpu.py)."""
    import torch

    wgth, comm, sampler = env
    pin = PINS["hetero_fanout_all"]
    srcs, dsts, eids, etps = (np.array(pin[k]) for k in ("srcs", "dsts", "eids", "etps"))
    order = np.lexsort((np.arange(len(srcs)), srcs))
    row_ptr = np.zeros(11, dtype=np.int64)
    np.cumsum(np.bincount(srcs, minlength=10), out=row_ptr[1:])
    wm_rp, wm_col = _wm(wgth, comm, row_ptr), _wm(wgth, comm, dsts[order].astype(np.int64))
    res = sampler.sample(wm_rp, wm_col, torch.tensor(pin["seeds"]).cuda(), torch.tensor([0, 2]).cuda(), [-1, -1], 0)
    m = res["renumber_map"].cpu().numpy()
    lho = res["label_hop_offsets"].cpu().numpy()
    for hop in (0, 1):
        a, b = lho[hop], lho[hop + 1]
        pos = order[res["edge_id"].cpu().numpy()[a:b]]
        for et in (0, 1):
            sel = etps[pos] == et
            exp = pin["expect"]["etype%d_hop%d" % (et, hop)]
            assert sorted(eids[pos][sel].tolist()) == exp["eids"]
            assert sorted(m[res["majors"].cpu().numpy()[a:b]][sel].tolist()) == exp["srcs"]
            assert sorted(m[res["minors"].cpu().numpy()[a:b]][sel].tolist()) == exp["dsts"]


def test_multihop_c1_karate_and_large_properties(env, oracle):
    import torch

    wgth, comm, sampler = env
    row_ptr, col = karate_csr(np.int64)
    wm_rp, wm_col = _wm(wgth, comm, row_ptr), _wm(wgth, comm, col)
    seeds = np.arange(34, dtype=np.int64)
    lo = np.array([0, 34], dtype=np.int64)
    got = sampler.sample(wm_rp, wm_col, torch.from_numpy(seeds).cuda(), torch.from_numpy(lo).cuda(), [5], 62)
    _assert_equal_coo(got, oracle.multihop_sample(row_ptr, col, seeds, lo, [5], 62))
    # larger call group: structural properties only
    nodes, edges = 300_000, 4_000_000
    row_ptr, col = random_csr(nodes, edges, seed=5)
    wm_rp, wm_col = _wm(wgth, comm, row_ptr), _wm(wgth, comm, col)
    B, per = 32, 1024
    seeds, lo = _labels(np.random.default_rng(0), nodes, [per] * B)
    got = sampler.sample(wm_rp, wm_col, torch.from_numpy(seeds).cuda(), torch.from_numpy(lo).cuda(), [25, 10], 62)
    lho = got["label_hop_offsets"].cpu().numpy()
    rmo = got["renumber_map_offsets"].cpu().numpy()
    m = got["renumber_map"].cpu().numpy()
    mj, mn, eid = (got[k].cpu().numpy() for k in ("majors", "minors", "edge_id"))
    assert lho[-1] == len(mj) and rmo[-1] == len(m)
    for b in range(B):
        mm = m[rmo[b]:rmo[b + 1]]
        assert np.array_equal(mm[:per], seeds[lo[b]:lo[b + 1]])
        assert len(np.unique(mm)) == len(mm)
        a, e = lho[2 * b], lho[2 * b + 2]
        assert np.array_equal(mm[mn[a:e]], col[eid[a:e]])
        s = mm[mj[a:e]]
        assert ((eid[a:e] >= row_ptr[s]) & (eid[a:e] < row_ptr[s + 1])).all()
        h0 = lho[2 * b + 1]
        assert mj[a:h0].max() < per and (mj[h0:e].min() >= per if e > h0 else True)
        assert mn[a:e].max() + 1 == len(mm)


def test_multihop_async_two_samplers_interleaved(env, oracle):
    """sample_async on two sampler objects, begun back to back and finished out of phase (the loader's software
    pipeline), gives exactly the synchronous results; an abandoned pending call does not poison the object."""
    import torch

    wgth, comm, sampler = env
    other = wgth.MultiHopSampler()
    row_ptr, col = random_csr(4001, 52000, seed=21)
    wm_rp, wm_col = _wm(wgth, comm, row_ptr), _wm(wgth, comm, col)
    calls = []
    rng = np.random.default_rng(5)
    for k in range(6):
        seeds, lo = _labels(rng, 4001, [40, 17, 0, 90])
        calls.append((seeds, lo, 700 + k))
    objs = [sampler, other]
    pend = objs[0].sample_async(wm_rp, wm_col, torch.from_numpy(calls[0][0]).cuda(), torch.from_numpy(calls[0][1]).cuda(), [7, 5], calls[0][2])
    for k in range(6):
        nxt = None
        if k + 1 < 6:
            s, l, seed = calls[k + 1]
            nxt = objs[(k + 1) & 1].sample_async(wm_rp, wm_col, torch.from_numpy(s).cuda(), torch.from_numpy(l).cuda(), [7, 5], seed)
        got = pend.result()
        assert pend.result() is got  # idempotent
        _assert_equal_coo(got, oracle.multihop_sample(row_ptr, col, calls[k][0], calls[k][1], [7, 5], calls[k][2]))
        pend = nxt
    # abandoned call, then a normal one on the same object
    s, l, seed = calls[0]
    other.sample_async(wm_rp, wm_col, torch.from_numpy(s).cuda(), torch.from_numpy(l).cuda(), [7, 5], seed)
    got = other.sample(wm_rp, wm_col, torch.from_numpy(s).cuda(), torch.from_numpy(l).cuda(), [7, 5], seed, compression="CSR")
    exp = oracle.multihop_sample(row_ptr, col, s, l, [7, 5], seed)
    assert np.array_equal(got["renumber_map"].cpu().numpy(), exp["renumber_map"])
    assert np.array_equal(got["minors"].cpu().numpy(), exp["minors"])
