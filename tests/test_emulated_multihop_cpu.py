"""The product's fused multi-hop sampler (csrc/multihop.cu: host code AND kernels) compiled for the CPU through the SIMT
shim in tests/emu and compared with the oracle, bit for bit.

What this is: a LOGIC check of the exact source the GPU runs (launch order, scratch sizing, indexing, barriers, hash
table protocol, random-stream geometry), available without a GPU.  It was written when the temporal path had to be
finished after the round's GPU minutes were spent; the emulated plain / heterogeneous calls below are the calibration --
that path is GPU-verified (tests/test_gpu_multihop.py, test_gpu_hetero.py), and the emulator reproduces it.
What this is not: evidence about memory ordering, races that need real parallelism, or speed.  The GPU parity tests stay
the gate (tests/test_gpu_temporal.py), and nothing here is linked into the product.
"""
import ctypes
import os
import sys

import numpy as np
import pytest

from graphs import random_typed_graph

HERE = os.path.dirname(os.path.abspath(__file__))
VP = ctypes.c_void_p
FLAG_CSR, FLAG_INT64 = 1, 2
COMPARISONS = ("strictly_increasing", "monotonically_increasing", "strictly_decreasing", "monotonically_decreasing")
HETERO = ("majors", "minors", "edge_id", "edge_type", "label_type_hop_offsets", "renumber_map", "renumber_map_offsets",
          "edge_renumber_map", "edge_renumber_map_offsets", "label_type_step_base")
HOMO = ("majors", "minors", "edge_id", "label_hop_offsets", "renumber_map", "renumber_map_offsets", "major_offsets", "label_step_base")


@pytest.fixture(scope="module")
def emu():
    sys.path.insert(0, os.path.join(HERE, "emu"))
    import build_emu

    if not build_emu.available():
        pytest.skip("CUDA headers not installed")
    return ctypes.CDLL(build_emu.build_multihop())


def _run(lib, row_ptrs, cols, vto, seeds, lo, fanout, random_state, *, hetero=True, times=None, seed_times=None, cmp=0, eids=None,
         flags=FLAG_INT64, expect_rc=0, reps=1, weights=None, seed_local_ids=False):
    T = len(row_ptrs)
    rp = [np.ascontiguousarray(r, dtype=np.int64) for r in row_ptrs]
    cl = [np.ascontiguousarray(c) for c in cols]
    is64 = cl[0].dtype == np.int64
    ne = np.array([c.shape[0] for c in cl], dtype=np.int64)
    tm = None if times is None else [np.ascontiguousarray(t, dtype=np.int64) for t in times]
    ei = None if eids is None else [None if e is None else np.ascontiguousarray(e, dtype=np.int64) for e in eids]
    wt = None if weights is None else [np.ascontiguousarray(w) for w in weights]
    vto = np.ascontiguousarray(vto, dtype=np.int64)
    seeds = np.ascontiguousarray(seeds, dtype=np.int64)
    st = np.zeros_like(seeds) if seed_times is None else np.ascontiguousarray(seed_times, dtype=np.int64)
    lo = np.ascontiguousarray(lo, dtype=np.int64)
    fo = np.ascontiguousarray(fanout, dtype=np.int32)

    def arr(xs):
        return (VP * T)(*[None if x is None else x.ctypes.data for x in xs])

    out_ptr, out_cnt, out_elt = (VP * 11)(), (ctypes.c_longlong * 11)(), (ctypes.c_int * 11)()
    lib.emu_multihop.restype = ctypes.c_int
    rc = lib.emu_multihop(T, arr(rp), ctypes.c_longlong(rp[0].shape[0] - 1), arr(cl), ne.ctypes.data_as(VP), int(is64),
                          None if tm is None else arr(tm), None if wt is None else arr(wt), int(wt is not None and wt[0].dtype == np.float64),
                          None if ei is None else arr(ei), vto.ctypes.data_as(VP), vto.shape[0] - 1,
                          int(hetero), seeds.ctypes.data_as(VP), st.ctypes.data_as(VP), ctypes.c_longlong(seeds.shape[0]), lo.ctypes.data_as(VP),
                          ctypes.c_longlong(lo.shape[0] - 1), fo.ctypes.data_as(VP), fo.shape[0] // T, ctypes.c_ulonglong(random_state), cmp,
                          flags, reps, int(seed_local_ids), None if tm is None else np.array([t.shape[0] for t in tm], dtype=np.int64).ctypes.data_as(VP),
                          out_ptr, out_cnt, out_elt)
    assert rc == expect_rc, rc
    outs = {}
    for k, name in enumerate((HETERO if hetero else HOMO + ("unused", "unused2")) + ("seed_local_ids",)):
        if not out_ptr[k]:
            continue
        dt = {4: np.int32, 8: np.int64}[out_elt[k]]
        n = int(out_cnt[k])
        outs[name] = np.frombuffer(ctypes.string_at(out_ptr[k], n * out_elt[k]), dtype=dt).copy() if n else np.empty(0, dt)
        lib.emu_free(VP(out_ptr[k]))
    return outs


def _same(got, exp, names):
    for name in names:
        e = np.asarray(exp[name]).reshape(-1)
        assert got[name].shape == e.shape, (name, got[name].shape, e.shape)
        assert np.array_equal(got[name].astype(np.int64), e.astype(np.int64)), name


def _majors_of_csr(out, L, B):
    """major_offsets -> majors, label by label (local ids restart per label; the CSR rows of a label are its sources)."""
    lho, moff = out["label_hop_offsets"], out["major_offsets"]
    majors = []
    for l in range(B):
        r0, r1 = lho[l * L], lho[(l + 1) * L]
        majors.append(np.repeat(np.arange(r1 - r0), np.diff(moff[r0:r1 + 1])))
    return np.concatenate(majors)


def _typed_case(col_dtype=np.int32, seed=6):
    edge_types = [(0, 1), (1, 0), (1, 1)]
    vto, row_ptrs, cols = random_typed_graph([700, 1500], edge_types, [9000, 14000, 60000], seed=seed, col_dtype=col_dtype)
    rng = np.random.default_rng(4)
    seeds = np.concatenate([rng.integers(0, 2200, 60), rng.integers(700, 2200, 33), rng.integers(0, 700, 1)]).astype(np.int64)
    seeds[7] = seeds[3]
    lo = np.array([0, 60, 60, 93, 94], dtype=np.int64)
    return vto, row_ptrs, cols, seeds, lo, rng


# ---- calibration: the GPU-verified paths through the emulator ----------------------------------------------------------------
@pytest.mark.parametrize("fanout,flags", [([25, 10], FLAG_INT64), ([4, 3, 2], FLAG_CSR | FLAG_INT64), ([40, 5], 0), ([-1, 2], FLAG_INT64)])
def test_emulated_plain_homogeneous_matches_oracle(emu, oracle, fanout, flags):
    vto, row_ptrs, cols = random_typed_graph([3000], [(0, 0)], [90000], seed=11)
    rng = np.random.default_rng(1)
    seeds = rng.integers(0, 3000, 200).astype(np.int64)
    lo = np.array([0, 64, 128, 200], dtype=np.int64)
    got = _run(emu, row_ptrs, cols, vto, seeds, lo, fanout, 62, hetero=False, flags=flags)
    exp = oracle.multihop_sample(row_ptrs[0], cols[0], seeds, lo, fanout, 62)
    _same(got, exp, ["minors", "edge_id", "renumber_map", "renumber_map_offsets"])
    if flags & FLAG_CSR:
        assert np.array_equal(_majors_of_csr(got, len(fanout), 3), exp["majors"])
    else:
        _same(got, exp, ["majors", "label_hop_offsets"])


@pytest.mark.parametrize("fanout", [[3, 2, 4, 2, 2, 2], [40, 33, 35, -1, 2, 50], [0, 0, 0, 4, 4, 4]])
def test_emulated_plain_heterogeneous_matches_oracle(emu, oracle, fanout):
    vto, row_ptrs, cols, seeds, lo, _ = _typed_case()
    got = _run(emu, row_ptrs, cols, vto, seeds, lo, fanout, 31, reps=2)
    _same(got, oracle.hetero_multihop_sample(row_ptrs, cols, vto, seeds, lo, fanout, 31), HETERO)


# ---- the temporal path (not yet run on a GPU) -----------------------------------------------------------------------------
@pytest.mark.parametrize("comparison", range(4))
@pytest.mark.parametrize("fanout", [[3, 2, 4, 2, 2, 2], [-1, 3, 0, 2, -1, 1], [5, 5, 5], [0, 0, 0, 4, 4, 4], [2, 40, 1], [33, 70, 8, 3, 3, 3]])
@pytest.mark.parametrize("col_dtype", [np.int32, np.int64])
def test_emulated_temporal_heterogeneous_matches_oracle(emu, oracle, comparison, fanout, col_dtype):
    """The configurations of tests/test_gpu_temporal.py::test_temporal_hetero_bit_exact_vs_oracle."""
    vto, row_ptrs, cols, seeds, lo, rng = _typed_case(col_dtype, seed=len(fanout))
    times = [rng.integers(0, 50, c.shape[0]).astype(np.int64) for c in cols]
    eids = [rng.permutation(c.shape[0]).astype(np.int64) + 1000000 * t for t, c in enumerate(cols)]
    mid = 10 if comparison < 2 else 40
    seed_times = (mid + rng.integers(-5, 6, seeds.shape[0])).astype(np.int64)
    if col_dtype is np.int64 and comparison != 0 and os.environ.get("WGB_EMU_FULL") != "1":
        pytest.skip("64-bit columns are swept with one comparison by default; WGB_EMU_FULL=1 runs all")
    got = _run(emu, row_ptrs, cols, vto, seeds, lo, fanout, 31, times=times, seed_times=seed_times, cmp=comparison, eids=eids,
               reps=2 if comparison == 0 else 1)
    exp = oracle.temporal_multihop_sample(row_ptrs, cols, times, vto, seeds, seed_times, lo, fanout, 31, COMPARISONS[comparison], edge_ids=eids)
    assert exp["majors"].shape[0] > 0 or not any(fanout[:3])  # hop 0 with fan-out 0 for every type: nothing is sampled
    _same(got, exp, HETERO)


@pytest.mark.parametrize("fanout,flags", [([4, 3], FLAG_INT64), ([40, 5], 0), ([-1, 2], FLAG_INT64)])
def test_emulated_temporal_homogeneous_matches_oracle(emu, oracle, fanout, flags):
    vto, row_ptrs, cols = random_typed_graph([3000], [(0, 0)], [90000], seed=11)
    rng = np.random.default_rng(1)
    times = [rng.integers(0, 1000, cols[0].shape[0]).astype(np.int64)]
    seeds = rng.integers(0, 3000, 200).astype(np.int64)
    seed_times = rng.integers(300, 700, 200).astype(np.int64)
    lo = np.array([0, 64, 128, 200], dtype=np.int64)
    got = _run(emu, row_ptrs, cols, vto, seeds, lo, fanout, 62, hetero=False, times=times, seed_times=seed_times, cmp=3, flags=flags)
    exp = oracle.temporal_multihop_sample(row_ptrs, cols, times, vto, seeds, seed_times, lo, fanout, 62, "monotonically_decreasing")
    assert np.array_equal(got["majors"], exp["majors"]) and np.array_equal(got["minors"], exp["minors"])
    assert np.array_equal(got["edge_id"], exp["edge_renumber_map"])
    assert np.array_equal(got["renumber_map"], exp["renumber_map"]) and np.array_equal(got["renumber_map_offsets"], exp["renumber_map_offsets"])
    assert np.array_equal(got["label_hop_offsets"], exp["label_type_hop_offsets"])
    assert np.array_equal(got["label_step_base"], exp["label_type_step_base"].reshape(-1))


def test_emulated_temporal_csr_output(emu, oracle):
    """Temporal + CSR compression (what NeighborLoader asks for on a homogeneous graph): major_offsets against the COO result."""
    vto, row_ptrs, cols = random_typed_graph([800], [(0, 0)], [20000], seed=3)
    rng = np.random.default_rng(5)
    times = [rng.integers(0, 100, cols[0].shape[0]).astype(np.int64)]
    seeds = rng.permutation(800)[:96].astype(np.int64)
    seed_times = rng.integers(30, 70, 96).astype(np.int64)
    lo = np.array([0, 32, 64, 96], dtype=np.int64)
    coo = _run(emu, row_ptrs, cols, vto, seeds, lo, [5, 3], 9, hetero=False, times=times, seed_times=seed_times, cmp=1)
    csr = _run(emu, row_ptrs, cols, vto, seeds, lo, [5, 3], 9, hetero=False, times=times, seed_times=seed_times, cmp=1, flags=FLAG_CSR | FLAG_INT64)
    assert np.array_equal(coo["minors"], csr["minors"]) and np.array_equal(coo["edge_id"], csr["edge_id"])
    assert np.array_equal(_majors_of_csr(csr, 2, 3), coo["majors"])
    assert csr["major_offsets"][-1] == coo["label_hop_offsets"][-1] == coo["majors"].shape[0] > 0


def test_emulated_temporal_open_window_equals_plain(emu):
    vto, row_ptrs, cols, seeds, lo, _ = _typed_case()
    times = [np.ones(c.shape[0], dtype=np.int64) for c in cols]
    for fanout in ([10, 10, 5, 5, 5, 5], [40, 3, 2, 2, 2, 2], [-1, 2, 1, 1, 1, 1]):
        plain = _run(emu, row_ptrs, cols, vto, seeds, lo, fanout, 5)
        temp = _run(emu, row_ptrs, cols, vto, seeds, lo, fanout, 5, times=times, seed_times=np.ones_like(seeds), cmp=1)
        _same(temp, plain, HETERO)


def test_emulated_temporal_argument_checks(emu):
    vto, row_ptrs, cols = random_typed_graph([50], [(0, 0)], [300], seed=2)
    seeds, lo = np.arange(10, dtype=np.int64), np.array([0, 10], dtype=np.int64)
    tm = [np.zeros(300, dtype=np.int64)]
    _run(emu, row_ptrs, cols, vto, seeds, lo, [2], 1, hetero=False, times=tm, seed_times=np.zeros(10, np.int64), cmp=7, expect_rc=6)  # INVALID_INPUT
    _run(emu, row_ptrs, cols, vto, seeds, lo, [2000], 1, hetero=False, times=tm, seed_times=np.zeros(10, np.int64), cmp=0, expect_rc=2)  # NOT_IMPLEMENTED


# ---- biased temporal --------------------------------------------------------------------------------------------------------
def _rows_as_sets(out, T, L, B):
    """{(label, type, hop, major): sorted original edge ids} of a heterogeneous result."""
    lto, groups = out["label_type_hop_offsets"], {}
    for l in range(B):
        for t in range(T):
            for h in range(L):
                a, b = lto[(l * T + t) * L + h], lto[(l * T + t) * L + h + 1]
                for m, e in zip(out["majors"][a:b].tolist(), out["edge_renumber_map"][a:b].tolist()):
                    groups.setdefault((l, t, h, m), []).append(e)
    return {k: sorted(v) for k, v in groups.items()}


@pytest.mark.parametrize("wdtype", [np.float32, np.float64])
def test_emulated_biased_temporal_open_window_equals_plain_biased(emu, wdtype):
    """Every edge eligible: the masked A-Res kernel must be weighted_kernel (same draws, same candidates, same order), over
    two hops and three edge types."""
    vto, row_ptrs, cols, seeds, lo, rng = _typed_case()
    wts = [(rng.random(c.shape[0]) + 0.01).astype(wdtype) for c in cols]
    times = [np.full(c.shape[0], 3, dtype=np.int64) for c in cols]
    for fanout in ([4, 3, 2, 2, 2, 2],) if wdtype is np.float64 else ([4, 3, 2, 2, 2, 2], [40, 3, 20, 1, 1, 1]):
        if fanout[0] == 40:  # the A-Res kernel sorts 2048 candidates per heavy row: keep the emulated call group small
            seeds, lo = seeds[:24], np.array([0, 12, 24], dtype=np.int64)
        plain = _run(emu, row_ptrs, cols, vto, seeds, lo, fanout, 5, weights=wts)
        temp = _run(emu, row_ptrs, cols, vto, seeds, lo, fanout, 5, weights=wts, times=times, seed_times=np.full_like(seeds, 3), cmp=3)
        assert plain["majors"].shape[0] > 0
        _same(temp, plain, HETERO)


@pytest.mark.parametrize("comparison", [0, 3])
def test_emulated_biased_temporal_one_hop_sets_match_oracle(emu, oracle, comparison):
    """One hop: per frontier row the SET of sampled edges equals the oracle's (the reference compares weighted samples as
    sets; the oracle lists a row in ascending key order, the kernel in descending order)."""
    vto, row_ptrs, cols, seeds, lo, rng = _typed_case()
    wts = [(rng.random(c.shape[0]) + 0.01).astype(np.float32) for c in cols]
    times = [rng.integers(0, 50, c.shape[0]).astype(np.int64) for c in cols]
    seed_times = ((10 if comparison < 2 else 40) + rng.integers(-5, 6, seeds.shape[0])).astype(np.int64)
    fanout = [4, 40, 3]
    got = _run(emu, row_ptrs, cols, vto, seeds, lo, fanout, 77, weights=wts, times=times, seed_times=seed_times, cmp=comparison)
    exp = oracle.temporal_multihop_sample(row_ptrs, cols, times, vto, seeds, seed_times, lo, fanout, 77, COMPARISONS[comparison], weights=wts)
    _same(got, exp, ["label_type_hop_offsets", "majors", "edge_type", "edge_id", "edge_renumber_map_offsets"])  # counts are data only
    a, b = _rows_as_sets(got, 3, 1, 4), _rows_as_sets(exp, 3, 1, 4)
    assert a == b and len(a) > 100
    assert sorted(got["renumber_map"].tolist()) == sorted(exp["renumber_map"].tolist())


def test_emulated_biased_temporal_two_hops_are_valid(emu):
    """Two hops, two edge types over ONE vertex type (local ids are then frontier rows, so the hop's edge order -- source row,
    edge type, slot -- can be rebuilt from the output): every sampled edge is eligible with respect to the time its source
    carries (seed time / time of the edge that reached it first), rows hold min(eligible, fan-out) distinct edges."""
    vto, row_ptrs, cols = random_typed_graph([1500], [(0, 0), (0, 0)], [20000, 45000], seed=21)
    rng = np.random.default_rng(8)
    T, L, B = 2, 2, 3
    seeds = rng.integers(0, 1500, 90).astype(np.int64)
    lo = np.array([0, 30, 60, 90], dtype=np.int64)
    wts = [(rng.random(c.shape[0]) + 0.01).astype(np.float32) for c in cols]
    times = [rng.integers(0, 50, c.shape[0]).astype(np.int64) for c in cols]
    seed_times = (10 + rng.integers(-5, 6, seeds.shape[0])).astype(np.int64)
    fanout = [3, 5, 4, 2]
    out = _run(emu, row_ptrs, cols, vto, seeds, lo, fanout, 9, weights=wts, times=times, seed_times=seed_times, cmp=1)
    lto, rmo = out["label_type_hop_offsets"], out["renumber_map_offsets"]
    checked = 0
    for l in range(B):
        gmap = out["renumber_map"][rmo[l]:rmo[l + 1]]
        vtime = {}
        for s in range(lo[l], lo[l + 1]):
            vtime.setdefault(int(seeds[s]), int(seed_times[s]))  # first occurrence of a repeated seed
        for h in range(L):
            arrivals = []
            for t in range(T):
                a, b = lto[(l * T + t) * L + h], lto[(l * T + t) * L + h + 1]
                src, dst = gmap[out["majors"][a:b]], gmap[out["minors"][a:b]]
                pos = out["edge_renumber_map"][a:b]  # no edge ids given: CSR positions of type t
                assert np.array_equal(cols[t][pos].astype(np.int64), dst)
                assert ((row_ptrs[t][src] <= pos) & (pos < row_ptrs[t][src + 1])).all()
                for u in np.unique(src):
                    sel = pos[src == u]
                    assert len(np.unique(sel)) == len(sel)
                    row_t = times[t][row_ptrs[t][u]:row_ptrs[t][u + 1]]
                    eligible = int((row_t >= vtime[int(u)]).sum())
                    assert len(sel) == min(eligible, fanout[h * T + t]) and (times[t][sel] >= vtime[int(u)]).all()
                    checked += 1
                arrivals += [(int(m), t, i, int(d), int(times[t][p])) for i, (m, d, p) in enumerate(zip(out["majors"][a:b], dst, pos))]
            for _, _, _, d, tm in sorted(arrivals):  # the hop's edge order: a new vertex keeps the time of its first arrival
                vtime.setdefault(d, tm)
    assert checked > 150


def test_emulated_temporal_edge_cases(emu, oracle):
    """Empty call, empty labels, nothing eligible anywhere, and the largest fan-out (1024) on rows heavier than it."""
    rng = np.random.default_rng(12)
    deg = np.concatenate([rng.integers(0, 6, 60), [1500, 2600, 0, 1100]])
    row_ptr = np.concatenate([[0], np.cumsum(deg)]).astype(np.int64)
    V = deg.shape[0]
    col = rng.integers(0, V, int(row_ptr[-1])).astype(np.int32)
    times = [rng.integers(0, 10, col.shape[0]).astype(np.int64)]
    vto = np.array([0, V], dtype=np.int64)
    # no seeds at all / labels without seeds
    out = _run(emu, [row_ptr], [col], vto, np.empty(0, np.int64), [0, 0, 0], [3, 2], 1, times=times, seed_times=np.empty(0, np.int64), cmp=1)
    assert out["majors"].shape[0] == 0 and out["renumber_map"].shape[0] == 0 and out["label_type_hop_offsets"].tolist() == [0] * 5
    seeds = np.array([60, 61, 62, 63, 5, 61], dtype=np.int64)
    lo = np.array([0, 0, 4, 4, 6], dtype=np.int64)
    # nothing eligible: every edge time is < 100
    out = _run(emu, [row_ptr], [col], vto, seeds, lo, [5, 5], 3, times=times, seed_times=np.full(6, 100, np.int64), cmp=0)
    exp = oracle.temporal_multihop_sample([row_ptr], [col], times, vto, seeds, np.full(6, 100, np.int64), lo, [5, 5], 3, "strictly_increasing")
    assert out["majors"].shape[0] == 0
    _same(out, exp, HETERO)
    # fan-out 1024 (the largest the samplers accept) on rows with 1500 / 2600 / 1100 edges, about half of them eligible
    st = np.array([4, 5, 5, 4, 5, 3], dtype=np.int64)
    for fanout in ([1024, 2], [700, 1]):
        out = _run(emu, [row_ptr], [col], vto, seeds, lo, fanout, 7, times=times, seed_times=st, cmp=1)
        exp = oracle.temporal_multihop_sample([row_ptr], [col], times, vto, seeds, st, lo, fanout, 7, "monotonically_increasing")
        assert exp["majors"].shape[0] > 1500
        _same(out, exp, HETERO)


def test_emulated_chunked_graph(emu, oracle):
    """CSR arrays presented as CHUNKED over 3 ranks (tests/emu/emu_runtime.cpp: emu_set_split_world): the CHUNKED template
    variants of every sampler kernel -- plain and temporal -- with their per-access owner lookup."""
    vto, row_ptrs, cols, seeds, lo, rng = _typed_case()
    times = [rng.integers(0, 50, c.shape[0]).astype(np.int64) for c in cols]
    eids = [rng.permutation(c.shape[0]).astype(np.int64) for c in cols]
    seed_times = (10 + rng.integers(-5, 6, seeds.shape[0])).astype(np.int64)
    emu.emu_set_split_world(3)
    try:
        got = _run(emu, row_ptrs, cols, vto, seeds, lo, [3, 40, -1, 2, 2, 2], 31, eids=eids)
        tgot = _run(emu, row_ptrs, cols, vto, seeds, lo, [3, 40, -1, 2, 2, 2], 31, times=times, seed_times=seed_times, cmp=1, eids=eids)
    finally:
        emu.emu_set_split_world(1)
    _same(got, oracle.hetero_multihop_sample(row_ptrs, cols, vto, seeds, lo, [3, 40, -1, 2, 2, 2], 31, edge_ids=eids), HETERO)
    _same(tgot, oracle.temporal_multihop_sample(row_ptrs, cols, times, vto, seeds, seed_times, lo, [3, 40, -1, 2, 2, 2], 31,
                                                "monotonically_increasing", edge_ids=eids), HETERO)


def test_emulated_c1_karate(emu, oracle):
    """BASELINE.json configs[0] ("karate.csv 1-hop fanout=[5] on CPU, bit-exact COO check, no GPU") with the PRODUCT's sampler
    source run on the CPU: every vertex a seed, one label, sampler seed 62 -- COO, renumber map and edge ids against the oracle;
    the same call is checked on the GPU in tests/test_gpu_multihop.py::test_multihop_c1_karate_and_large_properties."""
    from graphs import karate_csr

    row_ptr, col = karate_csr(np.int64)
    seeds = np.arange(34, dtype=np.int64)
    lo = np.array([0, 34], dtype=np.int64)
    got = _run(emu, [row_ptr], [col], [0, 34], seeds, lo, [5], 62, hetero=False)
    exp = oracle.multihop_sample(row_ptr, col, seeds, lo, [5], 62)
    _same(got, exp, ["majors", "minors", "edge_id", "label_hop_offsets", "renumber_map", "renumber_map_offsets"])
    deg = np.diff(row_ptr)
    assert got["majors"].shape[0] == int(np.minimum(deg, 5).sum())
    src = got["renumber_map"][got["majors"]]
    assert np.array_equal(col[got["edge_id"]], got["renumber_map"][got["minors"]]) and np.array_equal(np.searchsorted(row_ptr, got["edge_id"], "right") - 1, src)
