"""Kernel LOGIC on the CPU: the product's temporal one-hop kernels (csrc/temporal_device.cuh) compiled with g++ against
the SIMT shim in tests/emu/cuda_emu.h and compared with the oracle.

Why: the temporal path was written after the round's GPU minutes were spent, so its GPU parity tests
(tests/test_gpu_temporal.py) have not run yet.  This does not replace them -- it checks indexing, barrier placement and
the random-stream geometry, not memory ordering or anything about the hardware -- and it is test infrastructure only: the
emulated code is never linked into the product.  The already GPU-verified plain path (count_scan_kernel +
uniform_general_kernel) is run through the same shim first, which validates the shim itself against the oracle.
"""
import ctypes
import os

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
COMPARISONS = ("strictly_increasing", "monotonically_increasing", "strictly_decreasing", "monotonically_decreasing")


@pytest.fixture(scope="module")
def emu():
    import sys

    sys.path.insert(0, os.path.join(HERE, "emu"))
    import build_emu

    if not build_emu.available():
        pytest.skip("CUDA headers not installed")
    return ctypes.CDLL(build_emu.build_kernels())


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def _heavy_graph(seed=5):
    rng = np.random.default_rng(seed)
    deg = np.concatenate([rng.integers(0, 12, 150), rng.integers(30, 80, 40), rng.integers(130, 400, 10)])
    rng.shuffle(deg)
    row_ptr = np.concatenate([[0], np.cumsum(deg)]).astype(np.int64)
    col = rng.integers(0, deg.shape[0], int(row_ptr[-1])).astype(np.int64)
    return row_ptr, col


def _temporal_hop(emu, row_ptr, col, etime, centers, ftime, M, cmp, seed, grids=(3, 2, 5)):
    n = centers.shape[0]
    cap = int((row_ptr[1:] - row_ptr[:-1]).max()) * n + 16
    offsets = np.full(n + 1, -7, dtype=np.int32)
    elig = np.full(n + 1, -7, dtype=np.int32)
    dest = np.full(cap, -1, dtype=np.int64)
    lid = np.full(cap, -1, dtype=np.int32)
    gid = np.full(cap, -1, dtype=np.int64)
    emu.emu_temporal_hop.restype = ctypes.c_int
    tot = emu.emu_temporal_hop(_p(row_ptr), ctypes.c_longlong(row_ptr.shape[0] - 1), _p(col), _p(etime), ctypes.c_longlong(col.shape[0]), _p(centers),
                               _p(ftime), ctypes.c_int(n), ctypes.c_int(M), ctypes.c_int(cmp), ctypes.c_ulonglong(seed), ctypes.c_int(grids[0]),
                               ctypes.c_int(grids[1]), ctypes.c_int(grids[2]), _p(offsets), _p(elig), _p(dest), _p(lid), _p(gid))
    assert tot >= 0 and tot == offsets[n]
    assert (dest[tot:] == -1).all() and (gid[tot:] == -1).all(), "wrote past the end of the hop's edge list"
    return offsets, elig[:n], dest[:tot], lid[:tot], gid[:tot]


@pytest.mark.parametrize("M", [40, 100])
def test_shim_reproduces_the_verified_plain_path(emu, oracle, M):
    row_ptr, col = _heavy_graph()
    centers = np.random.default_rng(1).integers(0, row_ptr.shape[0] - 1, 1100).astype(np.int64)  # 2 scan tiles
    n = centers.shape[0]
    cap = M * n + 16
    offsets = np.zeros(n + 1, dtype=np.int32)
    dest, lid, gid = np.full(cap, -1, dtype=np.int64), np.full(cap, -1, dtype=np.int32), np.full(cap, -1, dtype=np.int64)
    emu.emu_plain_hop.restype = ctypes.c_int
    tot = emu.emu_plain_hop(_p(row_ptr), ctypes.c_longlong(row_ptr.shape[0] - 1), _p(col), ctypes.c_longlong(col.shape[0]), _p(centers), ctypes.c_int(n),
                            ctypes.c_int(M), ctypes.c_ulonglong(99), ctypes.c_int(2), ctypes.c_int(7), _p(offsets), _p(dest), _p(lid), _p(gid))
    eoff, edest, elid, egid = oracle.unweighted_sample(row_ptr, col, centers, M, 99)
    assert tot == eoff[-1]
    assert np.array_equal(offsets, eoff)
    assert np.array_equal(gid[:tot], egid) and np.array_equal(dest[:tot], edest) and np.array_equal(lid[:tot], elid)


@pytest.mark.parametrize("M", [1, 8, 32, 33, 100, -1])
def test_temporal_kernels_open_window_equal_plain_sampling(emu, oracle, M):
    """Every edge eligible: the temporal kernels must reproduce the plain S1 sampler bit for bit -- for fan-outs <= 32 this
    pins that the general chain with T = 32 threads and one draw each has the geometry of uniform_small_kernel."""
    row_ptr, col = _heavy_graph()
    centers = np.random.default_rng(2).integers(0, row_ptr.shape[0] - 1, 200).astype(np.int64)
    etime = np.full(col.shape[0], 7, dtype=np.int64)
    ftime = np.full(centers.shape[0], 7, dtype=np.int64)
    off, elig, dest, lid, gid = _temporal_hop(emu, row_ptr, col, etime, centers, ftime, M, 1, 1234)
    eoff, edest, elid, egid = oracle.unweighted_sample(row_ptr, col, centers, M, 1234)
    assert np.array_equal(elig, (row_ptr[centers + 1] - row_ptr[centers]).astype(np.int32))
    assert np.array_equal(off, eoff)
    assert np.array_equal(gid, egid) and np.array_equal(dest, edest) and np.array_equal(lid, elid)


@pytest.mark.parametrize("comparison", range(4))
@pytest.mark.parametrize("M", [3, 32, 40, -1])
def test_temporal_kernels_match_the_oracle(emu, oracle, comparison, M):
    """One hop, one label, distinct seeds: the oracle's temporal multi-hop restatement (pinned on the reference's tests)
    reduces to exactly what the three kernels compute."""
    row_ptr, col = _heavy_graph(seed=8)
    V = row_ptr.shape[0] - 1
    rng = np.random.default_rng(comparison * 10 + (M & 0xFF))
    centers = rng.permutation(V)[:160].astype(np.int64)
    etime = rng.integers(0, 20, col.shape[0]).astype(np.int64)
    ftime = rng.integers(5, 15, centers.shape[0]).astype(np.int64)
    off, elig, dest, lid, gid = _temporal_hop(emu, row_ptr, col, etime, centers, ftime, M, comparison, 4242)
    # hop 0 of the multi-hop oracle draws with hop_seed(random_state, 0) + type 0 offset = random_state itself
    exp = oracle.temporal_multihop_sample([row_ptr], [col], [etime], [0, V], centers, ftime, [0, centers.shape[0]], [M], 4242, COMPARISONS[comparison])
    ok = {0: lambda e, v: e > v, 1: lambda e, v: e >= v, 2: lambda e, v: e < v, 3: lambda e, v: e <= v}[comparison]
    want = np.array([int(ok(etime[row_ptr[c]:row_ptr[c + 1]], t).sum()) for c, t in zip(centers, ftime)], dtype=np.int32)
    assert np.array_equal(elig, want)
    assert off[-1] == exp["majors"].shape[0] > 0
    assert np.array_equal(lid, exp["majors"])  # seeds are distinct: local id of a seed = its row
    assert np.array_equal(gid, exp["edge_renumber_map"])  # no edge ids given: CSR positions
    assert np.array_equal(dest, exp["renumber_map"][exp["minors"]])
    assert ok(etime[gid], ftime[lid]).all()


def test_temporal_scan_spans_tiles(emu, oracle):
    """More frontier rows than one scan tile (1024): the ticketed look-back of scan_counts_kernel, take-all fan-out."""
    row_ptr, col = _heavy_graph(seed=9)
    V = row_ptr.shape[0] - 1
    rng = np.random.default_rng(3)
    centers = rng.integers(0, V, 1100).astype(np.int64)
    etime = rng.integers(0, 20, col.shape[0]).astype(np.int64)
    ftime = rng.integers(5, 15, centers.shape[0]).astype(np.int64)
    off, elig, dest, lid, gid = _temporal_hop(emu, row_ptr, col, etime, centers, ftime, -1, 3, 1, grids=(4, 3, 9))
    want = np.array([int((etime[row_ptr[c]:row_ptr[c + 1]] <= t).sum()) for c, t in zip(centers, ftime)], dtype=np.int64)
    assert np.array_equal(elig, want.astype(np.int32))
    assert np.array_equal(off, np.concatenate([[0], np.cumsum(want)]).astype(np.int32))
    exp_gid = np.concatenate([row_ptr[c] + np.flatnonzero(etime[row_ptr[c]:row_ptr[c + 1]] <= t) for c, t in zip(centers, ftime)])
    assert np.array_equal(gid, exp_gid) and np.array_equal(dest, col[exp_gid])
    assert np.array_equal(lid, np.repeat(np.arange(centers.shape[0]), want).astype(np.int32))
