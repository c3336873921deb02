"""S1 / S2, the one-hop samplers behind the reference's own C entry points (csrc/sample.cu: host code and kernels), through
the CPU emulator of tests/emu against the oracle: uniform bit-exact (offsets, neighbours, centre-local ids, edge positions),
weighted as per-row sets (the reference's own criterion; emulator and oracle share libm, so the sets are exact here).  Covers
every fan-out class (sub-warp kernels G = 8 / 16 / 32, the one-CTA-per-row kernel, take-all), 32- and 64-bit ids, the CHUNKED
presentation, and -- through exactly-sized fenced operands -- that nothing is read or written out of bounds.  These paths are
GPU-verified (tests/test_gpu_sample*.py); this adds the memory-safety angle and runs without a GPU.  Test infrastructure only.
"""
import ctypes
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
VP = ctypes.c_void_p


@pytest.fixture(scope="module")
def emu():
    sys.path.insert(0, os.path.join(HERE, "emu"))
    import build_emu

    if not build_emu.available():
        pytest.skip("CUDA headers not installed")
    lib = ctypes.CDLL(build_emu.build_onehop())
    lib.emu_one_hop.restype = ctypes.c_longlong
    return lib


def _graph(col_dtype, seed=5):
    rng = np.random.default_rng(seed)
    deg = np.concatenate([rng.integers(0, 12, 150), rng.integers(30, 80, 40), rng.integers(130, 1300, 10)])
    rng.shuffle(deg)
    row_ptr = np.concatenate([[0], np.cumsum(deg)]).astype(np.int64)
    col = rng.integers(0, deg.shape[0], int(row_ptr[-1])).astype(col_dtype)
    return row_ptr, col


def _one_hop(lib, row_ptr, col, centers, M, seed, weight=None):
    n = centers.shape[0]
    cap = int((row_ptr[1:] - row_ptr[:-1]).max()) * n + 8
    off = np.full(n + 1, -7, dtype=np.int32)
    dest = np.full(cap, -1, dtype=col.dtype)
    lid = np.full(cap, -1, dtype=np.int32)
    gid = np.full(cap, -1, dtype=np.int64)
    ll = ctypes.c_longlong
    p = lambda a: None if a is None else a.ctypes.data_as(VP)  # noqa: E731
    tot = lib.emu_one_hop(p(row_ptr), ll(row_ptr.shape[0] - 1), p(col), ll(col.shape[0]), int(col.dtype == np.int64), p(weight),
                          int(weight is not None and weight.dtype == np.float64), p(centers), ll(n), int(centers.dtype == np.int64), M,
                          ctypes.c_ulonglong(seed), p(off), p(dest), p(lid), p(gid), ll(cap))
    assert tot >= 0, tot
    return off, dest[:tot], lid[:tot], gid[:tot]


@pytest.mark.parametrize("M", [1, 5, 8, 9, 16, 25, 32, 33, 64, 300, 1024, -1])
@pytest.mark.parametrize("col_dtype,center_dtype,world", [(np.int32, np.int32, 1), (np.int64, np.int64, 1), (np.int32, np.int64, 3)])
def test_uniform_one_hop_bit_exact(emu, oracle, M, col_dtype, center_dtype, world):
    row_ptr, col = _graph(col_dtype)
    centers = np.random.default_rng(M & 0xFF).integers(0, row_ptr.shape[0] - 1, 333).astype(center_dtype)
    lib = emu
    lib.emu_set_split_world(world)
    try:
        off, dest, lid, gid = _one_hop(lib, row_ptr, col, centers, M, 1234)
    finally:
        lib.emu_set_split_world(1)
    eoff, edest, elid, egid = oracle.unweighted_sample(row_ptr, col, centers, M, 1234)
    assert np.array_equal(off, eoff)
    assert np.array_equal(gid, egid) and np.array_equal(dest, edest) and np.array_equal(lid, elid)


def test_uniform_fanout_above_1024_is_refused(emu):
    row_ptr, col = _graph(np.int32)
    centers = np.arange(10, dtype=np.int64)
    with pytest.raises(AssertionError, match="-1002"):  # WHOLEMEMORY_NOT_IMPLEMENTED, as the header documents
        _one_hop(emu, row_ptr, col, centers, 1025, 1)


@pytest.mark.parametrize("M", [3, 40, 300, -1])
@pytest.mark.parametrize("wdtype", [np.float32, np.float64])
def test_weighted_one_hop_row_sets(emu, oracle, M, wdtype):
    row_ptr, col = _graph(np.int64, seed=7)
    rng = np.random.default_rng(M & 0xFF)
    w = (rng.random(col.shape[0]) + 0.01).astype(wdtype)
    centers = rng.integers(0, row_ptr.shape[0] - 1, 120).astype(np.int64)
    off, dest, lid, gid = _one_hop(emu, row_ptr, col, centers, M, 99, weight=w)
    eoff, edest, elid, egid = oracle.weighted_sample(row_ptr, col, w, centers, M, 99)[:4]
    assert np.array_equal(off, eoff) and np.array_equal(lid, elid)
    for b in range(centers.shape[0]):
        a, e = off[b], off[b + 1]
        assert sorted(gid[a:e].tolist()) == sorted(egid[a:e].tolist()), b
    assert np.array_equal(col[gid], dest)
