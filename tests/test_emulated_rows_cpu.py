"""The gather / scatter kernels (csrc/gather_scatter.cu, G1/G2 -- the dominant kernel of the bench) and the CSR aggregation
(csrc/aggregate.cu, A1) through the CPU emulator of tests/emu, with every operand in an exactly-sized, guard-fenced
allocation: besides the values (bit-exact for same-dtype copies), this checks that no vector width / alignment / tail case
reads or writes outside its table, index array or output -- something a GPU run does not show.  `tests/emu/run_asan.sh`
repeats it under AddressSanitizer (reads included).  Test infrastructure only; these kernels are GPU-verified already
(tests/test_gpu_gather.py, test_gpu_aggregate.py).
"""
import ctypes
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
VP = ctypes.c_void_p
DT = {np.dtype(np.float32): 1, np.dtype(np.float16): 2, np.dtype(np.float64): 3, np.dtype(np.int32): 5, np.dtype(np.int64): 6,
      np.dtype(np.int16): 7, np.dtype(np.int8): 8}


@pytest.fixture(scope="module")
def emu():
    sys.path.insert(0, os.path.join(HERE, "emu"))
    import build_emu

    if not build_emu.available():
        pytest.skip("CUDA headers not installed")
    lib = ctypes.CDLL(build_emu.build_rows())
    lib.emu_rows_op.restype = ctypes.c_int
    lib.emu_csr_aggregate.restype = ctypes.c_int
    return lib


def _p(a):
    return None if a is None else a.ctypes.data_as(VP)


def _rows_op(lib, storage, rows, dim, stride, off, idx, dense, scatter=False, hot_slot=None, hot_rows=None, sms=-1):
    ll = ctypes.c_longlong
    return lib.emu_rows_op(_p(storage), ll(rows), ll(dim), ll(stride), ll(off), DT[storage.dtype], _p(idx), ll(idx.shape[0]),
                           int(idx.dtype == np.int64), _p(dense), ll(dense.shape[1]), DT[dense.dtype], int(scatter), _p(hot_slot), _p(hot_rows),
                           ll(0 if hot_rows is None else hot_rows.shape[0]), sms)


def _table(rng, rows, dim, stride, off, dtype):
    storage = np.zeros(off + rows * stride, dtype=dtype)
    view = storage[off:].reshape(rows, stride)[:, :dim]
    if np.issubdtype(dtype, np.floating):
        view[...] = rng.standard_normal((rows, dim)).astype(dtype)
    else:
        view[...] = rng.integers(-100, 100, (rows, dim)).astype(dtype)
    return storage, view


@pytest.mark.parametrize("dim,stride,off", [(1, 1, 0), (3, 3, 0), (4, 4, 0), (4, 6, 2), (32, 32, 0), (33, 33, 0), (33, 40, 1), (128, 128, 0),
                                            (128, 128, 4), (130, 136, 0), (257, 257, 3)])
@pytest.mark.parametrize("dtype", [np.float32, np.float16, np.int64, np.int8])
@pytest.mark.parametrize("world", [1, 3])
def test_gather_same_dtype_all_alignments(emu, dim, stride, off, dtype, world):
    """Same-dtype gather is byte movement: every (row bytes, stride, base offset) class that selects a different vector
    width, with 32- and 64-bit indices, repeated and boundary rows (first / last row of the table)."""
    rng = np.random.default_rng(dim * 7 + off)
    rows = 97
    storage, view = _table(rng, rows, dim, stride, off, dtype)
    for idt in (np.int32, np.int64):
        idx = np.concatenate([[0, rows - 1, rows - 1, 0], rng.integers(0, rows, 300)]).astype(idt)
        out = np.full((idx.shape[0], dim), 77, dtype=dtype)
        emu.emu_set_split_world(world)  # > 1: the table is presented as CHUNKED over `world` ranks (owner lookup per row)
        try:
            rc = _rows_op(emu, storage, rows, dim, stride, off, idx, out)
        finally:
            emu.emu_set_split_world(1)
        assert rc == 0
        assert np.array_equal(out, view[idx])


@pytest.mark.parametrize("src,dst", [(np.float16, np.float32), (np.float32, np.float16), (np.float32, np.float64), (np.int8, np.int32), (np.int64, np.int32)])
def test_gather_converting(emu, src, dst):
    rng = np.random.default_rng(5)
    rows, dim = 50, 19
    storage, view = _table(rng, rows, dim, dim + 1, 2, src)
    idx = rng.integers(0, rows, 120).astype(np.int64)
    out = np.zeros((120, dim), dtype=dst)
    assert _rows_op(emu, storage, rows, dim, dim + 1, 2, idx, out) == 0
    assert np.array_equal(out, view[idx].astype(dst))


@pytest.mark.parametrize("dim,stride,off", [(4, 4, 0), (33, 40, 1), (128, 128, 0)])
def test_scatter(emu, dim, stride, off):
    rng = np.random.default_rng(dim)
    rows = 64
    storage, view = _table(rng, rows, dim, stride, off, np.float32)
    before = storage.copy()
    idx = rng.permutation(rows)[:40].astype(np.int64)  # distinct rows: the result does not depend on the write order
    dense = rng.standard_normal((40, dim)).astype(np.float32)
    assert _rows_op(emu, storage, rows, dim, stride, off, idx, dense, scatter=True) == 0
    exp = before.copy()
    exp[off:].reshape(rows, stride)[idx, :dim] = dense
    assert np.array_equal(storage, exp)  # padding between rows and the leading offset are untouched


def test_gather_with_replicated_hot_rows(emu):
    """wholememory_embedding_set_hot_rows: rows with a slot are served from the replica, bit for bit the same output."""
    rng = np.random.default_rng(9)
    rows, dim = 200, 128
    storage, view = _table(rng, rows, dim, dim, 0, np.float32)
    hot_ids = rng.permutation(rows)[:30]
    slot = np.full(rows, -1, dtype=np.int32)
    slot[hot_ids] = np.arange(30, dtype=np.int32)
    hot = np.ascontiguousarray(view[hot_ids])
    poisoned = storage.copy()
    poisoned.reshape(rows, dim)[hot_ids] = np.nan  # a hot row must come from the replica, not from the table
    idx = np.concatenate([hot_ids[:10], rng.integers(0, rows, 500)]).astype(np.int64)
    out = np.zeros((idx.shape[0], dim), dtype=np.float32)
    emu.emu_set_split_world(2)  # the replica is only consulted for tables that span more than one rank
    try:
        rc = _rows_op(emu, poisoned, rows, dim, dim, 0, idx, out, hot_slot=slot, hot_rows=hot)
    finally:
        emu.emu_set_split_world(1)
    assert rc == 0
    assert np.array_equal(out, view[idx])


def test_gather_with_an_sm_budget(emu):
    """gather_sms (bench.py --gather-sms): a smaller grid walks the same rows."""
    rng = np.random.default_rng(3)
    storage, view = _table(rng, 300, 64, 64, 0, np.float32)
    idx = rng.integers(0, 300, 5000).astype(np.int64)
    out = np.zeros((5000, 64), dtype=np.float32)
    assert _rows_op(emu, storage, 300, 64, 64, 0, idx, out, sms=1) == 0
    assert np.array_equal(out, view[idx])


@pytest.mark.parametrize("F,xdtype", [(4, np.float32), (32, np.float32), (36, np.float32), (128, np.float32), (160, np.float32), (8, np.float16),
                                      (136, np.float16)])
@pytest.mark.parametrize("reduce", [0, 1])
@pytest.mark.parametrize("use_map,idx64", [(False, False), (True, True)])
def test_csr_aggregate(emu, F, xdtype, reduce, use_map, idx64):
    rng = np.random.default_rng(F + reduce)
    n_dst, n_src = 70, 150
    deg = rng.integers(0, 9, n_dst)
    deg[3] = 0
    deg[10] = 40
    indptr = np.concatenate([[0], np.cumsum(deg)]).astype(np.int64 if idx64 else np.int32)
    nnz = int(indptr[-1])
    n_x = 400 if use_map else n_src
    mp = rng.integers(0, n_x, n_src).astype(np.int64) if use_map else None
    indices = rng.integers(0, n_src, nnz).astype(np.int64 if idx64 else np.int32)
    x = rng.standard_normal((n_x, F)).astype(xdtype)
    out = np.full((n_dst, F), 5.0, dtype=np.float32)
    ll = ctypes.c_longlong
    rc = emu.emu_csr_aggregate(_p(indptr), int(idx64), ll(n_dst), _p(indices), int(idx64), ll(nnz), _p(mp), ll(n_src if use_map else 0), _p(x), ll(n_x),
                               ll(F), DT[x.dtype], reduce, _p(out))
    assert rc == 0
    rows = indices if mp is None else mp[indices]
    exp = np.zeros((n_dst, F), dtype=np.float64)
    for i in range(n_dst):
        seg = x[rows[indptr[i]:indptr[i + 1]]].astype(np.float64)
        if seg.shape[0]:
            exp[i] = seg.sum(0) / (seg.shape[0] if reduce == 1 else 1)
    assert np.allclose(out, exp, rtol=1e-3, atol=1e-4)  # north_star: within 1e-3 relative for fp32 aggregation


@pytest.mark.parametrize("dtype", [np.int32, np.int64])
@pytest.mark.parametrize("T,N,span", [(0, 0, 10), (5, 0, 10), (0, 7, 4), (300, 5000, 900), (1, 3000, 50), (700, 700, 100000)])
def test_append_unique(emu, oracle, dtype, T, N, span):
    """S3 (graph_append_unique): targets keep ids 0..T-1, every new neighbour gets the next id in first-occurrence order."""
    rng = np.random.default_rng(T + N)
    targets = rng.permutation(span)[:T].astype(dtype) if T <= span else rng.integers(0, span, T).astype(dtype)
    neighbors = rng.integers(0, span, N).astype(dtype)
    uniq = np.full(T + N + 1, -5, dtype=dtype)
    r2u = np.full(max(N, 1), -5, dtype=np.int32)
    emu.emu_append_unique.restype = ctypes.c_longlong
    cnt = emu.emu_append_unique(_p(targets), ctypes.c_longlong(T), _p(neighbors), ctypes.c_longlong(N), int(dtype is np.int64), _p(uniq), _p(r2u))
    e_uniq, e_r2u = oracle.append_unique(targets, neighbors)
    assert cnt == e_uniq.shape[0]
    assert np.array_equal(uniq[:cnt], e_uniq) and uniq[cnt] == -5
    assert np.array_equal(r2u[:N], e_r2u)
