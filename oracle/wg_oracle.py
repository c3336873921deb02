"""TEST INFRASTRUCTURE ONLY: numpy/ctypes front-end of the CPU oracle (oracle/wg_oracle.cpp).

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference``
legs may import this module; the product package (``cugraph-gnn_b200/``) never does.
Every function is a thin typed wrapper; the algorithm and its reference citations live in
``wg_oracle.cpp``.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

DT_FLOAT, DT_HALF, DT_DOUBLE, DT_BF16, DT_INT, DT_INT64, DT_INT16, DT_INT8 = range(1, 9)

_NP2DT = {
    np.dtype(np.float32): DT_FLOAT,
    np.dtype(np.float16): DT_HALF,
    np.dtype(np.float64): DT_DOUBLE,
    np.dtype(np.int32): DT_INT,
    np.dtype(np.int64): DT_INT64,
    np.dtype(np.int16): DT_INT16,
    np.dtype(np.int8): DT_INT8,
}


def build(force=False):
    so = os.path.join(_HERE, "libwg_oracle.so")
    src = os.path.join(_HERE, "wg_oracle.cpp")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return so


def lib():
    global _LIB
    if _LIB is None:
        _LIB = ctypes.CDLL(build())
        _LIB.wgo_multihop_sample.restype = ctypes.c_void_p
        _LIB.wgo_multihop_num_edges.restype = ctypes.c_int64
        _LIB.wgo_multihop_num_nodes.restype = ctypes.c_int64
        _LIB.wgo_hop_seed.restype = ctypes.c_uint64
    return _LIB


def _p(a):
    return None if a is None else ctypes.c_void_p(a.ctypes.data)


def _dt(a, bf16=False):
    if bf16:
        return DT_BF16
    return _NP2DT[a.dtype]


def num_threads():
    return int(lib().wgo_num_threads())


def set_num_threads(n):
    """Size the OpenMP pool explicitly (torchrun exports OMP_NUM_THREADS=1 into every rank)."""
    lib().wgo_set_num_threads(ctypes.c_int(int(n)))
    return num_threads()


def pcg32_reference_stream(initstate, initseq, n):
    out = np.empty(n, dtype=np.uint32)
    lib().wgo_pcg32_reference_stream(ctypes.c_uint64(initstate), ctypes.c_uint64(initseq), _p(out), ctypes.c_int(n))
    return out


def generate_random_positive_int(seed, subsequence, n, dtype=np.int32):
    out = np.empty(n, dtype=dtype)
    fn = lib().wgo_generate_random_positive_int if dtype == np.int32 else lib().wgo_generate_random_positive_int64
    fn(ctypes.c_int64(seed), ctypes.c_int64(subsequence), _p(out), ctypes.c_int64(n))
    return out


def generate_exponential_distribution_negative_float(seed, subsequence, n):
    out = np.empty(n, dtype=np.float32)
    lib().wgo_generate_exponential_distribution_negative_float(
        ctypes.c_int64(seed), ctypes.c_int64(subsequence), _p(out), ctypes.c_int64(n)
    )
    return out


def sample_offsets(row_ptr, centers, max_sample_count):
    row_ptr = np.ascontiguousarray(row_ptr, dtype=np.int64)
    centers = np.ascontiguousarray(centers)
    n = centers.shape[0]
    off = np.empty(n + 1, dtype=np.int32)
    lib().wgo_sample_offsets(_p(row_ptr), _p(centers), _dt(centers), ctypes.c_int64(n), ctypes.c_int(max_sample_count), _p(off))
    return off


def unweighted_sample(row_ptr, col, centers, max_sample_count, seed):
    """Returns (sample_offset int32[n+1], dest col-dtype[total], center_localid int32[total], edge_gid int64[total])."""
    row_ptr = np.ascontiguousarray(row_ptr, dtype=np.int64)
    col = np.ascontiguousarray(col)
    centers = np.ascontiguousarray(centers)
    n = centers.shape[0]
    off = sample_offsets(row_ptr, centers, max_sample_count)
    total = int(off[n])
    dest = np.empty(total, dtype=col.dtype)
    lid = np.empty(total, dtype=np.int32)
    gid = np.empty(total, dtype=np.int64)
    rc = lib().wgo_unweighted_sample(
        _p(row_ptr), _p(col), _dt(col), _p(centers), _dt(centers), ctypes.c_int64(n),
        ctypes.c_int(max_sample_count), ctypes.c_uint64(seed), _p(off), _p(dest), _p(lid), _p(gid),
    )
    if rc != 0:
        raise ValueError("oracle: unsupported max_sample_count")
    return off, dest, lid, gid


def weighted_sample(row_ptr, col, weights, centers, max_sample_count, seed, return_keys=False):
    row_ptr = np.ascontiguousarray(row_ptr, dtype=np.int64)
    col = np.ascontiguousarray(col)
    weights = np.ascontiguousarray(weights)
    centers = np.ascontiguousarray(centers)
    n = centers.shape[0]
    off = sample_offsets(row_ptr, centers, max_sample_count)
    total = int(off[n])
    dest = np.empty(total, dtype=col.dtype)
    lid = np.empty(total, dtype=np.int32)
    gid = np.empty(total, dtype=np.int64)
    keys = np.empty(total, dtype=np.float32)
    rc = lib().wgo_weighted_sample(
        _p(row_ptr), _p(col), _dt(col), _p(weights), _dt(weights), _p(centers), _dt(centers), ctypes.c_int64(n),
        ctypes.c_int(max_sample_count), ctypes.c_uint64(seed), _p(off), _p(dest), _p(lid), _p(gid), _p(keys),
    )
    if rc != 0:
        raise ValueError("oracle: unsupported max_sample_count")
    if return_keys:
        return off, dest, lid, gid, keys
    return off, dest, lid, gid


def weighted_row_keys(row_ptr, weights, v, b, max_sample_count, seed):
    row_ptr = np.ascontiguousarray(row_ptr, dtype=np.int64)
    weights = np.ascontiguousarray(weights)
    deg = int(row_ptr[v + 1] - row_ptr[v])
    keys = np.empty(deg, dtype=np.float32)
    lib().wgo_weighted_row_keys(_p(row_ptr), _p(weights), _dt(weights), ctypes.c_int64(v), ctypes.c_int64(b),
                                ctypes.c_int(max_sample_count), ctypes.c_uint64(seed), _p(keys))
    return keys


def append_unique(targets, neighbors):
    targets = np.ascontiguousarray(targets)
    neighbors = np.ascontiguousarray(neighbors, dtype=targets.dtype)
    T, N = targets.shape[0], neighbors.shape[0]
    uniq = np.empty(T + N, dtype=targets.dtype)
    cnt = np.zeros(1, dtype=np.int32)
    r2u = np.empty(N, dtype=np.int32)
    lib().wgo_append_unique(_p(targets), ctypes.c_int64(T), _p(neighbors), ctypes.c_int64(N), _dt(targets), _p(uniq), _p(cnt), _p(r2u))
    return uniq[: int(cnt[0])].copy(), r2u


def gather(table, indices, out_dtype=None, dim=None, stride=None, storage_offset=0, table_bf16=False, out_bf16=False,
           out=None, out_stride=None):
    """table: 2-D (or flat with explicit dim/stride) numpy array; bf16 tables are passed as uint16 + table_bf16."""
    indices = np.ascontiguousarray(indices)
    n = indices.shape[0]
    if dim is None:
        dim = table.shape[1] if table.ndim == 2 else 1
    if stride is None:
        stride = table.shape[1] if table.ndim == 2 else 1
    table = np.ascontiguousarray(table)
    tdt = DT_BF16 if table_bf16 else _dt(table)
    if out is None:
        odt_np = np.dtype(out_dtype) if out_dtype is not None else table.dtype
        if out_bf16:
            odt_np = np.dtype(np.uint16)
        out_stride = dim if out_stride is None else out_stride
        out = np.zeros((n, out_stride), dtype=odt_np)
    else:
        out_stride = out.shape[1] if out_stride is None else out_stride
    odt = DT_BF16 if out_bf16 else _dt(out)
    rc = lib().wgo_gather(_p(table), tdt, ctypes.c_int64(dim), ctypes.c_int64(stride), ctypes.c_int64(storage_offset),
                          _p(indices), _dt(indices), ctypes.c_int64(n), _p(out), odt, ctypes.c_int64(out_stride), ctypes.c_int64(0))
    if rc != 0:
        raise ValueError("oracle: embedding and output must both be floating or both integer")
    return out


def scatter(inp, indices, table, dim=None, stride=None, storage_offset=0, in_bf16=False, table_bf16=False):
    indices = np.ascontiguousarray(indices)
    inp = np.ascontiguousarray(inp)
    n = indices.shape[0]
    if dim is None:
        dim = inp.shape[1] if inp.ndim == 2 else 1
    if stride is None:
        stride = table.shape[1] if table.ndim == 2 else 1
    in_stride = inp.shape[1] if inp.ndim == 2 else 1
    idt = DT_BF16 if in_bf16 else _dt(inp)
    tdt = DT_BF16 if table_bf16 else _dt(table)
    rc = lib().wgo_scatter(_p(inp), idt, ctypes.c_int64(dim), ctypes.c_int64(in_stride), ctypes.c_int64(0), _p(indices), _dt(indices),
                           ctypes.c_int64(n), _p(table), tdt, ctypes.c_int64(stride), ctypes.c_int64(storage_offset))
    if rc != 0:
        raise ValueError("oracle: dtype class mismatch")
    return table


def csr_aggregate(indptr, indices, x, mean=True):
    indptr = np.ascontiguousarray(indptr, dtype=np.int64)
    indices = np.ascontiguousarray(indices, dtype=np.int32)
    x = np.ascontiguousarray(x, dtype=np.float32)
    n_dst = indptr.shape[0] - 1
    dim = x.shape[1]
    out = np.empty((n_dst, dim), dtype=np.float64)
    lib().wgo_csr_aggregate(_p(indptr), _p(indices), ctypes.c_int64(n_dst), _p(x), ctypes.c_int64(dim), ctypes.c_int64(dim), ctypes.c_int(1 if mean else 0), _p(out))
    return out


def hop_seed(random_state, hop):
    return int(lib().wgo_hop_seed(ctypes.c_uint64(random_state), ctypes.c_int(hop)))


def multihop_sample(row_ptr, col, seeds, label_offsets, fanout, random_state, weights=None, edge_ids=None):
    """pylibcugraph-shaped multi-hop sample (COO). Returns a dict of numpy arrays."""
    row_ptr = np.ascontiguousarray(row_ptr, dtype=np.int64)
    col = np.ascontiguousarray(col)
    seeds = np.ascontiguousarray(seeds, dtype=np.int64)
    label_offsets = np.ascontiguousarray(label_offsets, dtype=np.int64)
    fanout = np.ascontiguousarray(fanout, dtype=np.int32)
    if weights is not None:
        weights = np.ascontiguousarray(weights)
    if edge_ids is not None:
        edge_ids = np.ascontiguousarray(edge_ids, dtype=np.int64)
    B = label_offsets.shape[0] - 1
    L = fanout.shape[0]
    h = lib().wgo_multihop_sample(
        _p(row_ptr), _p(col), _dt(col), _p(weights), (_dt(weights) if weights is not None else 0), _p(edge_ids),
        _p(seeds), _p(label_offsets), ctypes.c_int64(B), _p(fanout), ctypes.c_int(L), ctypes.c_uint64(random_state),
    )
    h = ctypes.c_void_p(h)
    ne = int(lib().wgo_multihop_num_edges(h))
    nn = int(lib().wgo_multihop_num_nodes(h))
    out = {
        "majors": np.empty(ne, dtype=np.int32),
        "minors": np.empty(ne, dtype=np.int32),
        "edge_id": np.empty(ne, dtype=np.int64),
        "label_hop_offsets": np.empty(B * L + 1, dtype=np.int64),
        "renumber_map": np.empty(nn, dtype=np.int64),
        "renumber_map_offsets": np.empty(B + 1, dtype=np.int64),
    }
    lib().wgo_multihop_copy(h, _p(out["majors"]), _p(out["minors"]), _p(out["edge_id"]), _p(out["label_hop_offsets"]),
                            _p(out["renumber_map"]), _p(out["renumber_map_offsets"]))
    lib().wgo_multihop_free(h)
    return out


def hetero_multihop_sample(row_ptrs, cols, vertex_type_offsets, seeds, label_offsets, fanout, random_state, weights=None,
                           edge_ids=None):
    """Heterogeneous multi-hop sample: one CSR (global vertex ids) per edge type, fanout laid out [hop * T + etype].
    Returns the pylibcugraph-shaped dict (COO) as numpy arrays."""
    T = len(row_ptrs)
    row_ptrs = [np.ascontiguousarray(r, dtype=np.int64) for r in row_ptrs]
    cols = [np.ascontiguousarray(c) for c in cols]
    assert all(c.dtype == cols[0].dtype for c in cols)
    vto = np.ascontiguousarray(vertex_type_offsets, dtype=np.int64)
    Vt = vto.shape[0] - 1
    seeds = np.ascontiguousarray(seeds, dtype=np.int64)
    label_offsets = np.ascontiguousarray(label_offsets, dtype=np.int64)
    fanout = np.ascontiguousarray(fanout, dtype=np.int32).reshape(-1)
    assert fanout.shape[0] % T == 0
    L = fanout.shape[0] // T
    B = label_offsets.shape[0] - 1
    vp = ctypes.c_void_p

    def ptr_array(arrs):
        return (vp * T)(*[(a.ctypes.data if a is not None else None) for a in arrs])

    wdt = 0
    w_arr = None
    if weights is not None:
        weights = [np.ascontiguousarray(w) for w in weights]
        wdt = _dt(weights[0])
        w_arr = ptr_array(weights)
    e_arr = None
    if edge_ids is not None:
        edge_ids = [None if e is None else np.ascontiguousarray(e, dtype=np.int64) for e in edge_ids]
        e_arr = ptr_array(edge_ids)
    fn = lib().wgo_hetero_multihop_sample
    fn.restype = ctypes.c_void_p
    h = fn(ctypes.c_int(T), ptr_array(row_ptrs), ptr_array(cols), _dt(cols[0]), w_arr, ctypes.c_int(wdt), e_arr, _p(vto),
           ctypes.c_int(Vt), _p(seeds), _p(label_offsets), ctypes.c_int64(B), _p(fanout), ctypes.c_int(L),
           ctypes.c_uint64(random_state))
    h = vp(h)
    lib().wgo_hetero_num_edges.restype = ctypes.c_int64
    lib().wgo_hetero_num_nodes.restype = ctypes.c_int64
    ne = int(lib().wgo_hetero_num_edges(h))
    nn = int(lib().wgo_hetero_num_nodes(h))
    out = {
        "majors": np.empty(ne, dtype=np.int32),
        "minors": np.empty(ne, dtype=np.int32),
        "edge_type": np.empty(ne, dtype=np.int32),
        "edge_id": np.empty(ne, dtype=np.int64),
        "label_type_hop_offsets": np.empty(B * T * L + 1, dtype=np.int64),
        "renumber_map": np.empty(nn, dtype=np.int64),
        "renumber_map_offsets": np.empty(B * Vt + 1, dtype=np.int64),
        "edge_renumber_map": np.empty(ne, dtype=np.int64),
        "edge_renumber_map_offsets": np.empty(B * T + 1, dtype=np.int64),
        "label_type_step_base": np.empty((L + 1, Vt, B), dtype=np.int32),
    }
    lib().wgo_hetero_copy(h, _p(out["majors"]), _p(out["minors"]), _p(out["edge_type"]), _p(out["edge_id"]),
                          _p(out["label_type_hop_offsets"]), _p(out["renumber_map"]), _p(out["renumber_map_offsets"]),
                          _p(out["edge_renumber_map"]), _p(out["edge_renumber_map_offsets"]), _p(out["label_type_step_base"]))
    lib().wgo_hetero_free(h)
    return out


OPT_TYPES = {"sgd": 1, "adam": 2, "lazy_adam": 2, "rmsprop": 3, "adagrad": 4}


def embedding_gradient_apply(optimizer, params, emb, indices, grads, lr, states):
    """In place on `emb` (fp32 [N, dim]) and `states` (dict of fp32 arrays: adam m/v/beta12t, adagrad state_sum,
    rmsprop v): one optimizer step per distinct index with its summed gradient."""
    t = OPT_TYPES[optimizer]
    assert emb.dtype == np.float32 and emb.flags.c_contiguous
    indices = np.ascontiguousarray(indices, dtype=np.int64)
    grads = np.ascontiguousarray(grads, dtype=np.float32)
    a = states.get("m") if t == 2 else states.get("state_sum") if t == 4 else states.get("v") if t == 3 else None
    b = states.get("v") if t == 2 else None
    pr = states.get("beta12t") if t == 2 else None
    f = ctypes.c_float
    rc = lib().wgo_embedding_gradient_apply(
        ctypes.c_int(t), f(params.get("weight_decay", 0.0)), f(params.get("epsilon", 1e-8)), f(params.get("beta1", 0.9)),
        f(params.get("beta2", 0.999)), ctypes.c_int(1 if params.get("adam_w", 0.0) > 0.5 else 0), f(params.get("alpha", 0.99)),
        f(lr), _p(emb), ctypes.c_int64(emb.shape[1]), _p(indices), ctypes.c_int64(indices.shape[0]), _p(grads), _p(a), _p(b), _p(pr))
    assert rc == 0
    return emb


TIME_CMP = {"strictly_increasing": 0, "monotonically_increasing": 1, "strictly_decreasing": 2, "monotonically_decreasing": 3}


def temporal_multihop_sample(row_ptrs, cols, edge_times, vertex_type_offsets, seeds, seed_times, label_offsets, fanout,
                             random_state, comparison="strictly_increasing", edge_ids=None, weights=None):
    """Temporal multi-hop sample over typed CSRs: edge_times[t] int64 [E_t], seed_times int64 [S]; a vertex is sampled only
    through edges whose time compares as requested with the time it was reached at.  weights[t] (fp32 / fp64, all types):
    biased (A-Res over the eligible edges) instead of uniform."""
    T = len(row_ptrs)
    row_ptrs = [np.ascontiguousarray(r, dtype=np.int64) for r in row_ptrs]
    cols = [np.ascontiguousarray(c) for c in cols]
    edge_times = [np.ascontiguousarray(t, dtype=np.int64) for t in edge_times]
    vto = np.ascontiguousarray(vertex_type_offsets, dtype=np.int64)
    Vt = vto.shape[0] - 1
    seeds = np.ascontiguousarray(seeds, dtype=np.int64)
    seed_times = np.ascontiguousarray(seed_times, dtype=np.int64)
    label_offsets = np.ascontiguousarray(label_offsets, dtype=np.int64)
    fanout = np.ascontiguousarray(fanout, dtype=np.int32).reshape(-1)
    L = fanout.shape[0] // T
    B = label_offsets.shape[0] - 1
    vp = ctypes.c_void_p

    def ptr_array(arrs):
        return (vp * T)(*[(a.ctypes.data if a is not None else None) for a in arrs])

    e_arr = None
    if edge_ids is not None:
        edge_ids = [None if e is None else np.ascontiguousarray(e, dtype=np.int64) for e in edge_ids]
        e_arr = ptr_array(edge_ids)
    if weights is None:
        fn = lib().wgo_temporal_multihop_sample
        fn.restype = ctypes.c_void_p
        h = fn(ctypes.c_int(T), ptr_array(row_ptrs), ptr_array(cols), _dt(cols[0]), e_arr, ptr_array(edge_times), _p(vto), ctypes.c_int(Vt),
               _p(seeds), _p(seed_times), _p(label_offsets), ctypes.c_int64(B), _p(fanout), ctypes.c_int(L), ctypes.c_uint64(random_state),
               ctypes.c_int(TIME_CMP[comparison]))
    else:
        weights = [np.ascontiguousarray(w) for w in weights]
        assert all(w.dtype == weights[0].dtype and w.dtype in (np.float32, np.float64) for w in weights)
        fn = lib().wgo_temporal_biased_multihop_sample
        fn.restype = ctypes.c_void_p
        h = fn(ctypes.c_int(T), ptr_array(row_ptrs), ptr_array(cols), _dt(cols[0]), ptr_array(weights), _dt(weights[0]), e_arr,
               ptr_array(edge_times), _p(vto), ctypes.c_int(Vt), _p(seeds), _p(seed_times), _p(label_offsets), ctypes.c_int64(B), _p(fanout),
               ctypes.c_int(L), ctypes.c_uint64(random_state), ctypes.c_int(TIME_CMP[comparison]))
    h = vp(h)
    lib().wgo_hetero_num_edges.restype = ctypes.c_int64
    lib().wgo_hetero_num_nodes.restype = ctypes.c_int64
    ne, nn = int(lib().wgo_hetero_num_edges(h)), int(lib().wgo_hetero_num_nodes(h))
    out = {
        "majors": np.empty(ne, dtype=np.int32), "minors": np.empty(ne, dtype=np.int32), "edge_type": np.empty(ne, dtype=np.int32),
        "edge_id": np.empty(ne, dtype=np.int64), "label_type_hop_offsets": np.empty(B * T * L + 1, dtype=np.int64),
        "renumber_map": np.empty(nn, dtype=np.int64), "renumber_map_offsets": np.empty(B * Vt + 1, dtype=np.int64),
        "edge_renumber_map": np.empty(ne, dtype=np.int64), "edge_renumber_map_offsets": np.empty(B * T + 1, dtype=np.int64),
        "label_type_step_base": np.empty((L + 1, Vt, B), dtype=np.int32),
    }
    lib().wgo_hetero_copy(h, _p(out["majors"]), _p(out["minors"]), _p(out["edge_type"]), _p(out["edge_id"]),
                          _p(out["label_type_hop_offsets"]), _p(out["renumber_map"]), _p(out["renumber_map_offsets"]),
                          _p(out["edge_renumber_map"]), _p(out["edge_renumber_map_offsets"]), _p(out["label_type_step_base"]))
    lib().wgo_hetero_free(h)
    return out
