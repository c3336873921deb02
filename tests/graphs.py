"""Seeded synthetic inputs shared by the CPU and GPU tests (numpy only)."""
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))


def random_csr(num_nodes, num_edges, seed, col_dtype=np.int32, skew=True):
    """CSR with a heavy-tailed degree distribution (multigraph allowed), row_ptr int64."""
    rng = np.random.default_rng(seed)
    if skew:
        w = rng.pareto(1.5, num_nodes) + 0.05
        w[rng.random(num_nodes) < 0.05] = 0.0  # some isolated vertices
        p = w / w.sum()
        deg = rng.multinomial(num_edges, p)
    else:
        deg = rng.multinomial(num_edges, np.full(num_nodes, 1.0 / num_nodes))
    row_ptr = np.zeros(num_nodes + 1, dtype=np.int64)
    np.cumsum(deg, out=row_ptr[1:])
    col = rng.integers(0, num_nodes, size=num_edges).astype(col_dtype)
    return row_ptr, col


def karate_csr(col_dtype=np.int32):
    """Zachary karate club (tests/golden/karate.csv: 'src dst weight', 156 directed edges).

    CSR by source; the file is symmetric so this is also the PyG in-edge CSR."""
    data = np.loadtxt(os.path.join(HERE, "golden", "karate.csv"))
    src = data[:, 0].astype(np.int64)
    dst = data[:, 1].astype(np.int64)
    n = int(max(src.max(), dst.max())) + 1
    order = np.lexsort((np.arange(src.shape[0]), src))
    src, dst = src[order], dst[order]
    row_ptr = np.zeros(n + 1, dtype=np.int64)
    np.cumsum(np.bincount(src, minlength=n), out=row_ptr[1:])
    return row_ptr, dst.astype(col_dtype)


def typed_csrs(src, dst, etype, num_edge_types, num_vertices, col_dtype=np.int32):
    """One CSR by source per edge type over the global vertex id space (stable in the input order).
    Returns (row_ptrs, cols, positions) with positions[t][k] = index in the input arrays of CSR_t's k-th edge."""
    src, dst, etype = np.asarray(src, dtype=np.int64), np.asarray(dst, dtype=np.int64), np.asarray(etype)
    row_ptrs, cols, pos = [], [], []
    for t in range(num_edge_types):
        sel = np.nonzero(etype == t)[0]
        order = sel[np.argsort(src[sel], kind="stable")]
        rp = np.zeros(num_vertices + 1, dtype=np.int64)
        np.cumsum(np.bincount(src[order], minlength=num_vertices), out=rp[1:])
        row_ptrs.append(rp)
        cols.append(dst[order].astype(col_dtype))
        pos.append(order)
    return row_ptrs, cols, pos


def random_typed_graph(vertex_counts, edge_types, edges_per_type, seed, col_dtype=np.int32):
    """Heterogeneous multigraph: `vertex_counts` per vertex type (global ids are type-offset), `edge_types` =
    [(src_vtype, dst_vtype)], heavy-tailed out-degrees.  Returns (vertex_type_offsets, row_ptrs, cols)."""
    rng = np.random.default_rng(seed)
    vto = np.concatenate([[0], np.cumsum(vertex_counts)]).astype(np.int64)
    V = int(vto[-1])
    row_ptrs, cols = [], []
    for (sv, dv), ne in zip(edge_types, edges_per_type):
        ns, nd = vertex_counts[sv], vertex_counts[dv]
        w = rng.pareto(1.5, ns) + 0.05
        w[rng.random(ns) < 0.1] = 0.0
        deg = rng.multinomial(ne, w / w.sum())
        full = np.zeros(V, dtype=np.int64)
        full[vto[sv]:vto[sv + 1]] = deg
        rp = np.zeros(V + 1, dtype=np.int64)
        np.cumsum(full, out=rp[1:])
        row_ptrs.append(rp)
        cols.append((vto[dv] + rng.integers(0, nd, size=ne)).astype(col_dtype))
    return vto, row_ptrs, cols
