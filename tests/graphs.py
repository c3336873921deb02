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
