"""bench.py's synthetic-input helpers on the CPU (small sizes): the RMAT generator yields a valid CSR by destination, the
id scramble is a bijection for every |V| a config uses, seed shards differ per rank."""
import math
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def test_rmat_csr_is_valid_and_scrambled():
    n, e = 5000, 80000
    row_ptr, col = bench.rmat_csr(torch, n, e, 42, torch.device("cpu"))
    rp, c = row_ptr.numpy(), col.numpy()
    assert rp[0] == 0 and rp[-1] == e and (np.diff(rp) >= 0).all() and rp.shape[0] == n + 1
    assert c.dtype == np.int32 and c.min() >= 0 and c.max() < n
    deg = np.diff(rp)
    assert deg.max() > 20 * deg.mean()  # heavy tail survives the scramble
    # hubs are spread over the id range instead of sitting at the small ids (what balances contiguous row partitions)
    top = np.argsort(-deg)[:50]
    assert (top < n // 2).sum() > 10 and (top >= n // 2).sum() > 10
    # same seed, same graph
    row_ptr2, col2 = bench.rmat_csr(torch, n, e, 42, torch.device("cpu"))
    assert torch.equal(row_ptr, row_ptr2) and torch.equal(col, col2)


def test_scramble_is_a_bijection_for_every_config_size():
    for v in (bench.NUM_NODES, 2_400_000, 111_000_000, 100_000_000, 50_000_000, 5000):
        assert math.gcd(bench.SCRAMBLE_MUL, v) == 1
    v = 5000
    ids = (np.arange(v, dtype=np.int64) * bench.SCRAMBLE_MUL + bench.SCRAMBLE_ADD) % v
    assert np.array_equal(np.sort(ids), np.arange(v))
    assert bench.SCRAMBLE_MUL * bench.NUM_NODES < 2**63


def test_seed_sets_are_distinct_per_rank_and_within_a_call_group():
    old = bench.NUM_NODES
    try:
        bench.NUM_NODES = 100_000
        a = bench.seed_sets(torch, 2, 3, rank=0)
        b = bench.seed_sets(torch, 2, 3, rank=1)
    finally:
        bench.NUM_NODES = old
    assert len(a) == 2 and a[0].numel() == 3 * bench.BATCH and a[0].dtype == torch.int64
    assert not torch.equal(a[0], b[0]) and not torch.equal(a[0], a[1])
    assert torch.unique(a[0]).numel() == a[0].numel()  # a randperm prefix: distinct seeds inside a call group
