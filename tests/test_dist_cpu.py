"""world_size-2 gloo tests (CPU) of the multi-rank host logic: batch-id offsets, call-count equalisation, global
edge ids, vertex-count agreement, and the WholeMemory partition plan every rank derives independently.
The device side of N > 1 (peer-mapped gather, chunked sampling) is covered by tests/test_gpu_multirank.py."""
import os
import socket
import sys
import traceback

import pytest
import torch
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, fn_name, out):
    try:
        sys.path.insert(0, os.path.join(ROOT, "cugraph-gnn_b200"))
        os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
        torch.distributed.init_process_group("gloo", rank=rank, world_size=world)
        out[rank] = globals()[fn_name](rank, world)
        torch.distributed.barrier()
        torch.distributed.destroy_process_group()
    except Exception:
        out[rank] = "ERROR: " + traceback.format_exc()


def _run(fn_name, world=2):
    ctx = mp.get_context("spawn")
    out = ctx.Manager().dict()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, fn_name, out)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert not p.is_alive(), "gloo worker hung"
    res = [out.get(r) for r in range(world)]
    for r in res:
        assert not (isinstance(r, str) and r.startswith("ERROR")), r
    return res



def _dist_utils():
    from cugraph_pyg.utils import dist as du

    return du


def check_batch_offsets(rank, world):
    du = _dist_utils()
    # uneven: rank 0 has 5 batches, rank 1 has 3
    start, equal = du.batch_id_start(5 if rank == 0 else 3)
    start_eq, equal_eq = du.batch_id_start(4)
    start_assumed, _ = du.batch_id_start(7, assume_equal_input_size=True)
    calls = du.equalized_call_count(3 if rank == 0 else 1, equal=False)
    return dict(start=start, equal=equal, start_eq=start_eq, equal_eq=equal_eq, start_assumed=start_assumed, calls=calls)


def test_batch_id_offsets_and_call_equalisation_two_ranks():
    r0, r1 = _run("check_batch_offsets")
    assert (r0["start"], r1["start"]) == (0, 5) and not r0["equal"] and not r1["equal"]
    assert (r0["start_eq"], r1["start_eq"]) == (0, 4) and r0["equal_eq"] and r1["equal_eq"]
    assert (r0["start_assumed"], r1["start_assumed"]) == (0, 7)
    assert r0["calls"] == r1["calls"] == 3


def check_edge_ids(rank, world):
    du = _dist_utils()
    counts = [10, 0, 7] if rank == 0 else [4, 9, 1]  # three edge types
    starts = du.edge_id_starts(counts)
    nv = du.all_reduce_max(100 if rank == 0 else 250)
    table = du.all_gather_ints(counts)
    return dict(starts=starts, nv=nv, table=table)


def test_global_edge_ids_and_vertex_counts_two_ranks():
    r0, r1 = _run("check_edge_ids")
    assert r0["starts"] == [0, 0, 0] and r1["starts"] == [10, 0, 7]
    assert r0["nv"] == r1["nv"] == 250
    assert r0["table"] == r1["table"] == [[10, 0, 7], [4, 9, 1]]


def check_partition_plan(rank, world):
    """Every rank derives the same row partition of a WholeMemory table from (N, world) alone, and the seed shard of
    bench.py is disjoint across ranks: no collective is needed on the data path."""
    import ctypes
    import pylibwholegraph.binding.wholememory_binding as wmb

    lib = ctypes.CDLL(wmb.LIBRARY_PATH)
    fn = lib.wholememory_equal_entry_partition_plan
    fn.restype = ctypes.c_int
    per = ctypes.c_size_t()
    n = 10_000_003
    assert fn(ctypes.byref(per), ctypes.c_size_t(n), ctypes.c_int(world)) == 0
    lo, hi = min(rank * per.value, n), min((rank + 1) * per.value, n)
    sys.path.insert(0, ROOT)
    import bench

    seeds = bench.seed_sets(torch, 1, 2, rank)[0]
    t = torch.tensor([lo, hi, int(seeds[:8].sum())], dtype=torch.int64)
    all_t = [torch.zeros_like(t) for _ in range(world)]
    torch.distributed.all_gather(all_t, t)
    return [x.tolist() for x in all_t]


def test_partition_plan_and_seed_shards_two_ranks():
    r0, r1 = _run("check_partition_plan")
    assert r0 == r1
    (lo0, hi0, s0), (lo1, hi1, s1) = r0
    assert lo0 == 0 and hi0 == lo1 and hi1 == 10_000_003 and hi0 - lo0 == 5_000_002  # ceil(N / W) rows per rank
    assert s0 != s1  # different seed streams per rank
