"""Two processes, one GPU each, chunked table over cudaIpc: gather bandwidth for local / remote / mixed rows."""
import functools, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "cugraph-gnn_b200"))


def worker(rank, world, uid):
    import torch
    import pylibwholegraph.binding.wholememory_binding as wmb
    import pylibwholegraph.torch as wgth

    torch.cuda.set_device(rank)
    wgth.init(rank, world, rank, world)
    comm = wgth.WholeMemoryCommunicator(wmb.create_communicator(wmb.PyWholeMemoryUniqueID(uid), rank, world))
    rows, dim = 10_000_000, 128
    emb = wgth.create_wholememory_tensor(comm, "chunked", "cuda", [rows, dim], torch.float32, [dim, 1])
    local, start = emb.get_local_tensor()
    local.normal_()
    torch.cuda.synchronize(); comm.barrier()
    n = 2_500_000
    g = torch.Generator(device="cuda").manual_seed(rank)
    half = rows // 2
    other_lo = half if rank == 0 else 0
    idx = {
        "local": torch.randint(start, start + half, (n,), device="cuda", generator=g),
        "remote": torch.randint(other_lo, other_lo + half, (n,), device="cuda", generator=g),
        "mixed": torch.randint(0, rows, (n,), device="cuda", generator=g),
    }
    for solo in (True, False):
        for name, ix in idx.items():
            comm.barrier()
            if solo and rank != 0:
                comm.barrier()
                continue
            for _ in range(2):
                emb.gather(ix)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(5):
                emb.gather(ix)
            e1.record(); torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 5
            print("rank %d %s %-7s %8.3f ms %8.1f GB/s out" % (rank, "solo" if solo else "both", name, ms, n * dim * 4 / ms / 1e6), flush=True)
            if solo:
                comm.barrier()
    comm.barrier()


if __name__ == "__main__":
    import pylibwholegraph.binding.wholememory_binding as wmb
    from pylibwholegraph.utils.multiprocess import multiprocess_run
    uid = wmb.create_unique_id().get_bytes()
    multiprocess_run(2, functools.partial(worker, uid=uid))
