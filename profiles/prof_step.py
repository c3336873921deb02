"""Short driver for ncu: builds the bench workload (argv[3]: c4 default | c2 | headline) and runs a few sampler + gather
steps (no timing).  Same call shape as bench.py (int64 ids)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "cugraph-gnn_b200"))
import torch
import bench
import pylibwholegraph.torch as wgth

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
labels = int(sys.argv[2]) if len(sys.argv) > 2 else bench.LABELS_PER_STEP
workload = sys.argv[3] if len(sys.argv) > 3 else "c4"
if bench.WORKLOADS[workload] is not None:
    for k, v in bench.WORKLOADS[workload].items():
        setattr(bench, k, v)
torch.cuda.set_device(0)
dev = torch.device("cuda", 0)
wgth.init(0, 1, 0, 1)
comm = wgth.get_global_communicator()
row_ptr, col = bench.rmat_csr(torch, bench.NUM_NODES, bench.NUM_EDGES, 42, dev)
wm_rp = wgth.create_wholememory_tensor(comm, "chunked", "cuda", [bench.NUM_NODES + 1], torch.int64, [1])
wm_rp.get_local_tensor()[0].copy_(row_ptr)
wm_col = wgth.create_wholememory_tensor(comm, "chunked", "cuda", [col.numel()], torch.int32, [1])
wm_col.get_local_tensor()[0].copy_(col)
emb = wgth.create_embedding(comm, "chunked", "cuda", torch.float32, [bench.NUM_NODES, bench.FEAT_DIM])
emb.get_embedding_tensor().get_local_tensor()[0].fill_(1.0)
sampler = wgth.MultiHopSampler()
lo = (torch.arange(labels + 1, dtype=torch.int64) * bench.BATCH).to(dev)
seeds = [s.to(dev) for s in bench.seed_sets(torch, steps, labels)]
torch.cuda.synchronize()
for k in range(steps):
    res = sampler.sample(wm_rp, wm_col, seeds[k], lo, bench.FANOUT, 62 + k, int64_ids=True)
    x = emb.gather(res["renumber_map"])
torch.cuda.synchronize()
print("edges", res["minors"].numel(), "nodes", res["renumber_map"].numel())
