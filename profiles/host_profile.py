"""Where does the host time of one end-to-end step go?  (wall clock per phase, device idle between syncs)

    python profiles/host_profile.py [steps] [labels]
"""
import cProfile
import os
import pstats
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "cugraph-gnn_b200"))
import torch
import bench
import pylibwholegraph.torch as wgth

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 20
labels = int(sys.argv[2]) if len(sys.argv) > 2 else bench.LABELS_PER_STEP
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
host_seeds = [s.pin_memory() for s in bench.seed_sets(torch, steps + 3, labels)]


def one(k, T):
    t0 = time.perf_counter()
    sd = host_seeds[k].to(dev, non_blocking=True)
    t1 = time.perf_counter()
    res = sampler.sample(wm_rp, wm_col, sd, lo, bench.FANOUT, 62 + k)
    t2 = time.perf_counter()
    x = emb.gather(res["renumber_map"])
    t3 = time.perf_counter()
    metric = torch.cat([res["label_hop_offsets"].double(), res["renumber_map_offsets"].double(), x.sum(dtype=torch.float64).reshape(1)])
    t4 = time.perf_counter()
    host = metric.cpu()
    t5 = time.perf_counter()
    if T is not None:
        for i, d in enumerate((t1 - t0, t2 - t1, t3 - t2, t4 - t3, t5 - t4)):
            T[i] += d
    return host


for k in range(3):
    one(k, None)
torch.cuda.synchronize()
T = [0.0] * 5
t0 = time.perf_counter()
for k in range(steps):
    one(3 + k, T)
torch.cuda.synchronize()
wall = time.perf_counter() - t0
names = ["h2d seeds", "sampler.sample (incl. its host sync)", "emb.gather (enqueue)", "metric cat (enqueue)", "metric .cpu() (sync)"]
print("wall per step %.3f ms" % (1e3 * wall / steps))
for n, d in zip(names, T):
    print("  %-40s %.3f ms" % (n, 1e3 * d / steps))

pr = cProfile.Profile()
pr.enable()
for k in range(steps):
    one(3 + k, None)
torch.cuda.synchronize()
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(35)
