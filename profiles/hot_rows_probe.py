"""How much of the gathered rows would a replicated hot-row set capture?  (C2 workload, hotness = times a vertex appears
as a CSR column, i.e. how often the sampler can reach it.)"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "cugraph-gnn_b200"))
import torch
import bench
import pylibwholegraph.torch as wgth

torch.cuda.set_device(0)
dev = torch.device("cuda", 0)
wgth.init(0, 1, 0, 1)
row_ptr, col = bench.rmat_csr(torch, bench.NUM_NODES, bench.NUM_EDGES, 42, dev)
hot = torch.bincount(col.long(), minlength=bench.NUM_NODES)
order = torch.argsort(hot, descending=True)
sampler = wgth.MultiHopSampler()
lo = (torch.arange(65, dtype=torch.int64) * bench.BATCH).to(dev)
seeds = bench.seed_sets(torch, 1, 64)[0].to(dev)
res = sampler.sample(row_ptr, col, seeds, lo, bench.FANOUT, 62)
rows = res["renumber_map"]
print("rows gathered", rows.numel(), "distinct", torch.unique(rows).numel())
for H in (1000, 10_000, 100_000, 300_000, 1_000_000, 2_000_000):
    mask = torch.zeros(bench.NUM_NODES, dtype=torch.bool, device=dev)
    mask[order[:H]] = True
    frac = mask[rows].float().mean().item()
    print("top %8d rows (%6.1f MB replicated per GPU): %.1f %% of gathered rows" % (H, H * 512 / 1e6, 100 * frac))
