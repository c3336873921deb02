"""Scale check on one GPU, fan-out [25, 10], F = 128: C4 shape (|V| = 111 M, |E| = 1.6 B, 56.8 GB feature table) by
default; with |E| >= 2^31 (e.g. 2300000000) the CSR positions / edge ids need 64 bits and the feature table is skipped.
Samples call groups, checks the edge positions against the CSR, gathers the features and prints stage times.

    python profiles/scale_check.py [labels] [num_edges]
"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "cugraph-gnn_b200"))
import torch
import bench
import pylibwholegraph.torch as wgth

labels = int(sys.argv[1]) if len(sys.argv) > 1 else 64
V, E, F = 111_000_000, (int(sys.argv[2]) if len(sys.argv) > 2 else 1_600_000_000), 128
WITH_FEATURES = E < 2**31
torch.cuda.set_device(0)
dev = torch.device("cuda", 0)
wgth.init(0, 1, 0, 1)
comm = wgth.get_global_communicator()
t0 = time.time()
if WITH_FEATURES:
    row_ptr, col = bench.rmat_csr(torch, V, E, 42, dev)
else:
    # torch.sort refuses more than INT_MAX elements: build the CSR directly (degree 19..23 by vertex, random columns)
    deg = 19 + (torch.arange(V, device=dev) % 5)
    row_ptr = torch.zeros(V + 1, dtype=torch.int64, device=dev)
    row_ptr[1:] = deg.cumsum(0)
    E = int(row_ptr[-1])
    col = torch.empty(E, dtype=torch.int32, device=dev)
    gen = torch.Generator(device=dev).manual_seed(42)
    for lo in range(0, E, 1 << 28):
        hi = min(E, lo + (1 << 28))
        col[lo:hi] = torch.randint(0, V, (hi - lo,), device=dev, dtype=torch.int32, generator=gen)
    del deg
torch.cuda.synchronize()
print("CSR built in %.1f s: row_ptr[-1] = %d (> 2^31: %s), col %s, %.1f GB allocated" % (
    time.time() - t0, int(row_ptr[-1]), int(row_ptr[-1]) > 2**31, col.dtype, torch.cuda.memory_allocated() / 1e9), flush=True)
emb = None
if WITH_FEATURES:
    emb = wgth.create_embedding(comm, "chunked", "cuda", torch.float32, [V, F])  # 56.8 GB
    local = emb.get_embedding_tensor().get_local_tensor()[0]
    ar = torch.arange(F, device=dev)[None, :]
    for lo in range(0, V, 1 << 22):
        hi = min(V, lo + (1 << 22))
        local[lo:hi] = ((torch.arange(lo, hi, device=dev)[:, None] + ar) & 0xFFFF).float()
sampler = wgth.MultiHopSampler()
lo_t = (torch.arange(labels + 1, dtype=torch.int64) * bench.BATCH).to(dev)
g = torch.Generator().manual_seed(5)
for it in range(4):
    seeds = torch.randint(0, V, (labels * bench.BATCH,), generator=g).to(dev)
    a, b, c = (torch.cuda.Event(enable_timing=True) for _ in range(3))
    a.record()
    res = sampler.sample(row_ptr, col, seeds, lo_t, bench.FANOUT, 62 + it)
    b.record()
    x = emb.gather(res["renumber_map"]) if emb is not None else None
    c.record()
    torch.cuda.synchronize()
    print("call %d: %d edges, %d nodes, sample %.3f ms, gather %.3f ms" % (it, res["minors"].numel(), res["renumber_map"].numel(),
                                                                         a.elapsed_time(b), b.elapsed_time(c)), flush=True)
# every reported edge position lies in the CSR row of its (global) major and holds its (global) minor
lho, rmo = res["label_hop_offsets"].tolist(), res["renumber_map_offsets"].tolist()
eid, mj, mn, rmap = res["edge_id"], res["majors"].long(), res["minors"].long(), res["renumber_map"]
if not WITH_FEATURES:
    assert int(eid.max()) > 2**31, "this mode is meant to exercise positions above 2^31"
L = len(bench.FANOUT)
for l in (0, labels // 2, labels - 1):
    e0, e1 = lho[l * L], lho[(l + 1) * L]
    m = rmap[rmo[l]:rmo[l + 1]]
    src, dst, pos = m[mj[e0:e1]], m[mn[e0:e1]], eid[e0:e1]
    assert bool(((row_ptr[src] <= pos) & (pos < row_ptr[src + 1])).all())
    assert bool((col[pos].long() == dst).all())
    if x is None:
        continue
    feats = x[rmo[l]:rmo[l + 1]]
    assert bool((feats[:, 0] == (m & 0xFFFF).float()).all()) and bool((feats[:, 127] == ((m + 127) & 0xFFFF).float()).all())
print("scale check ok: edge positions up to %d verified against the CSR, features verified" % int(eid.max()))
