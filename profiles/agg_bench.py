"""A1 measurement on BASELINE config C3: products-shaped synthetic graph (|V| = 2.4 M, |E| = 123 M, RMAT), 3-layer
GraphSAGE aggregation over the sampler's CSR blocks, fan-out [15, 10, 5].

    python profiles/agg_bench.py [seeds_per_batch] [iters]

Per layer: device time (CUDA events around 8 back-to-back launches on the same block; blocks whose rows fit the 126 MB
L2 are served from it after the first launch and are marked "l2_resident"), algorithmic bytes nnz * (F * elt + 4) + n_dst * F * 4 and the
fraction of the measured HBM peak.  Layer 0 reads the fp32 feature table through the renumber map (fused gather),
layers 1-2 read the dense hidden activations (F = 256 fp32, and bf16 as the reduced-traffic variant)."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "cugraph-gnn_b200"))
import torch
import bench
import pylibwholegraph.torch as wgth
from pylibwholegraph.torch.aggregate import csr_aggregate_forward

seeds_per_batch = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 10
V, E, F, H = 2_400_000, 123_000_000, 128, 256
FANOUT = [15, 10, 5]
torch.cuda.set_device(0)
dev = torch.device("cuda", 0)
wgth.init(0, 1, 0, 1)
comm = wgth.get_global_communicator()
row_ptr, col = bench.rmat_csr(torch, V, E, 42, dev)
emb = wgth.create_embedding(comm, "chunked", "cuda", torch.float32, [V, F])
emb.get_embedding_tensor().get_local_tensor()[0].normal_()
sampler = wgth.MultiHopSampler()
peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))).get("hbm_gbs", 6650.0) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
g = torch.Generator().manual_seed(7)
acc = {}
REPS = 8
flush = torch.zeros(64 << 20, device=dev)  # 256 MB
for it in range(iters + 2):
    seeds = torch.randperm(V, generator=g)[:seeds_per_batch].to(dev)
    res = sampler.sample(row_ptr, col, seeds, torch.tensor([0, seeds_per_batch], device=dev), FANOUT, 62 + it, compression="CSR")
    mo, minors, lho = res["major_offsets"], res["minors"], res["label_hop_offsets"].tolist()
    n_id = res["renumber_map"]
    n_nodes = n_id.numel()
    hidden32 = torch.randn((n_nodes, H), device=dev)
    hidden16 = hidden32.to(torch.bfloat16)
    mo_host = mo.tolist()
    cases = []
    for k in range(3):
        n_rows = lho[3 - k]
        nnz = mo_host[n_rows]
        indptr, indices = mo[: n_rows + 1], minors[:nnz]
        if k == 0:
            cases.append(("layer0 fused-gather F=128 fp32 table", indptr, indices, emb, n_id, F, 4, n_rows, nnz))
        else:
            cases.append(("layer%d F=256 fp32" % k, indptr, indices, hidden32, None, H, 4, n_rows, nnz))
            cases.append(("layer%d F=256 bf16" % k, indptr, indices, hidden16, None, H, 2, n_rows, nnz))
    for name, indptr, indices, x, gmap, feat, elt, n_rows, nnz in cases:
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        flush.add_(1.0)  # > L2: the block's rows come from HBM on the first of the REPS launches
        a.record()
        for _ in range(REPS):  # back to back, so that the ~40 us of python per call hide behind the device queue
            out = csr_aggregate_forward(indptr, indices, x, "mean", gather_map=gmap)
        b.record()
        torch.cuda.synchronize()
        if it >= 2:
            ms = a.elapsed_time(b) / REPS
            alg = nnz * (feat * elt + 4 + (8 if gmap is not None else 0)) + n_rows * feat * 4
            r = acc.setdefault(name, [0.0, 0.0, 0, 0])
            r[0] += ms
            r[1] += alg
            r[2] += nnz
            r[3] += n_rows
for name, (ms, alg, nnz, rows) in acc.items():
    gbs = alg / (ms * 1e-3) / 1e9
    print(json.dumps({"kernel": "csr_aggregate_kernel", "case": name, "seeds": seeds_per_batch, "iters": iters,
                      "avg_ms": ms / iters, "avg_nnz": nnz / iters, "avg_rows": rows / iters, "alg_GBps": gbs,
                      "hbm_peak_GBps": peak, "frac_of_measured_peak": gbs / peak,
                      "l2_resident": bool(alg / iters < 100e6)}))
