"""Fused SAGE layer (csrc/sage_tile.cu: gather-mean -> smem operand tile -> tcgen05.mma, W by TMA) against the unfused path it
replaces (csr_aggregate kernel, then [agg || self] . W_cat^T in cuBLAS), on BASELINE config C3: products-shaped synthetic graph
(|V| = 2.4 M, |E| = 123 M, RMAT), first GraphSAGE layer (F_in = 128 -> 256) over the sampler's CSR block, fan-out [15, 10, 5].

    python profiles/sage_tile_bench.py [seeds_per_batch] [iters] [--once]

Every measurement: L2 flushed (256 MB write), then ONE launch timed with CUDA events on the launching stream; the mean over
`iters` different blocks is reported.  --once: a single launch of each path on one block (for ncu)."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "cugraph-gnn_b200"))
import torch
import bench
import pylibwholegraph.torch as wgth
from pylibwholegraph.torch.aggregate import csr_aggregate_forward, sage_layer_forward

argv = [a for a in sys.argv[1:] if not a.startswith("--")]
once = "--once" in sys.argv
seeds_per_batch = int(argv[0]) if len(argv) > 0 else 1024
iters = int(argv[1]) if len(argv) > 1 else 10
V, E, F, H = 2_400_000, 123_000_000, 128, 256
FANOUT = [15, 10, 5]
torch.cuda.set_device(0)
dev = torch.device("cuda", 0)
wgth.init(0, 1, 0, 1)
row_ptr, col = bench.rmat_csr(torch, V, E, 42, dev)
table = torch.randn((V, F), device=dev).to(torch.bfloat16)  # bf16 feature table (the fused layer's operand format)
sampler = wgth.MultiHopSampler()
peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {}
peak = peaks.get("hbm_gbs", 6650.0)
g = torch.Generator().manual_seed(7)
w_cat = (torch.randn((H, 2 * F), device=dev) / 16).to(torch.bfloat16)
bias = torch.randn(H, device=dev)
flush = torch.zeros(64 << 20, device=dev)  # 256 MB


def timed(fn):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    flush.add_(1.0)
    a.record()
    out = fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b), out


acc = {}
max_err = 0.0
for it in range((1 if once else iters) + 2):
    seeds = torch.randperm(V, generator=g)[:seeds_per_batch].to(dev)
    res = sampler.sample(row_ptr, col, seeds, torch.tensor([0, seeds_per_batch], device=dev), FANOUT, 62 + it, compression="CSR")
    mo, minors, lho = res["major_offsets"], res["minors"], res["label_hop_offsets"].tolist()
    n_rows = lho[3]
    nnz = int(mo[n_rows])
    indptr, indices = mo[: n_rows + 1], minors[:nnz]
    x = table[res["renumber_map"]].contiguous()  # the gathered block the model receives, bf16 [n_src, 128]
    n_src = x.shape[0]

    def unfused():
        agg = csr_aggregate_forward(indptr, indices, x, "mean")               # SIMT kernel, fp32 [n_rows, 128]
        a = torch.cat([agg.to(torch.bfloat16), x[:n_rows]], dim=1)             # operand of the dense part
        return torch.nn.functional.linear(a, w_cat).float() + bias            # cuBLAS bf16 GEMM, fp32 out + bias

    def unfused_two_gemms():  # how SAGEConv.forward spells it: two Linear layers
        agg = csr_aggregate_forward(indptr, indices, x, "mean")
        return (torch.nn.functional.linear(agg.to(torch.bfloat16), w_cat[:, :F]) + torch.nn.functional.linear(x[:n_rows], w_cat[:, F:])).float() + bias

    def fused():
        return sage_layer_forward(indptr, indices, x, w_cat, bias)

    def agg_only():
        return csr_aggregate_forward(indptr, indices, x, "mean")

    cases = [("fused sage_tile_kernel (tcgen05)", fused), ("unfused: csr_aggregate + cat + cuBLAS bf16 GEMM + bias", unfused),
             ("unfused: csr_aggregate + 2 cuBLAS bf16 GEMMs + bias", unfused_two_gemms), ("csr_aggregate kernel alone (SIMT, no dense part)", agg_only)]
    outs = {}
    for name, fn in cases:
        ms, out = timed(fn)
        outs[name] = out
        if it >= 2:
            r = acc.setdefault(name, [0.0, 0, 0, 0])
            r[0] += ms
            r[1] += nnz
            r[2] += n_rows
            r[3] += n_src
    if it >= 2:
        ref = outs[cases[1][0]]
        max_err = max(max_err, float(((outs[cases[0][0]] - ref).abs() / ref.abs().clamp(min=1.0)).max()))
n = 1 if once else iters
for name, (ms, nnz, rows, nsrc) in acc.items():
    # algorithmic bytes of the fused layer: every edge reads a 256 B bf16 row + a 4/8 B index, every destination row reads its own
    # 256 B row + 8 B of indptr and writes F_out fp32; flops of the dense part 2 * rows * 256 * F_out
    alg = nnz * (F * 2 + 4) + rows * (F * 2 + 8 + H * 4)
    print(json.dumps({"case": name, "seeds": seeds_per_batch, "iters": n, "avg_ms": ms / n, "avg_nnz": nnz / n, "avg_dst_rows": rows / n,
                      "avg_src_rows": nsrc / n, "alg_GBps_of_fused_layer": alg / (ms * 1e-3) / 1e9, "hbm_peak_GBps": peak,
                      "frac_of_measured_hbm_peak": alg / (ms * 1e-3) / 1e9 / peak, "dense_TFLOPs": 2.0 * rows * 2 * F * H / (ms * 1e-3) / 1e12,
                      "max_rel_diff_fused_vs_unfused": max_err}))
