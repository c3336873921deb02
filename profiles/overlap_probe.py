"""Does the feature gather run underneath the sampler?  Times, with CUDA events and no host work in between:
  A  K x sampler kernels alone (begin only: fused label kernel + meta)       B  K x gather alone
  C  K x (sampler on stream 1 || gather on stream 2), launched back to back
Perfect overlap: C = max(A, B); none: C = A + B.   argv: workload (c4 default), K, labels per call group (comma-separated list, default 64)"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "cugraph-gnn_b200"))
import torch
import bench
import pylibwholegraph.torch as wgth

workload = sys.argv[1] if len(sys.argv) > 1 else "c4"
K = int(sys.argv[2]) if len(sys.argv) > 2 else 10
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
del col
emb = wgth.create_embedding(comm, "chunked", "cuda", torch.float32, [bench.NUM_NODES, bench.FEAT_DIM])
label_list = [int(v) for v in sys.argv[3].split(",")] if len(sys.argv) > 3 else [64]
side = torch.cuda.Stream(device=dev)


def timed(fn):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    a.record()
    fn()
    torch.cuda.current_stream().wait_stream(side)
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / K


for labels in label_list:
  lo = (torch.arange(labels + 1, dtype=torch.int64) * bench.BATCH).to(dev)
  seeds = [s.to(dev) for s in bench.seed_sets(torch, K + 2, labels)]
  samplers = [wgth.MultiHopSampler() for _ in range(K + 2)]  # one object per call in flight: begin() only enqueues
  res = samplers[0].sample(wm_rp, wm_col, seeds[0], lo, bench.FANOUT, 62, int64_ids=True)
  ids = res["renumber_map"]
  x = emb.gather(ids)
  torch.cuda.synchronize()


  def sampler_only():
      pend = [samplers[k + 1].sample_async(wm_rp, wm_col, seeds[k + 1], lo, bench.FANOUT, 62 + k, int64_ids=True) for k in range(K)]
      return pend


  def gather_only():
      with torch.cuda.stream(side):
          for _ in range(K):
              emb.gather(ids)


  def both():
      pend = []
      for k in range(K):
          pend.append(samplers[k + 1].sample_async(wm_rp, wm_col, seeds[k + 1], lo, bench.FANOUT, 62 + k, int64_ids=True))
          with torch.cuda.stream(side):
              emb.gather(ids)
      return pend


  for name in ("warm", "run"):
      keep = []
      side.wait_stream(torch.cuda.current_stream())
      A = timed(lambda: keep.append(sampler_only()))
      [p.result() for p in keep[-1]]
      B = timed(gather_only)
      C = timed(lambda: keep.append(both()))
      [p.result() for p in keep[-1]]
      if name == "run":
          print("labels=%d bulk=%s  sampler alone %.3f ms   gather alone %.3f ms   both %.3f ms   (sum %.3f, max %.3f)  overlap efficiency %.2f  | sampler us per label %.2f" % (
              labels, os.environ.get("WGB_GATHER_BULK", "1"), A, B, C, A + B, max(A, B), (A + B - C) / min(A, B), 1e3 * A / labels))
