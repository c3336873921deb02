"""torch.profiler over two C5 steps (one GPU): which kernels the model part of the step spends its time in."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "cugraph-gnn_b200"))
import torch
from torch.profiler import profile, ProfilerActivity
import bench

sys.argv = ["bench.py", "--workload", "c5", "--steps", "2", "--warmup", "3", "--labels", "32"]
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    bench.main()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=25, max_name_column_width=70), file=sys.stderr)
