"""Single-process probe: gather kernel on cuda:0 reading a table that lives on cuda:1 (peer access)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "cugraph-gnn_b200"))
import torch
import pylibwholegraph.binding.wholememory_binding as wmb
from pylibwholegraph.torch.wholegraph_env import wrap_torch_tensor, get_wholegraph_env_fns, get_stream

rows, dim = 5_000_000, 128
remote = torch.empty((rows, dim), dtype=torch.float32, device="cuda:1").normal_()
local = torch.empty((rows, dim), dtype=torch.float32, device="cuda:0").normal_()
torch.cuda.set_device(0)
# enable peer access both ways by letting torch do one peer copy
import ctypes, glob
rt = ctypes.CDLL(glob.glob(os.path.join(os.path.dirname(torch.__file__), "..", "nvidia", "cuda_runtime", "lib", "libcudart.so*"))[0])
print("peer capable:", torch.cuda.can_device_access_peer(0, 1))
rt.cudaSetDevice(0); print("enable 0->1:", rt.cudaDeviceEnablePeerAccess(1, 0))
tmp = remote[:1024].to("cuda:0"); torch.cuda.synchronize()
n = 2_500_000
g = torch.Generator(device="cuda:0").manual_seed(0)
idx_rand = torch.randint(0, rows, (n,), device="cuda:0", generator=g)
idx_sorted = idx_rand.sort().values
idx_seq = torch.arange(n, device="cuda:0")
out = torch.empty((n, dim), dtype=torch.float32, device="cuda:0")

def run(table, idx, label):
    t = wrap_torch_tensor(table); i = wrap_torch_tensor(idx); o = wrap_torch_tensor(out)
    for _ in range(2):
        wmb.wholememory_gather_op(t, i, o, get_wholegraph_env_fns(), get_stream())
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        wmb.wholememory_gather_op(t, i, o, get_wholegraph_env_fns(), get_stream())
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    print("%-28s %8.3f ms  %8.1f GB/s out" % (label, ms, n * dim * 4 / ms / 1e6), flush=True)

run(local, idx_rand, "local random rows")
run(remote, idx_rand, "peer random rows")
run(remote, idx_sorted, "peer sorted rows")
run(remote, idx_seq, "peer sequential rows")
a = torch.empty((n, dim), dtype=torch.float32, device="cuda:0")
for _ in range(2): a.copy_(remote[:n])
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5): a.copy_(remote[:n])
e1.record(); torch.cuda.synchronize()
print("torch peer copy             %8.3f ms  %8.1f GB/s" % (e0.elapsed_time(e1)/5, n*dim*4/(e0.elapsed_time(e1)/5)/1e6))
