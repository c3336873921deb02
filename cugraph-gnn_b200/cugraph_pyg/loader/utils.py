"""Loader helpers (role of the reference's cugraph_pyg/loader/utils.py:9-20)."""
import torch


def generate_seed() -> int:
    """One random sampler seed per iterator, identical on every rank (broadcast from rank 0)."""
    t = torch.randint(0, 2**31 - 1, (1,), dtype=torch.int64)
    d = torch.distributed
    if d.is_available() and d.is_initialized() and d.get_world_size() > 1:
        t = t.cuda() if d.get_backend() == "nccl" else t
        d.broadcast(t, src=0)
    return int(t)
