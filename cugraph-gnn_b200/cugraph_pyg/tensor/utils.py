"""Helpers of the distributed tensors (role of the reference's cugraph_pyg/tensor/utils.py)."""
import warnings

import torch

import pylibwholegraph.torch as wgth

_warned_cpu = False


def has_nvlink_network() -> bool:
    """One NVSwitch box: every GPU reaches every peer by load/store."""
    return True


def is_empty(a) -> bool:
    return a is None or a.numel() == 0


def empty(dim: int = 1):
    return torch.tensor([], dtype=torch.float32).reshape((0,) * dim) if dim > 1 else torch.tensor([], dtype=torch.float32)


def backend_to_memory_type(backend: str) -> str:
    """reference mapping (tensor/utils.py:52-59): "vmm" -> continuous, "nccl" -> distributed, "chunked" -> chunked.
    Here every type is peer-mapped HBM read by P2P from inside the kernels; multi-rank continuous and distributed
    are served by the chunked layout."""
    table = {"vmm": "chunked", "chunked": "chunked", "nccl": "distributed", "continuous": "continuous", "distributed": "distributed"}
    if backend not in table:
        raise ValueError(f"Unsupported backend: {backend}")
    if backend == "nvshmem":
        raise ValueError("NVSHMEM backend is not supported")
    return table[backend]


def resolve_location(device: str) -> str:
    """The reference defaults to pinned host memory ('cpu'); on B200 the tables live in HBM (180 GB per GPU)."""
    global _warned_cpu
    dev = "cuda" if str(device).startswith("cuda") else "cpu"
    if dev == "cpu" and not _warned_cpu:
        warnings.warn("location='cpu' is kept for API compatibility; this build stores WholeMemory tables in HBM")
        _warned_cpu = True
    return "cuda"


def get_comm():
    world = torch.distributed.get_world_size() if torch.distributed.is_available() and torch.distributed.is_initialized() else 1
    rank = torch.distributed.get_rank() if world > 1 else 0
    if not getattr(get_comm, "_init", False):
        import os

        local_rank = int(os.environ.get("LOCAL_RANK", rank))
        wgth.init(rank, world, local_rank, int(os.environ.get("LOCAL_WORLD_SIZE", world)))
        get_comm._init = True
    return wgth.get_global_communicator()


def create_wg_dist_tensor(shape, dtype, location="cuda", backend="vmm", partition_book=None):
    """Collective: WholeMemory tensor (1-D) or embedding (2-D) over the global communicator."""
    comm = get_comm()
    mem_type = backend_to_memory_type(backend)
    resolve_location(location)
    part = None if partition_book is None else [int(x) for x in partition_book]
    if len(shape) == 2:
        return wgth.create_embedding(comm, mem_type, "cuda", dtype, [int(shape[0]), int(shape[1])], embedding_entry_partition=part)
    if len(shape) == 1:
        return wgth.create_wholememory_tensor(comm, mem_type, "cuda", [int(shape[0])], dtype, [1], part)
    raise ValueError("The shape of the tensor must be 1D or 2D.")
