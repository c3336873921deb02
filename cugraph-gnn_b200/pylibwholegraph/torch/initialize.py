"""Process bootstrap (role of the reference's pylibwholegraph/torch/initialize.py)."""
import os

import pylibwholegraph.binding.wholememory_binding as wmb
from pylibwholegraph.utils.imports import import_optional
from .comm import set_world_info, get_global_communicator, get_local_node_communicator, reset_communicators

torch = import_optional("torch")

_LOG_LEVELS = {
    "fatal": wmb.WholeMemoryLogLevel.LevFatal,
    "error": wmb.WholeMemoryLogLevel.LevError,
    "warn": wmb.WholeMemoryLogLevel.LevWarn,
    "info": wmb.WholeMemoryLogLevel.LevInfo,
    "debug": wmb.WholeMemoryLogLevel.LevDebug,
    "trace": wmb.WholeMemoryLogLevel.LevTrace,
}


def init(world_rank: int, world_size: int, local_rank: int, local_size: int, wm_log_level="warn"):
    wmb.init(0, _LOG_LEVELS[wm_log_level])
    set_world_info(world_rank, world_size, local_rank, local_size)


def init_torch_env(world_rank: int, world_size: int, local_rank: int, local_size: int, wm_log_level="warn"):
    """torch.distributed (NCCL when a GPU is present) + WholeMemory init, one process per GPU."""
    os.environ["RANK"] = str(world_rank)
    os.environ["WORLD_SIZE"] = str(world_size)
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("MASTER_PORT", "12335")
    has_gpu = torch.cuda.is_available()
    if has_gpu:
        torch.cuda.set_device(local_rank)
    if not torch.distributed.is_initialized():
        torch.distributed.init_process_group(backend="nccl" if has_gpu else "gloo", init_method="env://")
    init(world_rank, world_size, local_rank, local_size, wm_log_level)


def init_torch_env_and_create_wm_comm(world_rank: int, world_size: int, local_rank: int, local_size: int,
                                      distributed_backend_type="nccl", wm_log_level="warn"):
    init_torch_env(world_rank, world_size, local_rank, local_size, wm_log_level)
    global_comm = get_global_communicator(distributed_backend_type)
    local_comm = get_local_node_communicator()
    return global_comm, local_comm


def finalize():
    wmb.finalize()
    reset_communicators()
    if torch.distributed.is_initialized():
        torch.distributed.destroy_process_group()
