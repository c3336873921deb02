"""Communicators (role of the reference's pylibwholegraph/torch/comm.py).

The unique id is created on the group root and broadcast with torch.distributed exactly as the
reference does; the communicator itself is the shared-memory rendezvous of libwholegraph_b200
(one NVSwitch box), so no second NCCL communicator is created.
"""
import pylibwholegraph.binding.wholememory_binding as wmb
from pylibwholegraph.utils.imports import import_optional
from .utils import (
    str_to_wmb_wholememory_memory_type,
    str_to_wmb_wholememory_location,
    str_to_wmb_wholememory_distributed_backend,
    wholememory_distributed_backend_type_to_str,
)

torch = import_optional("torch")

global_communicators = {}
local_node_communicator = None
local_device_communicator = None
all_comm_world_rank = 0
all_comm_world_size = 1
all_comm_local_rank = 0
all_comm_local_size = 1


def reset_communicators():
    global global_communicators, local_node_communicator, local_device_communicator
    global_communicators = {}
    local_node_communicator = None
    local_device_communicator = None


def set_world_info(world_rank: int, world_size: int, local_rank: int, local_size: int):
    global all_comm_world_rank, all_comm_world_size, all_comm_local_rank, all_comm_local_size
    all_comm_world_rank, all_comm_world_size = world_rank, world_size
    all_comm_local_rank, all_comm_local_size = local_rank, local_size


class WholeMemoryCommunicator(object):
    def __init__(self, wmb_comm: wmb.PyWholeMemoryComm):
        self.wmb_comm = wmb_comm

    def get_rank(self):
        return self.wmb_comm.get_rank()

    def get_size(self):
        return self.wmb_comm.get_size()

    def barrier(self):
        return self.wmb_comm.barrier()

    def support_type_location(self, memory_type: str, memory_location: str):
        return self.wmb_comm.support_type_location(
            str_to_wmb_wholememory_memory_type(memory_type), str_to_wmb_wholememory_location(memory_location)
        )

    def destroy(self):
        wmb.destroy_communicator(self.wmb_comm)
        self.wmb_comm = None

    @property
    def distributed_backend(self):
        return wholememory_distributed_backend_type_to_str(self.wmb_comm.get_distributed_backend())

    @distributed_backend.setter
    def distributed_backend(self, value):
        self.wmb_comm.set_distributed_backend(str_to_wmb_wholememory_distributed_backend(value))


def _broadcast_unique_id(uid, root: int):
    """128 bytes from `root` to everybody, over whatever backend torch.distributed runs on."""
    dist = torch.distributed
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return uid
    dev = "cuda" if dist.get_backend() == "nccl" else "cpu"
    buf = torch.frombuffer(bytearray(uid.get_bytes()), dtype=torch.uint8).to(dev)
    dist.broadcast(buf, root)
    uid.set_bytes(bytes(buf.cpu().numpy().tobytes()))
    return uid


def create_group_communicator(group_size: int = -1, comm_stride: int = 1):
    """Groups of `group_size` ranks, members `comm_stride` apart (reference comm.py:132-171)."""
    dist = torch.distributed
    initialized = dist.is_available() and dist.is_initialized()
    world_size = dist.get_world_size() if initialized else 1
    world_rank = dist.get_rank() if initialized else 0
    if group_size == -1:
        group_size = world_size
    span = group_size * comm_stride
    assert world_size % span == 0
    my_block, in_block = divmod(world_rank, span)
    my_lane, my_index = in_block % comm_stride, in_block // comm_stride
    mine = None
    for block in range(world_size // span):
        for lane in range(comm_stride):
            root = block * span + lane
            uid = wmb.create_unique_id() if world_rank == root else wmb.PyWholeMemoryUniqueID()
            uid = _broadcast_unique_id(uid, root)
            if block == my_block and lane == my_lane:
                mine = uid
    return WholeMemoryCommunicator(wmb.create_communicator(mine, my_index, group_size))


def split_communicator(comm: WholeMemoryCommunicator, color: int, key: int = 0):
    if not isinstance(color, int) or not isinstance(key, int):
        raise TypeError("color and key must be int")
    if color < 0:
        return None
    return WholeMemoryCommunicator(wmb.split_communicator(comm.wmb_comm, color, key))


def destroy_communicator(wm_comm: WholeMemoryCommunicator):
    if wm_comm is not None and wm_comm.wmb_comm is not None:
        wmb.destroy_communicator(wm_comm.wmb_comm)
        wm_comm.wmb_comm = None


def comm_set_distributed_backend(wm_comm: WholeMemoryCommunicator, distributed_backend: str):
    wmb.communicator_set_distributed_backend(wm_comm.wmb_comm, str_to_wmb_wholememory_distributed_backend(distributed_backend))


def get_global_communicator(distributed_backend="nccl"):
    global local_node_communicator, local_device_communicator
    if distributed_backend not in global_communicators:
        comm = create_group_communicator()
        comm_set_distributed_backend(comm, distributed_backend)
        global_communicators[distributed_backend] = comm
        if distributed_backend == "nccl":
            if local_node_communicator is None and all_comm_local_size == all_comm_world_size:
                local_node_communicator = comm
            if local_device_communicator is None and all_comm_world_size == 1:
                local_device_communicator = comm
    return global_communicators[distributed_backend]


def get_local_node_communicator():
    global local_node_communicator
    if local_node_communicator is None:
        local_node_communicator = create_group_communicator(all_comm_local_size, 1)
    return local_node_communicator


def get_local_device_communicator():
    global local_device_communicator
    if local_device_communicator is None:
        local_device_communicator = create_group_communicator(1, 1)
    return local_device_communicator
