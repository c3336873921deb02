"""dtype / enum conversions (role of the reference's pylibwholegraph/torch/utils.py)."""
import pylibwholegraph.binding.wholememory_binding as wmb
from pylibwholegraph.utils.imports import import_optional

torch = import_optional("torch")

WholeMemoryDataType = wmb.WholeMemoryDataType

_PAIRS = None


def _pairs():
    global _PAIRS
    if _PAIRS is None:
        _PAIRS = [
            (torch.float32, WholeMemoryDataType.DtFloat),
            (torch.float16, WholeMemoryDataType.DtHalf),
            (torch.float64, WholeMemoryDataType.DtDouble),
            (torch.bfloat16, WholeMemoryDataType.DtBF16),
            (torch.int32, WholeMemoryDataType.DtInt),
            (torch.int64, WholeMemoryDataType.DtInt64),
            (torch.int16, WholeMemoryDataType.DtInt16),
            (torch.int8, WholeMemoryDataType.DtInt8),
        ]
    return _PAIRS


def torch_dtype_to_wholememory_dtype(torch_dtype):
    for t, w in _pairs():
        if t == torch_dtype:
            return w
    raise ValueError(f"torch dtype {torch_dtype} has no WholeMemory equivalent")


def wholememory_dtype_to_torch_dtype(wm_dtype):
    for t, w in _pairs():
        if int(w) == int(wm_dtype):
            return t
    raise ValueError(f"WholeMemory dtype {wm_dtype} has no torch equivalent")


def get_file_size(filename: str) -> int:
    import os

    if not os.path.isfile(filename):
        raise ValueError("File %s not found or not file" % (filename,))
    if not os.access(filename, os.R_OK):
        raise ValueError("File %s not readable" % (filename,))
    return os.path.getsize(filename)


def str_to_wmb_wholememory_memory_type(s: str):
    table = {
        "continuous": wmb.WholeMemoryMemoryType.MtContinuous,
        "chunked": wmb.WholeMemoryMemoryType.MtChunked,
        "distributed": wmb.WholeMemoryMemoryType.MtDistributed,
        "hierarchy": wmb.WholeMemoryMemoryType.MtHierarchy,
    }
    if s not in table:
        raise ValueError("WholeMemory type %s not supported, should be (continuous, chunked, distributed, hierarchy)" % (s,))
    return table[s]


def str_to_wmb_wholememory_location(s: str):
    table = {"cuda": wmb.WholeMemoryMemoryLocation.MlDevice, "cpu": wmb.WholeMemoryMemoryLocation.MlHost}
    if s not in table:
        raise ValueError("WholeMemory location %s not supported, should be (cuda, cpu)" % (s,))
    return table[s]


def str_to_wmb_wholememory_distributed_backend(s: str):
    table = {"nccl": wmb.WholeMemoryDistributedBackend.DbNCCL, "nvshmem": wmb.WholeMemoryDistributedBackend.DbNVSHMEM}
    if s not in table:
        raise ValueError("WholeMemory backend %s not supported, should be (nccl, nvshmem)" % (s,))
    return table[s]


def wholememory_distributed_backend_type_to_str(b):
    return {int(wmb.WholeMemoryDistributedBackend.DbNCCL): "nccl", int(wmb.WholeMemoryDistributedBackend.DbNVSHMEM): "nvshmem"}[int(b)]


def get_part_file_name(prefix: str, part_id: int, part_count: int) -> str:
    return "%s_part_%d_of_%d" % (prefix, part_id, part_count)


def get_part_file_list(prefix: str, part_count: int):
    return [get_part_file_name(prefix, i, part_count) for i in range(part_count)]


def view_as_torch(view):
    """torch tensor aliasing a wmb.DeviceArrayView (zero copy)."""
    shape = view.__cuda_array_interface__["shape"]
    if 0 in shape:
        dt = torch.bfloat16 if view.is_bf16 else None
        t = torch.empty(shape, device="cuda", dtype=dt if dt else torch.float32)
        return t
    t = torch.as_tensor(view, device="cuda")
    if view.is_bf16:
        t = t.view(torch.bfloat16)
    t._wm_owner = view.owner  # keep the allocation alive as long as the alias exists
    return t
