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


# ---- zero-copy torch views of raw device memory, through DLPack ---------------------------------------
# (torch.as_tensor on __cuda_array_interface__ asks the driver for pointer attributes, which fails for
#  cuMem peer mappings of another GPU's chunk; a DLPack capsule states device and layout explicitly.)
import ctypes as _ct


class _DLDevice(_ct.Structure):
    _fields_ = [("device_type", _ct.c_int32), ("device_id", _ct.c_int32)]


class _DLDataType(_ct.Structure):
    _fields_ = [("code", _ct.c_uint8), ("bits", _ct.c_uint8), ("lanes", _ct.c_uint16)]


class _DLTensor(_ct.Structure):
    _fields_ = [("data", _ct.c_void_p), ("device", _DLDevice), ("ndim", _ct.c_int32), ("dtype", _DLDataType),
                ("shape", _ct.POINTER(_ct.c_int64)), ("strides", _ct.POINTER(_ct.c_int64)), ("byte_offset", _ct.c_uint64)]


class _DLManagedTensor(_ct.Structure):
    pass


_DELETER = _ct.CFUNCTYPE(None, _ct.POINTER(_DLManagedTensor))
_DLManagedTensor._fields_ = [("dl_tensor", _DLTensor), ("manager_ctx", _ct.c_void_p), ("deleter", _DELETER)]

_DL_ALIVE = {}
# WholeMemory dtype -> (DLPack type code, bits): float 2, int 0, bfloat 4
_DL_TYPES = {1: (2, 32), 2: (2, 16), 3: (2, 64), 4: (4, 16), 5: (0, 32), 6: (0, 64), 7: (0, 16), 8: (0, 8)}


@_DELETER
def _dl_deleter(ptr):
    _DL_ALIVE.pop(_ct.addressof(ptr.contents), None)


def device_memory_as_torch(ptr: int, shape, strides_elts, wm_dtype: int, owner=None):
    """torch tensor aliasing `ptr` on the CURRENT cuda device; `owner` is kept alive with the tensor."""
    shape = tuple(int(x) for x in shape)
    if strides_elts is None:
        strides_elts, acc = [], 1
        for d in reversed(shape):
            strides_elts.insert(0, acc)
            acc *= max(d, 1)
    code, bits = _DL_TYPES[int(wm_dtype)]
    m = _DLManagedTensor()
    shp = (_ct.c_int64 * len(shape))(*shape)
    std = (_ct.c_int64 * len(shape))(*[int(x) for x in strides_elts])
    m.dl_tensor.data = _ct.c_void_p(ptr)
    m.dl_tensor.device = _DLDevice(2, torch.cuda.current_device())  # kDLCUDA
    m.dl_tensor.ndim = len(shape)
    m.dl_tensor.dtype = _DLDataType(code, bits, 1)
    m.dl_tensor.shape = shp
    m.dl_tensor.strides = std
    m.dl_tensor.byte_offset = 0
    m.manager_ctx = None
    m.deleter = _dl_deleter
    _DL_ALIVE[_ct.addressof(m)] = (m, shp, std, owner)
    _ct.pythonapi.PyCapsule_New.restype = _ct.py_object
    _ct.pythonapi.PyCapsule_New.argtypes = [_ct.c_void_p, _ct.c_char_p, _ct.c_void_p]
    capsule = _ct.pythonapi.PyCapsule_New(_ct.addressof(m), b"dltensor", None)
    return torch.utils.dlpack.from_dlpack(capsule)


def view_as_torch(view):
    """torch tensor aliasing a wmb.DeviceArrayView (zero copy)."""
    cai = view.__cuda_array_interface__
    shape = cai["shape"]
    if 0 in shape or cai["data"][0] == 0:
        return torch.empty(shape, device="cuda", dtype=wholememory_dtype_to_torch_dtype(view.wm_dtype))
    return device_memory_as_torch(cai["data"][0], shape, view.strides_elts, view.wm_dtype, owner=view.owner)


def str_to_wmb_wholememory_optimizer_type(str_wmb_optimizer_type: str):
    """reference: pylibwholegraph/torch/utils.py (sgd, adam = lazy adam, adagrad, rmsprop)."""
    table = {
        "sgd": wmb.WholeMemoryOptimizerType.OptSgd,
        "adam": wmb.WholeMemoryOptimizerType.OptLazyAdam,
        "adagrad": wmb.WholeMemoryOptimizerType.OptAdaGrad,
        "rmsprop": wmb.WholeMemoryOptimizerType.OptRmsProp,
    }
    if str_wmb_optimizer_type not in table:
        raise ValueError("WholeMemory optimizer type %s not supported, should be (sgd, adam, adagrad, rmsprop)" % (str_wmb_optimizer_type,))
    return table[str_wmb_optimizer_type]
