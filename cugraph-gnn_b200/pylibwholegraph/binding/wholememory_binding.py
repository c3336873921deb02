"""ctypes binding of the libwholegraph_b200 C ABI.

Plays the role of the reference's Cython module
``python/pylibwholegraph/pylibwholegraph/binding/wholememory_binding.pyx`` (same public names for
everything the hot path uses: enums, PyWholeMemoryComm/Handle/Tensor, WrappedLocalTensor,
GlobalContextWrapper, wholememory_gather_op / wholememory_scatter_op,
csr_{un}weighted_sample_without_replacement, append_unique, EmbeddingGatherForward, ...).
Error codes map to Python exceptions exactly as the .pyx does (:241-263).

There is deliberately NO fallback: if the shared library cannot be loaded the import fails.
"""
import ctypes
import enum
import os
import sys

_PKG_ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
_LIB_PATH = os.path.join(_PKG_ROOT, "lib", "libwholegraph_b200.so")


def _load_library():
    if not os.path.exists(_LIB_PATH):
        # building is cheap (nvcc cross-compiles without a GPU); do it once, in-tree
        sys.path.insert(0, _PKG_ROOT)
        try:
            import build as _b200_build  # type: ignore

            _b200_build.build()
        finally:
            sys.path.pop(0)
    try:
        return ctypes.CDLL(_LIB_PATH, mode=ctypes.RTLD_GLOBAL)
    except OSError as e:  # pragma: no cover - loud failure is the point
        raise ImportError(
            f"libwholegraph_b200.so could not be loaded from {_LIB_PATH}: {e}. "
            "Build it with `python cugraph-gnn_b200/build.py` (needs nvcc + the CUDA runtime)."
        ) from e


_lib = _load_library()
LIBRARY_PATH = _LIB_PATH


# ---------------------------------------------------------------------------------------------
# enums (values follow include/wholememory/*.h)
# ---------------------------------------------------------------------------------------------
class WholeMemoryErrorCode(enum.IntEnum):
    Success = 0
    UnknowError = 1
    NotImplemented = 2
    LogicError = 3
    CUDAError = 4
    CommunicationError = 5
    InvalidInput = 6
    InvalidValue = 7
    OutOfMemory = 8
    NotSupported = 9
    SystemError = 10


class WholeMemoryMemoryType(enum.IntEnum):
    MtNone = 0
    MtContinuous = 1
    MtChunked = 2
    MtDistributed = 3
    MtHierarchy = 4


class WholeMemoryMemoryLocation(enum.IntEnum):
    MlNone = 0
    MlDevice = 1
    MlHost = 2


class WholeMemoryDistributedBackend(enum.IntEnum):
    DbNone = 0
    DbNCCL = 1
    DbNVSHMEM = 2


class WholeMemoryLogLevel(enum.IntEnum):
    LevFatal = 0
    LevError = 1
    LevWarn = 2
    LevInfo = 3
    LevDebug = 4
    LevTrace = 5


class WholeMemoryMemoryAllocType(enum.IntEnum):
    MatNone = 0
    MatDevice = 1
    MatHost = 2
    MatPinned = 3


class WholeMemoryDataType(enum.IntEnum):
    DtUnknown = 0
    DtFloat = 1
    DtHalf = 2
    DtDouble = 3
    DtBF16 = 4
    DtInt = 5
    DtInt64 = 6
    DtInt16 = 7
    DtInt8 = 8
    DtCount = 9


class WholeMemoryAccessType(enum.IntEnum):
    AtNone = 0
    AtReadOnly = 1
    AtReadWrite = 2


_DT_SIZE = {1: 4, 2: 2, 3: 8, 4: 2, 5: 4, 6: 8, 7: 2, 8: 1}


def check_wholememory_error_code(err):
    """reference: wholememory_binding.pyx:241-263"""
    err = int(err)
    if err == WholeMemoryErrorCode.Success:
        return
    name = WholeMemoryErrorCode(err).name if err in WholeMemoryErrorCode._value2member_map_ else str(err)
    if err == WholeMemoryErrorCode.NotImplemented:
        raise NotImplementedError("libwholegraph_b200: not implemented")
    if err in (WholeMemoryErrorCode.InvalidInput, WholeMemoryErrorCode.InvalidValue):
        raise ValueError(f"libwholegraph_b200: {name}")
    if err == WholeMemoryErrorCode.OutOfMemory:
        raise MemoryError("libwholegraph_b200: out of memory")
    if err == WholeMemoryErrorCode.NotSupported:
        raise NotImplementedError("libwholegraph_b200: not supported on a single NVSwitch box build")
    if err == WholeMemoryErrorCode.SystemError:
        raise SystemError("libwholegraph_b200: system error")
    raise RuntimeError(f"libwholegraph_b200: {name}")


# ---------------------------------------------------------------------------------------------
# C structs
# ---------------------------------------------------------------------------------------------
MAX_DIM = 8


class _TensorDesc(ctypes.Structure):
    _fields_ = [
        ("sizes", ctypes.c_int64 * MAX_DIM),
        ("strides", ctypes.c_int64 * MAX_DIM),
        ("storage_offset", ctypes.c_int64),
        ("dim", ctypes.c_int),
        ("dtype", ctypes.c_int),
    ]


class _Gref(ctypes.Structure):
    _fields_ = [
        ("pointer", ctypes.c_void_p),
        ("rank_memory_offsets", ctypes.c_void_p),
        ("world_size", ctypes.c_int),
        ("stride", ctypes.c_size_t),
        ("same_chunk", ctypes.c_bool),
    ]


class _UniqueId(ctypes.Structure):
    _fields_ = [("internal", ctypes.c_char * 128)]


_CREATE_CTX = ctypes.CFUNCTYPE(None, ctypes.POINTER(ctypes.c_void_p), ctypes.c_void_p)
_DESTROY_CTX = ctypes.CFUNCTYPE(None, ctypes.c_void_p, ctypes.c_void_p)
_MALLOC = ctypes.CFUNCTYPE(ctypes.c_void_p, ctypes.POINTER(_TensorDesc), ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p)
_FREE = ctypes.CFUNCTYPE(None, ctypes.c_void_p, ctypes.c_void_p)


class _TempFns(ctypes.Structure):
    _fields_ = [
        ("create_memory_context_fn", _CREATE_CTX),
        ("destroy_memory_context_fn", _DESTROY_CTX),
        ("malloc_fn", _MALLOC),
        ("free_fn", _FREE),
        ("global_context", ctypes.c_void_p),
    ]


class _OutFns(ctypes.Structure):
    _fields_ = [("malloc_fn", _MALLOC), ("free_fn", _FREE), ("global_context", ctypes.c_void_p)]


class _EnvFns(ctypes.Structure):
    _fields_ = [("temporary_fns", _TempFns), ("output_fns", _OutFns)]


def _sig(name, restype, *argtypes):
    fn = getattr(_lib, name)
    fn.restype = restype
    fn.argtypes = list(argtypes)
    return fn


_vp = ctypes.c_void_p
_i = ctypes.c_int
_sz = ctypes.c_size_t
_psz = ctypes.POINTER(ctypes.c_size_t)

_c_init = _sig("wholememory_init", _i, ctypes.c_uint, _i)
_c_finalize = _sig("wholememory_finalize", _i)
_c_create_uid = _sig("wholememory_create_unique_id", _i, ctypes.POINTER(_UniqueId))
_c_create_comm = _sig("wholememory_create_communicator", _i, ctypes.POINTER(_vp), _UniqueId, _i, _i)
_c_destroy_comm = _sig("wholememory_destroy_communicator", _i, _vp)
_c_comm_rank = _sig("wholememory_communicator_get_rank", _i, ctypes.POINTER(_i), _vp)
_c_comm_size = _sig("wholememory_communicator_get_size", _i, ctypes.POINTER(_i), _vp)
_c_comm_barrier = _sig("wholememory_communicator_barrier", _i, _vp)
_c_comm_support = _sig("wholememory_communicator_support_type_location", _i, _vp, _i, _i)
_c_comm_set_backend = _sig("wholememory_communicator_set_distributed_backend", _i, _vp, _i)
_c_comm_get_backend = _sig("wholememory_communicator_get_distributed_backend", _i, _vp)
_c_malloc = _sig("wholememory_malloc", _i, ctypes.POINTER(_vp), _sz, _vp, _i, _i, _sz, _psz)
_c_free = _sig("wholememory_free", _i, _vp)
_c_get_comm = _sig("wholememory_get_communicator", _i, ctypes.POINTER(_vp), _vp)
_c_get_type = _sig("wholememory_get_memory_type", _i, _vp)
_c_get_loc = _sig("wholememory_get_memory_location", _i, _vp)
_c_get_total = _sig("wholememory_get_total_size", _sz, _vp)
_c_get_local = _sig("wholememory_get_local_memory", _i, ctypes.POINTER(_vp), _psz, _psz, _vp)
_c_get_rank_mem = _sig("wholememory_get_rank_memory", _i, ctypes.POINTER(_vp), _psz, _psz, _i, _vp)
_c_equal_plan = _sig("wholememory_equal_entry_partition_plan", _i, _psz, _sz, _i)
_c_load_file = _sig("wholememory_load_from_file", _i, _vp, _sz, _sz, _sz, ctypes.POINTER(ctypes.c_char_p), _i, _i)
_c_store_file = _sig("wholememory_store_to_file", _i, _vp, _sz, _sz, _sz, ctypes.c_char_p)
_c_fork_count = _sig("fork_get_device_count", _i)

_c_create_tensor = _sig("wholememory_create_tensor", _i, ctypes.POINTER(_vp), ctypes.POINTER(_TensorDesc), _vp, _i, _i, _psz)
_c_destroy_tensor = _sig("wholememory_destroy_tensor", _i, _vp)
_c_tensor_from_ptr = _sig("wholememory_make_tensor_from_pointer", _i, ctypes.POINTER(_vp), _vp, ctypes.POINTER(_TensorDesc))
_c_tensor_from_handle = _sig("wholememory_make_tensor_from_handle", _i, ctypes.POINTER(_vp), _vp, ctypes.POINTER(_TensorDesc))
_c_tensor_handle = _sig("wholememory_tensor_get_memory_handle", _vp, _vp)
_c_tensor_desc = _sig("wholememory_tensor_get_tensor_description", ctypes.POINTER(_TensorDesc), _vp)
_c_tensor_gref = _sig("wholememory_tensor_get_global_reference", _i, _vp, ctypes.POINTER(_Gref))
_c_tensor_sub = _sig("wholememory_tensor_get_subtensor", _i, _vp, ctypes.POINTER(ctypes.c_int64), ctypes.POINTER(ctypes.c_int64), ctypes.POINTER(_vp))
_c_tensor_local_count = _sig("wholememory_tensor_get_local_entry_count", _i, _psz, _vp)
_c_tensor_local_start = _sig("wholememory_tensor_get_local_entry_start", _i, _psz, _vp)
_c_tensor_entry_offsets = _sig("wholememory_tensor_get_entry_offsets", _i, _psz, _vp)
_c_tensor_count = _sig("get_wholememory_tensor_count", ctypes.c_int64)

_c_gather = _sig("wholememory_gather", _i, _vp, _vp, _vp, _vp, _vp, _i)
_c_scatter = _sig("wholememory_scatter", _i, _vp, _vp, _vp, _vp, _vp, _i)
_c_sample_u = _sig("wholegraph_csr_unweighted_sample_without_replacement", _i, _vp, _vp, _vp, _i, _vp, _vp, _vp, _vp, ctypes.c_ulonglong, _vp, _vp)
_c_sample_w = _sig("wholegraph_csr_weighted_sample_without_replacement", _i, _vp, _vp, _vp, _vp, _i, _vp, _vp, _vp, _vp, ctypes.c_ulonglong, _vp, _vp)
_c_append_unique = _sig("graph_append_unique", _i, _vp, _vp, _vp, _vp, _vp, _vp)
_c_rand_int = _sig("generate_random_positive_int_cpu", _i, ctypes.c_int64, ctypes.c_int64, _vp)
_c_rand_exp = _sig("generate_exponential_distribution_negative_float_cpu", _i, ctypes.c_int64, ctypes.c_int64, _vp)

_c_create_emb = _sig("wholememory_create_embedding", _i, ctypes.POINTER(_vp), ctypes.POINTER(_TensorDesc), _vp, _i, _i, _vp, _psz, _i, _i)
_c_destroy_emb = _sig("wholememory_destroy_embedding", _i, _vp)
_c_emb_tensor = _sig("wholememory_embedding_get_embedding_tensor", _vp, _vp)
_c_emb_gather = _sig("wholememory_embedding_gather", _i, _vp, _vp, _vp, ctypes.c_bool, _vp, ctypes.c_int64)
_c_emb_set_hot = _sig("wholememory_embedding_set_hot_rows", _i, _vp, _vp, _vp)
_c_emb_hot_count = _sig("wholememory_embedding_hot_row_count", ctypes.c_longlong, _vp)
_c_opt_create = _sig("wholememory_create_embedding_optimizer", _i, ctypes.POINTER(_vp), _i)
_c_opt_set = _sig("wholememory_optimizer_set_parameter", _i, _vp, ctypes.c_char_p, _vp)
_c_opt_destroy = _sig("wholememory_destroy_embedding_optimizer", None, _vp)
_c_emb_set_opt = _sig("wholememory_embedding_set_optimizer", _i, _vp, _vp)
_c_emb_apply = _sig("wholememory_embedding_gather_gradient_apply", _i, _vp, _vp, _vp, ctypes.c_bool, ctypes.c_float, _vp, ctypes.c_int64)
_c_emb_state_names = _sig("wholememory_embedding_get_optimizer_state_names", ctypes.POINTER(ctypes.c_char_p), _vp)
_c_emb_state = _sig("wholememory_embedding_get_optimizer_state", _vp, _vp, ctypes.c_char_p)


def native_symbol(name):
    """Raw ctypes access to any exported symbol (used by the B200 extension wrappers)."""
    return getattr(_lib, name)


# ---------------------------------------------------------------------------------------------
# allocation-callback bridge (reference: GlobalContextWrapper, wholememory_binding.pyx:352-548)
# ---------------------------------------------------------------------------------------------
_LIVE = {}  # id -> python memory-context objects created on behalf of the C side


class GlobalContextWrapper:
    """Holds the python callbacks and exposes a wholememory_env_func_t* for the ops."""

    def __init__(self):
        self._env = _EnvFns()
        self._keep = []

    def create_context(self, temp_create_context_fn, temp_destroy_context_fn, temp_malloc_fn, temp_free_fn,
                       temp_global_context, output_malloc_fn, output_free_fn, output_global_context):
        def _shape(desc):
            d = desc.contents
            return tuple(int(d.sizes[k]) for k in range(d.dim)), int(d.dtype)

        def c_create(pctx, _g):
            obj = temp_create_context_fn(temp_global_context)
            _LIVE[id(obj)] = obj
            pctx[0] = id(obj)

        def c_destroy(ctx, _g):
            obj = _LIVE.pop(ctx, None)
            if obj is not None:
                temp_destroy_context_fn(obj, temp_global_context)

        def c_temp_malloc(desc, alloc_type, ctx, _g):
            shape, dt = _shape(desc)
            return temp_malloc_fn(shape, dt, int(alloc_type), _LIVE[ctx], temp_global_context)

        def c_temp_free(ctx, _g):
            temp_free_fn(_LIVE[ctx], temp_global_context)

        def c_out_malloc(desc, alloc_type, ctx, _g):
            shape, dt = _shape(desc)
            return output_malloc_fn(shape, dt, int(alloc_type), _LIVE[ctx], output_global_context)

        def c_out_free(ctx, _g):
            output_free_fn(_LIVE[ctx], output_global_context)

        cbs = (_CREATE_CTX(c_create), _DESTROY_CTX(c_destroy), _MALLOC(c_temp_malloc), _FREE(c_temp_free),
               _MALLOC(c_out_malloc), _FREE(c_out_free))
        self._keep = [cbs, temp_global_context, output_global_context]
        t = self._env.temporary_fns
        t.create_memory_context_fn, t.destroy_memory_context_fn, t.malloc_fn, t.free_fn = cbs[0], cbs[1], cbs[2], cbs[3]
        t.global_context = None
        o = self._env.output_fns
        o.malloc_fn, o.free_fn = cbs[4], cbs[5]
        o.global_context = None

    def get_env_fns(self) -> int:
        return ctypes.addressof(self._env)


def register_output_context(obj) -> int:
    """Output contexts are python objects; the C side only carries their id around."""
    _LIVE[id(obj)] = obj
    return id(obj)


def unregister_output_context(obj):
    _LIVE.pop(id(obj), None)


# ---------------------------------------------------------------------------------------------
# python-visible wrappers
# ---------------------------------------------------------------------------------------------
class PyWholeMemoryUniqueID:
    def __init__(self, raw: bytes = None):
        self._c = _UniqueId()
        if raw is not None:
            self.set_bytes(raw)

    def get_bytes(self) -> bytes:
        return ctypes.string_at(ctypes.addressof(self._c), 128)

    def set_bytes(self, raw: bytes):
        assert len(raw) == 128
        ctypes.memmove(ctypes.addressof(self._c), raw, 128)

    def __len__(self):
        return 128


def init(flags: int = 0, log_level: WholeMemoryLogLevel = WholeMemoryLogLevel.LevWarn):
    check_wholememory_error_code(_c_init(flags, int(log_level)))


def finalize():
    check_wholememory_error_code(_c_finalize())


def create_unique_id() -> PyWholeMemoryUniqueID:
    uid = PyWholeMemoryUniqueID()
    check_wholememory_error_code(_c_create_uid(ctypes.byref(uid._c)))
    return uid


class PyWholeMemoryComm:
    def __init__(self, c_handle=None):
        self._h = c_handle

    def get_c_handle(self) -> int:
        return self._h

    def support_type_location(self, memory_type, memory_location) -> bool:
        return _c_comm_support(self._h, int(memory_type), int(memory_location)) == 0

    def get_rank(self) -> int:
        v = _i(-1)
        check_wholememory_error_code(_c_comm_rank(ctypes.byref(v), self._h))
        return v.value

    def get_size(self) -> int:
        v = _i(-1)
        check_wholememory_error_code(_c_comm_size(ctypes.byref(v), self._h))
        return v.value

    def barrier(self):
        check_wholememory_error_code(_c_comm_barrier(self._h))

    def get_distributed_backend(self):
        return WholeMemoryDistributedBackend(_c_comm_get_backend(self._h))

    def set_distributed_backend(self, backend):
        check_wholememory_error_code(_c_comm_set_backend(self._h, int(backend)))


def create_communicator(py_uid: PyWholeMemoryUniqueID, world_rank: int, world_size: int) -> PyWholeMemoryComm:
    h = _vp()
    check_wholememory_error_code(_c_create_comm(ctypes.byref(h), py_uid._c, world_rank, world_size))
    return PyWholeMemoryComm(h.value)


def destroy_communicator(py_comm: PyWholeMemoryComm):
    if py_comm is not None and py_comm._h:
        check_wholememory_error_code(_c_destroy_comm(py_comm._h))
        py_comm._h = None


def split_communicator(comm, color, key):
    raise NotImplementedError("split_communicator: multi-level communicators are outside the single-box hot path")


def communicator_set_distributed_backend(py_comm, backend):
    py_comm.set_distributed_backend(backend)


def equal_partition_plan(entry_count: int, world_size: int) -> int:
    v = ctypes.c_size_t(0)
    check_wholememory_error_code(_c_equal_plan(ctypes.byref(v), entry_count, world_size))
    return v.value


def fork_get_gpu_count() -> int:
    return int(_c_fork_count())


class PyWholeMemoryTensorDescription:
    def __init__(self):
        self._c = _TensorDesc()
        for k in range(MAX_DIM):
            self._c.sizes[k] = 1
            self._c.strides[k] = 1
        self._c.dim = 0
        self._c.dtype = 0
        self._c.storage_offset = 0

    def set_dtype(self, dtype):
        self._c.dtype = int(dtype)

    def set_shape(self, shape):
        assert 0 < len(shape) <= MAX_DIM
        self._c.dim = len(shape)
        for k, s in enumerate(shape):
            self._c.sizes[k] = int(s)

    def set_stride(self, strides):
        assert len(strides) == self._c.dim
        for k, s in enumerate(strides):
            self._c.strides[k] = int(s)

    def set_storage_offset(self, off):
        self._c.storage_offset = int(off)

    @property
    def dtype(self):
        return WholeMemoryDataType(self._c.dtype)

    def dim(self):
        return int(self._c.dim)

    @property
    def shape(self):
        return tuple(int(self._c.sizes[k]) for k in range(self._c.dim))

    def stride(self):
        return tuple(int(self._c.strides[k]) for k in range(self._c.dim))

    def storage_offset(self):
        return int(self._c.storage_offset)


class WrappedLocalTensor:
    """Non-owning wholememory_tensor_t view of caller memory (a torch tensor's data_ptr)."""

    def __init__(self):
        self._h = None

    def wrap_tensor(self, py_desc: PyWholeMemoryTensorDescription, data_ptr: int):
        if data_ptr == 0 and py_desc.dim() == 0:
            self._h = None  # "None" tensor
            return self
        h = _vp()
        check_wholememory_error_code(_c_tensor_from_ptr(ctypes.byref(h), _vp(data_ptr), ctypes.byref(py_desc._c)))
        self._h = h.value
        return self

    def get_c_handle(self):
        return self._h

    def __del__(self):
        if getattr(self, "_h", None):
            _c_destroy_tensor(self._h)
            self._h = None


class DeviceArrayView:
    """Exposes raw device memory through __cuda_array_interface__ so torch can alias it."""

    _TYPESTR = {1: "<f4", 2: "<f2", 3: "<f8", 4: "<i2", 5: "<i4", 6: "<i8", 7: "<i2", 8: "|i1"}

    def __init__(self, ptr, shape, dtype, strides_elts=None, owner=None):
        es = _DT_SIZE[int(dtype)]
        self.owner = owner
        self.wm_dtype = int(dtype)
        self.strides_elts = None if strides_elts is None else tuple(int(x) for x in strides_elts)
        self.is_bf16 = int(dtype) == WholeMemoryDataType.DtBF16
        self.__cuda_array_interface__ = {
            "shape": tuple(shape),
            "typestr": self._TYPESTR[int(dtype)],
            "data": (int(ptr) if ptr else 0, False),
            "version": 2,
            "strides": None if strides_elts is None else tuple(int(s) * es for s in strides_elts),
        }


class PyWholeMemoryHandle:
    def __init__(self, c_handle):
        self._h = c_handle

    def get_c_handle(self):
        return self._h

    def get_communicator(self):
        c = _vp()
        check_wholememory_error_code(_c_get_comm(ctypes.byref(c), self._h))
        return PyWholeMemoryComm(c.value)

    def get_memory_type(self):
        return WholeMemoryMemoryType(_c_get_type(self._h))

    def get_memory_location(self):
        return WholeMemoryMemoryLocation(_c_get_loc(self._h))

    def get_total_size(self):
        return int(_c_get_total(self._h))

    def get_local_memory(self):
        p, s, o = _vp(), ctypes.c_size_t(), ctypes.c_size_t()
        check_wholememory_error_code(_c_get_local(ctypes.byref(p), ctypes.byref(s), ctypes.byref(o), self._h))
        return p.value or 0, s.value, o.value

    def get_rank_memory(self, rank):
        p, s, o = _vp(), ctypes.c_size_t(), ctypes.c_size_t()
        check_wholememory_error_code(_c_get_rank_mem(ctypes.byref(p), ctypes.byref(s), ctypes.byref(o), rank, self._h))
        return p.value or 0, s.value, o.value

    def from_filelist(self, memory_offset, memory_entry_size, file_entry_size, round_robin_size, file_list):
        arr = (ctypes.c_char_p * len(file_list))(*[f.encode() for f in file_list])
        check_wholememory_error_code(
            _c_load_file(self._h, memory_offset, memory_entry_size, file_entry_size, arr, len(file_list), round_robin_size)
        )

    def to_file(self, memory_offset, memory_entry_size, file_entry_size, file_name):
        check_wholememory_error_code(_c_store_file(self._h, memory_offset, memory_entry_size, file_entry_size, file_name.encode()))


class PyWholeMemoryTensor:
    def __init__(self, c_handle, owner=True, parent=None):
        self._h = c_handle
        self._owner = owner
        self._parent = parent  # keeps the root alive for sub-tensors

    def get_c_handle(self):
        return self._h

    def _desc(self):
        return _c_tensor_desc(self._h).contents

    def get_wholememory_handle(self):
        return PyWholeMemoryHandle(_c_tensor_handle(self._h))

    @property
    def dtype(self):
        return WholeMemoryDataType(self._desc().dtype)

    def dim(self):
        return int(self._desc().dim)

    @property
    def shape(self):
        d = self._desc()
        return tuple(int(d.sizes[k]) for k in range(d.dim))

    def stride(self):
        d = self._desc()
        return tuple(int(d.strides[k]) for k in range(d.dim))

    def storage_offset(self):
        return int(self._desc().storage_offset)

    def get_local_entry_count(self):
        v = ctypes.c_size_t()
        check_wholememory_error_code(_c_tensor_local_count(ctypes.byref(v), self._h))
        return v.value

    def get_local_entry_start(self):
        v = ctypes.c_size_t()
        check_wholememory_error_code(_c_tensor_local_start(ctypes.byref(v), self._h))
        return v.value

    def get_entry_offsets(self):
        world = self.get_wholememory_handle().get_communicator().get_size()
        arr = (ctypes.c_size_t * (world + 1))()
        check_wholememory_error_code(_c_tensor_entry_offsets(arr, self._h))
        return [int(x) for x in arr]

    def get_sub_tensor(self, starts, ends):
        n = self.dim()
        assert len(starts) == n and len(ends) == n
        s = (ctypes.c_int64 * n)(*[int(x) for x in starts])
        e = (ctypes.c_int64 * n)(*[int(x) for x in ends])
        h = _vp()
        check_wholememory_error_code(_c_tensor_sub(self._h, s, e, ctypes.byref(h)))
        return PyWholeMemoryTensor(h.value, owner=True, parent=self)

    def get_global_reference(self):
        g = _Gref()
        check_wholememory_error_code(_c_tensor_gref(self._h, ctypes.byref(g)))
        return g

    def _view(self, ptr, rows):
        d = self._desc()
        es = _DT_SIZE[int(d.dtype)]
        base = ptr + int(d.storage_offset) * es
        if d.dim == 1:
            return DeviceArrayView(base, (rows,), d.dtype, None, owner=self)
        return DeviceArrayView(base, (rows, int(d.sizes[1])), d.dtype, (int(d.strides[0]), 1), owner=self)

    def get_local_view(self):
        """(DeviceArrayView of this rank's entries, first entry index)."""
        ptr, _, _ = self.get_wholememory_handle().get_local_memory()
        return self._view(ptr, self.get_local_entry_count()), self.get_local_entry_start()

    def get_rank_view(self, rank):
        ptr, size, _ = self.get_wholememory_handle().get_rank_memory(rank)
        offs = self.get_entry_offsets()
        return self._view(ptr, offs[rank + 1] - offs[rank]), offs[rank]

    def from_filelist(self, filelist, round_robin_size: int = 0):
        d = self._desc()
        es = _DT_SIZE[int(d.dtype)]
        if d.dim == 1:
            mem_entry, file_entry = es, es
        else:
            mem_entry, file_entry = int(d.strides[0]) * es, int(d.sizes[1]) * es
        self.get_wholememory_handle().from_filelist(int(d.storage_offset) * es, mem_entry, file_entry, round_robin_size, filelist)

    def to_file(self, filename):
        d = self._desc()
        es = _DT_SIZE[int(d.dtype)]
        if d.dim == 1:
            mem_entry, file_entry = es, es
        else:
            mem_entry, file_entry = int(d.strides[0]) * es, int(d.sizes[1]) * es
        self.get_wholememory_handle().to_file(int(d.storage_offset) * es, mem_entry, file_entry, filename)

    def destroy(self):
        if self._h and self._owner:
            check_wholememory_error_code(_c_destroy_tensor(self._h))
        self._h = None


def create_wholememory_tensor(tensor_description, comm, memory_type, memory_location, tensor_entry_partition=None):
    h = _vp()
    part = None
    if tensor_entry_partition is not None:
        part = (ctypes.c_size_t * len(tensor_entry_partition))(*[int(x) for x in tensor_entry_partition])
    check_wholememory_error_code(
        _c_create_tensor(ctypes.byref(h), ctypes.byref(tensor_description._c), comm.get_c_handle(), int(memory_type),
                         int(memory_location), part)
    )
    return PyWholeMemoryTensor(h.value)


def make_tensor_as_wholememory(tensor_description, data_ptr):
    h = _vp()
    check_wholememory_error_code(_c_tensor_from_ptr(ctypes.byref(h), _vp(data_ptr), ctypes.byref(tensor_description._c)))
    return PyWholeMemoryTensor(h.value)


def destroy_wholememory_tensor(t: PyWholeMemoryTensor):
    t.destroy()


def py_get_wholememory_tensor_count():
    return int(_c_tensor_count())


def _h(x):
    """c handle of a PyWholeMemoryTensor / WrappedLocalTensor / None"""
    return None if x is None else x.get_c_handle()


def wholememory_gather_op(wholememory_tensor, indices_tensor, output_tensor, p_env_fns_int, stream_int, gather_sms=-1):
    check_wholememory_error_code(
        _c_gather(_h(wholememory_tensor), _h(indices_tensor), _h(output_tensor), _vp(p_env_fns_int), _vp(stream_int), gather_sms)
    )


def wholememory_scatter_op(input_tensor, indices_tensor, wholememory_tensor, p_env_fns_int, stream_int, scatter_sms=-1):
    check_wholememory_error_code(
        _c_scatter(_h(input_tensor), _h(indices_tensor), _h(wholememory_tensor), _vp(p_env_fns_int), _vp(stream_int), scatter_sms)
    )


def csr_unweighted_sample_without_replacement(csr_row_ptr_tensor, csr_col_ptr_tensor, center_nodes_tensor,
                                              max_sample_count, output_sample_offset_tensor, output_dest_handle,
                                              output_center_localid_handle, output_edge_gid_handle, random_seed,
                                              p_env_fns_int, stream_int):
    check_wholememory_error_code(
        _c_sample_u(_h(csr_row_ptr_tensor), _h(csr_col_ptr_tensor), _h(center_nodes_tensor), max_sample_count,
                    _h(output_sample_offset_tensor), _vp(output_dest_handle or None),
                    _vp(output_center_localid_handle or None), _vp(output_edge_gid_handle or None),
                    ctypes.c_ulonglong(random_seed & 0xFFFFFFFFFFFFFFFF), _vp(p_env_fns_int), _vp(stream_int))
    )


def csr_weighted_sample_without_replacement(csr_row_ptr_tensor, csr_col_ptr_tensor, csr_weight_ptr_tensor,
                                            center_nodes_tensor, max_sample_count, output_sample_offset_tensor,
                                            output_dest_handle, output_center_localid_handle, output_edge_gid_handle,
                                            random_seed, p_env_fns_int, stream_int):
    check_wholememory_error_code(
        _c_sample_w(_h(csr_row_ptr_tensor), _h(csr_col_ptr_tensor), _h(csr_weight_ptr_tensor), _h(center_nodes_tensor),
                    max_sample_count, _h(output_sample_offset_tensor), _vp(output_dest_handle or None),
                    _vp(output_center_localid_handle or None), _vp(output_edge_gid_handle or None),
                    ctypes.c_ulonglong(random_seed & 0xFFFFFFFFFFFFFFFF), _vp(p_env_fns_int), _vp(stream_int))
    )


def append_unique(target_node_tensor, neighbor_node_tensor, output_unique_node_handle,
                  output_neighbor_raw_to_unique_mapping_tensor, p_env_fns_int, stream_int):
    check_wholememory_error_code(
        _c_append_unique(_h(target_node_tensor), _h(neighbor_node_tensor), _vp(output_unique_node_handle),
                         _h(output_neighbor_raw_to_unique_mapping_tensor), _vp(p_env_fns_int), _vp(stream_int))
    )


def host_generate_random_positive_int(random_seed, sub_sequence, output):
    check_wholememory_error_code(_c_rand_int(random_seed, sub_sequence, _h(output)))


def host_generate_exponential_distribution_negative_float(random_seed, sub_sequence, output):
    check_wholememory_error_code(_c_rand_exp(random_seed, sub_sequence, _h(output)))


# ---- embedding ------------------------------------------------------------------------------------
class WholeMemoryCachePolicy:
    """Only the 'no cache' policy exists on this path."""


def create_non_cache_policy():
    return WholeMemoryCachePolicy()


class PyWholeMemoryEmbedding:
    def __init__(self, c_handle=None):
        self._h = c_handle

    def get_c_handle(self):
        return self._h

    def get_embedding_tensor(self) -> PyWholeMemoryTensor:
        return PyWholeMemoryTensor(_c_emb_tensor(self._h), owner=False)

    def destroy_embedding(self):
        if self._h:
            check_wholememory_error_code(_c_destroy_emb(self._h))
            self._h = None

    def set_hot_rows(self, hot_indices, stream_int=0):
        """hot_indices: wrapped int64 device tensor (or None to drop the replica); see include/wholememory/b200_ops.h."""
        check_wholememory_error_code(_c_emb_set_hot(self._h, None if hot_indices is None else _h(hot_indices), _vp(stream_int or 0)))

    def hot_row_count(self) -> int:
        return int(_c_emb_hot_count(self._h))

    def get_optimizer_state_names(self):
        names, arr, i = [], _c_emb_state_names(self._h), 0
        while arr[i] is not None:
            names.append(arr[i].decode())
            i += 1
        return names

    def get_optimizer_state(self, state_name: str) -> PyWholeMemoryTensor:
        h = _c_emb_state(self._h, state_name.encode())
        if not h:
            raise ValueError(f"no optimizer state named {state_name!r}")
        return PyWholeMemoryTensor(h, owner=False)

    def writeback_all_cache(self, stream_int=0):
        pass

    def drop_all_cache(self, stream_int=0):
        pass


class WholeMemoryOptimizerType(enum.IntEnum):
    OptNone = 0
    OptSgd = 1
    OptLazyAdam = 2
    OptRmsProp = 3
    OptAdaGrad = 4


class WholeMemoryOptimizer:
    """wholememory_embedding_optimizer_t (reference: binding/wholememory_binding.pyx:672-708)."""

    def __init__(self):
        self._h = None
        self.optimizer_type = WholeMemoryOptimizerType.OptNone
        self.param_dict = {}

    def create_optimizer(self, optimizer_type, param_dict: dict):
        h = _vp()
        check_wholememory_error_code(_c_opt_create(ctypes.byref(h), int(optimizer_type)))
        self._h = h.value
        self.optimizer_type = WholeMemoryOptimizerType(int(optimizer_type))
        self.param_dict = dict(param_dict)
        for key, value in param_dict.items():
            v = ctypes.c_float(float(value))
            check_wholememory_error_code(_c_opt_set(self._h, key.encode(), ctypes.cast(ctypes.byref(v), _vp)))

    def add_embedding(self, embedding: PyWholeMemoryEmbedding):
        check_wholememory_error_code(_c_emb_set_opt(embedding.get_c_handle(), self._h))

    def destroy_optimizer(self):
        if self._h:
            _c_opt_destroy(self._h)
            self._h = None
        self.optimizer_type = WholeMemoryOptimizerType.OptNone


def create_optimizer(optimizer_type, param_dict: dict):
    o = WholeMemoryOptimizer()
    o.create_optimizer(optimizer_type, param_dict)
    return o


def create_non_optimizer():
    return WholeMemoryOptimizer()


def create_embedding(tensor_desc, comm, memory_type, memory_location, cache_policy=None,
                     embedding_entry_partition=None, user_defined_sms=-1, round_robin_size=0):
    h = _vp()
    part = None
    if embedding_entry_partition is not None:
        part = (ctypes.c_size_t * len(embedding_entry_partition))(*[int(x) for x in embedding_entry_partition])
    check_wholememory_error_code(
        _c_create_emb(ctypes.byref(h), ctypes.byref(tensor_desc._c), comm.get_c_handle(), int(memory_type),
                      int(memory_location), None, part, user_defined_sms, round_robin_size)
    )
    return PyWholeMemoryEmbedding(h.value)


def EmbeddingGatherGradientApply(wm_embedding, indice, grads, adjust_cache, lr, p_env_fns_int, stream_int):
    check_wholememory_error_code(
        _c_emb_apply(wm_embedding.get_c_handle(), _h(indice), _h(grads), bool(adjust_cache), float(lr), _vp(p_env_fns_int), stream_int or 0)
    )


def EmbeddingGatherForward(wm_embedding, indice, output, adjust_cache, p_env_fns_int, stream_int):
    check_wholememory_error_code(
        _c_emb_gather(wm_embedding.get_c_handle(), _h(indice), _h(output), bool(adjust_cache), _vp(p_env_fns_int), stream_int or 0)
    )
