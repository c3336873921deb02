"""Allocation environment: outputs and scratch of the native ops are torch tensors.

Same contract as the reference's pylibwholegraph/torch/wholegraph_env.py (TorchMemoryContext,
get_stream, get_wholegraph_env_fns, wrap_torch_tensor); the callbacks go through ctypes instead of a
JIT-compiled torch C++ extension, so nothing is compiled at import time.
"""
from typing import Union

import pylibwholegraph.binding.wholememory_binding as wmb
from pylibwholegraph.utils.imports import import_optional
from .utils import wholememory_dtype_to_torch_dtype, torch_dtype_to_wholememory_dtype

torch = import_optional("torch")

_default_env = None


def get_stream() -> int:
    """Current torch CUDA stream as an integer handle (0 = legacy default stream)."""
    return int(torch.cuda.current_stream().cuda_stream)


class TorchEmptyGlobalContext(object):
    pass


class TorchMemoryContext(object):
    """Receives the tensor a native op allocates through the env callbacks."""

    def __init__(self):
        self.tensor = None
        self._c = wmb.register_output_context(self)

    def get_c_context(self) -> int:
        return self._c

    def set_tensor(self, t):
        self.tensor = t

    def get_tensor(self):
        return self.tensor

    def free(self):
        self.tensor = None

    free_data = free

    def __del__(self):
        try:
            wmb.unregister_output_context(self)
        except Exception:  # interpreter shutdown
            pass


class _TempContext(object):
    def __init__(self):
        self.tensor = None

    def set_tensor(self, t):
        self.tensor = t

    def free(self):
        self.tensor = None

    free_data = free


def torch_create_memory_context_env_fn(global_context):
    return _TempContext()


def torch_destroy_memory_context_env_fn(memory_context, global_context):
    memory_context.free()


def torch_malloc_env_fn(shape, dtype_int, malloc_type_int, memory_context, global_context) -> int:
    if malloc_type_int == int(wmb.WholeMemoryMemoryAllocType.MatDevice):
        t = torch.empty(shape, dtype=wholememory_dtype_to_torch_dtype(dtype_int), device="cuda")
    else:
        pinned = malloc_type_int == int(wmb.WholeMemoryMemoryAllocType.MatPinned)
        t = torch.empty(shape, dtype=wholememory_dtype_to_torch_dtype(dtype_int), device="cpu", pin_memory=pinned)
    memory_context.set_tensor(t)
    return t.data_ptr()


def torch_free_env_fn(memory_context, global_context):
    memory_context.free_data()


def create_current_env_context():
    ctx = wmb.GlobalContextWrapper()
    g = TorchEmptyGlobalContext()
    ctx.create_context(
        torch_create_memory_context_env_fn,
        torch_destroy_memory_context_env_fn,
        torch_malloc_env_fn,
        torch_free_env_fn,
        g,
        torch_malloc_env_fn,
        torch_free_env_fn,
        g,
    )
    return ctx


def get_wholegraph_env_fns(use_default=True) -> int:
    global _default_env
    if not use_default:
        return create_current_env_context().get_env_fns()
    if _default_env is None:
        _default_env = create_current_env_context()
    return _default_env.get_env_fns()


def wrap_torch_tensor(t: Union["torch.Tensor", None]) -> wmb.WrappedLocalTensor:
    desc = wmb.PyWholeMemoryTensorDescription()
    wrapped = wmb.WrappedLocalTensor()
    if t is None:
        return wrapped.wrap_tensor(desc, 0)
    desc.set_dtype(torch_dtype_to_wholememory_dtype(t.dtype))
    desc.set_shape(tuple(t.shape))
    # empty tensors may report arbitrary strides (numpy gives 0): describe them as contiguous
    if t.numel() > 0:
        desc.set_stride(tuple(t.stride()))
    else:
        dense, acc = [], 1
        for s in reversed(t.shape):
            dense.insert(0, acc)
            acc *= max(int(s), 1)
        desc.set_stride(tuple(dense))
    desc.set_storage_offset(0)
    return wrapped.wrap_tensor(desc, t.data_ptr())
