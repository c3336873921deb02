"""Fused multi-hop, multi-label neighbour sampler (B200 extension, include/wholememory/b200_ops.h).

One native call replaces the reference's per-hop python loop of sample + append_unique
(pylibwholegraph/torch/graph_structure.py:136-196) and produces the pylibcugraph-shaped result that
cugraph-pyg's readers decode (cugraph_pyg/sampler/sampler.py:525-740).
"""
import ctypes
from typing import List, Optional

import pylibwholegraph.binding.wholememory_binding as wmb
from pylibwholegraph.utils.imports import import_optional
from .wholegraph_env import get_stream, TorchMemoryContext, get_wholegraph_env_fns, wrap_torch_tensor

torch = import_optional("torch")

FLAG_CSR = 1
FLAG_INT64_IDS = 2

_vp = ctypes.c_void_p
_create = wmb.native_symbol("wholegraph_create_multihop_sampler")
_create.restype = ctypes.c_int
_create.argtypes = [ctypes.POINTER(_vp)]
_destroy = wmb.native_symbol("wholegraph_destroy_multihop_sampler")
_destroy.restype = ctypes.c_int
_destroy.argtypes = [_vp]
_begin = wmb.native_symbol("wholegraph_multihop_neighbor_sample_begin")
_begin.restype = ctypes.c_int
_begin.argtypes = [_vp, _vp, _vp, _vp, _vp, _vp, _vp, ctypes.POINTER(ctypes.c_int), ctypes.c_int, ctypes.c_ulonglong,
                   ctypes.c_int, _vp]
_finish = wmb.native_symbol("wholegraph_multihop_neighbor_sample_finish")
_finish.restype = ctypes.c_int
_finish.argtypes = [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]

_hbegin = wmb.native_symbol("wholegraph_hetero_multihop_neighbor_sample_begin")
_hbegin.restype = ctypes.c_int
_hbegin.argtypes = [_vp, ctypes.c_int, ctypes.POINTER(_vp), ctypes.POINTER(_vp), ctypes.POINTER(_vp), ctypes.POINTER(_vp),
                    ctypes.POINTER(ctypes.c_longlong), ctypes.c_int, _vp, _vp, ctypes.POINTER(ctypes.c_int), ctypes.c_int,
                    ctypes.c_ulonglong, ctypes.c_int, _vp]
_hfinish = wmb.native_symbol("wholegraph_hetero_multihop_neighbor_sample_finish")
_hfinish.restype = ctypes.c_int
_hfinish.argtypes = [_vp] * 13

_tbegin = wmb.native_symbol("wholegraph_temporal_multihop_neighbor_sample_begin")
_tbegin.restype = ctypes.c_int
_tbegin.argtypes = [_vp, ctypes.c_int, ctypes.POINTER(_vp), ctypes.POINTER(_vp), ctypes.POINTER(_vp), ctypes.POINTER(_vp), ctypes.POINTER(_vp),
                    ctypes.POINTER(ctypes.c_longlong), ctypes.c_int, ctypes.c_int, _vp, _vp, _vp, ctypes.POINTER(ctypes.c_int),
                    ctypes.c_int, ctypes.c_ulonglong, ctypes.c_int, ctypes.c_int, _vp]

# pylibcugraph's temporal_sampling_comparison strings -> time_comparison of the C call (include/wholememory/b200_ops.h)
TIME_COMPARISONS = {"strictly_increasing": 0, "monotonically_increasing": 1, "strictly_decreasing": 2,
                    "monotonically_decreasing": 3}

_seed_ids = wmb.native_symbol("wholegraph_multihop_seed_local_ids")
_seed_ids.restype = ctypes.c_int
_seed_ids.argtypes = [_vp, _vp, _vp, _vp]

_HETERO_OUT_NAMES = ("majors", "minors", "edge_id", "edge_type", "label_type_hop_offsets", "renumber_map", "renumber_map_offsets",
                     "edge_renumber_map", "edge_renumber_map_offsets", "label_type_step_base")

_OUT_NAMES = ("majors", "minors", "edge_id", "label_hop_offsets", "renumber_map", "renumber_map_offsets", "major_offsets",
              "label_step_base")


def _handle(t):
    """wholememory_tensor_t of a WholeMemoryTensor / PyWholeMemoryTensor / torch tensor / None (+ keep-alive)."""
    if t is None:
        return None, None
    if hasattr(t, "wmb_tensor"):
        return t.wmb_tensor.get_c_handle(), t
    if hasattr(t, "get_c_handle"):
        return t.get_c_handle(), t
    w = wrap_torch_tensor(t)
    return w.get_c_handle(), (w, t)


class MultiHopSampler(object):
    """Owns the persistent device scratch of the fused sampler; reuse one instance per loader."""

    def __init__(self):
        h = _vp()
        wmb.check_wholememory_error_code(_create(ctypes.byref(h)))
        self._h = h.value

    def __del__(self):
        if getattr(self, "_h", None):
            _destroy(self._h)
            self._h = None

    def sample_async(self, csr_row_ptr, csr_col, seeds: "torch.Tensor", label_offsets: "torch.Tensor", fanout: List[int],
                     random_state: int, *, csr_weight=None, csr_edge_id=None, compression: str = "COO",
                     int64_ids: bool = False) -> "PendingSample":
        """Enqueues every hop on the current stream and returns at once; ``.result()`` of the returned object waits
        for the output sizes only and yields the dict ``sample`` returns.  One call may be pending per sampler
        object -- loaders alternate between two objects to keep the device busy (cugraph_pyg.sampler)."""
        assert seeds.is_cuda and seeds.dim() == 1 and seeds.dtype in (torch.int32, torch.int64)
        label_offsets = label_offsets.to(device=seeds.device, dtype=torch.int64)
        assert compression in ("COO", "CSR")
        csr = compression == "CSR"
        flags = (FLAG_CSR if csr else 0) | (FLAG_INT64_IDS if int64_ids else 0)
        keep = []
        handles = []
        for t in (csr_row_ptr, csr_col, csr_weight, csr_edge_id, seeds, label_offsets):
            h, k = _handle(t)
            handles.append(h)
            keep.append(k)
        fan = (ctypes.c_int * len(fanout))(*[int(f) for f in fanout])
        err = _begin(self._h, *handles, fan, len(fanout), ctypes.c_ulonglong(random_state & 0xFFFFFFFFFFFFFFFF), flags,
                     get_stream())
        wmb.check_wholememory_error_code(err)
        return PendingSample(self, keep, csr, len(fanout))

    def seed_local_ids(self) -> "torch.Tensor":
        """int32 [S]: local id of every input seed (duplicates included) of the call that was last finished on this
        object; for heterogeneous calls local to the seed's (label, vertex type).  Call before the next sample*()."""
        ctx = TorchMemoryContext()
        wmb.check_wholememory_error_code(_seed_ids(self._h, ctx.get_c_context(), get_wholegraph_env_fns(), get_stream()))
        return ctx.get_tensor()

    def sample_hetero_async(self, csr_row_ptrs, csr_cols, vertex_type_offsets, seeds: "torch.Tensor",
                            label_offsets: "torch.Tensor", fanout: List[int], random_state: int, *, csr_weights=None,
                            csr_edge_ids=None, int64_ids: bool = False) -> "PendingHeteroSample":
        """Heterogeneous call group: one CSR per edge type over one global id space (``csr_row_ptrs[t]``,
        ``csr_cols[t]``), ``vertex_type_offsets`` [Vt+1] (host ints), ``fanout`` laid out [hop * T + edge type]."""
        assert seeds.is_cuda and seeds.dim() == 1 and seeds.dtype in (torch.int32, torch.int64)
        label_offsets = label_offsets.to(device=seeds.device, dtype=torch.int64)
        T = len(csr_row_ptrs)
        assert T >= 1 and len(csr_cols) == T and len(fanout) % T == 0
        keep = []

        def handle_array(ts):
            if ts is None:
                return None
            arr = (_vp * T)()
            for i, t in enumerate(ts):
                h, k = _handle(t)
                arr[i] = h
                keep.append(k)
            return arr

        rp, col = handle_array(csr_row_ptrs), handle_array(csr_cols)
        wgt, eid = handle_array(csr_weights), handle_array(csr_edge_ids)
        hs, ks = _handle(seeds)
        hl, kl = _handle(label_offsets)
        keep += [ks, kl]
        vto = [int(v) for v in vertex_type_offsets]
        vto_c = (ctypes.c_longlong * len(vto))(*vto)
        fan = (ctypes.c_int * len(fanout))(*[int(f) for f in fanout])
        hops = len(fanout) // T
        err = _hbegin(self._h, T, rp, col, wgt, eid, vto_c, len(vto) - 1, hs, hl, fan, hops,
                      ctypes.c_ulonglong(random_state & 0xFFFFFFFFFFFFFFFF), FLAG_INT64_IDS if int64_ids else 0, get_stream())
        wmb.check_wholememory_error_code(err)
        return PendingHeteroSample(self, keep, hops, len(vto) - 1)

    def sample_temporal_async(self, csr_row_ptrs, csr_cols, csr_edge_times, seeds: "torch.Tensor", seed_times: "torch.Tensor",
                              label_offsets: "torch.Tensor", fanout: List[int], random_state: int, comparison: str, *,
                              vertex_type_offsets=None, csr_edge_ids=None, csr_weights=None, compression: str = "COO",
                              int64_ids: bool = False):
        """Temporal call group (include/wholememory/b200_ops.h: wholegraph_temporal_multihop_neighbor_sample_begin).
        ``csr_edge_times[t]`` int64 per edge type in CSR order, ``seed_times`` int64 [S], ``comparison`` one of
        TIME_COMPARISONS.  ``vertex_type_offsets`` given: heterogeneous arguments and result (PendingHeteroSample);
        None: one edge type, homogeneous result (PendingSample).  ``csr_weights[t]`` (fp32 / fp64): biased instead of
        uniform selection among the eligible edges."""
        assert seeds.is_cuda and seeds.dim() == 1 and seeds.dtype in (torch.int32, torch.int64)
        if comparison not in TIME_COMPARISONS:
            raise ValueError("temporal comparison must be one of %s, got %r" % (sorted(TIME_COMPARISONS), comparison))
        label_offsets = label_offsets.to(device=seeds.device, dtype=torch.int64)
        seed_times = seed_times.to(device=seeds.device, dtype=torch.int64).contiguous()
        assert seed_times.shape == seeds.shape
        hetero = vertex_type_offsets is not None
        T = len(csr_row_ptrs)
        assert T >= 1 and len(csr_cols) == T and len(csr_edge_times) == T and len(fanout) % T == 0
        assert hetero or T == 1, "more than one edge type needs vertex_type_offsets"
        assert compression in ("COO", "CSR") and not (hetero and compression == "CSR")
        csr = compression == "CSR"
        keep = []

        def handle_array(ts):
            if ts is None:
                return None
            arr = (_vp * T)()
            for i, t in enumerate(ts):
                h, k = _handle(t)
                arr[i] = h
                keep.append(k)
            return arr

        rp, col, tim = handle_array(csr_row_ptrs), handle_array(csr_cols), handle_array(csr_edge_times)
        eid, wgt = handle_array(csr_edge_ids), handle_array(csr_weights)
        hs, ks = _handle(seeds)
        ht, kt = _handle(seed_times)
        hl, kl = _handle(label_offsets)
        keep += [ks, kt, kl]
        vto = [int(v) for v in vertex_type_offsets] if hetero else [0, 0]
        vto_c = (ctypes.c_longlong * len(vto))(*vto)
        fan = (ctypes.c_int * len(fanout))(*[int(f) for f in fanout])
        hops = len(fanout) // T
        flags = (FLAG_CSR if csr else 0) | (FLAG_INT64_IDS if int64_ids else 0)
        err = _tbegin(self._h, T, rp, col, wgt, tim, eid, vto_c, len(vto) - 1, 1 if hetero else 0, hs, ht, hl, fan, hops,
                      ctypes.c_ulonglong(random_state & 0xFFFFFFFFFFFFFFFF), TIME_COMPARISONS[comparison], flags, get_stream())
        wmb.check_wholememory_error_code(err)
        if hetero:
            return PendingHeteroSample(self, keep, hops, len(vto) - 1)
        return PendingSample(self, keep, csr, hops)

    def sample_temporal(self, *args, **kwargs):
        return self.sample_temporal_async(*args, **kwargs).result()

    def sample_hetero(self, *args, **kwargs):
        """Returns a dict with majors, minors, edge_id, edge_type, label_type_hop_offsets, renumber_map,
        renumber_map_offsets, edge_renumber_map, edge_renumber_map_offsets, label_type_step_base [L+1, Vt, B]."""
        return self.sample_hetero_async(*args, **kwargs).result()

    def sample(self, csr_row_ptr, csr_col, seeds: "torch.Tensor", label_offsets: "torch.Tensor", fanout: List[int],
               random_state: int, *, csr_weight=None, csr_edge_id=None, compression: str = "COO",
               int64_ids: bool = False):
        """Returns a dict with majors|major_offsets, minors, edge_id, label_hop_offsets, renumber_map,
        renumber_map_offsets, label_step_base (all CUDA tensors)."""
        return self.sample_async(csr_row_ptr, csr_col, seeds, label_offsets, fanout, random_state, csr_weight=csr_weight,
                                 csr_edge_id=csr_edge_id, compression=compression, int64_ids=int64_ids).result()


class PendingSample(object):
    """A call group whose hops are running on the device (MultiHopSampler.sample_async)."""

    def __init__(self, sampler: MultiHopSampler, keep, csr: bool, hops: int):
        self._sampler = sampler
        self._keep = keep  # inputs stay alive until the outputs exist
        self._csr = csr
        self._hops = hops
        self._out = None
        self.want_seed_local_ids = False  # set before result(): adds out["seed_local_ids"]

    def result(self):
        if self._out is not None:
            return self._out
        ctx = {n: TorchMemoryContext() for n in _OUT_NAMES}
        csr = self._csr
        err = _finish(
            self._sampler._h,
            None if csr else ctx["majors"].get_c_context(), ctx["minors"].get_c_context(), ctx["edge_id"].get_c_context(),
            ctx["label_hop_offsets"].get_c_context(), ctx["renumber_map"].get_c_context(),
            ctx["renumber_map_offsets"].get_c_context(), ctx["major_offsets"].get_c_context() if csr else None,
            ctx["label_step_base"].get_c_context(),
            get_wholegraph_env_fns(), get_stream(),
        )
        wmb.check_wholememory_error_code(err)
        out = {n: ctx[n].get_tensor() for n in _OUT_NAMES if ctx[n].get_tensor() is not None}
        # [L+1, B]: first local id of the vertices each label discovered at step t (0 = seeds)
        out["label_step_base"] = out["label_step_base"].view(self._hops + 1, -1)
        if self.want_seed_local_ids:
            out["seed_local_ids"] = self._sampler.seed_local_ids()
        self._out = out
        self._keep = None
        return out


class PendingHeteroSample(object):
    """A heterogeneous call group whose hops are running on the device (MultiHopSampler.sample_hetero_async)."""

    def __init__(self, sampler: MultiHopSampler, keep, hops: int, num_vertex_types: int):
        self._sampler = sampler
        self._keep = keep
        self._hops = hops
        self._vt = num_vertex_types
        self._out = None
        self.want_seed_local_ids = False

    def result(self):
        if self._out is not None:
            return self._out
        ctx = {n: TorchMemoryContext() for n in _HETERO_OUT_NAMES}
        err = _hfinish(self._sampler._h, *[ctx[n].get_c_context() for n in _HETERO_OUT_NAMES], get_wholegraph_env_fns(),
                       get_stream())
        wmb.check_wholememory_error_code(err)
        out = {n: ctx[n].get_tensor() for n in _HETERO_OUT_NAMES}
        out["label_type_step_base"] = out["label_type_step_base"].view(self._hops + 1, self._vt, -1)
        if self.want_seed_local_ids:
            out["seed_local_ids"] = self._sampler.seed_local_ids()
        self._out = out
        self._keep = None
        return out


_default_sampler: Optional[MultiHopSampler] = None


def multihop_neighbor_sample(csr_row_ptr, csr_col, seeds, label_offsets, fanout, random_state, **kwargs):
    """Functional form using a process-wide sampler object."""
    global _default_sampler
    if _default_sampler is None:
        _default_sampler = MultiHopSampler()
    return _default_sampler.sample(csr_row_ptr, csr_col, seeds, label_offsets, fanout, random_state, **kwargs)
