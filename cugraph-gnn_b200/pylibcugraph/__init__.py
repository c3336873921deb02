"""pylibcugraph-shaped entry points on top of the B200 fused sampler.

cugraph-pyg's only use of pylibcugraph on the hot path is: build an SGGraph/MGGraph from COO arrays and
call ``{homogeneous,heterogeneous}_{uniform,biased}_neighbor_sample`` once per call group
(reference call sites: python/cugraph-pyg/cugraph_pyg/data/graph_store.py:263-329,
python/cugraph-pyg/cugraph_pyg/sampler/distributed_sampler.py:53-94, 784-819, 888-902).
libcugraph itself is not vendored in the reference tree; this module keeps the call signature and the
result dictionary so that code written against pylibcugraph runs unchanged, and routes the work to
``wholegraph_multihop_neighbor_sample`` (include/wholememory/b200_ops.h).

Differences that callers can observe:
  * arrays are torch CUDA tensors (the reference converts cupy -> torch right after the call anyway);
  * the random stream is this project's (S1/S2 geometry per hop), not libcugraph's -- unverifiable either way;
  * heterogeneous sampling keeps one CSR per edge type (built lazily from edge_type_array) and follows the definition in
    include/wholememory/b200_ops.h: edge_id indexes edge_renumber_map per (label, edge type);
  * temporal sampling (edge_start_time_array + the *_temporal_* entry points, uniform and biased) follows the definition in
    include/wholememory/b200_ops.h;
  * disjoint_sampling=True (COO) is the plain sample with cross-tree edges removed (_disjoint_filter, _disjoint_filter_hetero);
  * with_replacement=True raises NotImplementedError.
"""
from typing import Optional

import numpy as np
import torch

__version__ = "26.10.00+b200"


class ResourceHandle:
    """Placeholder for pylibcugraph.ResourceHandle (streams come from torch's current stream)."""

    def __init__(self, handle_ptr=None):
        self.handle_ptr = handle_ptr


class GraphProperties:
    def __init__(self, is_symmetric: bool = False, is_multigraph: bool = False):
        self.is_symmetric = is_symmetric
        self.is_multigraph = is_multigraph


def _as_cuda(a, dtype=None):
    if a is None:
        return None
    t = torch.as_tensor(a)
    if not t.is_cuda:
        t = t.cuda()
    if dtype is not None and t.dtype != dtype:
        t = t.to(dtype)
    return t.contiguous()


class SGGraph:
    """CSR by cuGraph source, resident on this GPU.

    The sampler follows out-edges of a seed, so rows are sources and columns destinations; cugraph-pyg passes
    src = PyG edge_index[1], dst = PyG edge_index[0] (graph_store.py:509-533), i.e. rows end up being PyG
    destinations and the sampled neighbours PyG sources (in-edges)."""

    def __init__(self, resource_handle, graph_properties, src_or_offset_array, dst_or_index_array, weight_array=None,
                 store_transposed=False, renumber=False, do_expensive_check=False, edge_id_array=None,
                 edge_type_array=None, input_array_format="COO", vertices_array=None, drop_self_loops=False,
                 drop_multi_edges=False, symmetrize=False, edge_start_time_array=None, edge_end_time_array=None,
                 num_vertices: Optional[int] = None):
        if edge_end_time_array is not None:
            raise NotImplementedError("edge end times are not supported (cugraph-pyg passes start times only)")
        if drop_self_loops or drop_multi_edges or symmetrize or renumber or store_transposed:
            raise NotImplementedError("graph transformations are not supported; pass the final COO")
        self.properties = graph_properties
        if input_array_format == "CSR":
            row_ptr = _as_cuda(src_or_offset_array, torch.int64)
            col = _as_cuda(dst_or_index_array)
            order = None
        elif input_array_format == "COO":
            src = _as_cuda(src_or_offset_array, torch.int64)
            dst = _as_cuda(dst_or_index_array, torch.int64)
            if num_vertices is None:
                if vertices_array is not None:
                    num_vertices = int(torch.as_tensor(vertices_array).max()) + 1 if len(vertices_array) else 0
                else:
                    num_vertices = int(max(src.max(), dst.max())) + 1 if src.numel() else 0
            order = torch.argsort(src, stable=True)
            col = dst[order]
            row_ptr = torch.zeros(num_vertices + 1, dtype=torch.int64, device=src.device)
            if src.numel():
                row_ptr[1:] = torch.bincount(src, minlength=num_vertices).cumsum(0)
        else:
            raise ValueError("input_array_format must be 'COO' or 'CSR'")
        self.num_vertices = int(row_ptr.numel() - 1)
        self.row_ptr = row_ptr
        # 32-bit columns halve the sampler's random-read traffic whenever the ids fit
        self.col = col.to(torch.int32) if self.num_vertices < 2**31 - 1 else col.to(torch.int64)
        pick = (lambda a, dt=None: None if a is None else (_as_cuda(a, dt)[order] if order is not None else _as_cuda(a, dt)))
        self.edge_id = pick(edge_id_array, torch.int64)
        if self.edge_id is None and order is not None:
            self.edge_id = order  # position in the caller's COO
        self.edge_type = pick(edge_type_array, torch.int32)
        self.weight = pick(weight_array)
        self.edge_time = pick(edge_start_time_array, torch.int64)
        if self.weight is not None and self.weight.dtype not in (torch.float32, torch.float64):
            self.weight = self.weight.float()
        self._sampler = None
        self._typed = {}

    def _typed_csrs(self, num_edge_types: int):
        """Per edge type: (row_ptr [V+1], col, weight|None, edge_id|None) over the global vertex id space, in the order
        of the combined CSR (stable), built once per num_edge_types."""
        if num_edge_types not in self._typed:
            if self.edge_type is None:
                if num_edge_types != 1:
                    raise ValueError("the graph was built without edge_type_array")
                self._typed[num_edge_types] = [(self.row_ptr, self.col, self.weight, self.edge_id, self.edge_time)]
            else:
                deg = self.row_ptr[1:] - self.row_ptr[:-1]
                rows = torch.repeat_interleave(torch.arange(self.num_vertices, device=self.col.device), deg)
                out = []
                for t in range(num_edge_types):
                    sel = torch.nonzero(self.edge_type == t).reshape(-1)
                    rp = torch.zeros(self.num_vertices + 1, dtype=torch.int64, device=self.col.device)
                    if sel.numel():
                        rp[1:] = torch.bincount(rows[sel], minlength=self.num_vertices).cumsum(0)
                    out.append((rp, self.col[sel].contiguous(),
                                None if self.weight is None else self.weight[sel].contiguous(),
                                None if self.edge_id is None else self.edge_id[sel].contiguous(),
                                None if self.edge_time is None else self.edge_time[sel].contiguous()))
                self._typed[num_edge_types] = out
        return self._typed[num_edge_types]

    @staticmethod
    def _drop_zero_weight(row_ptr, col, weight, edge_id, edge_time=None):
        """pylibcugraph never samples an edge whose bias is zero, even when the row has fewer candidates than the fan-out
        (pinned by the reference's test_neighbor_loader.py:97-133); the WholeGraph A-Res sampler takes every neighbour of
        a row with deg <= fan-out.  Sampling on the CSR without its zero-weight edges gives pylibcugraph's behaviour
        exactly; edge ids keep pointing at the caller's edges."""
        keep = weight > 0
        if bool(keep.all()):
            return row_ptr, col, weight, edge_id, edge_time
        n = row_ptr.numel() - 1
        rows = torch.repeat_interleave(torch.arange(n, device=col.device), row_ptr[1:] - row_ptr[:-1])
        sel = torch.nonzero(keep).reshape(-1)
        rp = torch.zeros(n + 1, dtype=torch.int64, device=col.device)
        if sel.numel():
            rp[1:] = torch.bincount(rows[sel], minlength=n).cumsum(0)
        if edge_id is None:
            edge_id = torch.arange(col.numel(), dtype=torch.int64, device=col.device)
        return (rp, col[sel].contiguous(), weight[sel].contiguous(), edge_id[sel].contiguous(),
                None if edge_time is None else edge_time[sel].contiguous())

    def _drop_zero_weight_cached(self):
        if "biased_all" not in self._typed:
            self._typed["biased_all"] = self._drop_zero_weight(self.row_ptr, self.col, self.weight, self.edge_id, self.edge_time)
        return self._typed["biased_all"]

    def _biased_csrs(self, num_edge_types: int):
        key = ("biased", num_edge_types)
        if key not in self._typed:
            self._typed[key] = [self._drop_zero_weight(*g) for g in self._typed_csrs(num_edge_types)]
        return self._typed[key]

    def _get_sampler(self):
        if self._sampler is None:
            from pylibwholegraph.torch.multihop import MultiHopSampler

            self._sampler = MultiHopSampler()
        return self._sampler


class MGGraph(SGGraph):
    """Multi-GPU graph: every rank contributes its edge partition, the CSR is REPLICATED on every GPU of the box
    (all-gather of the edge lists at construction).  Seeds are sharded by the callers, so sampling needs no
    communication at all; graphs that do not fit one B200 (180 GB) can be striped with WholeMemory tensors and
    sampled through pylibwholegraph.torch.MultiHopSampler directly."""

    def __init__(self, resource_handle, graph_properties, src_array, dst_array, weight_array=None, store_transposed=False,
                 do_expensive_check=False, edge_id_array=None, edge_type_array=None, vertices_array=None, num_arrays=1,
                 size=None, edge_start_time_array=None, **kwargs):
        import torch.distributed as dist

        multi = dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1
        world = dist.get_world_size() if multi else 1
        counts = None  # edges per rank, agreed on once: every per-edge array is laid out with the same offsets

        def local(a, dtype=None):
            if a is None:
                return None
            if isinstance(a, (list, tuple)):
                a = torch.cat([_as_cuda(x, dtype) for x in a]) if len(a) else torch.empty(0, dtype=dtype)
            return _as_cuda(a, dtype)

        def gather(a, dtype=None, per_edge=True):
            """all ranks' parts, concatenated in rank order, received straight into ONE buffer of the exact total size
            (no padding to the largest partition, no list of W padded copies, no torch.cat)"""
            nonlocal counts
            t = local(a, dtype)
            if t is None or not multi:
                return t
            if per_edge and counts is not None:
                sizes = counts
            else:
                n = torch.tensor([t.numel()], device=t.device, dtype=torch.int64)
                every = torch.empty(world, device=t.device, dtype=torch.int64)
                dist.all_gather_into_tensor(every, n)
                sizes = [int(x) for x in every.tolist()]
            if per_edge and counts is None:
                counts = sizes
                self._check_replicated_fit(sum(sizes), t.device, [weight_array, edge_id_array, edge_type_array, edge_start_time_array])
            if t.numel() != sizes[dist.get_rank()]:
                raise ValueError("MGGraph: every per-edge array of a rank must have as many entries as its src_array")
            out = torch.empty(sum(sizes), dtype=t.dtype, device=t.device)
            off = 0
            for r, n_r in enumerate(sizes):
                part = out[off:off + n_r]
                if r == dist.get_rank():
                    part.copy_(t)
                if n_r:
                    dist.broadcast(part, src=r)
                off += n_r
            return out

        nv = kwargs.pop("num_vertices", None)
        if nv is None and vertices_array is not None:
            v = gather(vertices_array, torch.int64, per_edge=False)
            nv = int(v.max()) + 1 if v.numel() else 0
        super().__init__(resource_handle, graph_properties, gather(src_array, torch.int64), gather(dst_array, torch.int64),
                         weight_array=gather(weight_array), edge_id_array=gather(edge_id_array, torch.int64),
                         edge_type_array=gather(edge_type_array, torch.int32), num_vertices=nv,
                         edge_start_time_array=gather(edge_start_time_array, torch.int64), **kwargs)


def _mg_check_replicated_fit(self, total_edges: int, device, optional_arrays):
    """The CSR of an MGGraph is REPLICATED on every GPU (limits): refuse up front, with the numbers, a
    graph whose replica cannot fit, instead of failing with an opaque CUDA out-of-memory half way through the build."""
    per_edge = 16 + sum(8 for a in optional_arrays if a is not None)  # src + dst (+ weight / id / type / time, <= 8 B each)
    need = int(total_edges) * (per_edge + 24)  # + sort keys, permutation and the CSR column array while building
    try:
        free, _total = torch.cuda.mem_get_info(device)
    except Exception:
        return
    if need > free:
        raise MemoryError(
            "MGGraph replicates the CSR on every GPU: %d edges need about %.1f GB per GPU to build, %.1f GB are free. "
            "Stripe the graph with WholeMemory tensors instead (pylibwholegraph.torch.MultiHopSampler samples a CHUNKED "
            "row_ptr / col_idx pair by in-kernel P2P), see INTEGRATION.md." % (total_edges, need / 1e9, free / 1e9))


MGGraph._check_replicated_fit = _mg_check_replicated_fit


def _disjoint_filter(out, num_hops: int):
    """disjoint_sampling=True on a homogeneous COO result (reference: `disjoint` of the loaders, pinned by
    tests/loader/test_neighbor_loader.py:138-187, 838-935): every seed of a label roots its own tree, a vertex belongs to the
    tree that reached it first (first edge in the hop's edge order -- the order the renumbering already uses), and an edge
    whose endpoints lie in different trees is dropped, so the trees of a mini-batch share no vertex.  The vertex set is
    unchanged (a vertex keeps the edge that discovered it); edges, label_hop_offsets and the per-node tree ids are rebuilt
    with a handful of vectorised torch ops per hop (host glue, not a hot path).  libcugraph's own choice among competing
    trees is not observable; the reference's tests check exactly the invariants above."""
    majors, minors = out["majors"], out["minors"]
    lho, rmo, base = out["label_hop_offsets"], out["renumber_map_offsets"], out["label_step_base"]
    dev = minors.device
    E, N, L = int(minors.numel()), int(out["renumber_map"].numel()), int(num_hops)
    B = int(rmo.numel()) - 1
    edge_pos = torch.arange(E, device=dev)
    group = torch.bucketize(edge_pos, lho[1:], right=True)  # (label, hop) group of every edge
    label, hop = group // L, group % L
    gmaj, gmin = majors.long() + rmo[label], minors.long() + rmo[label]  # positions in renumber_map
    node_pos = torch.arange(N, device=dev)
    node_label = torch.bucketize(node_pos, rmo[1:], right=True)
    local = node_pos - rmo[node_label]
    tree = torch.full((N,), -1, dtype=torch.int64, device=dev)
    is_seed = local < base[1].long()[node_label] if B else torch.zeros(0, dtype=torch.bool, device=dev)
    tree[is_seed] = local[is_seed]
    keep = torch.ones(E, dtype=torch.bool, device=dev)
    for h in range(L):
        sel = torch.nonzero(hop == h).reshape(-1)
        if sel.numel() == 0:
            continue
        mj, mn = gmaj[sel], gmin[sel]
        new = tree[mn] < 0
        first = torch.full((N,), E, dtype=torch.int64, device=dev)
        first.scatter_reduce_(0, mn[new], sel[new], reduce="amin")  # the edge that discovered each new vertex
        has = first < E
        tree[has] = tree[gmaj[first[has]]]
        keep[sel] = tree[mj] == tree[mn]
    counts = torch.bincount(group[keep], minlength=B * L) if E else torch.zeros(B * L, dtype=torch.int64, device=dev)
    out = dict(out)
    out["majors"], out["minors"], out["edge_id"] = majors[keep], minors[keep], out["edge_id"][keep]
    out["label_hop_offsets"] = torch.cat([torch.zeros(1, dtype=lho.dtype, device=dev), counts.cumsum(0).to(lho.dtype)])
    out["tree"] = tree  # extension: per renumber_map entry, the local id of the seed whose tree the vertex belongs to
    return out


def _disjoint_filter_hetero(out, num_edge_types: int, num_hops: int, src_vtype, dst_vtype):
    """_disjoint_filter for a heterogeneous result: vertices are addressed by their position in renumber_map (segments
    [label][vertex type]), an edge of type t joins a src_vtype[t] vertex (major) to a dst_vtype[t] vertex (minor).  Among the
    edges of one hop that reach the same new vertex, the first in OUTPUT order (label, edge type, position) defines its tree
    (the interleaving of edge types inside a hop is not recoverable from the typed result; any one choice gives vertex-disjoint
    trees).  edge_id is renumbered inside every (label, edge type) group, as the native call defines it."""
    majors, minors = out["majors"], out["minors"]
    lto, rmo, base = out["label_type_hop_offsets"], out["renumber_map_offsets"], out["label_type_step_base"]
    dev = minors.device
    T, L = int(num_edge_types), int(num_hops)
    Vt = int(base.shape[1])
    B = (int(rmo.numel()) - 1) // Vt
    E, N = int(minors.numel()), int(out["renumber_map"].numel())
    svt = torch.as_tensor(src_vtype, dtype=torch.int64, device=dev)
    dvt = torch.as_tensor(dst_vtype, dtype=torch.int64, device=dev)
    edge_pos = torch.arange(E, device=dev)
    group = torch.bucketize(edge_pos, lto[1:], right=True)  # (label * T + type) * L + hop
    label, etype, hop = group // (T * L), (group // L) % T, group % L
    gmaj = majors.long() + rmo[label * Vt + svt[etype]]
    gmin = minors.long() + rmo[label * Vt + dvt[etype]]
    node_pos = torch.arange(N, device=dev)
    seg = torch.bucketize(node_pos, rmo[1:], right=True)  # label * Vt + vertex type
    local = node_pos - rmo[seg]
    tree = torch.full((N,), -1, dtype=torch.int64, device=dev)
    is_seed = local < base[1].long()[seg % Vt, seg // Vt] if N else torch.zeros(0, dtype=torch.bool, device=dev)
    tree[is_seed] = node_pos[is_seed]
    keep = torch.ones(E, dtype=torch.bool, device=dev)
    for h in range(L):
        sel = torch.nonzero(hop == h).reshape(-1)
        if sel.numel() == 0:
            continue
        mj, mn = gmaj[sel], gmin[sel]
        new = tree[mn] < 0
        first = torch.full((N,), E, dtype=torch.int64, device=dev)
        first.scatter_reduce_(0, mn[new], sel[new], reduce="amin")
        has = first < E
        tree[has] = tree[gmaj[first[has]]]
        keep[sel] = tree[mj] == tree[mn]
    counts = torch.bincount(group[keep], minlength=B * T * L) if E else torch.zeros(B * T * L, dtype=torch.int64, device=dev)
    new_lto = torch.cat([torch.zeros(1, dtype=lto.dtype, device=dev), counts.cumsum(0).to(lto.dtype)])
    out = dict(out)
    for k in ("majors", "minors", "edge_type", "edge_renumber_map"):
        out[k] = out[k][keep]
    kept_group = group[keep]
    type_start = new_lto[(kept_group // L) * L]  # first edge of the (label, edge type) group of every kept edge
    out["edge_id"] = (torch.arange(int(keep.sum()), device=dev) - type_start).to(out["edge_id"].dtype)
    out["label_type_hop_offsets"] = new_lto
    out["edge_renumber_map_offsets"] = new_lto[::L].contiguous()
    out["tree"] = tree
    return out


def _neighbor_sample(input_graph, start_vertex_list, starting_vertex_label_offsets, h_fan_out, biased, *,
                     with_replacement=False, do_expensive_check=False, prior_sources_behavior=None,
                     deduplicate_sources=False, return_hops=False, renumber=False, retain_seeds=False,
                     compression="COO", compress_per_hop=False, random_state=None, disjoint_sampling=False,
                     return_dict=True, return_seed_local_ids=False, **unused):
    if with_replacement:
        raise NotImplementedError("sampling with replacement is not on the B200 hot path")
    if disjoint_sampling and compression != "COO":
        raise NotImplementedError("disjoint sampling returns COO")
    if compress_per_hop:
        raise NotImplementedError("compress_per_hop=True is not supported")
    if not renumber:
        raise NotImplementedError("the fused sampler always renumbers (cugraph-pyg calls with renumber=True)")
    if prior_sources_behavior not in (None, "exclude") or (prior_sources_behavior is None and deduplicate_sources is False and len(h_fan_out) > 1):
        # the fused sampler implements exactly deduplicate_sources=True + prior_sources_behavior="exclude"
        raise NotImplementedError("only deduplicate_sources=True with prior_sources_behavior='exclude' is supported")
    if biased and input_graph.weight is None:
        raise ValueError("biased sampling needs a graph with edge weights")
    seeds = _as_cuda(start_vertex_list)
    if seeds.dtype not in (torch.int32, torch.int64):
        seeds = seeds.long()
    if starting_vertex_label_offsets is None:
        offsets = torch.tensor([0, seeds.numel()], dtype=torch.int64, device=seeds.device)
    else:
        offsets = _as_cuda(starting_vertex_label_offsets, torch.int64)
    fanout = [int(f) for f in np.asarray(h_fan_out).reshape(-1)]
    if random_state is None:
        random_state = int(np.random.randint(0, 2**62))
    row_ptr, col, weight, edge_id = input_graph.row_ptr, input_graph.col, None, input_graph.edge_id
    if biased:
        # one CSR (edge types, if any, are ignored by the homogeneous entry point), zero-bias edges removed
        row_ptr, col, weight, edge_id, _ = input_graph._drop_zero_weight_cached()
    pend = input_graph._get_sampler().sample_async(row_ptr, col, seeds, offsets, fanout, int(random_state), csr_weight=weight,
                                                   csr_edge_id=edge_id, compression=compression, int64_ids=True)
    pend.want_seed_local_ids = bool(return_seed_local_ids)
    res = pend.result()
    out = {
        "majors": res.get("majors"),
        "minors": res["minors"],
        "major_offsets": res.get("major_offsets"),
        "edge_id": res["edge_id"],
        "edge_type": None,
        "weight": None,
        "hop_id": None,
        "renumber_map": res["renumber_map"],
        "renumber_map_offsets": res["renumber_map_offsets"],
        "label_hop_offsets": res["label_hop_offsets"],
        # extension: [L+1, B] first local id of the vertices each label discovered at step t (0 = seeds)
        "label_step_base": res["label_step_base"],
    }
    if return_seed_local_ids:
        out["seed_local_ids"] = res["seed_local_ids"]  # extension: local id of every input seed (link-prediction loaders)
    if disjoint_sampling:
        out = _disjoint_filter(out, len(fanout))
    return out


def homogeneous_uniform_neighbor_sample(resource_handle, input_graph, start_vertex_list, starting_vertex_label_offsets,
                                        h_fan_out, **kwargs):
    return _neighbor_sample(input_graph, start_vertex_list, starting_vertex_label_offsets, h_fan_out, False, **kwargs)


def homogeneous_biased_neighbor_sample(resource_handle, input_graph, start_vertex_list, starting_vertex_label_offsets,
                                       h_fan_out, **kwargs):
    return _neighbor_sample(input_graph, start_vertex_list, starting_vertex_label_offsets, h_fan_out, True, **kwargs)


def _edge_type_endpoints(typed, vto):
    """(vertex type of the rows, vertex type of the columns) of every per-type CSR, read off its first edge."""
    bounds = torch.as_tensor(vto[1:-1], dtype=torch.int64)
    src, dst = [], []
    for g in typed:
        row_ptr, col = g[0], g[1]
        if col.numel() == 0:
            src.append(0)
            dst.append(0)
            continue
        first_row = int(torch.nonzero(row_ptr[1:] > row_ptr[:-1])[0])
        src.append(int(torch.bucketize(torch.tensor([first_row]), bounds, right=True)))
        dst.append(int(torch.bucketize(col[:1].long().cpu(), bounds, right=True)))
    return src, dst


def _hetero_neighbor_sample(input_graph, start_vertex_list, starting_vertex_label_offsets, h_fan_out, num_edge_types,
                            vertex_type_offsets, biased, *, with_replacement=False, do_expensive_check=False,
                            prior_sources_behavior=None, deduplicate_sources=False, return_hops=False, renumber=False,
                            retain_seeds=False, compression="COO", compress_per_hop=False, random_state=None,
                            disjoint_sampling=False, return_dict=True, return_seed_local_ids=False, **unused):
    if with_replacement:
        raise NotImplementedError("sampling with replacement is not on the B200 hot path")
    if compress_per_hop or compression != "COO":
        raise NotImplementedError("heterogeneous sampling returns COO (as the reference's reader requires)")
    if not renumber:
        raise NotImplementedError("the fused sampler always renumbers (cugraph-pyg calls with renumber=True)")
    if prior_sources_behavior not in (None, "exclude") or (prior_sources_behavior is None and deduplicate_sources is False):
        raise NotImplementedError("only deduplicate_sources=True with prior_sources_behavior='exclude' is supported")
    if biased and input_graph.weight is None:
        raise ValueError("biased sampling needs a graph with edge weights")
    seeds = _as_cuda(start_vertex_list)
    if seeds.dtype not in (torch.int32, torch.int64):
        seeds = seeds.long()
    if starting_vertex_label_offsets is None:
        offsets = torch.tensor([0, seeds.numel()], dtype=torch.int64, device=seeds.device)
    else:
        offsets = _as_cuda(starting_vertex_label_offsets, torch.int64)
    fanout = [int(f) for f in np.asarray(h_fan_out).reshape(-1)]
    T = int(num_edge_types)
    if T < 1 or len(fanout) % T != 0:
        raise ValueError(f"Illegal fanout for {T} edge types.")
    vto = [int(v) for v in torch.as_tensor(vertex_type_offsets).reshape(-1).tolist()]
    if random_state is None:
        random_state = int(np.random.randint(0, 2**62))
    typed = input_graph._biased_csrs(T) if biased else input_graph._typed_csrs(T)
    pend = input_graph._get_sampler().sample_hetero_async(
        [g[0] for g in typed], [g[1] for g in typed], vto, seeds, offsets, fanout, int(random_state),
        csr_weights=[g[2] for g in typed] if biased else None,
        csr_edge_ids=[g[3] for g in typed] if any(g[3] is not None for g in typed) else None, int64_ids=True,
    )
    pend.want_seed_local_ids = bool(return_seed_local_ids)
    res = pend.result()
    if disjoint_sampling:
        key = ("endpoints", T, bool(biased))
        if key not in input_graph._typed:  # host syncs: once per graph, not per call group
            input_graph._typed[key] = _edge_type_endpoints(typed, vto)
        res = _disjoint_filter_hetero(res, T, len(fanout) // T, *input_graph._typed[key])
    return {
        **({"seed_local_ids": res["seed_local_ids"]} if return_seed_local_ids else {}),
        "majors": res["majors"],
        "minors": res["minors"],
        "major_offsets": None,
        "edge_id": res["edge_id"],
        "edge_type": res["edge_type"],
        "weight": None,
        "hop_id": None,
        "renumber_map": res["renumber_map"],
        "renumber_map_offsets": res["renumber_map_offsets"],
        "label_type_hop_offsets": res["label_type_hop_offsets"],
        "edge_renumber_map": res["edge_renumber_map"],
        "edge_renumber_map_offsets": res["edge_renumber_map_offsets"],
        # extension: [L+1, Vt, B] first local id of the type-vt vertices each label discovered at step s
        "label_type_step_base": res["label_type_step_base"],
    }


def heterogeneous_uniform_neighbor_sample(resource_handle, input_graph, start_vertex_list, starting_vertex_label_offsets,
                                          vertex_type_offsets=None, h_fan_out=None, num_edge_types=1, **kwargs):
    return _hetero_neighbor_sample(input_graph, start_vertex_list, starting_vertex_label_offsets, h_fan_out, num_edge_types,
                                   vertex_type_offsets, False, **kwargs)


def heterogeneous_biased_neighbor_sample(resource_handle, input_graph, start_vertex_list, starting_vertex_label_offsets,
                                         vertex_type_offsets=None, h_fan_out=None, num_edge_types=1, **kwargs):
    return _hetero_neighbor_sample(input_graph, start_vertex_list, starting_vertex_label_offsets, h_fan_out, num_edge_types,
                                   vertex_type_offsets, True, **kwargs)


def _temporal_neighbor_sample(input_graph, start_vertex_list, starting_vertex_label_offsets, h_fan_out, *, heterogeneous, biased=False,
                              num_edge_types=1, vertex_type_offsets=None, starting_vertex_times=None,
                              temporal_property_name=None, temporal_sampling_comparison="strictly_increasing",
                              with_replacement=False, do_expensive_check=False, prior_sources_behavior=None,
                              deduplicate_sources=False, return_hops=False, renumber=False, retain_seeds=False,
                              compression="COO", compress_per_hop=False, random_state=None, disjoint_sampling=False,
                              return_dict=True, return_seed_local_ids=False, **unused):
    """pylibcugraph.{homogeneous,heterogeneous}_{uniform,biased}_temporal_neighbor_sample (reference call site:
    python/cugraph-pyg/cugraph_pyg/sampler/distributed_sampler.py:56-80, 808-810, 897-900).  Edge times are the graph's
    edge_start_time_array; without starting_vertex_times the first hop is unconstrained."""
    if with_replacement:
        raise NotImplementedError("sampling with replacement is not on the B200 hot path")
    if disjoint_sampling and compression != "COO":
        raise NotImplementedError("disjoint sampling returns COO")
    if compress_per_hop or (heterogeneous and compression != "COO"):
        raise NotImplementedError("compress_per_hop / heterogeneous CSR output are not supported")
    if not renumber:
        raise NotImplementedError("the fused sampler always renumbers (cugraph-pyg calls with renumber=True)")
    if prior_sources_behavior not in (None, "exclude") or (prior_sources_behavior is None and deduplicate_sources is False):
        raise NotImplementedError("only deduplicate_sources=True with prior_sources_behavior='exclude' is supported")
    if input_graph.edge_time is None:
        raise ValueError("temporal sampling needs a graph built with edge_start_time_array")
    if biased and input_graph.weight is None:
        raise ValueError("biased sampling needs a graph with edge weights")
    from pylibwholegraph.torch.multihop import TIME_COMPARISONS

    if temporal_sampling_comparison not in TIME_COMPARISONS:
        raise ValueError("temporal_sampling_comparison must be one of %s" % sorted(TIME_COMPARISONS))
    seeds = _as_cuda(start_vertex_list)
    if seeds.dtype not in (torch.int32, torch.int64):
        seeds = seeds.long()
    if starting_vertex_label_offsets is None:
        offsets = torch.tensor([0, seeds.numel()], dtype=torch.int64, device=seeds.device)
    else:
        offsets = _as_cuda(starting_vertex_label_offsets, torch.int64)
    if starting_vertex_times is None:
        info = torch.iinfo(torch.int64)
        open_end = info.min if "increasing" in temporal_sampling_comparison else info.max
        times = torch.full((seeds.numel(),), open_end, dtype=torch.int64, device=seeds.device)
    else:
        times = _as_cuda(starting_vertex_times, torch.int64)
        if times.numel() != seeds.numel():
            raise ValueError("starting_vertex_times must have one entry per start vertex")
    fanout = [int(f) for f in np.asarray(h_fan_out).reshape(-1)]
    T = int(num_edge_types) if heterogeneous else 1
    if T < 1 or len(fanout) % T != 0:
        raise ValueError(f"Illegal fanout for {T} edge types.")
    if random_state is None:
        random_state = int(np.random.randint(0, 2**62))
    if heterogeneous:
        typed = input_graph._biased_csrs(T) if biased else input_graph._typed_csrs(T)
    elif biased:
        typed = [input_graph._drop_zero_weight_cached()]  # zero-bias edges are never sampled (see _drop_zero_weight)
    else:
        typed = [(input_graph.row_ptr, input_graph.col, None, input_graph.edge_id, input_graph.edge_time)]
    vto = None
    if heterogeneous:
        vto = [int(v) for v in torch.as_tensor(vertex_type_offsets).reshape(-1).tolist()]
    pend = input_graph._get_sampler().sample_temporal_async(
        [g[0] for g in typed], [g[1] for g in typed], [g[4] for g in typed], seeds, times, offsets, fanout, int(random_state),
        temporal_sampling_comparison, vertex_type_offsets=vto,
        csr_edge_ids=[g[3] for g in typed] if all(g[3] is not None for g in typed) else None,
        csr_weights=[g[2] for g in typed] if biased else None, compression=compression, int64_ids=True)
    pend.want_seed_local_ids = bool(return_seed_local_ids)
    res = pend.result()
    extra = {"seed_local_ids": res["seed_local_ids"]} if return_seed_local_ids else {}
    if heterogeneous:
        if disjoint_sampling:
            key = ("endpoints", T, bool(biased))
            if key not in input_graph._typed:
                input_graph._typed[key] = _edge_type_endpoints(typed, vto)
            res = _disjoint_filter_hetero(res, T, len(fanout) // T, *input_graph._typed[key])
        return {
            **extra,
            "majors": res["majors"], "minors": res["minors"], "major_offsets": None, "edge_id": res["edge_id"],
            "edge_type": res["edge_type"], "weight": None, "hop_id": None, "renumber_map": res["renumber_map"],
            "renumber_map_offsets": res["renumber_map_offsets"], "label_type_hop_offsets": res["label_type_hop_offsets"],
            "edge_renumber_map": res["edge_renumber_map"], "edge_renumber_map_offsets": res["edge_renumber_map_offsets"],
            "label_type_step_base": res["label_type_step_base"],
        }
    out = {
        **extra,
        "majors": res.get("majors"), "minors": res["minors"], "major_offsets": res.get("major_offsets"),
        "edge_id": res["edge_id"], "edge_type": None, "weight": None, "hop_id": None, "renumber_map": res["renumber_map"],
        "renumber_map_offsets": res["renumber_map_offsets"], "label_hop_offsets": res["label_hop_offsets"],
        "label_step_base": res["label_step_base"],
    }
    return _disjoint_filter(out, len(fanout)) if disjoint_sampling else out


def homogeneous_uniform_temporal_neighbor_sample(resource_handle, input_graph, start_vertex_list,
                                                 starting_vertex_label_offsets, h_fan_out, **kwargs):
    return _temporal_neighbor_sample(input_graph, start_vertex_list, starting_vertex_label_offsets, h_fan_out,
                                     heterogeneous=False, **kwargs)


def heterogeneous_uniform_temporal_neighbor_sample(resource_handle, input_graph, start_vertex_list,
                                                   starting_vertex_label_offsets, vertex_type_offsets=None, h_fan_out=None,
                                                   num_edge_types=1, **kwargs):
    return _temporal_neighbor_sample(input_graph, start_vertex_list, starting_vertex_label_offsets, h_fan_out,
                                     heterogeneous=True, num_edge_types=num_edge_types, vertex_type_offsets=vertex_type_offsets,
                                     **kwargs)


def homogeneous_biased_temporal_neighbor_sample(resource_handle, input_graph, start_vertex_list,
                                                starting_vertex_label_offsets, h_fan_out, **kwargs):
    return _temporal_neighbor_sample(input_graph, start_vertex_list, starting_vertex_label_offsets, h_fan_out,
                                     heterogeneous=False, biased=True, **kwargs)


def heterogeneous_biased_temporal_neighbor_sample(resource_handle, input_graph, start_vertex_list,
                                                  starting_vertex_label_offsets, vertex_type_offsets=None, h_fan_out=None,
                                                  num_edge_types=1, **kwargs):
    return _temporal_neighbor_sample(input_graph, start_vertex_list, starting_vertex_label_offsets, h_fan_out,
                                     heterogeneous=True, biased=True, num_edge_types=num_edge_types,
                                     vertex_type_offsets=vertex_type_offsets, **kwargs)


def negative_sampling(resource_handle, graph, num_samples, random_state=None, vertices=None, src_bias=None, dst_bias=None,
                      remove_duplicates=False, remove_false_negatives=False, exact_number_of_samples=False,
                      do_expensive_check=False):
    """pylibcugraph.negative_sampling: `num_samples` random (source, destination) pairs, endpoints drawn independently
    from `vertices` (default: every vertex) in proportion to src_bias / dst_bias (default: uniform).
    remove_false_negatives drops pairs that are edges of `graph`, remove_duplicates repeated pairs; with
    exact_number_of_samples the draw is repeated (a bounded number of times) until num_samples pairs remain.
    Reference call site: python/cugraph-pyg/cugraph_pyg/sampler/sampler_utils.py:66-92."""
    dev = graph.col.device
    n = int(num_samples)
    gen = None
    if random_state is not None:
        gen = torch.Generator(device=dev).manual_seed(int(random_state) & 0x7FFFFFFFFFFFFFFF)
    verts = None if vertices is None else _as_cuda(vertices, torch.int64)
    count = graph.num_vertices if verts is None else int(verts.numel())

    def draw(bias, k):
        if bias is None:
            ix = torch.randint(0, count, (k,), device=dev, generator=gen)
        else:
            b = _as_cuda(bias).double()
            if b.numel() != count:
                raise ValueError("bias arrays must have one entry per candidate vertex")
            cdf = torch.cumsum(b, 0)
            u = torch.rand(k, device=dev, dtype=torch.float64, generator=gen) * cdf[-1]
            ix = torch.searchsorted(cdf, u, right=True).clamp_(max=count - 1)
        return ix if verts is None else verts[ix]

    edge_keys = None
    src = torch.empty(0, dtype=torch.int64, device=dev)
    dst = torch.empty(0, dtype=torch.int64, device=dev)
    for attempt in range(8):
        need = n - int(src.numel())
        if need <= 0:
            break
        s, d = draw(src_bias, need), draw(dst_bias, need)
        if remove_false_negatives:
            if edge_keys is None:
                rows = torch.repeat_interleave(torch.arange(graph.num_vertices, device=dev), graph.row_ptr[1:] - graph.row_ptr[:-1])
                edge_keys = torch.sort(rows * graph.num_vertices + graph.col.long()).values
            key = s * graph.num_vertices + d
            pos = torch.searchsorted(edge_keys, key).clamp_(max=max(int(edge_keys.numel()) - 1, 0))
            ok = edge_keys[pos] != key if edge_keys.numel() else torch.ones_like(key, dtype=torch.bool)
            s, d = s[ok], d[ok]
        src, dst = torch.cat([src, s]), torch.cat([dst, d])
        if remove_duplicates:
            key = torch.unique(src * graph.num_vertices + dst)
            src, dst = key // graph.num_vertices, key % graph.num_vertices
        if not exact_number_of_samples:
            break
    return {"sources": src[:n], "destinations": dst[:n]}
