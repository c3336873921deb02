"""pylibcugraph-shaped entry points (cugraph-gnn_b200/pylibcugraph) on the CPU: the Python code is the product's, tensors stay
on the CPU, and the native sampler underneath is the CPU emulation of csrc/multihop.cu (tests/emu) instead of the CUDA
library.  Checks the argument plumbing (CSR construction from COO, per-type CSRs, zero-weight removal, edge times, default
seed times, result dictionaries) of the plain and the temporal functions, and replays the reference's deterministic
temporal expectations (python/cugraph-pyg/cugraph_pyg/tests/loader/test_neighbor_loader.py:943-1058) at this level, for
the uniform and the biased variant (the reference parametrises both).  Test infrastructure only.
"""
import numpy as np
import pytest

torch = pytest.importorskip("torch")

from test_emulated_multihop_cpu import FLAG_CSR, FLAG_INT64, _run, emu  # noqa: E402,F401  (emu: fixture)

CMP = {"strictly_increasing": 0, "monotonically_increasing": 1, "strictly_decreasing": 2, "monotonically_decreasing": 3}


class _Pending:
    """The emulated call runs when result() is asked for (the callers set want_seed_local_ids in between)."""

    def __init__(self, run):
        self._run = run
        self.want_seed_local_ids = False

    def result(self):
        return self._run(self.want_seed_local_ids)


class EmulatedSampler:
    """Stands in for pylibwholegraph.torch.MultiHopSampler: same methods, CPU tensors in and out."""

    def __init__(self, lib):
        self.lib = lib
        self.calls = []

    @staticmethod
    def _np(t):
        return None if t is None else t.detach().cpu().numpy()

    def _pending(self, call, hops, hetero, vt=1):
        def run(want_seed_ids):
            out = call(want_seed_ids)
            res = {k: torch.from_numpy(v) for k, v in out.items()}
            if hetero:
                res["label_type_step_base"] = res["label_type_step_base"].view(hops + 1, vt, -1)
            else:
                res["label_step_base"] = res["label_step_base"].view(hops + 1, -1)
            return res

        return _Pending(run)

    def sample_async(self, csr_row_ptr, csr_col, seeds, label_offsets, fanout, random_state, *, csr_weight=None, csr_edge_id=None,
                     compression="COO", int64_ids=False):
        self.calls.append("plain")
        V = csr_row_ptr.numel() - 1
        flags = (FLAG_CSR if compression == "CSR" else 0) | (FLAG_INT64 if int64_ids else 0)
        return self._pending(lambda want: _run(
            self.lib, [self._np(csr_row_ptr)], [self._np(csr_col)], [0, V], self._np(seeds), self._np(label_offsets), fanout, random_state,
            hetero=False, eids=None if csr_edge_id is None else [self._np(csr_edge_id)],
            weights=None if csr_weight is None else [self._np(csr_weight)], flags=flags, seed_local_ids=want), len(fanout), False)

    def sample_hetero_async(self, csr_row_ptrs, csr_cols, vertex_type_offsets, seeds, label_offsets, fanout, random_state, *, csr_weights=None,
                            csr_edge_ids=None, int64_ids=False):
        self.calls.append("hetero")
        T = len(csr_row_ptrs)
        return self._pending(lambda want: _run(
            self.lib, [self._np(r) for r in csr_row_ptrs], [self._np(c) for c in csr_cols], vertex_type_offsets, self._np(seeds),
            self._np(label_offsets), fanout, random_state, eids=None if csr_edge_ids is None else [self._np(e) for e in csr_edge_ids],
            weights=None if csr_weights is None else [self._np(w) for w in csr_weights], flags=FLAG_INT64 if int64_ids else 0,
            seed_local_ids=want), len(fanout) // T, True, len(vertex_type_offsets) - 1)

    def sample_temporal_async(self, csr_row_ptrs, csr_cols, csr_edge_times, seeds, seed_times, label_offsets, fanout, random_state, comparison, *,
                              vertex_type_offsets=None, csr_edge_ids=None, csr_weights=None, compression="COO", int64_ids=False):
        if comparison not in CMP:  # as MultiHopSampler.sample_temporal_async
            raise ValueError("temporal comparison must be one of %s, got %r" % (sorted(CMP), comparison))
        self.calls.append("temporal")
        T = len(csr_row_ptrs)
        hetero = vertex_type_offsets is not None
        V = csr_row_ptrs[0].numel() - 1
        flags = (FLAG_CSR if compression == "CSR" else 0) | (FLAG_INT64 if int64_ids else 0)
        return self._pending(lambda want: _run(
            self.lib, [self._np(r) for r in csr_row_ptrs], [self._np(c) for c in csr_cols], vertex_type_offsets if hetero else [0, V],
            self._np(seeds), self._np(label_offsets), fanout, random_state, hetero=hetero, times=[self._np(t) for t in csr_edge_times],
            seed_times=self._np(seed_times), cmp=CMP[comparison], eids=None if csr_edge_ids is None else [self._np(e) for e in csr_edge_ids],
            weights=None if csr_weights is None else [self._np(w) for w in csr_weights], flags=flags, seed_local_ids=want),
            len(fanout) // T, hetero, (len(vertex_type_offsets) - 1) if hetero else 1)


@pytest.fixture()
def plc(monkeypatch, emu):  # noqa: F811
    import pylibcugraph

    def as_cpu(a, dtype=None):
        if a is None:
            return None
        t = torch.as_tensor(a)
        if dtype is not None and t.dtype != dtype:
            t = t.to(dtype)
        return t.contiguous()

    monkeypatch.setattr(pylibcugraph, "_as_cuda", as_cpu)
    sampler = EmulatedSampler(emu)
    monkeypatch.setattr(pylibcugraph.SGGraph, "_get_sampler", lambda self: sampler)
    return pylibcugraph, sampler


def _graph(plc_mod, src, dst, *, etype=None, weight=None, time=None, num_vertices=None):
    n = len(src)
    return plc_mod.SGGraph(plc_mod.ResourceHandle(), plc_mod.GraphProperties(is_multigraph=True), torch.tensor(src), torch.tensor(dst),
                           weight_array=None if weight is None else torch.tensor(weight, dtype=torch.float32),
                           edge_id_array=torch.arange(n) if etype is None else _per_type_ids(etype),
                           edge_type_array=None if etype is None else torch.tensor(etype, dtype=torch.int32),
                           edge_start_time_array=None if time is None else torch.tensor(time), num_vertices=num_vertices)


def _per_type_ids(etype):
    """running index per edge type, as GraphStore assigns edge ids (graph_store.py:578-607)"""
    etype = np.asarray(etype)
    ids = np.zeros(len(etype), dtype=np.int64)
    for t in np.unique(etype):
        ids[etype == t] = np.arange(int((etype == t).sum()))
    return torch.from_numpy(ids)


KW = dict(renumber=True, return_hops=True, prior_sources_behavior="exclude", deduplicate_sources=True, retain_seeds=True, random_state=62)


@pytest.mark.parametrize("biased", [False, True])
def test_temporal_pin_homogeneous(plc, biased):
    """test_neighbor_loader_temporal_simple: path 3 -> 2 -> 1 -> 0 along strictly increasing edge times."""
    mod, sampler = plc
    src_cite, dst_cite, tme = [3, 2, 1, 2], [2, 1, 0, 0], [0, 1, 2, 0]
    # GraphStore hands cuGraph src = PyG edge_index[1], dst = PyG edge_index[0]; the test puts [dst_cite, src_cite] as edge_index
    g = _graph(mod, src_cite, dst_cite, weight=[1.0] * 4 if biased else None, time=tme, num_vertices=4)
    fn = mod.homogeneous_biased_temporal_neighbor_sample if biased else mod.homogeneous_uniform_temporal_neighbor_sample
    out = fn(mod.ResourceHandle(), g, torch.tensor([3]), torch.tensor([0, 1]), np.array([2, 2, 2], dtype=np.int32),
             starting_vertex_times=torch.tensor([-1]), temporal_property_name="time", temporal_sampling_comparison="strictly_increasing", **KW)
    assert sampler.calls == ["temporal"]
    assert out["renumber_map"].tolist() == [3, 2, 1, 0]
    assert out["edge_id"].tolist() == [0, 1, 2]
    assert np.diff(out["label_hop_offsets"].numpy()).tolist() == [1, 1, 1]
    assert out["majors"].tolist() == [0, 1, 2] and out["minors"].tolist() == [1, 2, 3]
    # CSR compression (what NeighborLoader asks for on a homogeneous graph)
    out = fn(mod.ResourceHandle(), g, torch.tensor([3]), torch.tensor([0, 1]), np.array([2, 2, 2], dtype=np.int32), compression="CSR",
             starting_vertex_times=torch.tensor([-1]), temporal_sampling_comparison="strictly_increasing", **KW)
    assert out["majors"] is None and out["major_offsets"].tolist() == [0, 1, 2, 3] and out["minors"].tolist() == [1, 2, 3]


@pytest.mark.parametrize("biased", [False, True])
def test_temporal_pin_heterogeneous(plc, biased):
    """test_neighbor_loader_temporal_hetero.  Vertex types sorted: author (offset 0), paper (offset 3); edge types sorted:
    (author, writes, paper) = 0, (paper, cites, paper) = 1."""
    mod, sampler = plc
    src_cite, dst_cite, tme_cite = [3, 2, 1, 2], [2, 1, 0, 0], [0, 1, 2, 0]
    src_author, dst_author, tme_author = [3, 2, 2, 1, 3, 2, 0], [0, 0, 1, 1, 2, 2, 2], [0, 0, 1, 0, 2, 1, 1]
    # edge_index = [dst_*, src_*]: PyG sources are row 0.  cuGraph src = PyG destination (+ its type offset), dst = PyG source
    src = [p + 3 for p in src_author] + [p + 3 for p in src_cite]
    dst = [a + 0 for a in dst_author] + [p + 3 for p in dst_cite]
    etype = [0] * 7 + [1] * 4
    g = _graph(mod, src, dst, etype=etype, weight=[1.0] * 11 if biased else None, time=tme_author + tme_cite, num_vertices=7)
    fn = mod.heterogeneous_biased_temporal_neighbor_sample if biased else mod.heterogeneous_uniform_temporal_neighbor_sample
    fanout = np.array([2, 2, 2, 2, 0, 2], dtype=np.int32)  # [hop * T + etype]: writes [2, 2, 0], cites [2, 2, 2]
    out = fn(mod.ResourceHandle(), g, torch.tensor([3 + 3]), torch.tensor([0, 1]), vertex_type_offsets=torch.tensor([0, 3, 7]), h_fan_out=fanout,
             num_edge_types=2, starting_vertex_times=torch.tensor([-1]), temporal_sampling_comparison="strictly_increasing", **KW)
    rmo, lto, ermo = out["renumber_map_offsets"].numpy(), out["label_type_hop_offsets"].numpy(), out["edge_renumber_map_offsets"].numpy()
    rm = out["renumber_map"].numpy()
    assert sorted(rm[rmo[0]:rmo[1]].tolist()) == [0, 1, 2] and (rm[rmo[1]:rmo[2]] - 3).tolist() == [3, 2, 1, 0]
    assert sorted(out["edge_renumber_map"].numpy()[ermo[0]:ermo[1]].tolist()) == [0, 2, 4, 5]
    assert np.diff(lto[0:4]).tolist() == [2, 2, 0]


def test_default_seed_times_leave_the_first_hop_open(plc):
    mod, _ = plc
    rng = np.random.default_rng(0)
    src, dst = rng.integers(0, 60, 900), rng.integers(0, 60, 900)
    tme = rng.integers(0, 100, 900)
    g = _graph(mod, src.tolist(), dst.tolist(), time=tme.tolist(), num_vertices=60)
    seeds = torch.arange(20)
    for comparison in CMP:
        out = mod.homogeneous_uniform_temporal_neighbor_sample(mod.ResourceHandle(), g, seeds, torch.tensor([0, 20]), np.array([-1], dtype=np.int32),
                                                               temporal_sampling_comparison=comparison, **KW)
        deg = np.bincount(src, minlength=60)[:20].sum()
        assert out["minors"].numel() == deg, comparison  # no starting times: every edge of every seed is eligible


def test_plain_paths_still_route_and_drop_zero_weights(plc, oracle):
    """The GPU-verified plain functions through the same stand-in: guards the shared helpers the temporal work touched
    (_typed_csrs / _drop_zero_weight now carry edge times)."""
    mod, sampler = plc
    rng = np.random.default_rng(3)
    src, dst = rng.integers(0, 80, 2000), rng.integers(0, 80, 2000)
    etype = rng.integers(0, 2, 2000)
    w = rng.random(2000).astype(np.float32)
    w[rng.random(2000) < 0.3] = 0.0
    g = _graph(mod, src.tolist(), dst.tolist(), etype=etype.tolist(), weight=w.tolist(), time=rng.integers(0, 9, 2000).tolist(), num_vertices=80)
    seeds, lo = torch.arange(40), torch.tensor([0, 20, 40])
    out = mod.homogeneous_biased_neighbor_sample(mod.ResourceHandle(), g, seeds, lo, np.array([3, 2], dtype=np.int32), **KW)
    ids = _per_type_ids(etype).numpy()
    assert out["minors"].numel() > 0
    # every sampled edge has a positive weight: recover the COO position from (source vertex, edge id, destination)
    rm, rmo, lho = out["renumber_map"].numpy(), out["renumber_map_offsets"].numpy(), out["label_hop_offsets"].numpy()
    for l in range(2):
        a, b = lho[l * 2], lho[l * 2 + 2]
        s_g = rm[rmo[l]:rmo[l + 1]][out["majors"].numpy()[a:b]]
        d_g = rm[rmo[l]:rmo[l + 1]][out["minors"].numpy()[a:b]]
        for s, d, e in zip(s_g, d_g, out["edge_id"].numpy()[a:b]):
            cand = np.flatnonzero((src == s) & (dst == d) & (ids == e))
            assert len(cand) >= 1 and (w[cand] > 0).any()
    out = mod.heterogeneous_biased_neighbor_sample(mod.ResourceHandle(), g, seeds, lo, vertex_type_offsets=torch.tensor([0, 80]),
                                                   h_fan_out=np.array([3, 3, 2, 2], dtype=np.int32), num_edge_types=2, **KW)
    assert out["minors"].numel() > 0 and set(out["edge_type"].tolist()) == {0, 1}
    out = mod.homogeneous_uniform_neighbor_sample(mod.ResourceHandle(), g, seeds, lo, np.array([3, 2], dtype=np.int32), **KW)
    exp = oracle.multihop_sample(g.row_ptr.numpy(), g.col.numpy(), seeds.numpy(), lo.numpy(), [3, 2], 62, edge_ids=g.edge_id.numpy())
    for k in ("majors", "minors", "edge_id", "renumber_map", "renumber_map_offsets", "label_hop_offsets"):
        assert np.array_equal(out[k].numpy(), exp[k]), k
    assert sampler.calls == ["plain", "hetero", "plain"]
