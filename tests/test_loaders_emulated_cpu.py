"""The cugraph_pyg mirror's loader stack (GraphStore -> pylibcugraph graph -> DistributedNeighborSampler -> readers ->
SampleIterator -> NeighborLoader) on the CPU: every line of Python is the product's; underneath, "cuda" tensors are CPU
tensors (torch factory functions / Tensor.cuda / Tensor.to are patched for the duration of a test), the feature store is a
dictionary, and the native sampler is the CPU emulation of csrc/multihop.cu (tests/emu).

Purpose: the temporal path (time_attr / input_time / temporal_comparison) reached the loaders after the round's GPU minutes
were spent; this replays the reference's loader-level temporal tests (tests/loader/test_neighbor_loader.py:943-1058,
uniform and biased) and guards the non-temporal loader paths the change touched.  Test infrastructure only.
"""
import numpy as np
import pytest

torch = pytest.importorskip("torch")

from test_emulated_multihop_cpu import emu  # noqa: E402,F401  (fixture)
from test_pylibcugraph_emulated_cpu import EmulatedSampler  # noqa: E402


def _cpu_device(d):
    if d is None:
        return None
    return "cpu" if str(d).startswith("cuda") else d


@pytest.fixture()
def stack(monkeypatch, emu):  # noqa: F811
    """Patches torch so that device='cuda' means the CPU, and pylibcugraph so that graphs sample through the emulator."""
    for name in ("arange", "full", "empty", "zeros", "ones", "tensor", "as_tensor", "randint", "rand", "randperm", "full_like", "empty_like"):
        real = getattr(torch, name)

        def wrapped(*a, __real=real, **k):
            if "device" in k:
                k["device"] = _cpu_device(k["device"])
            return __real(*a, **k)

        monkeypatch.setattr(torch, name, wrapped)
    real_to = torch.Tensor.to

    def to(self, *a, **k):
        a = tuple(_cpu_device(x) if isinstance(x, (str, torch.device)) else x for x in a)
        if "device" in k:
            k["device"] = _cpu_device(k["device"])
        return real_to(self, *a, **k)

    monkeypatch.setattr(torch.Tensor, "to", to)
    monkeypatch.setattr(torch.Tensor, "cuda", lambda self, *a, **k: self)
    monkeypatch.setattr(torch.Tensor, "pin_memory", lambda self, *a, **k: self)
    import pylibcugraph

    sampler = EmulatedSampler(emu)
    monkeypatch.setattr(pylibcugraph.SGGraph, "_get_sampler", lambda self: sampler)
    import cugraph_pyg
    from cugraph_pyg._pyg_compat import FeatureStoreBase, TensorAttr

    class DictFeatureStore(FeatureStoreBase):
        """FeatureStore interface over a dictionary of CPU tensors (the product's FeatureStore lives on WholeMemory)."""

        def __init__(self):
            super().__init__()
            self._t = {}

        def _put_tensor(self, tensor, attr):
            self._t[(attr.group_name, attr.attr_name)] = torch.as_tensor(tensor)
            return True

        def _get_tensor(self, attr):
            t = self._t.get((attr.group_name, attr.attr_name))
            if t is None:
                return None
            if attr.is_set("index") and attr.index is not None:
                return t[attr.index]
            return t

        def _remove_tensor(self, attr):
            return self._t.pop((attr.group_name, attr.attr_name), None) is not None

        def _get_tensor_size(self, attr):
            return tuple(self._t[(attr.group_name, attr.attr_name)].shape)

        def get_all_tensor_attrs(self):
            return [TensorAttr(group_name=g, attr_name=a) for g, a in self._t]

    return cugraph_pyg, DictFeatureStore, sampler


def _cite_graph(cugraph_pyg, FS, biased):
    src_cite, dst_cite, tme_cite = torch.tensor([3, 2, 1, 2]), torch.tensor([2, 1, 0, 0]), torch.tensor([0, 1, 2, 0])
    graph_store, feature_store = cugraph_pyg.data.GraphStore(), FS()
    graph_store[("paper", "cites", "paper"), "coo", False, (4, 4)] = [dst_cite, src_cite]
    feature_store[("paper", "cites", "paper"), "time", None] = tme_cite
    if biased:
        feature_store[("paper", "cites", "paper"), "bias", None] = torch.tensor([1.0] * 4)
    return graph_store, feature_store


@pytest.mark.parametrize("biased", [False, True])
def test_neighbor_loader_temporal_simple(stack, biased):
    """tests/loader/test_neighbor_loader.py:943-988 of the reference, verbatim expectations."""
    cugraph_pyg, FS, sampler = stack
    graph_store, feature_store = _cite_graph(cugraph_pyg, FS, biased)
    loader = cugraph_pyg.loader.NeighborLoader((feature_store, graph_store), num_neighbors=[2, 2, 2], batch_size=1,
                                               input_nodes=torch.tensor([3]), input_time=torch.tensor([-1]), time_attr="time", shuffle=False,
                                               weight_attr="bias" if biased else None, temporal_comparison="strictly_increasing",
                                               local_seeds_per_call=64)
    out = next(iter(loader))
    assert sampler.calls == ["temporal"]
    assert out.n_id.tolist() == [3, 2, 1, 0]
    assert out.e_id.tolist() == [0, 1, 2]
    assert out.num_sampled_nodes.tolist() == [1, 1, 1, 1]
    assert out.num_sampled_edges.tolist() == [1, 1, 1]


@pytest.mark.parametrize("biased", [False, True])
def test_neighbor_loader_temporal_hetero(stack, biased):
    """tests/loader/test_neighbor_loader.py:991-1056 of the reference, verbatim expectations."""
    cugraph_pyg, FS, sampler = stack
    graph_store, feature_store = _cite_graph(cugraph_pyg, FS, biased)
    src_author = torch.tensor([3, 2, 2, 1, 3, 2, 0])
    dst_author = torch.tensor([0, 0, 1, 1, 2, 2, 2])
    graph_store[("author", "writes", "paper"), "coo", False, (3, 4)] = [dst_author, src_author]
    feature_store[("author", "writes", "paper"), "time", None] = torch.tensor([0, 0, 1, 0, 2, 1, 1])
    if biased:
        feature_store[("author", "writes", "paper"), "bias", None] = torch.tensor([1.0] * 7)
    loader = cugraph_pyg.loader.NeighborLoader(
        (feature_store, graph_store),
        num_neighbors={("paper", "cites", "paper"): [2, 2, 2], ("author", "writes", "paper"): [2, 2, 0]},
        batch_size=1, input_nodes=("paper", torch.tensor([3])), input_time=torch.tensor([-1]), time_attr="time",
        weight_attr="bias" if biased else None, shuffle=False, temporal_comparison="strictly_increasing", local_seeds_per_call=64)
    out = next(iter(loader))
    assert sampler.calls == ["temporal"]
    assert sorted(out["author"].n_id.tolist()) == [0, 1, 2]
    assert out["paper"].n_id.tolist() == [3, 2, 1, 0]
    assert sorted(out["author", "writes", "paper"].e_id.tolist()) == [0, 2, 4, 5]
    assert out["author", "writes", "paper"].num_sampled_edges.tolist() == [2, 2, 0]


def test_neighbor_loader_temporal_batches_carry_their_own_times(stack):
    """Several batches and call groups: every batch is sampled with ITS seeds' times (times follow the shuffle / the
    call-group split), checked through the hop-0 edges: they respect the comparison with the seed's time."""
    cugraph_pyg, FS, sampler = stack
    rng = np.random.default_rng(0)
    n, e = 60, 1500
    src, dst = torch.from_numpy(rng.integers(0, n, e)), torch.from_numpy(rng.integers(0, n, e))
    tme = torch.from_numpy(rng.integers(0, 100, e))
    graph_store, feature_store = cugraph_pyg.data.GraphStore(), FS()
    graph_store[("n", "to", "n"), "coo", False, (n, n)] = [src, dst]
    feature_store[("n", "to", "n"), "time", None] = tme
    feature_store["n", "x", None] = torch.arange(n * 4, dtype=torch.float32).reshape(n, 4)
    seeds = torch.from_numpy(rng.permutation(n)[:40])
    seed_time = torch.from_numpy(rng.integers(20, 80, 40))
    loader = cugraph_pyg.loader.NeighborLoader((feature_store, graph_store), num_neighbors=[4, 2], batch_size=8, input_nodes=seeds,
                                               input_time=seed_time, time_attr="time", shuffle=True, temporal_comparison="monotonically_decreasing",
                                               local_seeds_per_call=16)  # 2 batches per call group, 3 call groups
    time_of = dict(zip(seeds.tolist(), seed_time.tolist()))
    batches = 0
    for out in loader:
        batches += 1
        assert out.x.shape == (out.n_id.numel(), 4) and torch.equal(out.x[:, 0], out.n_id.float() * 4)
        k = int(out.num_sampled_edges[0])
        assert k > 0
        hop0_dst_local = out.edge_index[1][:k]  # PyG: messages flow source -> destination; hop-0 destinations are the seeds
        assert int(hop0_dst_local.max()) < int(out.num_sampled_nodes[0])
        seed_ids = out.n_id[hop0_dst_local].tolist()
        edge_times = tme[out.e_id[:k]].tolist()
        assert all(t <= time_of[s] for s, t in zip(seed_ids, edge_times))
        # the sampled edges are edges of the graph: e_id indexes the put order
        assert torch.equal(src[out.e_id], out.n_id[out.edge_index[0]]) and torch.equal(dst[out.e_id], out.n_id[out.edge_index[1]])
    assert batches == 5 and sampler.calls == ["temporal"] * 3


def test_plain_loaders_unchanged(stack, oracle):
    """Non-temporal homogeneous (CSR decode) and heterogeneous loaders through the same stack: the paths the temporal work
    touched (NodeLoader's input data, call-group splitting, GraphStore's edge list) still produce valid mini-batches."""
    cugraph_pyg, FS, sampler = stack
    rng = np.random.default_rng(1)
    n, e = 80, 2500
    src, dst = torch.from_numpy(rng.integers(0, n, e)), torch.from_numpy(rng.integers(0, n, e))
    graph_store, feature_store = cugraph_pyg.data.GraphStore(), FS()
    graph_store[("n", "to", "n"), "coo", False, (n, n)] = [src, dst]
    feature_store["n", "x", None] = torch.arange(n * 2, dtype=torch.float32).reshape(n, 2)
    loader = cugraph_pyg.loader.NeighborLoader((feature_store, graph_store), num_neighbors=[5, 3], batch_size=16, input_nodes=torch.arange(40),
                                               shuffle=False, local_seeds_per_call=32)
    seen = 0
    for i, out in enumerate(loader):
        assert out.input_id.tolist() == list(range(16 * i, min(16 * i + 16, 40)))
        assert out.n_id[:out.batch_size].tolist() == out.input_id.tolist()  # seeds first
        assert torch.equal(src[out.e_id], out.n_id[out.edge_index[0]]) and torch.equal(dst[out.e_id], out.n_id[out.edge_index[1]])
        assert int(out.num_sampled_nodes.sum()) == out.n_id.numel() and int(out.num_sampled_edges.sum()) == out.e_id.numel()
        assert torch.equal(out.x[:, 0], out.n_id.float() * 2)
        seen += out.batch_size
    assert seen == 40 and sampler.calls == ["plain", "plain"]
    with pytest.raises(ValueError):
        cugraph_pyg.loader.NeighborLoader((feature_store, graph_store), num_neighbors=[5], input_nodes=torch.arange(4), input_time=torch.arange(4))


def test_call_group_feature_prefetch_equals_per_batch_fetch(stack):
    """SampleIterator fetches node features once per call group and hands every mini-batch a row slice; switched off it
    fetches per mini-batch as the reference does (sampler.py:51-76).  Every mini-batch carries the rows of its own n_id either way,
    and the per-call-group form asks the feature store once per call group."""
    cugraph_pyg, FS, sampler = stack
    from cugraph_pyg.sampler.sampler import SampleIterator

    rng = np.random.default_rng(5)
    n, e = 120, 3000
    src, dst = torch.from_numpy(rng.integers(0, n, e)), torch.from_numpy(rng.integers(0, n, e))

    class CountingFS(FS):
        fetches = 0

        def multi_get_tensor(self, attrs):
            CountingFS.fetches += 1
            return super().multi_get_tensor(attrs)

    def run(prefetch):
        graph_store, feature_store = cugraph_pyg.data.GraphStore(), CountingFS()
        graph_store[("n", "to", "n"), "coo", False, (n, n)] = [src, dst]
        feature_store["n", "x", None] = torch.arange(n * 3, dtype=torch.float32).reshape(n, 3)
        feature_store["n", "y", None] = torch.arange(n, dtype=torch.int64)
        CountingFS.fetches = 0
        old = SampleIterator.prefetch_call_group_features
        SampleIterator.prefetch_call_group_features = prefetch
        try:
            loader = cugraph_pyg.loader.NeighborLoader((feature_store, graph_store), num_neighbors=[4, 2], batch_size=8, input_nodes=torch.arange(64),
                                                       shuffle=False, local_seeds_per_call=32)
            return [(b.n_id.clone(), b.x.clone(), b.y.clone(), b.edge_index.clone()) for b in loader], CountingFS.fetches
        finally:
            SampleIterator.prefetch_call_group_features = old

    a, fetch_a = run(True)
    b, fetch_b = run(False)
    assert len(a) == len(b) == 8
    for batches in (a, b):  # (the two loaders draw different samples: every batch is checked against the closed form of its own n_id)
        for n_id, x, y, ei in batches:
            assert torch.equal(x, torch.arange(n * 3, dtype=torch.float32).reshape(n, 3)[n_id]) and torch.equal(y, n_id)
            assert x.shape[0] == n_id.numel() and int(ei.max()) < n_id.numel()
    for (n1, _, _, _), (n2, _, _, _) in zip(a, b):
        assert torch.equal(n1[:8], n2[:8])  # same seeds, seeds first
    assert fetch_a == 2 and fetch_b == 8  # 2 call groups of 4 mini-batches


@pytest.mark.parametrize("prefetch", [True, False])
def test_hetero_loader_features_per_call_group(stack, prefetch):
    """Heterogeneous mini-batches carry, per vertex type, the feature rows of their own n_id -- with the per-call-group feature
    fetch (one gather per (type, feature) and call group, row slices per mini-batch) and with the per-mini-batch fetch."""
    cugraph_pyg, FS, sampler = stack
    from cugraph_pyg.sampler.sampler import SampleIterator

    rng = np.random.default_rng(11)
    na, nb = 50, 70
    graph_store, feature_store = cugraph_pyg.data.GraphStore(), FS()
    ab = torch.stack([torch.from_numpy(rng.integers(0, na, 600)), torch.from_numpy(rng.integers(0, nb, 600))])
    ba = torch.stack([torch.from_numpy(rng.integers(0, nb, 500)), torch.from_numpy(rng.integers(0, na, 500))])
    aa = torch.stack([torch.from_numpy(rng.integers(0, na, 400)), torch.from_numpy(rng.integers(0, na, 400))])
    graph_store[("a", "ab", "b"), "coo", False, (na, nb)] = ab
    graph_store[("b", "ba", "a"), "coo", False, (nb, na)] = ba
    graph_store[("a", "aa", "a"), "coo", False, (na, na)] = aa
    xa = torch.arange(na * 2, dtype=torch.float32).reshape(na, 2)
    xb = 1000 + torch.arange(nb * 3, dtype=torch.float32).reshape(nb, 3)
    feature_store["a", "x", None] = xa
    feature_store["b", "x", None] = xb
    feature_store["a", "y", None] = torch.arange(na)
    old = SampleIterator.prefetch_call_group_features
    SampleIterator.prefetch_call_group_features = prefetch
    try:
        loader = cugraph_pyg.loader.NeighborLoader((feature_store, graph_store), num_neighbors={("a", "ab", "b"): [3, 2], ("b", "ba", "a"): [3, 2],
                                                                                             ("a", "aa", "a"): [2, 2]},
                                                   input_nodes=("a", torch.arange(40)), batch_size=8, shuffle=False, local_seeds_per_call=16)
        seen = 0
        for i, batch in enumerate(loader):
            assert batch["a"].n_id[:8].tolist() == list(range(8 * i, 8 * i + 8))  # seeds first
            assert torch.equal(batch["a"].x, xa[batch["a"].n_id]) and torch.equal(batch["a"].y, batch["a"].n_id)
            assert torch.equal(batch["b"].x, xb[batch["b"].n_id])
            for et, full in ((("a", "ab", "b"), ab), (("b", "ba", "a"), ba), (("a", "aa", "a"), aa)):
                ei, e_id = batch[et].edge_index, batch[et].e_id
                assert ei.shape[1] == e_id.numel() == int(batch[et].num_sampled_edges.sum())
                assert torch.equal(full[0][e_id], batch[et[0]].n_id[ei[0]]) and torch.equal(full[1][e_id], batch[et[2]].n_id[ei[1]])
            assert int(batch["a"].num_sampled_nodes.sum()) == batch["a"].n_id.numel()
            seen += 1
        assert seen == 5
    finally:
        SampleIterator.prefetch_call_group_features = old


@pytest.mark.parametrize("biased", [False, True])
def test_link_neighbor_loader_temporal_homogeneous(stack, biased):
    """tests/loader/test_neighbor_loader.py:1059-1101 of the reference (seed edge 3 -> 3 at time -1)."""
    cugraph_pyg, FS, sampler = stack
    graph_store, feature_store = _cite_graph(cugraph_pyg, FS, biased)
    loader = cugraph_pyg.loader.LinkNeighborLoader((feature_store, graph_store), num_neighbors=[2, 2, 2], batch_size=1,
                                                   edge_label_index=torch.tensor([[3], [3]]), edge_label_time=torch.tensor([-1]), time_attr="time",
                                                   weight_attr="bias" if biased else None, shuffle=False, temporal_comparison="strictly_increasing",
                                                   local_seeds_per_call=64)
    out = next(iter(loader))
    assert sampler.calls == ["temporal"]
    assert out.n_id.tolist() == [3, 2, 1, 0]
    assert out.edge_label_index.tolist() == [[0], [0]]


@pytest.mark.parametrize("biased", [False, True])
def test_link_neighbor_loader_temporal_heterogeneous(stack, biased):
    """tests/loader/test_neighbor_loader.py:1106-1170 of the reference (seed edge author 0 -> paper 3 at time -1)."""
    cugraph_pyg, FS, sampler = stack
    graph_store, feature_store = _cite_graph(cugraph_pyg, FS, biased)
    graph_store[("author", "writes", "paper"), "coo", False, (3, 4)] = [torch.tensor([0, 0, 1, 1, 2, 2, 2]), torch.tensor([3, 2, 2, 1, 3, 2, 0])]
    feature_store[("author", "writes", "paper"), "time", None] = torch.tensor([0, 0, 1, 0, 2, 1, 1])
    if biased:
        feature_store[("author", "writes", "paper"), "bias", None] = torch.tensor([1.0] * 7)
    loader = cugraph_pyg.loader.LinkNeighborLoader(
        (feature_store, graph_store),
        num_neighbors={("paper", "cites", "paper"): [2, 2, 2], ("author", "writes", "paper"): [2, 2, 0]},
        batch_size=1, edge_label_index=(("author", "writes", "paper"), torch.tensor([[0], [3]])), edge_label_time=torch.tensor([-1]),
        time_attr="time", weight_attr="bias" if biased else None, shuffle=False, temporal_comparison="strictly_increasing",
        local_seeds_per_call=64)
    out = next(iter(loader))
    assert sorted(out["author"].n_id.tolist()) == [0, 1, 2]
    assert out["paper"].n_id.tolist() == [3, 2, 1, 0]
    assert sorted(out["author", "writes", "paper"].e_id.tolist()) == [0, 2, 4, 5]
    assert out["author", "writes", "paper"].num_sampled_edges.tolist() == [2, 2, 0]


def test_plain_link_loader_unchanged(stack):
    """Non-temporal link prediction (GPU-verified) through the same stack: edge_label_index maps the seed edges' endpoints."""
    cugraph_pyg, FS, sampler = stack
    rng = np.random.default_rng(2)
    n, e = 70, 1800
    src, dst = torch.from_numpy(rng.integers(0, n, e)), torch.from_numpy(rng.integers(0, n, e))
    graph_store, feature_store = cugraph_pyg.data.GraphStore(), FS()
    graph_store[("n", "to", "n"), "coo", False, (n, n)] = [src, dst]
    feature_store["n", "x", None] = torch.arange(n, dtype=torch.float32).reshape(n, 1)
    eli = torch.stack([src[:24], dst[:24]])
    loader = cugraph_pyg.loader.LinkNeighborLoader((feature_store, graph_store), num_neighbors=[3, 2], batch_size=8, edge_label_index=eli,
                                                   shuffle=False, local_seeds_per_call=16)
    for i, out in enumerate(loader):
        lo = 8 * i
        assert torch.equal(out.n_id[out.edge_label_index[0]], eli[0, lo:lo + 8]) and torch.equal(out.n_id[out.edge_label_index[1]], eli[1, lo:lo + 8])
        assert torch.equal(out.x[:, 0], out.n_id.float())
    # local_seeds_per_call counts seed VERTICES: 16 = one batch of 8 seed edges per native call -> three call groups
    assert i == 2 and sampler.calls == ["plain", "plain", "plain"]


# ---- disjoint sampling (homogeneous): the reference's three loader tests -----------------------------------------------------
def test_neighbor_loader_disjoint(stack):
    """tests/loader/test_neighbor_loader.py:838-885 of the reference, verbatim expectations."""
    cugraph_pyg, FS, sampler = stack
    graph_store, feature_store = cugraph_pyg.data.GraphStore(), FS()
    graph_store.put_edge_index(torch.stack([torch.tensor([2, 2]), torch.tensor([0, 1])]), ("node", "connects", "node"), "coo", False, (3, 3))
    feature_store["node", "feat", None] = torch.randint(128, (3, 8))
    NeighborLoader = cugraph_pyg.loader.NeighborLoader
    batch_nd = next(iter(NeighborLoader((feature_store, graph_store), [1], input_nodes=torch.tensor([0, 1]), batch_size=2, disjoint=False,
                                        local_seeds_per_call=64)))
    assert batch_nd.e_id.numel() == 2
    batch_d = next(iter(NeighborLoader((feature_store, graph_store), [1], input_nodes=torch.tensor([0, 1]), batch_size=2, disjoint=True,
                                       local_seeds_per_call=64)))
    assert batch_d.e_id.numel() == 1
    assert batch_d.input_id.min() >= 0
    assert batch_d.input_id.max() == batch_d.batch_size - 1
    assert sorted(batch_d.input_id.tolist()) == [0, 1]
    assert sorted(batch_d.n_id.tolist()) == [0, 1, 2]


@pytest.mark.parametrize("batch_size", [1, 2, 4, 8, 16])
def test_neighbor_loader_disjoint_batch_structure(stack, batch_size):
    """tests/loader/test_neighbor_loader.py:888-935 of the reference on the karate graph: the vertex sets of the trees grown
    from the seeds of a mini-batch are pairwise disjoint."""
    import os

    cugraph_pyg, FS, sampler = stack
    el = np.loadtxt(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "karate.csv"), delimiter=" ", usecols=(0, 1)).astype(np.int64)
    src, dst = torch.from_numpy(el[:, 0].copy()), torch.from_numpy(el[:, 1].copy())
    num_nodes = int(el.max()) + 1
    graph_store, feature_store = cugraph_pyg.data.GraphStore(), FS()
    graph_store.put_edge_index(torch.stack([dst, src]), ("person", "knows", "person"), "coo", False, (num_nodes, num_nodes))
    feature_store["person", "feat", None] = torch.randint(128, (num_nodes, 16))
    loader = cugraph_pyg.loader.NeighborLoader((feature_store, graph_store), [5, 5], input_nodes=torch.arange(num_nodes), batch_size=batch_size,
                                               disjoint=True, local_seeds_per_call=64)
    batches = edges = 0
    for batch in loader:
        batches += 1
        edges += int(batch.e_id.numel())
        assert int(batch.num_sampled_edges.sum()) == batch.edge_index.shape[1] == batch.e_id.numel()
        tree_vertices = {}
        for n_id in range(int(batch.num_sampled_nodes[0])):
            tree_vertices[n_id] = {n_id}
            edge_offset = 0
            for hop in range(len(batch.num_sampled_edges)):
                edges_hop = int(batch.num_sampled_edges[hop])
                e_h = batch.edge_index[:, edge_offset:edge_offset + edges_hop]
                e_in = torch.isin(e_h[1], torch.tensor(list(tree_vertices[n_id])))
                tree_vertices[n_id].update(e_h[0][e_in].tolist())
                edge_offset += edges_hop
        tv = list(tree_vertices.values())
        for i in range(len(tv)):
            for j in range(i + 1, len(tv)):
                assert (tv[i] & tv[j]) == set()
        # every vertex of the mini-batch belongs to exactly one tree, and the sampled edges are edges of the graph
        assert sorted(set().union(*tv)) == list(range(batch.n_id.numel()))
        assert torch.equal(dst[batch.e_id], batch.n_id[batch.edge_index[0]]) and torch.equal(src[batch.e_id], batch.n_id[batch.edge_index[1]])
    assert batches == (num_nodes + batch_size - 1) // batch_size and edges > 0


def test_link_neighbor_loader_disjoint(stack):
    """tests/loader/test_neighbor_loader.py:138-187 of the reference, verbatim expectations."""
    cugraph_pyg, FS, sampler = stack
    graph_store, feature_store = cugraph_pyg.data.GraphStore(), FS()
    graph_store[("node", "connects", "node"), "coo", False, (5, 5)] = torch.stack([torch.tensor([4, 4, 4, 4]), torch.tensor([0, 1, 2, 3])])
    eli = torch.tensor([[0, 2], [1, 3]])
    L = cugraph_pyg.loader.LinkNeighborLoader
    batch_nd = next(iter(L((feature_store, graph_store), num_neighbors=[1], edge_label_index=eli, batch_size=2, shuffle=False, disjoint=False,
                           local_seeds_per_call=64)))
    assert batch_nd.e_id.numel() == 4
    batch_d = next(iter(L((feature_store, graph_store), num_neighbors=[1], edge_label_index=eli, batch_size=2, shuffle=False, disjoint=True,
                          local_seeds_per_call=64)))
    assert batch_d.e_id.numel() == 1
    assert batch_d.edge_label_index.shape == (2, 2)
    assert batch_d.edge_label_index.min() >= 0 and batch_d.edge_label_index.max() < batch_d.n_id.numel()
    assert batch_d.n_id[batch_d.edge_label_index[0]].tolist() == [0, 2] and batch_d.n_id[batch_d.edge_label_index[1]].tolist() == [1, 3]


def test_neighbor_loader_disjoint_heterogeneous(stack):
    """Disjoint sampling on a typed graph (no reference test exists; same invariants as the homogeneous ones): the trees grown
    from the seeds of a mini-batch share no vertex, every sampled vertex lies in exactly one tree, every kept edge is an edge of
    the graph, and cross-tree edges are gone (fewer edges than the plain loader returns for the same seeds)."""
    cugraph_pyg, FS, sampler = stack
    rng = np.random.default_rng(4)
    na, nb = 30, 40
    ets = {("a", "x", "b"): (na, nb, 300), ("b", "y", "a"): (nb, na, 300), ("b", "z", "b"): (nb, nb, 400)}
    edges = {et: (torch.from_numpy(rng.integers(0, ns, m)), torch.from_numpy(rng.integers(0, nd, m))) for et, (ns, nd, m) in ets.items()}

    def stores():
        graph_store, feature_store = cugraph_pyg.data.GraphStore(), FS()
        for et, (s, d) in edges.items():
            graph_store[et, "coo", False, (ets[et][0], ets[et][1])] = [s, d]
        feature_store["a", "x", None] = torch.arange(na, dtype=torch.float32).reshape(na, 1)
        feature_store["b", "x", None] = torch.arange(nb, dtype=torch.float32).reshape(nb, 1)
        return feature_store, graph_store

    fan = {et: [3, 2] for et in ets}
    seeds = torch.from_numpy(rng.permutation(nb)[:24])
    totals = {}
    for disjoint in (False, True):
        loader = cugraph_pyg.loader.NeighborLoader(stores(), num_neighbors=fan, input_nodes=("b", seeds), batch_size=8, shuffle=False,
                                                   disjoint=disjoint, local_seeds_per_call=16)
        total = 0
        for i, batch in enumerate(loader):
            assert batch["b"].n_id[:8].tolist() == seeds[8 * i:8 * i + 8].tolist()
            for et, (s, d) in edges.items():
                ei, eid = batch[et].edge_index, batch[et].e_id
                total += int(eid.numel())
                assert torch.equal(s[eid], batch[et[0]].n_id[ei[0]]) and torch.equal(d[eid], batch[et[2]].n_id[ei[1]])
                assert int(batch[et].num_sampled_edges.sum()) == eid.numel()
            if not disjoint:
                continue
            trees = [{("b", k)} for k in range(8)]
            offset = {et: 0 for et in ets}
            for hop in range(2):
                grown = [set() for _ in trees]
                for et in ets:
                    k = int(batch[et].num_sampled_edges[hop])
                    e = batch[et].edge_index[:, offset[et]:offset[et] + k]
                    offset[et] += k
                    for s_l, d_l in zip(e[0].tolist(), e[1].tolist()):
                        for j, tr in enumerate(trees):
                            if (et[2], d_l) in tr:
                                grown[j].add((et[0], s_l))
                for tr, g in zip(trees, grown):
                    tr |= g
            for a in range(8):
                for b in range(a + 1, 8):
                    assert not (trees[a] & trees[b])
            every = set().union(*trees)
            assert every == {("a", k) for k in range(batch["a"].n_id.numel())} | {("b", k) for k in range(batch["b"].n_id.numel())}
        totals[disjoint] = total
    assert 0 < totals[True] < totals[False]


def test_temporal_and_disjoint_combine(stack):
    """time_attr + disjoint on a typed graph: the temporal call's result goes through the same cross-tree filter."""
    cugraph_pyg, FS, sampler = stack
    graph_store, feature_store = _cite_graph(cugraph_pyg, FS, False)
    graph_store[("author", "writes", "paper"), "coo", False, (3, 4)] = [torch.tensor([0, 0, 1, 1, 2, 2, 2]), torch.tensor([3, 2, 2, 1, 3, 2, 0])]
    feature_store[("author", "writes", "paper"), "time", None] = torch.tensor([0, 0, 1, 0, 2, 1, 1])
    # fan-outs above every eligible count: both loaders take all eligible edges, whatever seed they draw for the epoch
    kw = dict(num_neighbors={("paper", "cites", "paper"): [8, 8], ("author", "writes", "paper"): [8, 8]}, batch_size=2,
              input_nodes=("paper", torch.tensor([3, 2])), input_time=torch.tensor([-1, -1]), time_attr="time", shuffle=False,
              temporal_comparison="strictly_increasing", local_seeds_per_call=64)
    plain = next(iter(cugraph_pyg.loader.NeighborLoader((feature_store, graph_store), disjoint=False, **kw)))
    graph_store, feature_store = _cite_graph(cugraph_pyg, FS, False)
    graph_store[("author", "writes", "paper"), "coo", False, (3, 4)] = [torch.tensor([0, 0, 1, 1, 2, 2, 2]), torch.tensor([3, 2, 2, 1, 3, 2, 0])]
    feature_store[("author", "writes", "paper"), "time", None] = torch.tensor([0, 0, 1, 0, 2, 1, 1])
    dis = next(iter(cugraph_pyg.loader.NeighborLoader((feature_store, graph_store), disjoint=True, **kw)))
    assert sampler.calls == ["temporal", "temporal"]
    for t in ("author", "paper"):
        assert dis[t].n_id.tolist() == plain[t].n_id.tolist()  # the vertex set does not change
    n_plain = sum(int(plain[et].e_id.numel()) for et in (("paper", "cites", "paper"), ("author", "writes", "paper")))
    n_dis = sum(int(dis[et].e_id.numel()) for et in (("paper", "cites", "paper"), ("author", "writes", "paper")))
    assert 0 < n_dis < n_plain  # seed 3 reaches paper 2 (the other seed) and shared authors: those edges cross trees


@pytest.mark.parametrize("mode", ["binary", "triplet"])
def test_link_loader_negative_sampling_unchanged(stack, mode):
    """Negative sampling (GPU-verified) through the same stack: guards BaseSampler.sample_from_edges, which the temporal link
    path touched."""
    cugraph_pyg, FS, sampler = stack
    from cugraph_pyg._pyg_compat import NegativeSampling

    rng = np.random.default_rng(6)
    n, e = 50, 900
    src, dst = torch.from_numpy(rng.integers(0, n, e)), torch.from_numpy(rng.integers(0, n, e))
    graph_store, feature_store = cugraph_pyg.data.GraphStore(), FS()
    graph_store[("n", "to", "n"), "coo", False, (n, n)] = [src, dst]
    feature_store["n", "x", None] = torch.arange(n, dtype=torch.float32).reshape(n, 1)
    eli = torch.stack([src[:16], dst[:16]])
    loader = cugraph_pyg.loader.LinkNeighborLoader((feature_store, graph_store), num_neighbors=[3], batch_size=8, edge_label_index=eli,
                                                   neg_sampling=NegativeSampling(mode, amount=1.0), shuffle=False, local_seeds_per_call=64)
    seen = 0
    for out in loader:
        seen += 1
        assert torch.equal(out.x[:, 0], out.n_id.float())
        if mode == "binary":
            lab = out.edge_label
            assert out.edge_label_index.shape[1] == lab.numel() and int((lab == 1).sum()) == 8 and int((lab == 0).sum()) >= 1
            pos = lab == 1
            assert torch.equal(out.n_id[out.edge_label_index[0][pos]], eli[0, 8 * (seen - 1):8 * seen])
            assert torch.equal(out.n_id[out.edge_label_index[1][pos]], eli[1, 8 * (seen - 1):8 * seen])
        else:  # as tests/test_gpu_loader.py::test_link_neighbor_loader_negative_sampling checks the triplet form
            lab = out.edge_label
            pos = int((lab == 1.0).sum())
            assert pos == out.input_id.numel() == 8 and lab.numel() > pos and bool((lab[pos:] == 0.0).all())
            got = out.n_id[out.edge_label_index]
            assert got.shape[1] == lab.numel() and torch.equal(got[1, :pos], eli[1, 8 * (seen - 1):8 * seen])
    assert seen == 2
    with pytest.raises(NotImplementedError):
        cugraph_pyg.loader.LinkNeighborLoader((feature_store, graph_store), num_neighbors=[3], edge_label_index=eli, edge_label_time=torch.zeros(16),
                                              time_attr="x", neg_sampling=NegativeSampling("binary"))
