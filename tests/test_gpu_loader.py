"""cugraph_pyg layer on the GPU: GraphStore / FeatureStore / NeighborLoader invariants taken from the reference's
own tests (python/cugraph-pyg/cugraph_pyg/tests/loader/test_neighbor_loader.py:20-97, data/test_feature_store.py,
sampler/test_distributed_sampler.py) plus exact agreement of the loader's mini-batches with the oracle."""
import os

import numpy as np
import pytest

from graphs import karate_csr, random_csr

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def pyg():
    import torch

    torch.cuda.set_device(0)
    os.environ.setdefault("LOCAL_WORLD_SIZE", "1")
    import cugraph_pyg
    from cugraph_pyg.data import GraphStore, FeatureStore
    from cugraph_pyg.loader import NeighborLoader

    return torch, GraphStore, FeatureStore, NeighborLoader


def _karate_edge_index(torch):
    row_ptr, col = karate_csr(np.int64)
    dst = np.repeat(np.arange(34), np.diff(row_ptr))
    # PyG convention: edge_index[0] = source, edge_index[1] = destination; CSR rows are destinations
    return torch.from_numpy(np.stack([col, dst]))


def test_neighbor_loader_karate_features_follow_n_id(pyg):
    """reference: test_neighbor_loader (:20-49): feat[n_id] == batch.feat for every batch."""
    torch, GraphStore, FeatureStore, NeighborLoader = pyg
    ei = _karate_edge_index(torch)
    graph_store = GraphStore()
    graph_store.put_edge_index(ei, ("person", "knows", "person"), "coo", False, (34, 34))
    feature_store = FeatureStore()
    feat = torch.randint(128, (34, 16))
    feature_store["person", "feat", None] = feat
    loader = NeighborLoader((feature_store, graph_store), [5, 5], input_nodes=torch.arange(34))
    n_batches = 0
    for batch in loader:
        n_batches += 1
        assert (feature_store["person", "feat", None][batch.n_id].cpu() == batch.feat.cpu()).all()
        assert torch.equal(batch.feat.cpu(), feat[batch.n_id.cpu()])
        # every sampled edge exists: PyG edge (src -> dst) with src = n_id[row], dst = n_id[col]
        src = batch.n_id[batch.edge_index[0]].cpu()
        dst = batch.n_id[batch.edge_index[1]].cpu()
        have = set(zip(ei[0].tolist(), ei[1].tolist()))
        assert all((int(s), int(d)) in have for s, d in zip(src, dst))
        assert batch.batch_size == batch.input_id.numel() <= 16
        assert int(batch.num_sampled_nodes.sum()) == batch.n_id.numel()
        assert int(batch.num_sampled_edges.sum()) == batch.edge_index.shape[1]
        assert torch.equal(batch.n_id[: batch.batch_size].cpu(), batch.batch.cpu())
    assert n_batches == len(loader) == 3


def test_neighbor_loader_len(pyg):
    """reference: test_neighbor_loader_len (:52-96)."""
    torch, GraphStore, FeatureStore, NeighborLoader = pyg
    src, dst = torch.tensor([1, 2, 3, 4]), torch.tensor([0, 1, 2, 3])
    graph_store = GraphStore()
    graph_store.put_edge_index(torch.stack([src, dst]), ("person", "knows", "person"), "coo", False, (5, 5))
    feature_store = FeatureStore()
    feature_store["person", "feat", None] = torch.randint(128, (5, 16))
    assert len(NeighborLoader((feature_store, graph_store), [1], input_nodes=torch.arange(5), batch_size=2)) == 3
    assert len(NeighborLoader((feature_store, graph_store), [1], input_nodes=torch.arange(5), batch_size=2, drop_last=True)) == 2
    with pytest.raises(ValueError, match="input_nodes"):
        len(NeighborLoader((feature_store, graph_store), [1], input_nodes="person", batch_size=2))


@pytest.mark.parametrize("compression", ["CSR", "COO"])
def test_neighbor_loader_matches_oracle(pyg, oracle, compression):
    """Loader mini-batches == oracle multihop on the same seeds / seed / fan-out, through every python layer."""
    torch, GraphStore, FeatureStore, NeighborLoader = pyg
    import cugraph_pyg.loader.node_loader as nl

    nodes, edges, dim = 5003, 60000, 32
    row_ptr, col = random_csr(nodes, edges, seed=29, col_dtype=np.int64)
    dst = np.repeat(np.arange(nodes), np.diff(row_ptr))
    ei = torch.from_numpy(np.stack([col, dst]))
    graph_store = GraphStore()
    graph_store[("n", "e", "n"), "coo", False, (nodes, nodes)] = ei
    feature_store = FeatureStore(location="cuda")
    x = torch.randn((nodes, dim))
    y = torch.randint(0, 10, (nodes,))
    feature_store["n", "x", None] = x
    feature_store["n", "y", None] = y
    seeds = torch.from_numpy(np.random.default_rng(0).permutation(nodes)[:1000])
    nl.generate_seed = lambda: 4321  # pin the per-iterator seed
    loader = NeighborLoader((feature_store, graph_store), [10, 5], input_nodes=seeds, batch_size=128,
                            compression=compression, local_seeds_per_call=384)  # 3 batches per call group
    lo = np.concatenate([np.arange(0, 1000, 128), [1000]])
    batches = list(loader)
    assert len(batches) == 8
    for b, batch in enumerate(batches):
        group, within = divmod(b, 3)
        g_lo = lo[group * 3: min(group * 3 + 3, 8) + 1] - lo[group * 3]
        g_seeds = seeds.numpy()[lo[group * 3]: lo[min(group * 3 + 3, 8)]]
        exp = oracle.multihop_sample(row_ptr, col, g_seeds, g_lo, [10, 5], 4321 + group)
        a, e = exp["label_hop_offsets"][within * 2], exp["label_hop_offsets"][within * 2 + 2]
        m0, m1 = exp["renumber_map_offsets"][within], exp["renumber_map_offsets"][within + 1]
        assert np.array_equal(batch.n_id.cpu().numpy(), exp["renumber_map"][m0:m1])
        assert np.array_equal(batch.edge_index[0].cpu().numpy(), exp["minors"][a:e])
        assert np.array_equal(batch.edge_index[1].cpu().numpy(), exp["majors"][a:e])
        # the CSR was built from (src sorted stable) so CSR position == position after the store's sort
        assert np.array_equal(np.sort(batch.e_id.cpu().numpy()), np.sort(exp["edge_id"][a:e]))
        assert batch.num_sampled_edges.tolist() == np.diff(exp["label_hop_offsets"][within * 2: within * 2 + 3]).tolist()
        assert torch.equal(batch.x.cpu(), x[batch.n_id.cpu()])
        assert torch.equal(batch.y.cpu(), y[batch.n_id.cpu()])
        assert torch.equal(batch.input_id.cpu(), torch.arange(lo[b], lo[b + 1]))
        if compression == "CSR":
            assert batch.csr_indptr[-1] == batch.edge_index.shape[1]


def test_feature_store_roundtrip_and_graph_store_layouts(pyg):
    torch, GraphStore, FeatureStore, NeighborLoader = pyg
    fs = FeatureStore()
    a = torch.randn((100, 8))
    fs["v", "a", None] = a
    assert torch.equal(fs["v", "a", None].get_local_tensor().cpu(), a)       # index=None returns the DistEmbedding itself
    assert torch.equal(fs["v", "a", torch.tensor([5, 1, 99])].cpu(), a[[5, 1, 99]])
    assert fs.get_tensor_size("v", "a") == (100, 8)
    assert [(t.group_name, t.attr_name) for t in fs.get_all_tensor_attrs()] == [("v", "a")]
    del fs["v", "a", None]
    assert fs.get_all_tensor_attrs() == []
    gs = GraphStore()
    ei = torch.tensor([[0, 1, 1, 2], [1, 0, 2, 1]])
    gs.put_edge_index(ei, ("v", "e", "v"), "coo", False, (3, 3))
    row, col = gs.get_edge_index(("v", "e", "v"), "coo")
    assert torch.equal(torch.stack([row, col]).cpu(), ei)
    assert gs.is_homogeneous and gs._vertex_offsets == {"v": 0}
    assert gs._graph.num_vertices == 3
    gs.finalize()
    with pytest.raises(RuntimeError):
        gs.finalize()


# ---- heterogeneous loaders (reference: test_neighbor_loader.py:355-451) -------------------------------------
def _paper_author(torch, GraphStore):
    src = torch.tensor([0, 1, 2, 4, 3, 4, 5, 5])  # paper
    dst = torch.tensor([4, 5, 4, 3, 2, 1, 0, 1])  # paper
    asrc = torch.tensor([0, 1, 2, 3, 3, 0])  # author
    adst = torch.tensor([0, 1, 2, 3, 4, 5])  # paper
    graph_store = GraphStore()
    graph_store[("paper", "cites", "paper"), "coo", False, (6, 6)] = [src, dst]
    graph_store[("author", "writes", "paper"), "coo", False, (4, 6)] = [asrc, adst]
    return graph_store, src, dst, asrc, adst


def test_neighbor_loader_hetero_basic(pyg):
    torch, GraphStore, FeatureStore, NeighborLoader = pyg
    graph_store, src, dst, asrc, adst = _paper_author(torch, GraphStore)
    feature_store = FeatureStore()
    pfeat, afeat = torch.arange(6 * 4).reshape(6, 4).float(), 100 + torch.arange(4 * 3).reshape(4, 3).float()
    feature_store["paper", "x", None] = pfeat
    feature_store["author", "x", None] = afeat
    loader = NeighborLoader((feature_store, graph_store),
                            num_neighbors={("paper", "cites", "paper"): [1, 1], ("author", "writes", "paper"): [1, 1]},
                            input_nodes=("paper", torch.tensor([0, 1])), batch_size=2)
    out = next(iter(loader))
    pc, aw = out["paper", "cites", "paper"], out["author", "writes", "paper"]
    ei_out = out["paper"].n_id.cpu()[pc.edge_index.cpu()]
    assert (src[pc.e_id.cpu()] == ei_out[0]).all() and (dst[pc.e_id.cpu()] == ei_out[1]).all()
    ej_out = torch.stack([out["author"].n_id.cpu()[aw.edge_index[0].cpu()], out["paper"].n_id.cpu()[aw.edge_index[1].cpu()]])
    assert (asrc[aw.e_id.cpu()] == ej_out[0]).all() and (adst[aw.e_id.cpu()] == ej_out[1]).all()
    # seeds first, features follow n_id, counts add up
    assert out["paper"].n_id[:2].tolist() == [0, 1] and out["paper"].batch_size == 2
    assert torch.equal(out["paper"].x.cpu(), pfeat[out["paper"].n_id.cpu()])
    assert torch.equal(out["author"].x.cpu(), afeat[out["author"].n_id.cpu()])
    assert int(out["paper"].num_sampled_nodes.sum()) == out["paper"].n_id.numel()
    assert int(out["author"].num_sampled_nodes.sum()) == out["author"].n_id.numel()
    assert int(pc.num_sampled_edges.sum()) == pc.edge_index.shape[1] and int(aw.num_sampled_edges.sum()) == aw.edge_index.shape[1]
    # fan-out 1 per type: every seed with an in-edge of that type gets exactly one
    assert pc.num_sampled_edges.tolist()[0] == 2 and aw.num_sampled_edges.tolist()[0] == 2


def test_neighbor_loader_hetero_single_etype(pyg):
    """An edge type absent from num_neighbors has fan-out 0: empty outputs, [0, 0] counts (:414-451)."""
    torch, GraphStore, FeatureStore, NeighborLoader = pyg
    graph_store, src, dst, asrc, adst = _paper_author(torch, GraphStore)
    loader = NeighborLoader((FeatureStore(), graph_store), num_neighbors={("paper", "cites", "paper"): [1, 1]},
                            input_nodes=("paper", torch.tensor([0, 1])), batch_size=2)
    out = next(iter(loader))
    assert out["author"].n_id.numel() == 0
    assert out["author", "writes", "paper"].edge_index.numel() == 0
    assert out["author", "writes", "paper"].num_sampled_edges.tolist() == [0, 0]
    assert out["paper", "cites", "paper"].edge_index.shape[1] > 0


def test_neighbor_loader_hetero_random_graph_matches_oracle(pyg, oracle):
    """Mini-batches of the heterogeneous loader against the oracle run on the same typed CSRs, batch by batch."""
    torch, GraphStore, FeatureStore, NeighborLoader = pyg
    from graphs import typed_csrs

    rng = np.random.default_rng(3)
    n_user, n_item = 300, 500
    # PyG edge types (sorted: ("item","rev","user"), ("user","buys","item"), ("user","follows","user"))
    e = {
        ("user", "buys", "item"): (rng.integers(0, n_user, 4000), rng.integers(0, n_item, 4000)),
        ("item", "rev", "user"): (rng.integers(0, n_item, 3000), rng.integers(0, n_user, 3000)),
        ("user", "follows", "user"): (rng.integers(0, n_user, 2500), rng.integers(0, n_user, 2500)),
    }
    size = {"user": n_user, "item": n_item}
    graph_store = GraphStore()
    for k, (s, d) in e.items():
        graph_store[k, "coo", False, (size[k[0]], size[k[2]])] = [torch.from_numpy(s), torch.from_numpy(d)]
    fanout = {("user", "buys", "item"): [3, 2], ("item", "rev", "user"): [2, 2], ("user", "follows", "user"): [4, 0]}
    seeds = torch.from_numpy(rng.permutation(n_user)[:96])
    loader = NeighborLoader((FeatureStore(), graph_store), num_neighbors=fanout, input_nodes=("user", seeds), batch_size=32,
                            local_seeds_per_call=64)
    batches = list(loader)
    assert len(batches) == 3
    # oracle on the same graph: vertex types sorted (item, user) -> item ids [0, 500), user ids [500, 800)
    off = {"item": 0, "user": n_item}
    keys = sorted(e.keys())
    srcs, dsts, etps = [], [], []
    for t, k in enumerate(keys):
        s, d = e[k]
        srcs.append(d + off[k[2]])  # cuGraph src = PyG destination
        dsts.append(s + off[k[0]])
        etps.append(np.full(len(s), t))
    row_ptrs, cols, pos = typed_csrs(np.concatenate(srcs), np.concatenate(dsts), np.concatenate(etps), 3, n_user + n_item, np.int32)
    fan = [fanout[k][h] for h in range(2) for k in keys]
    for call, (lo, hi) in enumerate([(0, 64), (64, 96)]):
        s = seeds[lo:hi].numpy() + off["user"]
        label_offsets = np.arange(0, hi - lo + 1, 32)
        if label_offsets[-1] != hi - lo:
            label_offsets = np.append(label_offsets, hi - lo)
        # the loader draws random_state per epoch; recover it from the first batch is not possible -> compare structure
        exp = oracle.hetero_multihop_sample(row_ptrs, cols, [0, n_item, n_item + n_user], s, label_offsets, fan, 0)
        for b in range(len(label_offsets) - 1):
            batch = batches[lo // 32 + b]
            # seeds first in the user map; take-all-free counts: hop-0 edge counts are min(deg, fanout) -> identical
            assert batch["user"].n_id[: label_offsets[b + 1] - label_offsets[b]].tolist() == seeds[lo + label_offsets[b]: lo + label_offsets[b + 1]].tolist()
            for t, k in enumerate(keys):
                g = (b * 3 + t) * 2
                assert batch[k].num_sampled_edges.tolist()[0] == exp["label_type_hop_offsets"][g + 1] - exp["label_type_hop_offsets"][g]
    # every sampled edge exists with the reported original id
    for batch in batches:
        for k, (s, d) in e.items():
            ei = batch[k].edge_index.cpu()
            eid = batch[k].e_id.cpu().numpy()
            assert np.array_equal(s[eid], batch[k[0]].n_id.cpu().numpy()[ei[0].numpy()])
            assert np.array_equal(d[eid], batch[k[2]].n_id.cpu().numpy()[ei[1].numpy()])


def test_neighbor_loader_biased_reference_pin(pyg):
    """reference: test_neighbor_loader_biased (test_neighbor_loader.py:97-133) verbatim: a zero-bias edge is never
    sampled, even though its row has fewer candidates than the fan-out."""
    torch, GraphStore, FeatureStore, NeighborLoader = pyg
    eix = torch.tensor([[3, 4, 5], [0, 1, 2]])
    graph_store = GraphStore()
    graph_store.put_edge_index(eix, ("person", "knows", "person"), "coo", False, (6, 6))
    feature_store = FeatureStore()
    feature_store["person", "feat", None] = torch.randint(128, (6, 12))
    feature_store[("person", "knows", "person"), "bias", None] = torch.tensor([0, 12, 14], dtype=torch.float32)
    loader = NeighborLoader((feature_store, graph_store), [1], input_nodes=torch.tensor([0, 1, 2], dtype=torch.int64),
                            batch_size=3, weight_attr="bias")
    out = list(iter(loader))
    assert len(out) == 1
    out = out[0]
    assert out.edge_index.shape[1] == 2
    assert (out.edge_index.cpu() == torch.tensor([[3, 4], [1, 2]])).all()
    assert out.e_id.cpu().tolist() == [1, 2]


# ---- link prediction loaders (reference: test_neighbor_loader.py:196-352, 455-527) -----------------------------
@pytest.mark.parametrize("num_nodes,num_edges,select_edges,batch_size", [(7, 29, 17, 1), (19, 62, 17, 3), (120, 1500, 700, 32)])
@pytest.mark.parametrize("depth,num_neighbors", [(1, 1), (3, 4)])
def test_link_neighbor_loader_basic(pyg, num_nodes, num_edges, select_edges, batch_size, depth, num_neighbors):
    torch, GraphStore, FeatureStore, NeighborLoader = pyg
    from cugraph_pyg.loader import LinkNeighborLoader

    g = torch.Generator().manual_seed(num_edges)
    graph_store = GraphStore()
    ei = torch.stack([torch.randint(0, num_nodes, (num_edges,), generator=g), torch.randint(0, num_nodes, (num_edges,), generator=g)])
    graph_store[("n", "e", "n"), "coo", False, (num_nodes, num_nodes)] = ei
    eix = torch.randperm(num_edges, generator=g)[:select_edges]
    elx = ei[:, eix]
    loader = LinkNeighborLoader((FeatureStore(), graph_store), num_neighbors=[num_neighbors] * depth, edge_label_index=elx,
                                batch_size=batch_size, shuffle=False)
    assert len(loader) == (select_edges + batch_size - 1) // batch_size
    have = set(zip(ei[0].tolist(), ei[1].tolist()))
    seen = 0
    for i, batch in enumerate(loader):
        lo, hi = i * batch_size, min((i + 1) * batch_size, select_edges)
        assert (batch.input_id.cpu() == torch.arange(lo, hi)).all()
        assert (elx[:, lo:hi] == batch.n_id.cpu()[batch.edge_label_index.cpu()]).all()
        # seeds (the batch's distinct endpoints) come first and sorted, as in the reference
        uniq = torch.unique(elx[:, lo:hi])
        assert batch.n_id[: uniq.numel()].cpu().tolist() == uniq.tolist()
        src, dst = batch.n_id.cpu()[batch.edge_index[0].cpu()], batch.n_id.cpu()[batch.edge_index[1].cpu()]
        assert all((int(s), int(d)) in have for s, d in zip(src, dst))
        seen += 1
    assert seen == len(loader)


def test_link_neighbor_loader_len(pyg):
    torch, GraphStore, FeatureStore, NeighborLoader = pyg
    from cugraph_pyg.loader import LinkNeighborLoader

    eli = torch.tensor([[0, 1, 2, 3, 4], [1, 2, 3, 4, 0]])
    graph_store = GraphStore()
    graph_store[("n", "e", "n"), "coo", False, (5, 5)] = eli
    data = (FeatureStore(), graph_store)
    assert len(LinkNeighborLoader(data, num_neighbors=[1], edge_label_index=eli, batch_size=2)) == 3
    assert len(LinkNeighborLoader(data, num_neighbors=[1], edge_label_index=eli, batch_size=2, drop_last=True)) == 2
    loader = LinkNeighborLoader(data, num_neighbors=[1], edge_label_index=("n", "e", "n"), batch_size=2)
    with pytest.raises(ValueError, match="edge_label_index"):
        len(loader)
    assert sum(1 for _ in loader) == 3  # all edges of the type are the seeds


@pytest.mark.parametrize("batch_size", [1, 2])
@pytest.mark.parametrize("mode,amount", [("binary", 1), ("binary", 0.1), ("triplet", 2)])
def test_link_neighbor_loader_negative_sampling(pyg, batch_size, mode, amount):
    torch, GraphStore, FeatureStore, NeighborLoader = pyg
    from cugraph_pyg.loader import LinkNeighborLoader
    from cugraph_pyg._pyg_compat import NegativeSampling

    num_edges, num_nodes, select_edges = 62, 19, 17
    g = torch.Generator().manual_seed(5)
    ei = torch.stack([torch.randint(0, num_nodes, (num_edges,), generator=g), torch.randint(0, num_nodes, (num_edges,), generator=g)])
    graph_store = GraphStore()
    graph_store[("n", "e", "n"), "coo", False, (num_nodes, num_nodes)] = ei
    elx = ei[:, torch.randperm(num_edges, generator=g)[:select_edges]]
    loader = LinkNeighborLoader((FeatureStore(), graph_store), num_neighbors=[3, 3, 3], edge_label_index=elx,
                                batch_size=batch_size, neg_sampling=NegativeSampling(mode, amount), shuffle=False)
    n_pos = 0
    for i, batch in enumerate(loader):
        assert batch.edge_label[0] == 1.0
        pos = int((batch.edge_label == 1.0).sum())
        assert pos == batch.input_id.numel() and batch.edge_label.numel() > pos  # at least one negative per batch
        assert (batch.edge_label[pos:] == 0.0).all()
        lo = i * batch_size
        got = batch.n_id.cpu()[batch.edge_label_index.cpu()]
        assert got.shape[1] == batch.edge_label.numel()
        if mode == "binary":
            assert (got[:, :pos] == elx[:, lo:lo + pos]).all()
        else:
            assert (got[1, :pos] == elx[1, lo:lo + pos]).all()
        assert int(got.max()) < num_nodes
        n_pos += pos
    assert n_pos == select_edges


def test_neighbor_loader_hetero_linkpred(pyg):
    """reference: test_neighbor_loader_hetero_linkpred (:455-527) with its value-level expectations."""
    torch, GraphStore, FeatureStore, NeighborLoader = pyg
    from cugraph_pyg.loader import LinkNeighborLoader

    graph_store, src, dst, asrc, adst = _paper_author(torch, GraphStore)
    loader = LinkNeighborLoader((FeatureStore(), graph_store),
                                num_neighbors={("paper", "cites", "paper"): [2, 2], ("author", "writes", "paper"): [2, 2]},
                                edge_label_index=(("author", "writes", "paper"), torch.stack([asrc, adst])), batch_size=5)
    out = next(iter(loader))
    assert out["paper"].n_id.tolist() == [0, 1, 2, 3, 4, 5]
    assert out["author"].n_id.tolist() == [0, 1, 2, 3]
    assert out["paper"].num_sampled_nodes.tolist() == [5, 1, 0]
    assert out["author"].num_sampled_nodes.tolist() == [4, 0, 0]
    assert out["paper", "cites", "paper"].edge_index.shape == torch.Size([2, 8])
    assert out["paper", "cites", "paper"].num_sampled_edges.tolist() == [7, 1]
    assert "edge_label_index" not in out["paper", "cites", "paper"]
    assert out["author", "writes", "paper"].edge_index.shape == torch.Size([2, 6])
    assert out["author", "writes", "paper"].num_sampled_edges.tolist() == [5, 1]
    assert list(out["author", "writes", "paper"].edge_label_index.shape) == [2, 5]
    assert out["author", "writes", "paper"].edge_label_index.tolist()[0] == [0, 1, 2, 3, 3]
    assert out["author", "writes", "paper"].edge_label_index.tolist()[1] == [0, 1, 2, 3, 4]


def test_feature_store_replicate_hot_rows_keeps_results(pyg):
    """The hot-row replica (B200 extension) never changes what a loader returns."""
    torch, GraphStore, FeatureStore, NeighborLoader = pyg
    ei = _karate_edge_index(torch)
    graph_store = GraphStore()
    graph_store.put_edge_index(ei, ("person", "knows", "person"), "coo", False, (34, 34))
    feature_store = FeatureStore()
    feat = torch.arange(34 * 8).reshape(34, 8).float()
    feature_store["person", "feat", None] = feat
    done = feature_store.replicate_hot_rows(graph_store, ratio=0.25)
    assert done == {("person", "feat"): 8}
    assert torch.equal(feature_store["person", "feat", None][torch.arange(34).cuda()].cpu(), feat)
    for batch in NeighborLoader((feature_store, graph_store), [5, 5], input_nodes=torch.arange(34)):
        assert torch.equal(batch.feat.cpu(), feat[batch.n_id.cpu()])
