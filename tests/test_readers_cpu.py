"""The vectorised mini-batch decode of cugraph_pyg's readers against a line-by-line restatement of the reference's
per-batch decode (python/cugraph-pyg/cugraph_pyg/sampler/sampler.py:642-740 COO, :525-640 CSR, :280-490 hetero), fed
with the ORACLE's sampler output on CPU tensors -- no GPU involved."""
import numpy as np
import pytest
import torch

from graphs import random_csr, random_typed_graph


def _step_base(seeds, lo, minors, lho, L):
    """[L+1, B] first local id discovered at step t, from the hop-monotone renumbering contract."""
    B = len(lo) - 1
    base = np.zeros((L + 1, B), dtype=np.int32)
    for l in range(B):
        cur = len(np.unique(seeds[lo[l]:lo[l + 1]]))
        base[1 if L >= 1 else 0, l] = cur
        for h in range(L):
            if h + 1 <= L:
                base[h + 1, l] = cur
            e0, e1 = lho[l * L + h], lho[l * L + h + 1]
            if e1 > e0:
                cur = max(cur, int(minors[e0:e1].max()) + 1)
        base[0, l] = 0
    return base


def _reference_decode_coo(raw, index):
    """sampler.py:642-740 restated (homogeneous COO)."""
    L = (raw["label_hop_offsets"].numel() - 1) // (raw["renumber_map_offsets"].numel() - 1)
    s = int(raw["label_hop_offsets"][index * L])
    e = int(raw["label_hop_offsets"][(index + 1) * L])
    majors, minors, edge_id = raw["majors"][s:e], raw["minors"][s:e], raw["edge_id"][s:e]
    rmap = raw["map"][int(raw["renumber_map_offsets"][index]):int(raw["renumber_map_offsets"][index + 1])]
    num_sampled_edges = raw["label_hop_offsets"][index * L:(index + 1) * L + 1].diff()
    num_seeds = (majors[: int(num_sampled_edges[0])].max() + 1).reshape((1,))
    hops = torch.tensor([int(minors[: int(num_sampled_edges[:i].sum())].max()) + 1 for i in range(1, L + 1)])
    num_sampled_nodes = torch.cat([num_seeds, hops.diff(prepend=num_seeds)])
    return dict(node=rmap, row=minors, col=majors, edge=edge_id, num_sampled_nodes=num_sampled_nodes,
                num_sampled_edges=num_sampled_edges, num_seeds=int(num_seeds))


@pytest.mark.parametrize("fanout", [[5, 3], [4, 4, 2], [6]])
def test_homogeneous_reader_equals_reference_decode(oracle, fanout):
    from cugraph_pyg.sampler import HomogeneousSampleReader

    nodes = 3000
    row_ptr, col = random_csr(nodes, 40000, seed=2, skew=False)  # every vertex has neighbours: the reference's max()+1 is exact
    assert (np.diff(row_ptr) > 0).all()
    rng = np.random.default_rng(0)
    sizes = [16, 16, 16, 7]
    seeds = np.concatenate([rng.permutation(nodes)[:s] for s in sizes]).astype(np.int64)
    lo = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)
    L = len(fanout)
    out = oracle.multihop_sample(row_ptr, col, seeds, lo, fanout, 7)
    raw = {
        "majors": torch.from_numpy(out["majors"]), "minors": torch.from_numpy(out["minors"]),
        "edge_id": torch.from_numpy(out["edge_id"]), "label_hop_offsets": torch.from_numpy(out["label_hop_offsets"]),
        "map": torch.from_numpy(out["renumber_map"]), "renumber_map_offsets": torch.from_numpy(out["renumber_map_offsets"]),
        "label_step_base": torch.from_numpy(_step_base(seeds, lo, out["minors"], out["label_hop_offsets"], L)),
        "input_index": torch.arange(len(seeds)), "input_offsets": torch.from_numpy(lo),
    }
    reader = HomogeneousSampleReader(iter([(raw, 0, len(sizes) - 1)]))
    for index, got in enumerate(reader):
        exp = _reference_decode_coo(raw, index)
        assert torch.equal(got.node, exp["node"]) and torch.equal(got.row, exp["row"]) and torch.equal(got.col, exp["col"])
        assert torch.equal(got.edge, exp["edge"])
        assert got.num_sampled_edges.tolist() == exp["num_sampled_edges"].tolist()
        assert got.num_sampled_nodes.tolist() == exp["num_sampled_nodes"].tolist()
        assert torch.equal(got.batch, exp["node"][: exp["num_seeds"]])
        assert got.metadata[0].tolist() == list(range(int(lo[index]), int(lo[index + 1])))
        # what the loaders rely on: seeds first, node features follow n_id, every edge endpoint is a known node
        assert got.node[: sizes[index]].tolist() == seeds[lo[index]:lo[index + 1]].tolist()
        assert int(got.row.max()) < got.node.numel() and int(got.col.max()) < got.node.numel()
    assert index == len(sizes) - 1


def test_homogeneous_reader_counts_isolated_seeds_exactly(oracle):
    """Where the two differ on purpose: a batch whose LAST seed has no neighbour.  The reference infers the seed count from
    `majors.max() + 1` and under-counts it (sampler.py:676-683); the native sampler reports it."""
    from cugraph_pyg.sampler import HomogeneousSampleReader

    row_ptr = np.array([0, 2, 4, 4], dtype=np.int64)  # vertex 2 is isolated
    col = np.array([1, 2, 0, 2], dtype=np.int32)
    seeds, lo = np.array([0, 1, 2], dtype=np.int64), np.array([0, 3], dtype=np.int64)
    out = oracle.multihop_sample(row_ptr, col, seeds, lo, [2], 1)
    raw = {"majors": torch.from_numpy(out["majors"]), "minors": torch.from_numpy(out["minors"]), "edge_id": torch.from_numpy(out["edge_id"]),
           "label_hop_offsets": torch.from_numpy(out["label_hop_offsets"]), "map": torch.from_numpy(out["renumber_map"]),
           "renumber_map_offsets": torch.from_numpy(out["renumber_map_offsets"]),
           "label_step_base": torch.from_numpy(_step_base(seeds, lo, out["minors"], out["label_hop_offsets"], 1)),
           "input_index": torch.arange(3), "input_offsets": torch.from_numpy(lo)}
    got = next(HomogeneousSampleReader(iter([(raw, 0, 0)])))
    assert got.num_sampled_nodes.tolist() == [3, 0] and got.batch.tolist() == [0, 1, 2]
    assert _reference_decode_coo(raw, 0)["num_seeds"] == 2  # the reference's estimate


def test_heterogeneous_reader_equals_reference_semantics(oracle):
    """Hetero decode (sampler.py:280-490): per edge type row/col/edge slices, per vertex type de-offset node ids; counts
    from the native step bases instead of max()+1 over the edge arrays."""
    from cugraph_pyg.sampler import HeterogeneousSampleReader

    edge_types_num = [(0, 1), (1, 0), (1, 1)]  # (CSR row type = PyG destination type, column type = PyG source type)
    vto, row_ptrs, cols = random_typed_graph([300, 500], edge_types_num, [4000, 5000, 6000], seed=4)
    rng = np.random.default_rng(1)
    seeds = (300 + rng.permutation(500)[:40]).astype(np.int64)  # type-1 seeds, global ids
    lo = np.array([0, 25, 40], dtype=np.int64)
    fanout = [3, 2, 2, 2, 2, 1]
    out = oracle.hetero_multihop_sample(row_ptrs, cols, vto, seeds, lo, fanout, 3)
    names = ["a", "b"]  # vertex types in sorted order
    pyg_edge_types = [("b", "r0", "a"), ("a", "r1", "b"), ("b", "r2", "b")]  # (PyG src, rel, PyG dst): src = column type
    raw = {k: torch.from_numpy(v) for k, v in out.items() if k != "renumber_map"}
    raw.update(map=torch.from_numpy(out["renumber_map"]), input_index=torch.arange(40), input_offsets=torch.from_numpy(lo), input_type="b")
    reader = HeterogeneousSampleReader(iter([(raw, 0, 1)]), src_types=torch.tensor([1, 0, 1]), dst_types=torch.tensor([0, 1, 1]),
                                       vertex_offsets=torch.from_numpy(vto), edge_types=pyg_edge_types, vertex_types=names)
    T, Vt, L = 3, 2, 2
    lto, rmo, ermo = out["label_type_hop_offsets"], out["renumber_map_offsets"], out["edge_renumber_map_offsets"]
    for index, got in enumerate(reader):
        for vt, name in enumerate(names):
            m = out["renumber_map"][rmo[index * Vt + vt]:rmo[index * Vt + vt + 1]] - vto[vt]
            assert got.node[name].tolist() == m.tolist()  # sampler.py:304-320
            assert int(got.num_sampled_nodes[name].sum()) == len(m)
        assert got.num_sampled_nodes["b"].tolist()[0] == int(lo[index + 1] - lo[index]) and got.num_sampled_nodes["a"].tolist()[0] == 0
        for t, et in enumerate(pyg_edge_types):
            a, b = lto[(index * T + t) * L], lto[(index * T + t + 1) * L]
            emap = out["edge_renumber_map"][ermo[index * T + t]:ermo[index * T + t + 1]]
            assert got.edge[et].tolist() == emap[out["edge_id"][a:b]].tolist()  # sampler.py:334-341
            assert got.col[et].tolist() == out["majors"][a:b].tolist() and got.row[et].tolist() == out["minors"][a:b].tolist()
            assert got.num_sampled_edges[et].tolist() == np.diff(lto[(index * T + t) * L:(index * T + t) * L + L + 1]).tolist()
            # local ids address the right node lists: PyG source = minors -> et[0], destination = majors -> et[2]
            if b > a:
                assert int(got.row[et].max()) < got.node[et[0]].numel() and int(got.col[et].max()) < got.node[et[2]].numel()
        assert got.metadata[0][0] == "b" and got.metadata[0][1].tolist() == list(range(int(lo[index]), int(lo[index + 1])))
    assert index == 1
