"""Temporal and disjoint sampling on the GPU (first run on hardware in round 2: 59 passed, gpurun_out/r2a_temporal_tests.log):
  * S0 temporal (wholegraph_temporal_multihop_neighbor_sample_begin, uniform and biased): bit-exact against the oracle, the
    reference's deterministic pins (tests/loader/test_neighbor_loader.py:943-1170), "open window = plain sampling";
  * the temporal node / link loaders;
  * disjoint sampling.
The same configurations also run on the CPU through the SIMT emulator in tests/emu (tests/test_emulated_*_cpu.py).
"""
import numpy as np
import pytest

from graphs import random_typed_graph

pytestmark = [pytest.mark.gpu]

HETERO_KEYS = ("label_type_hop_offsets", "renumber_map_offsets", "renumber_map", "majors", "minors", "edge_id", "edge_type",
               "edge_renumber_map", "edge_renumber_map_offsets", "label_type_step_base")
COMPARISONS = ("strictly_increasing", "monotonically_increasing", "strictly_decreasing", "monotonically_decreasing")


@pytest.fixture(scope="module")
def env():
    import torch
    import pylibwholegraph.torch as wgth

    torch.cuda.set_device(0)
    wgth.init(0, 1, 0, 1)
    return wgth, wgth.get_global_communicator(), wgth.MultiHopSampler()


def _dev(arrs):
    import torch

    return [torch.from_numpy(np.ascontiguousarray(a)).cuda() for a in arrs]


def _assert_equal(got, exp, keys=HETERO_KEYS):
    for k in keys:
        g = got[k].cpu().numpy()
        assert g.shape == exp[k].shape, (k, g.shape, exp[k].shape)
        assert np.array_equal(g, exp[k]), k


@pytest.mark.parametrize("comparison", COMPARISONS)
@pytest.mark.parametrize("fanout", [[3, 2, 4, 2, 2, 2], [-1, 3, 0, 2, -1, 1], [5, 5, 5], [0, 0, 0, 4, 4, 4], [2, 40, 1], [33, 70, 8, 3, 3, 3]])
@pytest.mark.parametrize("col_dtype", [np.int32, np.int64])
def test_temporal_hetero_bit_exact_vs_oracle(env, oracle, comparison, fanout, col_dtype):
    import torch

    wgth, comm, sampler = env
    edge_types = [(0, 1), (1, 0), (1, 1)]
    vto, row_ptrs, cols = random_typed_graph([700, 1500], edge_types, [9000, 14000, 60000], seed=len(fanout), col_dtype=col_dtype)
    rng = np.random.default_rng(4)
    times = [rng.integers(0, 50, c.shape[0]).astype(np.int64) for c in cols]
    eids = [rng.permutation(c.shape[0]).astype(np.int64) + 1000000 * t for t, c in enumerate(cols)]
    seeds = np.concatenate([rng.integers(0, 2200, 60), rng.integers(700, 2200, 33), rng.integers(0, 700, 1)]).astype(np.int64)
    seeds[7] = seeds[3]  # a repeated seed inside a label: the first occurrence's time counts
    lo = np.array([0, 60, 60, 93, 94], dtype=np.int64)
    mid = 10 if "increasing" in comparison else 40
    seed_times = (mid + rng.integers(-5, 6, seeds.shape[0])).astype(np.int64)
    d_rp, d_col, d_tm, d_eid = _dev(row_ptrs), _dev(cols), _dev(times), _dev(eids)
    for rep in range(2):
        got = sampler.sample_temporal(d_rp, d_col, d_tm, torch.from_numpy(seeds).cuda(), torch.from_numpy(seed_times).cuda(),
                                      torch.from_numpy(lo).cuda(), fanout, 31 + rep, comparison, vertex_type_offsets=vto.tolist(),
                                      csr_edge_ids=d_eid)
        exp = oracle.temporal_multihop_sample(row_ptrs, cols, times, vto, seeds, seed_times, lo, fanout, 31 + rep, comparison, edge_ids=eids)
        assert exp["majors"].shape[0] > 0 or not any(fanout[:3])  # hop 0 with fan-out 0 for every type: nothing is sampled
        _assert_equal(got, exp)


@pytest.mark.parametrize("fanout", [[4, 3], [40, 5], [-1, 2]])
def test_temporal_homogeneous_matches_oracle(env, oracle, fanout):
    """One edge type through the homogeneous finish: COO output against the oracle's T = 1, Vt = 1 result."""
    import torch

    wgth, comm, sampler = env
    vto, row_ptrs, cols = random_typed_graph([3000], [(0, 0)], [90000], seed=11)
    rng = np.random.default_rng(1)
    times = [rng.integers(0, 1000, cols[0].shape[0]).astype(np.int64)]
    seeds = rng.integers(0, 3000, 200).astype(np.int64)
    seed_times = rng.integers(300, 700, 200).astype(np.int64)
    lo = np.array([0, 64, 128, 200], dtype=np.int64)
    got = sampler.sample_temporal(_dev(row_ptrs), _dev(cols), _dev(times), torch.from_numpy(seeds).cuda(),
                                  torch.from_numpy(seed_times).cuda(), torch.from_numpy(lo).cuda(), fanout, 62, "monotonically_decreasing")
    exp = oracle.temporal_multihop_sample(row_ptrs, cols, times, vto, seeds, seed_times, lo, fanout, 62, "monotonically_decreasing")
    L, B = len(fanout), 3
    assert np.array_equal(got["majors"].cpu().numpy(), exp["majors"])
    assert np.array_equal(got["minors"].cpu().numpy(), exp["minors"])
    assert np.array_equal(got["edge_id"].cpu().numpy(), exp["edge_renumber_map"])  # homogeneous edge_id = original edge id
    assert np.array_equal(got["renumber_map"].cpu().numpy(), exp["renumber_map"])
    assert np.array_equal(got["renumber_map_offsets"].cpu().numpy(), exp["renumber_map_offsets"])
    assert np.array_equal(got["label_hop_offsets"].cpu().numpy(), exp["label_type_hop_offsets"])
    assert np.array_equal(got["label_step_base"].cpu().numpy(), exp["label_type_step_base"].reshape(L + 1, B))
    # hop-0 edges respect the comparison with their seed's time (majors of hop 0 are seed-local ids in first-occurrence order)
    lho = exp["label_type_hop_offsets"]
    for l in range(B):
        uniq = list(dict.fromkeys(seeds[lo[l]:lo[l + 1]].tolist()))
        first_time = {v: seed_times[lo[l] + seeds[lo[l]:lo[l + 1]].tolist().index(v)] for v in uniq}
        a, b = lho[l * L], lho[l * L + 1]
        maj = got["majors"].cpu().numpy()[a:b]
        eid = got["edge_id"].cpu().numpy()[a:b]
        assert all(times[0][e] <= first_time[uniq[m]] for m, e in zip(maj, eid))


def test_temporal_open_window_equals_plain_sampling(env):
    """With every edge eligible the temporal kernels must return exactly what the plain kernels return (same streams,
    same order) -- at a size the oracle would not finish quickly."""
    import torch

    wgth, comm, sampler = env
    edge_types = [(0, 1), (1, 0)]
    vto, row_ptrs, cols = random_typed_graph([20000, 30000], edge_types, [600000, 900000], seed=6)
    times = [np.ones(c.shape[0], dtype=np.int64) for c in cols]
    rng = np.random.default_rng(0)
    seeds = rng.integers(0, 50000, 4096).astype(np.int64)
    lo = np.arange(0, 4097, 256, dtype=np.int64)
    d_rp, d_col, d_tm = _dev(row_ptrs), _dev(cols), _dev(times)
    d_seeds, d_lo = torch.from_numpy(seeds).cuda(), torch.from_numpy(lo).cuda()
    for fanout in ([10, 10, 5, 5], [40, 3, 2, 2], [-1, 2, 1, 1]):
        plain = sampler.sample_hetero(d_rp, d_col, vto.tolist(), d_seeds, d_lo, fanout, 5)
        temp = sampler.sample_temporal(d_rp, d_col, d_tm, d_seeds, torch.ones(4096, dtype=torch.int64).cuda(), d_lo, fanout, 5,
                                       "monotonically_increasing", vertex_type_offsets=vto.tolist())
        for k in HETERO_KEYS:
            assert torch.equal(plain[k], temp[k]), (fanout, k)


def test_biased_temporal_open_window_equals_plain_biased(env):
    """Every edge eligible: temporal_weighted_kernel must be weighted_kernel (same draws, candidates and order), two hops."""
    import torch

    wgth, comm, sampler = env
    edge_types = [(0, 1), (1, 0), (1, 1)]
    vto, row_ptrs, cols = random_typed_graph([700, 1500], edge_types, [9000, 14000, 60000], seed=6)
    rng = np.random.default_rng(4)
    seeds = np.concatenate([rng.integers(0, 2200, 60), rng.integers(700, 2200, 34)]).astype(np.int64)
    lo = np.array([0, 60, 60, 93, 94], dtype=np.int64)
    d_rp, d_col = _dev(row_ptrs), _dev(cols)
    d_seeds, d_lo = torch.from_numpy(seeds).cuda(), torch.from_numpy(lo).cuda()
    d_tm = _dev([np.full(c.shape[0], 3, dtype=np.int64) for c in cols])
    for wdtype in (np.float32, np.float64):
        d_w = _dev([(rng.random(c.shape[0]) + 0.01).astype(wdtype) for c in cols])
        for fanout in ([4, 3, 2, 2, 2, 2], [40, 3, 300, 2, 2, 2]):
            plain = sampler.sample_hetero(d_rp, d_col, vto.tolist(), d_seeds, d_lo, fanout, 5, csr_weights=d_w)
            temp = sampler.sample_temporal(d_rp, d_col, d_tm, d_seeds, torch.full((94,), 3, dtype=torch.int64).cuda(), d_lo, fanout, 5,
                                           "monotonically_decreasing", vertex_type_offsets=vto.tolist(), csr_weights=d_w)
            for k in HETERO_KEYS:
                assert torch.equal(plain[k], temp[k]), (wdtype, fanout, k)


def test_biased_temporal_one_hop_sets_vs_oracle(env, oracle):
    """One hop, random times: per frontier row the set of sampled edges against the oracle (the reference compares weighted
    samples as sets; a device / glibc log1pf ulp difference may flip a near-tie in < 1 % of the rows)."""
    import torch

    wgth, comm, sampler = env
    edge_types = [(0, 1), (1, 0), (1, 1)]
    vto, row_ptrs, cols = random_typed_graph([700, 1500], edge_types, [9000, 14000, 60000], seed=6)
    rng = np.random.default_rng(4)
    seeds = rng.permutation(2200)[:94].astype(np.int64)
    lo = np.array([0, 60, 60, 93, 94], dtype=np.int64)
    wts = [(rng.random(c.shape[0]) + 0.01).astype(np.float32) for c in cols]
    times = [rng.integers(0, 50, c.shape[0]).astype(np.int64) for c in cols]
    seed_times = (40 + rng.integers(-5, 6, 94)).astype(np.int64)
    fanout = [4, 40, 3]
    got = sampler.sample_temporal(_dev(row_ptrs), _dev(cols), _dev(times), torch.from_numpy(seeds).cuda(), torch.from_numpy(seed_times).cuda(),
                                  torch.from_numpy(lo).cuda(), fanout, 77, "monotonically_decreasing", vertex_type_offsets=vto.tolist(),
                                  csr_weights=_dev(wts))
    exp = oracle.temporal_multihop_sample(row_ptrs, cols, times, vto, seeds, seed_times, lo, fanout, 77, "monotonically_decreasing", weights=wts)
    got = {k: v.cpu().numpy() for k, v in got.items()}
    for k in ("label_type_hop_offsets", "majors", "edge_type", "edge_id", "edge_renumber_map_offsets"):
        assert np.array_equal(got[k].reshape(-1), exp[k].reshape(-1)), k  # counts are min(eligible, fan-out): independent of the draws

    def rows(out):
        d, lto = {}, out["label_type_hop_offsets"]
        for g in range(4 * 3):
            for m, e in zip(out["majors"][lto[g]:lto[g + 1]].tolist(), out["edge_renumber_map"][lto[g]:lto[g + 1]].tolist()):
                d.setdefault((g, m), []).append(e)
        return {k: sorted(v) for k, v in d.items()}

    a, b = rows(got), rows(exp)
    assert a.keys() == b.keys() and len(a) > 100
    bad = sum(a[k] != b[k] for k in a)
    assert bad <= max(1, len(a) // 100)


def test_temporal_rejects_bad_arguments(env):
    import torch

    wgth, comm, sampler = env
    vto, row_ptrs, cols = random_typed_graph([50], [(0, 0)], [300], seed=2)
    d_rp, d_col = _dev(row_ptrs), _dev(cols)
    tm = [torch.zeros(300, dtype=torch.int64).cuda()]
    seeds, lo = torch.arange(10).cuda(), torch.tensor([0, 10]).cuda()
    with pytest.raises(ValueError):
        sampler.sample_temporal(d_rp, d_col, tm, seeds, torch.zeros(10, dtype=torch.int64).cuda(), lo, [2], 1, "sometimes")
    with pytest.raises(Exception):  # edge times shorter than the edge list
        sampler.sample_temporal(d_rp, d_col, [tm[0][:100]], seeds, torch.zeros(10, dtype=torch.int64).cuda(), lo, [2], 1, "strictly_increasing")


def test_temporal_loader_reference_pin_homogeneous():
    """tests/loader/test_neighbor_loader.py:943-990 (uniform variant) through GraphStore / NeighborLoader."""
    import torch
    import cugraph_pyg
    from cugraph_pyg.data import GraphStore, FeatureStore

    src_cite, dst_cite, tme_cite = torch.tensor([3, 2, 1, 2]), torch.tensor([2, 1, 0, 0]), torch.tensor([0, 1, 2, 0])
    graph_store, feature_store = GraphStore(), FeatureStore()
    graph_store[("paper", "cites", "paper"), "coo", False, (4, 4)] = [dst_cite, src_cite]
    feature_store[("paper", "cites", "paper"), "time", None] = tme_cite
    loader = cugraph_pyg.loader.NeighborLoader((feature_store, graph_store), num_neighbors=[2, 2, 2], batch_size=1,
                                               input_nodes=torch.tensor([3]), input_time=torch.tensor([-1]), time_attr="time",
                                               shuffle=False, temporal_comparison="strictly_increasing")
    out = next(iter(loader))
    assert out.n_id.tolist() == [3, 2, 1, 0]
    assert out.e_id.tolist() == [0, 1, 2]
    assert out.num_sampled_nodes.tolist() == [1, 1, 1, 1]
    assert out.num_sampled_edges.tolist() == [1, 1, 1]


def test_temporal_loader_reference_pin_heterogeneous():
    """tests/loader/test_neighbor_loader.py:993-1058 (uniform variant)."""
    import torch
    import cugraph_pyg
    from cugraph_pyg.data import GraphStore, FeatureStore

    src_cite, dst_cite, tme_cite = torch.tensor([3, 2, 1, 2]), torch.tensor([2, 1, 0, 0]), torch.tensor([0, 1, 2, 0])
    src_author = torch.tensor([3, 2, 2, 1, 3, 2, 0])
    dst_author = torch.tensor([0, 0, 1, 1, 2, 2, 2])
    tme_author = torch.tensor([0, 0, 1, 0, 2, 1, 1])
    graph_store, feature_store = GraphStore(), FeatureStore()
    graph_store[("paper", "cites", "paper"), "coo", False, (4, 4)] = [dst_cite, src_cite]
    graph_store[("author", "writes", "paper"), "coo", False, (3, 4)] = [dst_author, src_author]
    feature_store[("paper", "cites", "paper"), "time", None] = tme_cite
    feature_store[("author", "writes", "paper"), "time", None] = tme_author
    loader = cugraph_pyg.loader.NeighborLoader(
        (feature_store, graph_store),
        num_neighbors={("paper", "cites", "paper"): [2, 2, 2], ("author", "writes", "paper"): [2, 2, 0]},
        batch_size=1, input_nodes=("paper", torch.tensor([3])), input_time=torch.tensor([-1]), time_attr="time", shuffle=False,
        temporal_comparison="strictly_increasing")
    out = next(iter(loader))
    assert sorted(out["author"].n_id.tolist()) == [0, 1, 2]
    assert out["paper"].n_id.tolist() == [3, 2, 1, 0]
    assert sorted(out["author", "writes", "paper"].e_id.tolist()) == [0, 2, 4, 5]
    assert out["author", "writes", "paper"].num_sampled_edges.tolist() == [2, 2, 0]


def test_temporal_link_loader_reference_pins():
    """tests/loader/test_neighbor_loader.py:1059-1170 (uniform variants): seed edges with edge_label_time."""
    import torch
    import cugraph_pyg
    from cugraph_pyg.data import GraphStore, FeatureStore

    src_cite, dst_cite, tme_cite = torch.tensor([3, 2, 1, 2]), torch.tensor([2, 1, 0, 0]), torch.tensor([0, 1, 2, 0])
    graph_store, feature_store = GraphStore(), FeatureStore()
    graph_store[("paper", "cites", "paper"), "coo", False, (4, 4)] = [dst_cite, src_cite]
    feature_store[("paper", "cites", "paper"), "time", None] = tme_cite
    loader = cugraph_pyg.loader.LinkNeighborLoader((feature_store, graph_store), num_neighbors=[2, 2, 2], batch_size=1,
                                                   edge_label_index=torch.tensor([[3], [3]]), edge_label_time=torch.tensor([-1]), time_attr="time",
                                                   shuffle=False, temporal_comparison="strictly_increasing")
    out = next(iter(loader))
    assert out.n_id.tolist() == [3, 2, 1, 0]

    graph_store, feature_store = GraphStore(), FeatureStore()
    graph_store[("paper", "cites", "paper"), "coo", False, (4, 4)] = [dst_cite, src_cite]
    graph_store[("author", "writes", "paper"), "coo", False, (3, 4)] = [torch.tensor([0, 0, 1, 1, 2, 2, 2]), torch.tensor([3, 2, 2, 1, 3, 2, 0])]
    feature_store[("paper", "cites", "paper"), "time", None] = tme_cite
    feature_store[("author", "writes", "paper"), "time", None] = torch.tensor([0, 0, 1, 0, 2, 1, 1])
    loader = cugraph_pyg.loader.LinkNeighborLoader(
        (feature_store, graph_store),
        num_neighbors={("paper", "cites", "paper"): [2, 2, 2], ("author", "writes", "paper"): [2, 2, 0]},
        batch_size=1, edge_label_index=(("author", "writes", "paper"), torch.tensor([[0], [3]])), edge_label_time=torch.tensor([-1]),
        time_attr="time", shuffle=False, temporal_comparison="strictly_increasing")
    out = next(iter(loader))
    assert sorted(out["author"].n_id.tolist()) == [0, 1, 2]
    assert out["paper"].n_id.tolist() == [3, 2, 1, 0]
    assert sorted(out["author", "writes", "paper"].e_id.tolist()) == [0, 2, 4, 5]
    assert out["author", "writes", "paper"].num_sampled_edges.tolist() == [2, 2, 0]


def test_disjoint_loader_reference_pins():
    """Disjoint sampling (pylibcugraph._disjoint_filter: torch ops on top of the plain call; checked on the CPU in
    tests/test_loaders_emulated_cpu.py, gated here with the rest of what has not run on a GPU):
    tests/loader/test_neighbor_loader.py:838-885 and :138-187 of the reference."""
    import torch
    import cugraph_pyg
    from cugraph_pyg.data import GraphStore, FeatureStore

    graph_store, feature_store = GraphStore(), FeatureStore()
    graph_store.put_edge_index(torch.stack([torch.tensor([2, 2]), torch.tensor([0, 1])]), ("node", "connects", "node"), "coo", False, (3, 3))
    feature_store["node", "feat", None] = torch.randint(128, (3, 8)).float()
    batch_d = next(iter(cugraph_pyg.loader.NeighborLoader((feature_store, graph_store), [1], input_nodes=torch.tensor([0, 1]), batch_size=2,
                                                          disjoint=True)))
    assert batch_d.e_id.numel() == 1 and sorted(batch_d.input_id.tolist()) == [0, 1] and sorted(batch_d.n_id.tolist()) == [0, 1, 2]

    graph_store, feature_store = GraphStore(), FeatureStore()
    graph_store[("node", "connects", "node"), "coo", False, (5, 5)] = torch.stack([torch.tensor([4, 4, 4, 4]), torch.tensor([0, 1, 2, 3])])
    feature_store["node", "feat", None] = torch.zeros(5, 2)
    eli = torch.tensor([[0, 2], [1, 3]])
    batch_d = next(iter(cugraph_pyg.loader.LinkNeighborLoader((feature_store, graph_store), num_neighbors=[1], edge_label_index=eli, batch_size=2,
                                                              shuffle=False, disjoint=True)))
    assert batch_d.e_id.numel() == 1
    assert batch_d.n_id[batch_d.edge_label_index[0]].tolist() == [0, 2] and batch_d.n_id[batch_d.edge_label_index[1]].tolist() == [1, 3]
