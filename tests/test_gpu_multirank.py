"""(e) Multi-GPU path on whatever devices exist: world_size-2 with one process per rank.  With a single
visible GPU both ranks share device 0 -- cudaIpc mapping, the chunk partition plan and the in-kernel
address translation are exactly those of two GPUs (peer loads then stay on one device)."""
import functools
import os
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _setup(rank, world):
    for p in (os.path.join(ROOT, "cugraph-gnn_b200"), os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
        if p not in sys.path:
            sys.path.insert(0, p)
    import torch

    torch.cuda.set_device(rank % torch.cuda.device_count())
    return torch


def _comm(uid, rank, world):
    import pylibwholegraph.binding.wholememory_binding as wmb
    import pylibwholegraph.torch as wgth

    wgth.init(rank, world, rank, world)
    return wgth, wgth.WholeMemoryCommunicator(wmb.create_communicator(wmb.PyWholeMemoryUniqueID(uid), rank, world))


def _gather_scatter_worker(rank, world, uid, partition):
    torch = _setup(rank, world)
    wgth, comm = _comm(uid, rank, world)
    rows, dim = 10007, 128
    part = None
    if partition == "uneven":
        part = [rows // 3, rows - rows // 3] if world == 2 else None
    emb = wgth.create_wholememory_tensor(comm, "chunked", "cuda", [rows, dim], torch.float32, [dim, 1], part)
    local, start = emb.get_local_tensor()
    # reference partition plan: ceil(N / W) rows per rank (memory_handle.cpp:1597-1629)
    if part is None:
        per = (rows + world - 1) // world
        assert start == min(rank * per, rows) and local.shape[0] == min((rank + 1) * per, rows) - start
    else:
        assert start == sum(part[:rank]) and local.shape[0] == part[rank]
    ids = torch.arange(start, start + local.shape[0], device="cuda")
    local.copy_(((ids[:, None] * 7 + torch.arange(dim, device="cuda")[None, :]) % 4096).float())
    torch.cuda.synchronize()
    comm.barrier()
    g = torch.Generator().manual_seed(5)  # same indices on every rank: both read local AND remote rows
    idx = torch.randint(0, rows, (5000,), generator=g)
    idx[::11] = -1
    got = emb.gather(idx.cuda())
    exp = ((idx[:, None] * 7 + torch.arange(dim)[None, :]) % 4096).float()
    keep = idx >= 0
    assert torch.equal(got.cpu()[keep], exp[keep]), "rank %d: chunked gather mismatch" % rank
    # int32 indices + fp16 output over the remote chunk only
    other = (rank + 1) % world
    lo = start if other == rank else (0 if rank == 1 else local.shape[0])
    remote_idx = torch.arange(lo, lo + 100, dtype=torch.int32)
    got16 = emb.gather(remote_idx.cuda(), force_dtype=torch.float16)
    assert torch.equal(got16.cpu(), ((remote_idx.long()[:, None] * 7 + torch.arange(dim)[None, :]) % 4096).half())
    comm.barrier()
    # scatter: rank 0 writes rows that live on every rank, owners verify their local memory
    new_rows = torch.arange(0, rows, 97)
    if rank == 0:
        emb.scatter(torch.full((new_rows.shape[0], dim), -3.0).cuda(), new_rows.cuda())
        torch.cuda.synchronize()
    comm.barrier()
    mine = new_rows[(new_rows >= start) & (new_rows < start + local.shape[0])] - start
    assert bool((local[mine.cuda()] == -3.0).all()), "rank %d: scattered rows did not arrive" % rank
    # per-rank views of every chunk are P2P mapped
    chunks, starts = emb.get_all_chunked_tensor()
    assert len(chunks) == world and starts[rank] == start
    assert float(chunks[other][0, 0]) == float(chunks[other][0, 0])
    comm.barrier()
    wgth.destroy_wholememory_tensor(emb)


def _sampler_worker(rank, world, uid):
    torch = _setup(rank, world)
    wgth, comm = _comm(uid, rank, world)
    import wg_oracle as oracle
    from graphs import random_csr
    from pylibwholegraph.torch import wholegraph_ops

    nodes, edges = 9703, 104323
    row_ptr, col = random_csr(nodes, edges, seed=77)
    # CSR striped over the ranks by element range, as the reference's tests do
    # (cpp/tests/wholegraph_ops/wholegraph_csr_unweighted_sample_without_replacement_tests.cu:182-195)
    def striped(arr):
        t = torch.from_numpy(arr)
        wm = wgth.create_wholememory_tensor(comm, "chunked", "cuda", [arr.shape[0]], t.dtype, [1])
        local, start = wm.get_local_tensor()
        local.copy_(t[start:start + local.shape[0]].cuda())
        return wm

    wm_rp, wm_col = striped(row_ptr), striped(col)
    torch.cuda.synchronize()
    comm.barrier()
    centers = np.random.default_rng(rank).integers(0, nodes, 700).astype(np.int64)
    for M in (10, 25, 50, -1):
        off, dest, lid, gid = wholegraph_ops.unweighted_sample_without_replacement(
            wm_rp.wmb_tensor, wm_col.wmb_tensor, torch.from_numpy(centers).cuda(), M, 1234 + rank, True, True)
        eoff, edest, elid, egid = oracle.unweighted_sample(row_ptr, col, centers, M, 1234 + rank)
        assert np.array_equal(off.cpu().numpy(), eoff) and np.array_equal(gid.cpu().numpy(), egid)
        assert np.array_equal(dest.cpu().numpy(), edest)
    # fused multi-hop sampler over the striped graph
    sampler = wgth.MultiHopSampler()
    seeds = np.random.default_rng(10 + rank).permutation(nodes)[:300].astype(np.int64)
    lo = np.array([0, 100, 300], dtype=np.int64)
    got = sampler.sample(wm_rp, wm_col, torch.from_numpy(seeds).cuda(), torch.from_numpy(lo).cuda(), [5, 3], 99)
    exp = oracle.multihop_sample(row_ptr, col, seeds, lo, [5, 3], 99)
    for k in ("majors", "minors", "edge_id", "renumber_map", "label_hop_offsets"):
        assert np.array_equal(got[k].cpu().numpy(), exp[k]), k
    comm.barrier()


def _hot_rows_worker(rank, world, uid):
    """Replicated hot rows: gathers return exactly the table's rows whether a row is served by the replica, the local
    chunk or a peer; the replica is a snapshot until it is rebuilt."""
    torch = _setup(rank, world)
    wgth, comm = _comm(uid, rank, world)
    rows, dim = 20011, 96
    emb = wgth.create_embedding(comm, "chunked", "cuda", torch.float32, [rows, dim])
    local, start = emb.get_embedding_tensor().get_local_tensor()
    ids = torch.arange(start, start + local.shape[0], device="cuda")
    local.copy_(((ids[:, None] * 5 + torch.arange(dim, device="cuda")[None, :]) % 8191).float())
    torch.cuda.synchronize()
    comm.barrier()
    g = torch.Generator().manual_seed(3 + rank)
    idx = torch.randint(0, rows, (30000,), generator=g)
    idx[::13] = -1
    exp = ((idx[:, None] * 5 + torch.arange(dim)[None, :]) % 8191).float()
    keep = idx >= 0
    plain = emb.gather(idx.cuda()).cpu()
    hot = torch.randperm(rows, generator=torch.Generator().manual_seed(99))[:4000]  # rows of both ranks, same set on both
    emb.set_hot_rows(hot.cuda())
    assert emb.hot_row_count() == 4000
    got = emb.gather(idx.cuda()).cpu()
    assert torch.equal(got[keep], exp[keep]) and torch.equal(got[keep], plain[keep])
    got32 = emb.gather(idx.int().cuda()).cpu()  # int32 indices
    assert torch.equal(got32[keep], exp[keep])
    got16 = emb.gather(idx.cuda(), force_dtype=torch.float16).cpu()  # converting path ignores the replica, same values
    assert torch.equal(got16[keep], exp[keep].half())
    comm.barrier()
    # a scatter into the table drops the replica of the rank that scatters (it would serve stale rows otherwise); replicas
    # on other ranks are snapshots until their owners rebuild them
    if rank == 0:
        emb.get_embedding_tensor().scatter(torch.full((1, dim), -7.0).cuda(), hot[:1].cuda())
        torch.cuda.synchronize()
        assert emb.hot_row_count() == 0
    comm.barrier()
    one = emb.gather(hot[:1].cuda()).cpu()
    if rank == 0:
        assert float(one[0, 0]) == -7.0
    else:
        assert emb.hot_row_count() == 4000
        assert torch.equal(one, ((hot[:1, None] * 5 + torch.arange(dim)[None, :]) % 8191).float())
    emb.set_hot_rows(hot.cuda())
    assert float(emb.gather(hot[:1].cuda())[0, 0]) == -7.0
    emb.set_hot_rows(None)
    assert emb.hot_row_count() == 0 and float(emb.gather(hot[:1].cuda())[0, 0]) == -7.0
    comm.barrier()
    wgth.destroy_embedding(emb)


def test_hot_row_replica_two_ranks():
    _run(_hot_rows_worker)


def _optimizer_worker(rank, world, uid):
    """Trainable embedding on two ranks: every rank contributes gradients for rows of BOTH ranks (with repeats inside
    and across ranks); owners pull them from the peers' mailboxes.  Result = the oracle applied to the concatenation
    of the ranks' contributions in rank order."""
    torch = _setup(rank, world)
    wgth, comm = _comm(uid, rank, world)
    import wg_oracle as oracle

    rows, dim = 3001, 48
    table = np.random.default_rng(0).standard_normal((rows, dim)).astype(np.float32)
    for opt, params in (("sgd", {"weight_decay": 0.01}), ("adam", {"weight_decay": 0.01}), ("adagrad", {}), ("rmsprop", {"alpha": 0.9})):
        emb = wgth.create_embedding(comm, "chunked", "cuda", torch.float32, [rows, dim])
        local, start = emb.get_embedding_tensor().get_local_tensor()
        local.copy_(torch.from_numpy(table[start:start + local.shape[0]]).cuda())
        torch.cuda.synchronize()
        comm.barrier()
        optimizer = wgth.create_wholememory_optimizer(emb, opt, params)
        ref = table.copy()
        states = {"adam": {"m": np.zeros_like(ref), "v": np.zeros_like(ref), "beta12t": np.ones((rows, 2), np.float32)},
                  "adagrad": {"state_sum": np.zeros_like(ref)}, "rmsprop": {"v": np.zeros_like(ref)}, "sgd": {}}[opt]
        for step in range(3):
            per_rank = []
            for r in range(world):
                g = np.random.default_rng(100 * step + r)
                n = 700 + 300 * r  # uneven contributions; step 2: rank 0 contributes nothing
                if step == 2 and r == 0:
                    n = 0
                idx = g.integers(0, rows, n)
                idx[: n // 3] = g.integers(0, 25, n // 3)
                per_rank.append((idx, g.standard_normal((n, dim)).astype(np.float32)))
            idx, grads = per_rank[rank]
            emb.add_gradients(torch.from_numpy(idx).cuda(), torch.from_numpy(grads).cuda())
            emb.need_apply = True
            optimizer.step(0.03)
            oracle.embedding_gradient_apply(opt, params, ref, np.concatenate([p[0] for p in per_rank]),
                                            np.concatenate([p[1] for p in per_rank]), 0.03, states)
            got = local.cpu().numpy()
            np.testing.assert_allclose(got, ref[start:start + local.shape[0]], rtol=3e-5, atol=3e-6)
        comm.barrier()
        wgth.destroy_wholememory_optimizer(optimizer)
        wgth.destroy_embedding(emb)


def test_trainable_embedding_two_ranks():
    _run(_optimizer_worker)


def _file_io_worker(rank, world, uid, tmpdir):
    """(f3) two ranks, uneven parts: every rank stores its own part file (to_file_prefix), a tensor with ANOTHER partition
    is loaded from the two files (from_file_prefix: each rank reads exactly the byte range of its chunk out of the
    concatenation), and gathers over both chunks return the closed form."""
    torch = _setup(rank, world)
    wgth, comm = _comm(uid, rank, world)
    rows, dim, stride = 10_007, 48, 64
    part = [rows // 3, rows - rows // 3]
    t = wgth.create_wholememory_tensor(comm, "chunked", "cuda", [rows, dim], torch.float32, [stride, 1], part)
    local, start = t.get_local_tensor()
    ids = torch.arange(start, start + local.shape[0], device="cuda")
    local.copy_(((ids[:, None] * 11 + torch.arange(dim, device="cuda")[None, :]) % 2039).float())
    torch.cuda.synchronize()
    comm.barrier()
    prefix = os.path.join(tmpdir, "emb")
    t.to_file_prefix(prefix)
    comm.barrier()
    sizes = [os.path.getsize("%s_part_%d_of_2" % (prefix, r)) for r in range(2)]
    assert sizes == [part[0] * dim * 4, part[1] * dim * 4], sizes  # rows without their stride padding
    back = wgth.create_wholememory_tensor(comm, "chunked", "cuda", [rows, dim], torch.float32, [dim, 1])  # equal partition, no padding
    back.get_local_tensor()[0].fill_(-1.0)
    back.from_file_prefix(prefix)
    comm.barrier()
    idx = torch.randint(0, rows, (4000,), generator=torch.Generator().manual_seed(9))
    exp = ((idx[:, None] * 11 + torch.arange(dim)[None, :]) % 2039).float()
    assert torch.equal(back.gather(idx.cuda()).cpu(), exp), "rank %d: rows loaded from the part files differ" % rank
    lb, sb = back.get_local_tensor()
    want = ((torch.arange(sb, sb + lb.shape[0])[:, None] * 11 + torch.arange(dim)[None, :]) % 2039).float()
    assert torch.equal(lb.cpu(), want)
    comm.barrier()
    wgth.destroy_wholememory_tensor(t)
    wgth.destroy_wholememory_tensor(back)


def test_file_round_trip_two_ranks_uneven_parts(tmp_path):
    _run(_file_io_worker, tmpdir=str(tmp_path))


def _run(fn, **kw):
    import pylibwholegraph.binding.wholememory_binding as wmb
    from pylibwholegraph.utils.multiprocess import multiprocess_run

    uid = wmb.create_unique_id().get_bytes()
    multiprocess_run(2, functools.partial(fn, uid=uid, **kw))


@pytest.mark.parametrize("partition", ["equal", "uneven"])
def test_chunked_gather_scatter_two_ranks(partition):
    _run(_gather_scatter_worker, partition=partition)


def test_chunked_csr_sampling_two_ranks():
    _run(_sampler_worker)


# ---- cugraph_pyg loaders on two ranks (reference: tests/loader/test_neighbor_loader_mg.py:52-190) ---------------
def _pyg_worker(rank, world, port, case):
    torch = _setup(rank, world)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), LOCAL_WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    torch.distributed.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    from cugraph_pyg.data import GraphStore, FeatureStore
    from cugraph_pyg.loader import NeighborLoader
    from graphs import karate_csr

    if case == "karate":
        row_ptr, col = karate_csr(np.int64)
        dst = np.repeat(np.arange(34), np.diff(row_ptr))
        ei_all = torch.from_numpy(np.stack([col, dst]))  # PyG (source, destination)
        ei = torch.tensor_split(ei_all.clone(), world, dim=1)[rank].cuda()
        graph_store = GraphStore()
        graph_store.put_edge_index(ei, ("person", "knows", "person"), "coo", False, (34, 34))
        feat = (torch.arange(34)[:, None] * 3 + torch.arange(16)[None, :]).float()
        feature_store = FeatureStore()
        feature_store["person", "feat", None] = torch.tensor_split(feat, world)[rank]  # every rank puts its slice
        ix_train = torch.tensor_split(torch.arange(34), world)[rank]
        loader = NeighborLoader((feature_store, graph_store), [5, 5], input_nodes=ix_train, batch_size=8)
        have = set(zip(ei_all[0].tolist(), ei_all[1].tolist()))
        seen = []
        for batch in loader:
            assert torch.equal(batch.feat.cpu(), feat[batch.n_id.cpu()])  # rows of both ranks, read by P2P
            src, dst_ = batch.n_id.cpu()[batch.edge_index[0].cpu()], batch.n_id.cpu()[batch.edge_index[1].cpu()]
            assert all((int(s), int(d)) in have for s, d in zip(src, dst_))  # edges contributed by either rank
            assert int(batch.num_sampled_nodes.sum()) == batch.n_id.numel()
            seen += batch.n_id[: batch.batch_size].cpu().tolist()
        assert seen == ix_train.tolist()
        # global edge ids: this rank's partition starts after the lower ranks' edges
        assert len(loader) == (len(ix_train) + 7) // 8
    else:
        # reference: run_test_neighbor_loader_biased_mg -- the zero-bias edge of every rank is never sampled
        eix = torch.stack([torch.arange(3 * (world + rank), 3 * (world + rank + 1)), torch.arange(3 * rank, 3 * (rank + 1))]).cuda()
        graph_store = GraphStore()
        graph_store.put_edge_index(eix, ("person", "knows", "person"), "coo")
        feature_store = FeatureStore()
        feature_store["person", "feat", None] = torch.randint(128, (6 * world, 12))
        feature_store[("person", "knows", "person"), "bias", None] = torch.cat([torch.tensor([0, 1, 1], dtype=torch.float32) for _ in range(world)])
        loader = NeighborLoader((feature_store, graph_store), [1], input_nodes=torch.arange(3 * rank, 3 * (rank + 1)).cuda(),
                                batch_size=3, weight_attr="bias")
        out = list(iter(loader))
        assert len(out) == 1
        assert (out[0].edge_index.cpu() == torch.tensor([[3, 4], [1, 2]])).all()
    torch.distributed.barrier()
    torch.distributed.destroy_process_group()


@pytest.mark.parametrize("case", ["karate", "biased"])
def test_neighbor_loader_two_ranks(case):
    import socket
    import torch
    from pylibwholegraph.utils.multiprocess import multiprocess_run

    if torch.cuda.device_count() < 2:
        pytest.skip("NCCL needs one GPU per rank")
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    multiprocess_run(2, functools.partial(_pyg_worker, port=port, case=case))
