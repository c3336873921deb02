"""(f3) Binary on-disk formats: wholememory_store_to_file / wholememory_load_from_file and their Python wrappers
(reference: cpp/src/wholememory/file_io.cpp:1849-2165, python/pylibwholegraph/pylibwholegraph/torch/tensor.py:153-197),
and the file triple the reference's converter writes (python/pylibwholegraph/examples/ogbn_papers100m_convert.py:12-74:
node_feat.bin row-major, homograph_csr_row_ptr int64, homograph_csr_col_idx int32) loaded into WholeMemory and used by the
hot path (gather + multi-hop sampling) against the oracle."""
import os

import numpy as np
import pytest

from graphs import random_csr

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def env():
    import torch
    import pylibwholegraph.torch as wgth

    torch.cuda.set_device(0)
    wgth.init(0, 1, 0, 1)
    return wgth, wgth.get_global_communicator()


@pytest.mark.parametrize("dtype_name,dim,stride", [("float32", 128, 128), ("float32", 100, 128), ("float16", 37, 40), ("int64", 1, 1), ("int32", 0, 0)])
def test_to_file_prefix_from_file_prefix_round_trip(env, tmp_path, dtype_name, dim, stride):
    import torch

    wgth, comm = env
    dtype = getattr(torch, dtype_name)
    rows = 50_003
    sizes, strides = ([rows, dim], [stride, 1]) if dim > 0 else ([rows], [1])
    t = wgth.create_wholememory_tensor(comm, "chunked", "cuda", sizes, dtype, strides)
    local, start = t.get_local_tensor()
    assert start == 0 and local.shape[0] == rows
    g = torch.Generator().manual_seed(3)
    vals = torch.randint(-1000, 1000, tuple(sizes), generator=g).to(dtype)
    local.copy_(vals.cuda())
    prefix = str(tmp_path / "tensor")
    t.to_file_prefix(prefix)
    part = prefix + "_part_0_of_1"
    assert os.path.exists(part)
    # the file holds the rows WITHOUT the stride padding, row-major: numpy reads it back directly
    elt = torch.tensor([], dtype=dtype).element_size()
    assert os.path.getsize(part) == rows * max(dim, 1) * elt
    raw = np.fromfile(part, dtype=np.dtype(dtype_name)).reshape(tuple(sizes))
    assert np.array_equal(raw, vals.numpy())
    back = wgth.create_wholememory_tensor(comm, "chunked", "cuda", sizes, dtype, strides)
    back.get_local_tensor()[0].fill_(0)
    back.from_file_prefix(prefix)
    assert torch.equal(back.get_local_tensor()[0].cpu(), vals)
    wgth.destroy_wholememory_tensor(t)
    wgth.destroy_wholememory_tensor(back)


def test_from_filelist_concatenates_uneven_files_and_sizes_itself(env, tmp_path):
    import torch

    wgth, comm = env
    dim = 24
    parts = [np.random.default_rng(k).standard_normal((n, dim)).astype(np.float32) for k, n in enumerate((1000, 1, 7777, 0, 312))]
    files = []
    for k, a in enumerate(parts):
        f = str(tmp_path / ("feat_%d.bin" % k))
        a.tofile(f)
        files.append(f)
    t = wgth.create_wholememory_tensor_from_filelist(comm, "chunked", "cuda", files, torch.float32, dim)
    full = np.concatenate(parts)
    assert tuple(t.shape) == full.shape
    assert np.array_equal(t.get_local_tensor()[0].cpu().numpy(), full)
    idx = torch.randint(0, full.shape[0], (5000,), generator=torch.Generator().manual_seed(1))
    assert np.array_equal(t.gather(idx.cuda()).cpu().numpy(), full[idx.numpy()])
    wgth.destroy_wholememory_tensor(t)
    with pytest.raises(Exception):  # a file that is not a whole number of rows
        np.zeros(dim + 1, np.float32).tofile(str(tmp_path / "bad.bin"))
        wgth.create_wholememory_tensor_from_filelist(comm, "chunked", "cuda", [str(tmp_path / "bad.bin")], torch.float32, dim)


def test_converter_triple_feeds_the_hot_path(env, tmp_path, oracle):
    """what ogbn_papers100m_convert.py writes -> WholeMemory tensors -> gather and fused sampling == oracle on the arrays"""
    import torch

    wgth, comm = env
    nodes, edges, dim = 30_011, 400_000, 128
    row_ptr, col = random_csr(nodes, edges, seed=12)
    feat = np.random.default_rng(5).standard_normal((nodes, dim)).astype(np.float32)
    feat.tofile(str(tmp_path / "node_feat.bin"))                         # :46-47
    row_ptr.astype("int64").tofile(str(tmp_path / "homograph_csr_row_ptr"))   # :69-72
    col.astype("int32").tofile(str(tmp_path / "homograph_csr_col_idx"))
    wm_feat = wgth.create_wholememory_tensor_from_filelist(comm, "chunked", "cuda", str(tmp_path / "node_feat.bin"), torch.float32, dim)
    wm_rp = wgth.create_wholememory_tensor_from_filelist(comm, "chunked", "cuda", str(tmp_path / "homograph_csr_row_ptr"), torch.int64, 0)
    wm_col = wgth.create_wholememory_tensor_from_filelist(comm, "chunked", "cuda", str(tmp_path / "homograph_csr_col_idx"), torch.int32, 0)
    assert tuple(wm_feat.shape) == (nodes, dim) and tuple(wm_rp.shape) == (nodes + 1,) and tuple(wm_col.shape) == (edges,)
    seeds = np.random.default_rng(0).permutation(nodes)[:512].astype(np.int64)
    lo = np.array([0, 128, 512], dtype=np.int64)
    got = wgth.MultiHopSampler().sample(wm_rp, wm_col, torch.from_numpy(seeds).cuda(), torch.from_numpy(lo).cuda(), [10, 5], 62)
    exp = oracle.multihop_sample(row_ptr, col, seeds, lo, [10, 5], 62)
    for k in ("majors", "minors", "edge_id", "renumber_map", "renumber_map_offsets", "label_hop_offsets"):
        assert np.array_equal(got[k].cpu().numpy(), exp[k]), k
    x = wm_feat.gather(got["renumber_map"])
    assert np.array_equal(x.cpu().numpy(), feat[exp["renumber_map"]])
    for t in (wm_feat, wm_rp, wm_col):
        wgth.destroy_wholememory_tensor(t)
