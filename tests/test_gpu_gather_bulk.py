"""G1 through the copy engine (rows_bulk_gather_kernel: cp.async.bulk global -> shared ring -> global) against the register
path (WGB_GATHER_BULK=0) and the closed form: every byte equal, for every shape class the dispatcher sends there (16-byte
aligned rows up to 2 KB, any float / integer dtype, int32 / int64 indices, negative indices skipped, ragged last tile,
strided table and strided output) and for the shapes it must leave to the register path."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def env():
    import torch
    import pylibwholegraph.torch as wgth

    torch.cuda.set_device(0)
    wgth.init(0, 1, 0, 1)
    return wgth, wgth.get_global_communicator()


class _bulk(object):
    def __init__(self, on):
        self.on = on

    def __enter__(self):
        self.old = os.environ.get("WGB_GATHER_BULK")
        os.environ["WGB_GATHER_BULK"] = "1" if self.on else "0"

    def __exit__(self, *a):
        if self.old is None:
            os.environ.pop("WGB_GATHER_BULK", None)
        else:
            os.environ["WGB_GATHER_BULK"] = self.old


def _gather_into(t, idx, out):
    """wholememory_gather into a caller-provided (possibly strided) output, as the C ABI allows"""
    import pylibwholegraph.binding.wholememory_binding as wmb
    from pylibwholegraph.torch.wholegraph_env import get_stream, get_wholegraph_env_fns, wrap_torch_tensor

    wmb.wholememory_gather_op(t.wmb_tensor, wrap_torch_tensor(idx), wrap_torch_tensor(out), get_wholegraph_env_fns(), get_stream())


def _table(torch, wgth, comm, rows, dim, dtype, stride):
    t = wgth.create_wholememory_tensor(comm, "chunked", "cuda", [rows, dim], dtype, [stride, 1])
    local, _ = t.get_local_tensor()
    vals = (torch.arange(rows, device="cuda")[:, None] * 3 + torch.arange(dim, device="cuda")[None, :]) % 120
    local.copy_(vals.to(dtype))
    return t, vals.to(dtype)


@pytest.mark.parametrize("dim,dtype_name", [(4, "float32"), (32, "float32"), (128, "float32"), (256, "float32"), (512, "float32"),
                                            (516, "float32"), (127, "float32"), (128, "float16"), (64, "bfloat16"), (256, "int8"),
                                            (16, "int64"), (24, "float64")])
@pytest.mark.parametrize("idx_dtype", ["int32", "int64"])
def test_bulk_gather_equals_register_path(env, dim, dtype_name, idx_dtype):
    import torch

    wgth, comm = env
    dtype = getattr(torch, dtype_name)
    rows, n = 200_003, 100_001
    t, vals = _table(torch, wgth, comm, rows, dim, dtype, dim)
    g = torch.Generator().manual_seed(dim)
    idx = torch.randint(0, rows, (n,), generator=g)
    idx[::17] = -1  # skipped rows: output untouched
    idx = idx.to(getattr(torch, idx_dtype)).cuda()
    outs = []
    for on in (True, False):
        with _bulk(on):
            out = torch.full((n, dim), 7, dtype=dtype, device="cuda")
            _gather_into(t, idx, out)
            outs.append(out)
    assert torch.equal(outs[0].view(torch.uint8), outs[1].view(torch.uint8)), "bulk and register paths differ"
    keep = idx >= 0
    assert torch.equal(outs[0][keep], vals[idx[keep].long()])
    assert bool((outs[0][~keep] == 7).all())
    wgth.destroy_wholememory_tensor(t)


def test_bulk_gather_all_valid_contiguous_tiles_and_small_calls(env):
    """no negative index: whole tiles leave with one bulk store; calls below the bulk threshold take the register path"""
    import torch

    wgth, comm = env
    rows, dim = 150_000, 128
    t, vals = _table(torch, wgth, comm, rows, dim, torch.float32, dim)
    for n in (4096, 65_536, 65_537, 300_000, 100, 1):
        idx = torch.randint(0, rows, (n,), generator=torch.Generator().manual_seed(n)).cuda()
        with _bulk(True):
            out = t.gather(idx)
        assert torch.equal(out, vals[idx])
    wgth.destroy_wholememory_tensor(t)


def test_bulk_gather_strided_table_and_strided_output(env):
    import torch

    wgth, comm = env
    rows, dim, tstride, ostride = 50_000, 96, 128, 160  # all multiples of 16 bytes
    t, vals = _table(torch, wgth, comm, rows, dim, torch.float32, tstride)
    n = 40_000
    idx = torch.randint(0, rows, (n,), generator=torch.Generator().manual_seed(1)).cuda()
    for on in (True, False):
        with _bulk(on):
            buf = torch.full((n, ostride), -1.0, device="cuda")
            out = buf[:, 8:8 + dim]  # storage offset 32 bytes, row stride 640 bytes
            _gather_into(t, idx, out)
            assert torch.equal(out, vals[idx]) and bool((buf[:, :8] == -1).all()) and bool((buf[:, 8 + dim:] == -1).all())
    wgth.destroy_wholememory_tensor(t)
