"""No-GPU checks of the C-ABI library: it loads, exports every symbol include/*.h declares, the
host-only entry points (descriptors, RNG twins, error paths) behave like the reference's."""
import ctypes
import glob
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def wmb():
    import pylibwholegraph.binding.wholememory_binding as wmb

    return wmb


def _declared_functions():
    names = set()
    pat = re.compile(r"^\s*(?:[A-Za-z_][\w\s\*]*?)\b([a-z_][a-z0-9_]*)\s*\($", re.M)
    for h in glob.glob(os.path.join(ROOT, "include", "wholememory", "*.h")):
        text = open(h).read()
        text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
        text = re.sub(r"#define.*?(?<!\\)\n", "\n", text, flags=re.S)
        # declarations look like: <ret type> name(<args>);  possibly over several lines
        for m in re.finditer(r"\b([a-z_][a-z0-9_]*)\s*\(", text):
            name = m.group(1)
            tail = text[m.end():]
            depth, i = 1, 0
            while depth and i < len(tail):
                depth += tail[i] == "("
                depth -= tail[i] == ")"
                i += 1
            after = tail[i:i + 3].strip()
            before = text[max(0, m.start() - 80):m.start()]
            if after.startswith(";") and "typedef" not in before.split(";")[-1] and "(*" not in before.split(";")[-1]:
                names.add(name)
    return sorted(n for n in names if n not in ("defined", "sizeof", "fprintf"))


def test_library_exports_every_declared_symbol(wmb):
    lib = ctypes.CDLL(wmb.LIBRARY_PATH)
    declared = _declared_functions()
    assert len(declared) > 60, declared
    missing = [n for n in declared if not hasattr(lib, n)]
    assert not missing, missing


def test_hot_path_entry_points_present(wmb):
    lib = ctypes.CDLL(wmb.LIBRARY_PATH)
    for name in (
        "wholememory_gather", "wholememory_scatter", "wholememory_embedding_gather",
        "wholegraph_csr_unweighted_sample_without_replacement",
        "wholegraph_csr_weighted_sample_without_replacement", "graph_append_unique",
        "wholememory_create_tensor", "wholememory_make_tensor_from_pointer", "wholememory_tensor_get_global_reference",
        "wholememory_malloc", "wholememory_create_communicator",
    ):
        assert hasattr(lib, name), name


def test_dtype_helpers(wmb):
    lib = ctypes.CDLL(wmb.LIBRARY_PATH)
    lib.wholememory_dtype_get_element_size.restype = ctypes.c_size_t
    sizes = {1: 4, 2: 2, 3: 8, 4: 2, 5: 4, 6: 8, 7: 2, 8: 1}
    for dt, s in sizes.items():
        assert lib.wholememory_dtype_get_element_size(dt) == s
    assert lib.wholememory_dtype_get_element_size(0) == ctypes.c_size_t(-1).value
    lib.wholememory_dtype_is_floating_number.restype = ctypes.c_bool
    lib.wholememory_dtype_is_integer_number.restype = ctypes.c_bool
    assert [lib.wholememory_dtype_is_floating_number(d) for d in range(1, 9)] == [True] * 4 + [False] * 4
    assert [lib.wholememory_dtype_is_integer_number(d) for d in range(1, 9)] == [False] * 4 + [True] * 4


def test_equal_partition_plan(wmb):
    # reference: cpp/src/wholememory/memory_handle.cpp:2116-2122 -> ceil(N / W)
    assert wmb.equal_partition_plan(10, 4) == 3
    assert wmb.equal_partition_plan(8, 4) == 2
    assert wmb.equal_partition_plan(1024 * 256 * 8 + 3, 8) == 1024 * 256 + 1


def test_host_random_twins_match_oracle(wmb, oracle):
    """the library's host twin of the device stream == the oracle's restatement."""
    import torch
    from pylibwholegraph.torch import wholegraph_ops

    for seed, sub in [(62, 0), (62, 31), (12345678901234, 987654321), (0, 1 << 40)]:
        got = wholegraph_ops.generate_random_positive_int_cpu(seed, sub, 9).numpy()
        assert got.tolist() == oracle.generate_random_positive_int(seed, sub, 9).tolist()
        gf = wholegraph_ops.generate_exponential_distribution_negative_float_cpu(seed, sub, 9).numpy()
        assert np.array_equal(gf, oracle.generate_exponential_distribution_negative_float(seed, sub, 9))
    assert torch.int32 == wholegraph_ops.generate_random_positive_int_cpu(1, 1, 1).dtype


def test_error_codes_map_to_exceptions(wmb):
    # .pyx:241-263: InvalidInput -> ValueError, OutOfMemory -> MemoryError, NotImplemented -> NotImplementedError
    with pytest.raises(ValueError):
        wmb.check_wholememory_error_code(wmb.WholeMemoryErrorCode.InvalidInput)
    with pytest.raises(MemoryError):
        wmb.check_wholememory_error_code(wmb.WholeMemoryErrorCode.OutOfMemory)
    with pytest.raises(NotImplementedError):
        wmb.check_wholememory_error_code(wmb.WholeMemoryErrorCode.NotImplemented)
    with pytest.raises(RuntimeError):
        wmb.check_wholememory_error_code(wmb.WholeMemoryErrorCode.CUDAError)
    wmb.check_wholememory_error_code(0)


def test_wrapped_tensor_descriptor_roundtrip(wmb):
    import torch
    from pylibwholegraph.torch.wholegraph_env import wrap_torch_tensor

    before = wmb.py_get_wholememory_tensor_count()
    t = torch.zeros((7, 33), dtype=torch.float16)[:, :32]
    w = wrap_torch_tensor(t)
    view = wmb.PyWholeMemoryTensor(w.get_c_handle(), owner=False)
    assert view.shape == (7, 32) and view.stride() == (33, 1) and view.dtype == wmb.WholeMemoryDataType.DtHalf
    assert wmb.py_get_wholememory_tensor_count() == before + 1
    del w
    assert wmb.py_get_wholememory_tensor_count() == before
    # last stride must be 1 (reference: wholememory_tensor.cpp:112-156)
    with pytest.raises(ValueError):
        wrap_torch_tensor(torch.zeros((4, 4)).t())


def test_communicator_rendezvous_two_processes(wmb):
    """world_size-2 bootstrap over shared memory, no GPU involved (barrier + rank/size)."""
    from pylibwholegraph.utils.multiprocess import multiprocess_run
    import functools

    uid = wmb.create_unique_id().get_bytes()
    multiprocess_run(2, functools.partial(_comm_worker, uid=uid))


def _comm_worker(rank, world, uid):
    import sys

    sys.path.insert(0, os.path.join(ROOT, "cugraph-gnn_b200"))
    import pylibwholegraph.binding.wholememory_binding as wmb

    comm = wmb.create_communicator(wmb.PyWholeMemoryUniqueID(uid), rank, world)
    assert comm.get_rank() == rank and comm.get_size() == world
    for _ in range(50):
        comm.barrier()
    wmb.destroy_communicator(comm)


def test_example_support_modules_options_and_launch(monkeypatch):
    """common_options / distributed_launch keep the reference's flags and environment contract (no GPU needed)."""
    import argparse

    import pylibwholegraph.torch as wgth

    p = argparse.ArgumentParser()
    for f in (wgth.add_training_options, wgth.add_common_graph_options, wgth.add_common_model_options, wgth.add_common_sampler_options,
              wgth.add_node_classfication_options, wgth.add_dataloader_options, wgth.add_distributed_launch_options):
        f(p)
    a = p.parse_args(["--launch-agent", "pytorch", "-n", "10,5", "-l", "2", "--train-embedding"])
    assert (a.epochs, a.batchsize, a.hiddensize, a.model, a.framework, a.classnum, a.cache_type) == (24, 1024, 256, "sage", "wg", 172, "none")
    assert wgth.parse_max_neighbors(2, a.neighbors) == [10, 5] and wgth.parse_max_neighbors(3, "30") == [30, 30, 30]
    for k, v in (("RANK", "3"), ("WORLD_SIZE", "8"), ("LOCAL_RANK", "3"), ("LOCAL_WORLD_SIZE", "8"), ("MASTER_ADDR", "127.0.0.1"), ("MASTER_PORT", "29999")):
        monkeypatch.setenv(k, v)
    seen = {}
    wgth.distributed_launch(a, lambda: seen.update(rank=wgth.get_rank(), world=wgth.get_world_size(), local=wgth.get_local_rank(), lsize=wgth.get_local_size()))
    assert seen == {"rank": 3, "world": 8, "local": 3, "lsize": 8}


def test_missing_library_fails_loudly_and_product_never_imports_the_oracle():
    """No CPU fallback: if the shared library cannot be loaded, importing the binding raises; and no product module
    (package, bench's own arm, entry points) reaches into oracle/."""
    import subprocess
    import sys

    pkg = os.path.join(ROOT, "cugraph-gnn_b200")
    code = ("import ctypes, sys\n"
            "def boom(*a, **k):\n    raise OSError('libwholegraph_b200.so: cannot open shared object file')\n"
            "ctypes.CDLL = boom\n"
            "sys.path.insert(0, %r)\n"
            "import pylibwholegraph.binding.wholememory_binding\n" % pkg)
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True)
    assert r.returncode != 0 and "ImportError" in r.stderr and "could not be loaded" in r.stderr
    offenders = []
    for d, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                text = open(os.path.join(d, f)).read()
                if re.search(r"^\s*(import|from)\s+wg_oracle|oracle[/\\.]wg_oracle|sys\.path.*oracle", text, re.M):
                    offenders.append(os.path.join(d, f))
    assert not offenders, offenders
