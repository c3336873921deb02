"""The drop-in boundary without Python in the loop: tests/c_client/abi_client.cpp is a plain C++ program that includes the
reference-compatible headers of include/wholememory, links libwholegraph_b200.so and libcudart, and drives gather, the one-hop
sampler and the allocation callbacks with cudaMalloc'd buffers -- what the reference's own C++ callers (and its Cython module)
do (cpp/include/wholememory/*.h).  CPU: it compiles, links and its host-only calls run.  GPU: it runs the kernels and checks
the results against closed forms."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "c_client", "abi_client.cpp")
LIBDIR = os.path.join(ROOT, "cugraph-gnn_b200", "lib")
CUDA = os.environ.get("CUDA_HOME", "/usr/local/cuda")


def _build(tmp_path):
    if shutil.which("g++") is None or not os.path.exists(os.path.join(CUDA, "include", "cuda_runtime.h")):
        pytest.skip("g++ or the CUDA headers are not installed")
    assert os.path.exists(os.path.join(LIBDIR, "libwholegraph_b200.so")), "build the library first (python cugraph-gnn_b200/build.py)"
    exe = str(tmp_path / "abi_client")
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-I", os.path.join(ROOT, "include"), "-I", os.path.join(CUDA, "include"), SRC, "-o", exe,
                           "-L", LIBDIR, "-lwholegraph_b200", "-L", os.path.join(CUDA, "lib64"), "-lcudart", "-Wl,-rpath," + LIBDIR,
                           "-Wl,-rpath," + os.path.join(CUDA, "lib64")])
    return exe


def test_cpp_client_compiles_links_and_runs_host_calls(tmp_path):
    exe = _build(tmp_path)
    out = subprocess.run([exe, "link"], capture_output=True, text=True, timeout=60)
    assert out.returncode == 0 and "abi_client link ok" in out.stdout, out.stdout + out.stderr


@pytest.mark.gpu
def test_cpp_client_runs_gather_and_sampler_on_the_gpu(tmp_path):
    exe = _build(tmp_path)
    out = subprocess.run([exe, "gpu"], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and "abi_client gpu ok" in out.stdout, out.stdout + out.stderr
