"""TEST INFRASTRUCTURE ONLY: builds the CPU emulation libraries under tests/emu/_build (git-ignored) with g++.

    libtemporal_emu.so   device kernels of csrc/temporal_device.cuh / sample_device.cuh behind small C drivers
    libmultihop_emu.so   the product's csrc/multihop.cu (host code + kernels, launches rewritten by emu_preprocess.py),
                         the runtime stand-ins of emu_runtime.cpp and the C driver multihop_driver.cpp
"""
import os
import subprocess

from emu_preprocess import rewrite_launches

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
CSRC = os.path.join(ROOT, "cugraph-gnn_b200", "csrc")
ASAN = os.environ.get("WGB_EMU_ASAN") == "1"  # tests/emu/run_asan.sh: same tests, libraries built with -fsanitize=address
OUT = os.path.join(HERE, "_build_asan" if ASAN else "_build")
CUDA_INC = os.environ.get("CUDA_INC", "/usr/local/cuda/include")
# -Bsymbolic: the stand-ins for the CUDA runtime must win over a real libcudart that torch may have loaded into the process
FLAGS = ["g++", "-O2", "-std=c++17", "-w", "-DWGB_HOST_EMULATION", "-DWGB_BUILDING_LIB", "-fPIC", "-shared", "-Wl,-Bsymbolic", "-I", HERE, "-I", CSRC,
         "-I", os.path.join(ROOT, "include"), "-I", CUDA_INC]
if ASAN:
    FLAGS = [f for f in FLAGS if f != "-O2"] + ["-O1", "-g", "-fsanitize=address", "-fno-omit-frame-pointer"]


def available() -> bool:
    return os.path.exists(os.path.join(CUDA_INC, "cuda_runtime.h"))


def _stale(target, deps):
    return not os.path.exists(target) or any(os.path.getmtime(d) > os.path.getmtime(target) for d in deps)


def _headers():
    return [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cuh")] + [os.path.join(HERE, "cuda_emu.h"), __file__]


def build_kernels() -> str:
    os.makedirs(OUT, exist_ok=True)
    lib = os.path.join(OUT, "libtemporal_emu.so")
    src = os.path.join(HERE, "temporal_emu.cpp")
    if _stale(lib, [src] + _headers()):
        subprocess.check_call(FLAGS + [src, "-o", lib])
    return lib


def build_multihop() -> str:
    os.makedirs(OUT, exist_ok=True)
    lib = os.path.join(OUT, "libmultihop_emu.so")
    cu = os.path.join(CSRC, "multihop.cu")
    srcs = [os.path.join(HERE, "emu_runtime.cpp"), os.path.join(HERE, "multihop_driver.cpp")]
    if _stale(lib, [cu, os.path.join(HERE, "emu_preprocess.py")] + srcs + _headers()):
        text, n = rewrite_launches(open(cu).read())
        assert n > 0
        gen = os.path.join(OUT, "multihop_emu.cpp")
        with open(gen, "w") as f:
            f.write(text)
        subprocess.check_call(FLAGS + ["-include", "cuda_emu.h", gen] + srcs + ["-o", lib])
    return lib


def build_rows() -> str:
    """gather / scatter (gather_scatter.cu), CSR aggregation (aggregate.cu) and append-unique (append_unique.cu) behind rows_driver.cpp"""
    os.makedirs(OUT, exist_ok=True)
    lib = os.path.join(OUT, "librows_emu.so")
    cus = [os.path.join(CSRC, "gather_scatter.cu"), os.path.join(CSRC, "aggregate.cu"), os.path.join(CSRC, "append_unique.cu")]
    srcs = [os.path.join(HERE, "emu_runtime.cpp"), os.path.join(HERE, "rows_driver.cpp")]
    if _stale(lib, cus + [os.path.join(HERE, "emu_preprocess.py")] + srcs + _headers()):
        gens = []
        for cu in cus:
            text, n = rewrite_launches(open(cu).read())
            assert n > 0
            gen = os.path.join(OUT, os.path.basename(cu)[:-3] + "_emu.cpp")
            with open(gen, "w") as f:
                f.write(text)
            gens.append(gen)
        subprocess.check_call(FLAGS + ["-include", "cuda_emu.h"] + gens + srcs + ["-o", lib])
    return lib


def build_onehop() -> str:
    """the one-hop samplers S1 / S2 (sample.cu, host code and kernels) behind onehop_driver.cpp"""
    os.makedirs(OUT, exist_ok=True)
    lib = os.path.join(OUT, "libonehop_emu.so")
    cu = os.path.join(CSRC, "sample.cu")
    srcs = [os.path.join(HERE, "emu_runtime.cpp"), os.path.join(HERE, "onehop_driver.cpp")]
    if _stale(lib, [cu, os.path.join(HERE, "emu_preprocess.py")] + srcs + _headers()):
        text, n = rewrite_launches(open(cu).read())
        assert n > 0
        gen = os.path.join(OUT, "sample_emu.cpp")
        with open(gen, "w") as f:
            f.write(text)
        subprocess.check_call(FLAGS + ["-DEMU_PRODUCT_SKIP_TABLE", "-include", "cuda_emu.h", gen] + srcs + ["-o", lib])
    return lib
