#!/bin/bash
# TEST INFRASTRUCTURE ONLY: the emulated sampler tests under AddressSanitizer (out-of-bounds reads and writes of kernels and
# host code on "device" memory).  Usage: tests/emu/run_asan.sh [pytest args]      (from the repository root)
set -e
export WGB_EMU_ASAN=1
export ASAN_OPTIONS=detect_leaks=0:detect_stack_use_after_return=0
LD_PRELOAD=$(gcc -print-file-name=libasan.so) python -m pytest tests/test_emulated_multihop_cpu.py tests/test_emulated_kernels_cpu.py tests/test_emulated_rows_cpu.py tests/test_emulated_onehop_cpu.py -x -q -p no:cacheprovider "$@"
