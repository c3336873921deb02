#!/bin/bash
# compute-sanitizer memcheck over the whole single-GPU suite (multi-rank test file excluded: it spawns its own processes), synccheck on the new kernels
out=gpurun_out
CS=/usr/local/cuda/bin/compute-sanitizer
timeout 1500 $CS --tool memcheck --error-exitcode 9 --print-limit 30 python -m pytest tests -m gpu -x -q --ignore=tests/test_gpu_multirank.py > $out/r2au_sanitizer_memcheck_full.log 2>&1
echo "== memcheck full rc=$?"; grep -E "ERROR SUMMARY|passed|failed|Invalid" $out/r2au_sanitizer_memcheck_full.log | tail -5
timeout 600 $CS --tool synccheck --error-exitcode 9 --print-limit 20 python -m pytest tests/test_gpu_multihop_fused.py tests/test_gpu_sage_tile.py tests/test_gpu_gather_bulk.py -m gpu -x -q -k "B9 or B100 or ragged or vs_fp64 or register_path" > $out/r2au_sanitizer_synccheck.log 2>&1
echo "== synccheck rc=$?"; grep -E "ERROR SUMMARY|passed|failed|Barrier|error" $out/r2au_sanitizer_synccheck.log | tail -5
