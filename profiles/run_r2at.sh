#!/bin/bash
# compute-sanitizer over the round's new kernels: memcheck (out-of-bounds / misaligned global + shared accesses) and racecheck
# (shared-memory hazards) on a subset of the parity tests
out=gpurun_out
CS=/usr/local/cuda/bin/compute-sanitizer
SEL_FUSED='tests/test_gpu_multihop_fused.py::test_fused_equals_chain_equals_oracle[25x10-B9] tests/test_gpu_multihop_fused.py::test_fused_equals_chain_equals_oracle[25x10-B100] tests/test_gpu_multihop_fused.py::test_fused_equals_chain_equals_oracle[15x10x5-ragged] tests/test_gpu_multihop_fused.py::test_fused_dtypes_edge_ids_csr_int64_ids[int64-int64]'
SEL_BULK='tests/test_gpu_gather_bulk.py::test_bulk_gather_equals_register_path[int32-128-float32] tests/test_gpu_gather_bulk.py::test_bulk_gather_equals_register_path[int64-32-float32] tests/test_gpu_gather_bulk.py::test_bulk_gather_equals_register_path[int32-127-float32]'
SEL_SAGE='tests/test_gpu_sage_tile.py::test_sage_tile_vs_fp64[48-129] tests/test_gpu_sage_tile.py::test_sage_tile_vs_fp64[256-3001] tests/test_gpu_sage_tile.py::test_sage_tile_index_types'
for tool in memcheck racecheck; do
  timeout 900 $CS --tool $tool --error-exitcode 9 --print-limit 20 python -m pytest $SEL_FUSED $SEL_BULK $SEL_SAGE -m gpu -x -q > $out/r2at_sanitizer_$tool.log 2>&1
  echo "== $tool rc=$?"; grep -E "ERROR SUMMARY|passed|failed|hazard|Invalid|error" $out/r2at_sanitizer_$tool.log | tail -6
done
