#!/bin/bash
# sampler with the prefetched row extents; register vs bulk (TMA) gather at 148 labels per call group
out=gpurun_out
timeout 900 python -m pytest tests/test_gpu_multihop_fused.py tests/test_gpu_multihop.py -m gpu -x -q 2>&1 | tail -3
timeout 300 python profiles/overlap_probe.py c4 10 64,148 2>&1 | tail -2
for bulk in 0 1; do
  WGB_GATHER_BULK=$bulk timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-parity-check > $out/r2s_bench_c4_bulk$bulk.json 2> $out/r2s_bench_c4_bulk$bulk.err
  echo "== bulk=$bulk: $(grep -o '"value": [0-9.e+]*\|"ms_per_step": [0-9.e+]*\|"sample_renumber_ms_per_step": [0-9.e+]*\|"gather_ms_per_step": [0-9.e+]*\|"gather_alone_ms_per_step": [0-9.e+]*' $out/r2s_bench_c4_bulk$bulk.json | tr '\n' ' ')"
done
