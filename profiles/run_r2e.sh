#!/bin/bash
# bulk-async (TMA) gather on hardware for the first time + co-residency with the fused sampler
tag=${1:-r2e}
out=gpurun_out
mkdir -p $out
timeout 600 python -m pytest tests/test_gpu_gather_bulk.py -m gpu -x -q > $out/${tag}_bulk_tests.log 2>&1
echo "bulk tests exit code $?" >> $out/${tag}_bulk_tests.log
tail -12 $out/${tag}_bulk_tests.log
timeout 900 python -m pytest tests -m gpu -x -q > $out/${tag}_gpu_tests.log 2>&1
tail -5 $out/${tag}_gpu_tests.log
for bulk in 1 0; do
  WGB_GATHER_BULK=$bulk timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-parity-check > $out/${tag}_bench_c4_bulk${bulk}.json 2> $out/${tag}_bench_c4_bulk${bulk}.err
  echo "bulk=$bulk"; grep -o '"value": [0-9.e+]*\|"ms_per_step": [0-9.e+]*\|"sample_renumber_ms_per_step": [0-9.e+]*\|"gather_ms_per_step": [0-9.e+]*\|"gather_alone_ms_per_step": [0-9.e+]*' $out/${tag}_bench_c4_bulk${bulk}.json | head -6
  tail -2 $out/${tag}_bench_c4_bulk${bulk}.err
done
WGB_MH_TIMING=1 timeout 300 python profiles/prof_step.py 15 64 c4 > $out/${tag}_stage_times.txt 2>&1
tail -24 $out/${tag}_stage_times.txt
