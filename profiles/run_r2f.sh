#!/bin/bash
tag=${1:-r2f}
out=gpurun_out
mkdir -p $out
timeout 600 python -m pytest tests/test_gpu_gather_bulk.py tests/test_gpu_multihop_fused.py -m gpu -x -q > $out/${tag}_tests.log 2>&1
tail -3 $out/${tag}_tests.log
for bulk in 1 0; do
  WGB_GATHER_BULK=$bulk timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-parity-check > $out/${tag}_bench_c4_bulk${bulk}.json 2> $out/${tag}_bench_c4_bulk${bulk}.err
  echo "bulk=$bulk"; grep -o '"value": [0-9.e+]*\|"ms_per_step": [0-9.e+]*\|"sample_renumber_ms_per_step": [0-9.e+]*\|"gather_ms_per_step": [0-9.e+]*\|"gather_alone_ms_per_step": [0-9.e+]*' $out/${tag}_bench_c4_bulk${bulk}.json | head -6
done
