#!/bin/bash
out=gpurun_out
timeout 600 python -m pytest tests/test_gpu_multihop_fused.py tests/test_gpu_parity.py tests/test_gpu_gather_bulk.py tests/test_gpu_multirank.py -m gpu -x -q 2>&1 | tail -2
timeout 300 python profiles/overlap_probe.py c4 10 2>&1 | tail -1
WGB_MH_CARVEOUT=0 timeout 300 python profiles/overlap_probe.py c4 10 2>&1 | tail -1
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-parity-check > $out/r2j_bench_c4.json 2> $out/r2j_bench_c4.err
grep -o '"value": [0-9.e+]*\|"ms_per_step": [0-9.e+]*\|"sample_renumber_ms_per_step": [0-9.e+]*\|"gather_ms_per_step": [0-9.e+]*\|"gather_alone_ms_per_step": [0-9.e+]*' $out/r2j_bench_c4.json | head -6
