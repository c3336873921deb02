#!/bin/bash
out=gpurun_out
timeout 600 python -m pytest tests/test_gpu_file_io.py tests/test_gpu_multirank.py -m gpu -x -q 2>&1 | tail -4
timeout 600 python bench.py --loader --workload c2 --steps 10 --warmup 3 > $out/r2m_loader_c2.json 2> $out/r2m_loader_c2.err
tail -3 $out/r2m_loader_c2.err; cat $out/r2m_loader_c2.json | cut -c1-900
timeout 600 python bench.py --workload c2 --steps 20 --warmup 3 --no-cpu-baseline > $out/r2m_bench_c2.json 2> $out/r2m_bench_c2.err
grep -o '"value": [0-9.e+]*\|"ms_per_step": [0-9.e+]*\|"sample_renumber_ms_per_step": [0-9.e+]*\|"gather_ms_per_step": [0-9.e+]*' $out/r2m_bench_c2.json | head -5
