#!/bin/bash
out=gpurun_out
BENCH_E2E_DEBUG=1 timeout 900 python bench.py --workload headline --steps 10 --warmup 3 --no-cpu-baseline --no-parity-check > $out/r2aw_bench_headline.json 2> $out/r2aw_bench_headline.err
echo "== headline: $(grep -o '"value": [0-9.e+]*\|"ms_per_step": [0-9.e+]*\|"gather_alone_ms_per_step": [0-9.e+]*' $out/r2aw_bench_headline.json | tr '\n' ' ')"
grep -E "e2e device phases|e2e [0-9.]+ ms" $out/r2aw_bench_headline.err | cut -c1-600
grep "step [0-9] begin" $out/r2aw_bench_headline.err | tail -10
timeout 900 python bench.py --workload c2 --steps 10 --warmup 3 --no-cpu-baseline --no-parity-check > $out/r2aw_bench_c2.json 2> $out/r2aw_bench_c2.err
echo "== c2: $(grep -o '"value": [0-9.e+]*\|"ms_per_step": [0-9.e+]*\|"gather_alone_ms_per_step": [0-9.e+]*' $out/r2aw_bench_c2.json | tr '\n' ' ')"
