#!/bin/bash
out=gpurun_out
for w in c2 headline; do
  timeout 900 python bench.py --workload $w --steps 10 --warmup 3 > $out/r2av_bench_$w.json 2> $out/r2av_bench_$w.err
  echo "== $w rc=$?: $(grep -o '"value": [0-9.e+]*\|"ms_per_step": [0-9.e+]*\|"sample_renumber_ms_per_step": [0-9.e+]*\|"gather_alone_ms_per_step": [0-9.e+]*\|"gather_gbs": [0-9.e+]*\|"parity_checked": [a-z]*\|"frac_alone": [0-9.e+]*' $out/r2av_bench_$w.json | tr '\n' ' ')"
done
