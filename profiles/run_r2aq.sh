#!/bin/bash
# how many SMs' worth of bulk-gather CTAs saturate HBM?  gather alone and the pipelined step at several --gather-sms
out=gpurun_out
for g in 24 40 64 100; do
  timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-parity-check --gather-sms $g > $out/r2aq_bench_gsms$g.json 2> $out/r2aq_bench_gsms$g.err
  echo "== gather-sms $g: $(grep -o '"value": [0-9.e+]*\|"ms_per_step": [0-9.e+]*\|"gather_ms_per_step": [0-9.e+]*\|"gather_alone_ms_per_step": [0-9.e+]*' $out/r2aq_bench_gsms$g.json | head -4 | tr '\n' ' ')"
done
