#!/bin/bash
# spatial split: sampler kernel on at most n SMs, the gather of the previous call group on the rest
out=gpurun_out
for n in 56 66 76 90 110; do
  WGB_MH_SMS=$n timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-parity-check > $out/r2ar_bench_sms$n.json 2> $out/r2ar_bench_sms$n.err
  echo "== sampler SMs $n: $(grep -o '"value": [0-9.e+]*\|"ms_per_step": [0-9.e+]*\|"sample_renumber_ms_per_step": [0-9.e+]*\|"gather_ms_per_step": [0-9.e+]*' $out/r2ar_bench_sms$n.json | tr '\n' ' ')"
done
