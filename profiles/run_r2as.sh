#!/bin/bash
# spatial split, enforced: the gather as ONE 4-warp CTA with ~200 KB of tile rings per SM on g SMs, the sampler on s SMs (s + g = 148)
out=gpurun_out
run() { # s g
  WGB_MH_SMS=$1 WGB_GATHER_BULK_KB=200 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-parity-check --gather-sms $2 > $out/r2as_bench_s$1_g$2.json 2> $out/r2as_bench_s$1_g$2.err
  echo "== sampler SMs $1, gather SMs $2: $(grep -o '"value": [0-9.e+]*\|"ms_per_step": [0-9.e+]*\|"sample_renumber_ms_per_step": [0-9.e+]*\|"gather_ms_per_step": [0-9.e+]*\|"gather_alone_ms_per_step": [0-9.e+]*' $out/r2as_bench_s$1_g$2.json | tr '\n' ' ')"
}
run 148 148
run 60 88
run 68 80
run 76 72
