#!/bin/bash
out=gpurun_out
for kb in 96 200 96 200; do
  WGB_GATHER_BULK_KB=$kb timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-parity-check > $out/r2ba_kb$kb.json 2> $out/r2ba_kb$kb.err
  echo "== ring KB $kb: $(grep -o '"value": [0-9.e+]*\|"ms_per_step": [0-9.e+]*\|"gather_alone_ms_per_step": [0-9.e+]*' $out/r2ba_kb$kb.json | tr '\n' ' ')"
done
