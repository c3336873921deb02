#!/bin/bash
out=gpurun_out
timeout 300 python -m pytest tests/test_gpu_sage_tile.py -m gpu -x -q 2>&1 | tail -2
for s in 1024 16384; do
  timeout 300 python profiles/sage_tile_bench.py $s 10 > $out/r2ao_sage_c3_$s.jsonl 2> $out/r2ao_sage_c3_$s.err
  echo "== sage bench seeds=$s rc=$?"; cut -c1-110 $out/r2ao_sage_c3_$s.jsonl | head -2; tail -1 $out/r2ao_sage_c3_$s.err
done
