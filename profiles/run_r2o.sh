#!/bin/bash
# fused SAGE tile: parity first (short timeout: a wrong barrier hangs), then the C3 A/B; then the driver's two bench commands as the driver runs them
out=gpurun_out
timeout 300 python -m pytest tests/test_gpu_sage_tile.py -m gpu -x -q 2>&1 | tail -15
rc=${PIPESTATUS[0]}
echo "sage tests rc=$rc"
if [ "$rc" = "0" ]; then
  for s in 1024 16384; do
    timeout 300 python profiles/sage_tile_bench.py $s 10 > $out/r2o_sage_c3_$s.jsonl 2> $out/r2o_sage_c3_$s.err
    echo "== sage bench seeds=$s rc=$?"; cut -c1-330 $out/r2o_sage_c3_$s.jsonl; tail -2 $out/r2o_sage_c3_$s.err
  done
fi
t0=$(date +%s)
timeout 900 python bench.py --impl reference > $out/r2o_bench_reference.json 2> $out/r2o_bench_reference.err
echo "== reference arm rc=$? wall $(( $(date +%s) - t0 )) s"; cut -c1-600 $out/r2o_bench_reference.json
t0=$(date +%s)
timeout 900 python bench.py > $out/r2o_bench.json 2> $out/r2o_bench.err
echo "== default bench rc=$? wall $(( $(date +%s) - t0 )) s"; cut -c1-300 $out/r2o_bench.json; tail -3 $out/r2o_bench.err | cut -c1-300
