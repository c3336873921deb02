#!/bin/bash
# final state: full GPU suite, smoke, the driver's two commands
out=gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > $out/r2ak_gpu_tests.log 2>&1
tail -3 $out/r2ak_gpu_tests.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 900 python bench.py --impl reference > $out/r2ak_bench_reference.json 2> $out/r2ak_bench_reference.err
echo "== reference arm rc=$?"; cut -c1-300 $out/r2ak_bench_reference.json
timeout 900 python bench.py > $out/r2ak_bench.json 2> $out/r2ak_bench.err
echo "== default bench rc=$?"; cut -c1-300 $out/r2ak_bench.json; tail -1 $out/r2ak_bench.err | cut -c1-300
