#!/bin/bash
# full GPU suite on the current build, then the driver's two commands (reference arm, default bench)
out=gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > $out/r2t_gpu_tests.log 2>&1
tail -4 $out/r2t_gpu_tests.log
timeout 900 python bench.py --impl reference > $out/r2t_bench_reference.json 2> $out/r2t_bench_reference.err
echo "== reference arm rc=$?"; cut -c1-400 $out/r2t_bench_reference.json
timeout 900 python bench.py > $out/r2t_bench.json 2> $out/r2t_bench.err
echo "== default bench rc=$?"; cut -c1-300 $out/r2t_bench.json; tail -2 $out/r2t_bench.err | cut -c1-300
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
