#!/bin/bash
# First gpurun call of the next round (run from the repo root on the GPU box):
#   /usr/local/graft/bin/gpurun --timeout 1500 -- 'bash profiles/run_round2_first_pass.sh r2a'
# 1. the temporal path on hardware for the first time (written after round 1's GPU minutes were spent; logic-checked on the CPU
#    through tests/emu): under `timeout`, so that a kernel that hangs cannot take the box with it
# 2. the regular GPU suite (regression)
# 3. the overlap experiment of DESIGN.md §10.1: gather at half / three-quarter occupancy against the default
# 4. the larger shapes BASELINE.json names (papers100M shape, north-star shape) at N = 1
# Outputs land in gpurun_out/<tag>_*.
tag=${1:-r2a}
out=gpurun_out
mkdir -p $out
WGB_RUN_UNVERIFIED=1 timeout 300 python -m pytest tests/test_gpu_temporal.py -m gpu -x -q > $out/${tag}_temporal_tests.log 2>&1
echo "temporal tests exit code $?" >> $out/${tag}_temporal_tests.log
tail -5 $out/${tag}_temporal_tests.log
timeout 900 python -m pytest tests -m gpu -x -q > $out/${tag}_gpu_tests.log 2>&1
tail -3 $out/${tag}_gpu_tests.log
for sms in -1 74 111; do
  timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --gather-sms $sms > $out/${tag}_bench_gather_sms_${sms}.json 2> $out/${tag}_bench_gather_sms_${sms}.err
done
grep -h -o '"value": [0-9.e+]*\|"gather_ms_per_step": [0-9.e+]*' $out/${tag}_bench_gather_sms_*.json
timeout 600 python bench.py --workload c4 --steps 10 --warmup 3 --no-cpu-baseline > $out/${tag}_bench_c4.json 2> $out/${tag}_bench_c4.err
timeout 600 python bench.py --workload headline --steps 10 --warmup 3 --no-cpu-baseline > $out/${tag}_bench_headline.json 2> $out/${tag}_bench_headline.err
ls -la $out | tail -12
