#!/bin/bash
# Round 2, third GPU pass: the fused per-label sampler (csrc/multihop_fused.cuh) on hardware for the first time.
tag=${1:-r2c}
out=gpurun_out
mkdir -p $out
timeout 600 python -m pytest tests/test_gpu_multihop_fused.py -m gpu -x -q > $out/${tag}_fused_tests.log 2>&1
echo "fused tests exit code $?" >> $out/${tag}_fused_tests.log
tail -15 $out/${tag}_fused_tests.log
timeout 900 python -m pytest tests -m gpu -x -q > $out/${tag}_gpu_tests.log 2>&1
tail -5 $out/${tag}_gpu_tests.log
for fused in 1 0; do
  WGB_MH_FUSED=$fused WGB_MH_TIMING=1 timeout 300 python profiles/prof_step.py 15 64 c4 > $out/${tag}_stage_times_fused${fused}.txt 2>&1
  tail -8 $out/${tag}_stage_times_fused${fused}.txt
done
for sms in -1 111; do
  timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --gather-sms $sms > $out/${tag}_bench_c4_sms${sms}.json 2> $out/${tag}_bench_c4_sms${sms}.err
  tail -2 $out/${tag}_bench_c4_sms${sms}.err
done
timeout 300 python bench.py --workload c2 --steps 20 --warmup 3 --no-cpu-baseline --gather-sms 111 > $out/${tag}_bench_c2.json 2> $out/${tag}_bench_c2.err
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:"fz_|mh_|rows_copy" -s 8 -c 12 --csv \
    --log-file $out/${tag}_step_traffic.csv python profiles/prof_step.py 3 64 c4 > $out/${tag}_step_traffic.log 2>&1
ls -la $out | tail -12
