#!/bin/bash
# Round 2, second GPU pass: the new default bench line (C4, parity-checked), the reference arm, the gather-sms A/B on C4,
# sampler stage times and the ncu launch list + full capture of one C4 step.
tag=${1:-r2b}
out=gpurun_out
mkdir -p $out
timeout 600 python bench.py --steps 10 --warmup 3 > $out/${tag}_bench.json 2> $out/${tag}_bench.err
tail -2 $out/${tag}_bench.err
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > $out/${tag}_bench_reference.json 2> $out/${tag}_bench_reference.err
for sms in 111 96; do
  timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-parity-check --gather-sms $sms > $out/${tag}_bench_gather_sms_${sms}.json 2> $out/${tag}_bench_gather_sms_${sms}.err
done
WGB_MH_TIMING=1 timeout 300 python profiles/prof_step.py 15 > $out/${tag}_stage_times.txt 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/${tag}_launches_bench.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-parity-check > $out/${tag}_bench_under_ncu.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on \
    -k regex:"rows_copy|mh_|uniform_small|count_scan" -s 20 -c 24 \
    -o $out/${tag}_full python profiles/prof_step.py 2 > $out/${tag}_full.log 2>&1
ls -la $out | tail -12
