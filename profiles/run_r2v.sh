#!/bin/bash
# DRAM traffic + launch list of one C4 step at 148 labels per call group (current build), full-section capture of the bulk gather kernel
out=gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file $out/r2v_ncu_step_traffic_c4_l148.csv python profiles/prof_step.py 3 148 c4 > $out/r2v_step.log 2>&1
tail -2 $out/r2v_step.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:rows_bulk_gather -s 1 -c 1 -f -o $out/r2v_gather_bulk python profiles/prof_step.py 3 148 c4 > $out/r2v_gather.log 2>&1
tail -2 $out/r2v_gather.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/r2v_launches_bench.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-parity-check > $out/r2v_bench_under_ncu.log 2>&1
tail -1 $out/r2v_bench_under_ncu.log | cut -c1-200
