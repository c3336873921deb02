#!/bin/bash
# One gpurun call that regenerates every measurement kept under profiles/ (run from the repo root on the GPU box):
#   bash profiles/run_profile_pass.sh <tag>
# Outputs land in gpurun_out/<tag>_*; copy the summaries you want judged into profiles/.
tag=${1:-pass}
out=gpurun_out
mkdir -p $out
python bench.py --steps 10 --warmup 3 > $out/${tag}_bench.json 2> $out/${tag}_bench.err
python bench.py --impl reference --steps 5 --warmup 1 > $out/${tag}_bench_reference.json 2> $out/${tag}_bench_reference.err
# every launch of the bench command with its device time (cold-cache, serialised: compare shares)
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/${tag}_launches_bench.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $out/${tag}_bench_under_ncu.log 2>&1
# full sections for the kernels of one step (second step of the driver script)
ncu --set full --clock-control none --import-source on \
    -k regex:"rows_copy|mh_insert|uniform_small|mh_compact|mh_emit|count_scan|mh_reinsert|mh_seed" -s 12 -c 18 \
    -o $out/${tag}_full python profiles/prof_step.py 2 > $out/${tag}_full.log 2>&1
WGB_MH_TIMING=1 python profiles/prof_step.py 15 > $out/${tag}_stage_times.txt 2>&1
python profiles/agg_bench.py 1024 6 > $out/${tag}_agg_c3.jsonl 2> $out/${tag}_agg.err
python profiles/agg_bench.py 16384 6 >> $out/${tag}_agg_c3.jsonl 2>> $out/${tag}_agg.err
ncu --set full --clock-control none -k regex:csr_aggregate -s 40 -c 5 -o $out/${tag}_agg_full python profiles/agg_bench.py 16384 1 > $out/${tag}_agg_full.log 2>&1
ls -la $out | tail -20
