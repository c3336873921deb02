#!/bin/bash
# is the per-label time at one CTA per label a function of how many labels (how much scratch) are in flight?
out=gpurun_out
for l in 76 100 124 148; do
  WGB_MH_TIMING=1 timeout 300 python profiles/prof_step.py 12 $l c4 > $out/r2aj_stage_times_l$l.txt 2>&1
  echo "== labels $l"; grep -E "kernel span|hop1: (sample|insert|first)|label total" $out/r2aj_stage_times_l$l.txt
done
