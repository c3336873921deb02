#!/bin/bash
# does the insert rate per SM depend on the call group's working set?  phase clock at 16 / 32 / 64 labels per call group
out=gpurun_out
for l in 16 32 64; do
  WGB_MH_TIMING=1 timeout 300 python profiles/prof_step.py 12 $l c4 > $out/r2ag_stage_times_l$l.txt 2>&1
  echo "== labels $l"; grep -E "kernel span|hop1: (sample|insert|first)|label total|fused labels" $out/r2ag_stage_times_l$l.txt
done
