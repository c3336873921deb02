#!/bin/bash
out=gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:fz_label -s 2 -c 1 -f -o $out/r2ai_fused_c4_l148 python profiles/prof_step.py 4 148 c4 > $out/r2ai_fused.log 2>&1
tail -2 $out/r2ai_fused.log
