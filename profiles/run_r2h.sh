#!/bin/bash
out=gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:fz_label -s 2 -c 1 -o $out/r2h_fused python profiles/prof_step.py 4 64 c2 > $out/r2h_fused.log 2>&1
tail -3 $out/r2h_fused.log
ls -la $out/r2h_fused.ncu-rep
