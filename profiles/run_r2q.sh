#!/bin/bash
# ncu --set full with source import: the fused sampler kernel (C4, 64 labels) and the fused SAGE tile kernel (C3, 16 K seeds)
out=gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:fz_label -s 2 -c 1 -f -o $out/r2q_fused_c4 python profiles/prof_step.py 4 64 c4 > $out/r2q_fused.log 2>&1
tail -2 $out/r2q_fused.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:sage_tile -s 2 -c 1 -f -o $out/r2q_sage_tile python profiles/sage_tile_bench.py 16384 1 --once > $out/r2q_sage.log 2>&1
tail -2 $out/r2q_sage.log
ls -la $out/*.ncu-rep
