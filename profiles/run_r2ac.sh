#!/bin/bash
# sampler: insert loops and the sampling loop as functions of their own (no spill reloads inside the loops).  Parity, phase clock, call-group time
out=gpurun_out
timeout 600 python -m pytest tests/test_gpu_multihop_fused.py tests/test_gpu_multihop.py tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -2
WGB_MH_TIMING=1 timeout 300 python profiles/prof_step.py 15 148 c4 > $out/r2ac_stage_times_l148.txt 2>&1
tail -24 $out/r2ac_stage_times_l148.txt
timeout 300 python profiles/overlap_probe.py c4 10 64,148 2>&1 | tail -2
