#!/bin/bash
# sampler: copy rows walked edge by edge across the warp (coalesced stores).  Parity, phase clock, call-group time
out=gpurun_out
timeout 600 python -m pytest tests/test_gpu_multihop_fused.py tests/test_gpu_multihop.py tests/test_gpu_parity.py tests/test_gpu_hetero.py -m gpu -x -q 2>&1 | tail -2
WGB_MH_TIMING=1 timeout 300 python profiles/prof_step.py 15 148 c4 > $out/r2ah_stage_times_l148.txt 2>&1
tail -24 $out/r2ah_stage_times_l148.txt
timeout 300 python profiles/overlap_probe.py c4 10 64,148 2>&1 | tail -2
