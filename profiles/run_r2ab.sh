#!/bin/bash
# sampler: lane-decoupled inserts, copy passes of 12, scans with 2 chunks in flight.  Parity, phase clock, call-group time; A/B scan ILP 1
out=gpurun_out
timeout 600 python -m pytest tests/test_gpu_multihop_fused.py tests/test_gpu_multihop.py tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -2
WGB_MH_TIMING=1 timeout 300 python profiles/prof_step.py 15 148 c4 > $out/r2ab_stage_times_l148.txt 2>&1
tail -24 $out/r2ab_stage_times_l148.txt
timeout 300 python profiles/overlap_probe.py c4 10 64,148 2>&1 | tail -2
cp cugraph-gnn_b200/lib/libwholegraph_b200.so /tmp/lib_shipped.so
touch cugraph-gnn_b200/csrc/multihop.cu
WGB_EXTRA_NVCC_FLAGS="-DWGB_FZ_ILP_SCAN=1" python cugraph-gnn_b200/build.py > /dev/null 2>&1
echo "== WGB_FZ_ILP_SCAN=1"
timeout 300 python profiles/overlap_probe.py c4 10 148 2>&1 | tail -1
touch cugraph-gnn_b200/csrc/multihop.cu
WGB_EXTRA_NVCC_FLAGS="-DWGB_FZ_ILP_SCAN=4" python cugraph-gnn_b200/build.py > /dev/null 2>&1
echo "== WGB_FZ_ILP_SCAN=4"
timeout 300 python profiles/overlap_probe.py c4 10 148 2>&1 | tail -1
cp /tmp/lib_shipped.so cugraph-gnn_b200/lib/libwholegraph_b200.so
