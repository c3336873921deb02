#!/bin/bash
# sampler: evict-last L2 policy on every table access, evict-first on col_idx reads.  Parity, phase clock, call-group time, with / without the pipelined hash loops
out=gpurun_out
timeout 600 python -m pytest tests/test_gpu_multihop_fused.py tests/test_gpu_multihop.py tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -2
WGB_MH_TIMING=1 timeout 300 python profiles/prof_step.py 15 148 c4 > $out/r2aa_stage_times_l148.txt 2>&1
tail -24 $out/r2aa_stage_times_l148.txt
timeout 300 python profiles/overlap_probe.py c4 10 64,148 2>&1 | tail -2
cp cugraph-gnn_b200/lib/libwholegraph_b200.so /tmp/lib_shipped.so
touch cugraph-gnn_b200/csrc/multihop.cu
WGB_EXTRA_NVCC_FLAGS="-DWGB_FZ_PIPE=0" python cugraph-gnn_b200/build.py > /dev/null 2>&1
echo "== WGB_FZ_PIPE=0"
timeout 300 python profiles/overlap_probe.py c4 10 148 2>&1 | tail -1
cp /tmp/lib_shipped.so cugraph-gnn_b200/lib/libwholegraph_b200.so
