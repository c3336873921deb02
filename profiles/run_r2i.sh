#!/bin/bash
# sampler tuning A/B: CTA width x ILP, rebuilt on the box; one correctness pass on the shipped build first
out=gpurun_out
timeout 600 python -m pytest tests/test_gpu_multihop_fused.py tests/test_gpu_parity.py tests/test_gpu_gather_bulk.py -m gpu -x -q 2>&1 | tail -2
cp cugraph-gnn_b200/lib/libwholegraph_b200.so /tmp/lib_shipped.so
for v in "1024 1" "1024 2" "896 1" "896 2"; do
  set -- $v
  touch cugraph-gnn_b200/csrc/multihop.cu
  WGB_EXTRA_NVCC_FLAGS="-DWGB_FZ_THREADS=$1 -DWGB_FZ_ILP=$2" python cugraph-gnn_b200/build.py > /dev/null 2>&1
  echo "== threads $1 ilp $2"
  timeout 300 python profiles/overlap_probe.py c4 10 2>&1 | tail -1
done
cp /tmp/lib_shipped.so cugraph-gnn_b200/lib/libwholegraph_b200.so
