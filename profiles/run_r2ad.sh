#!/bin/bash
# sampler variants (all parity-tested forms): which combination is the fastest at 148 labels per call group
out=gpurun_out
timeout 600 python -m pytest tests/test_gpu_multihop_fused.py -m gpu -x -q 2>&1 | tail -1
echo "== default (stream inserts, sample loop not inlined, copy width 4, scan ILP 1)"
timeout 300 python profiles/overlap_probe.py c4 10 148 2>&1 | tail -1
cp cugraph-gnn_b200/lib/libwholegraph_b200.so /tmp/lib_shipped.so
for fl in "-DWGB_FZ_INSERT_STREAM=0 -DWGB_FZ_SAMPLE_INLINE=1" "-DWGB_FZ_SAMPLE_INLINE=1" "-DWGB_FZ_INSERT_STREAM=0"; do
  touch cugraph-gnn_b200/csrc/multihop.cu
  WGB_EXTRA_NVCC_FLAGS="$fl" python cugraph-gnn_b200/build.py > /dev/null 2>&1
  echo "== $fl"
  timeout 300 python profiles/overlap_probe.py c4 10 148 2>&1 | tail -1
done
cp /tmp/lib_shipped.so cugraph-gnn_b200/lib/libwholegraph_b200.so
