#!/bin/bash
# (1) sage tile v2 (quarter-warp gather, deferred epilogue): parity + C3 A/B;  (2) sampler with the copy-row path: parity on every
# sampler test file, phase clock, then CTA width A/B (rebuilt on the box) at 64 / 74 / 148 labels per call group;  (3) bench at 64 / 148 labels
out=gpurun_out
timeout 300 python -m pytest tests/test_gpu_sage_tile.py -m gpu -x -q 2>&1 | tail -4
if [ "${PIPESTATUS[0]}" = "0" ]; then
  for s in 1024 16384; do
    timeout 300 python profiles/sage_tile_bench.py $s 10 > $out/r2p_sage_c3_$s.jsonl 2> $out/r2p_sage_c3_$s.err
    echo "== sage bench seeds=$s rc=$?"; cut -c1-130 $out/r2p_sage_c3_$s.jsonl; tail -2 $out/r2p_sage_c3_$s.err
  done
fi
timeout 900 python -m pytest tests/test_gpu_multihop_fused.py tests/test_gpu_parity.py tests/test_gpu_multihop.py tests/test_gpu_hetero.py tests/test_gpu_temporal.py tests/test_gpu_loader.py -m gpu -x -q 2>&1 | tail -4
WGB_MH_TIMING=1 timeout 300 python profiles/prof_step.py 15 64 c4 > $out/r2p_stage_times.txt 2>&1
tail -25 $out/r2p_stage_times.txt
cp cugraph-gnn_b200/lib/libwholegraph_b200.so /tmp/lib_shipped.so
for v in "896 1024" "768 768" "640 640" "1024 1024"; do
  set -- $v
  touch cugraph-gnn_b200/csrc/multihop.cu
  WGB_EXTRA_NVCC_FLAGS="-DWGB_FZ_THREADS=$1 -DWGB_FZ_BOUND=$2" python cugraph-gnn_b200/build.py > /dev/null 2>&1
  echo "== threads $1 bound $2"
  timeout 300 python profiles/overlap_probe.py c4 10 64,74,148 2>&1 | tail -3
done
cp /tmp/lib_shipped.so cugraph-gnn_b200/lib/libwholegraph_b200.so
for l in 64 148; do
  timeout 600 python bench.py --steps 20 --warmup 3 --labels $l --no-cpu-baseline --no-parity-check > $out/r2p_bench_c4_l$l.json 2> $out/r2p_bench_c4_l$l.err
  echo "== labels $l: $(grep -o '"value": [0-9.e+]*\|"ms_per_step": [0-9.e+]*\|"sample_renumber_ms_per_step": [0-9.e+]*\|"gather_ms_per_step": [0-9.e+]*\|"gather_alone_ms_per_step": [0-9.e+]*' $out/r2p_bench_c4_l$l.json | tr '\n' ' ')"
  tail -2 $out/r2p_bench_c4_l$l.err | cut -c1-200
done
