#!/bin/bash
# sampler: 1024-thread CTAs, CTA barrier for one-CTA clusters; hash-loop ILP A/B; default bench at 148 labels with the e2e host/device breakdown
out=gpurun_out
timeout 900 python -m pytest tests/test_gpu_multihop_fused.py tests/test_gpu_multihop.py -m gpu -x -q 2>&1 | tail -3
WGB_MH_TIMING=1 timeout 300 python profiles/prof_step.py 15 148 c4 > $out/r2r_stage_times_l148.txt 2>&1
tail -25 $out/r2r_stage_times_l148.txt
echo "== shipped build (ILP_HASH 1)"
timeout 300 python profiles/overlap_probe.py c4 10 64,148 2>&1 | tail -2
cp cugraph-gnn_b200/lib/libwholegraph_b200.so /tmp/lib_shipped.so
touch cugraph-gnn_b200/csrc/multihop.cu
WGB_EXTRA_NVCC_FLAGS="-DWGB_FZ_ILP_HASH=2" python cugraph-gnn_b200/build.py > /dev/null 2>&1
echo "== ILP_HASH 2"
timeout 300 python profiles/overlap_probe.py c4 10 64,148 2>&1 | tail -2
cp /tmp/lib_shipped.so cugraph-gnn_b200/lib/libwholegraph_b200.so
BENCH_E2E_DEBUG=1 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $out/r2r_bench_c4.json 2> $out/r2r_bench_c4.err
echo "== bench: $(grep -o '"value": [0-9.e+]*\|"ms_per_step": [0-9.e+]*\|"sample_renumber_ms_per_step": [0-9.e+]*\|"gather_ms_per_step": [0-9.e+]*\|"gather_alone_ms_per_step": [0-9.e+]*' $out/r2r_bench_c4.json | tr '\n' ' ')"
grep -v "^\[rank 0\] step" $out/r2r_bench_c4.err | tail -6 | cut -c1-900
grep "^\[rank 0\] step" $out/r2r_bench_c4.err | tail -10
# sage tile v3: coalesced epilogue (warp transpose), four neighbour rows in flight
timeout 300 python -m pytest tests/test_gpu_sage_tile.py -m gpu -x -q 2>&1 | tail -3
for s in 1024 16384; do
  timeout 300 python profiles/sage_tile_bench.py $s 10 > $out/r2r_sage_c3_$s.jsonl 2> $out/r2r_sage_c3_$s.err
  echo "== sage bench seeds=$s rc=$?"; cut -c1-130 $out/r2r_sage_c3_$s.jsonl; tail -2 $out/r2r_sage_c3_$s.err
done
