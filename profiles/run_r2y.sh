#!/bin/bash
# sampler: copy rows prefetched to L2 and stored after the sampled rows; software-pipelined hash loops.  Parity, phase clock, call-group time.
out=gpurun_out
timeout 600 python -m pytest tests/test_gpu_multihop_fused.py tests/test_gpu_multihop.py tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -3
WGB_MH_TIMING=1 timeout 300 python profiles/prof_step.py 15 148 c4 > $out/r2y_stage_times_l148.txt 2>&1
tail -24 $out/r2y_stage_times_l148.txt
timeout 300 python profiles/overlap_probe.py c4 10 64,148 2>&1 | tail -2
cp cugraph-gnn_b200/lib/libwholegraph_b200.so /tmp/lib_shipped.so
touch cugraph-gnn_b200/csrc/multihop.cu
WGB_EXTRA_NVCC_FLAGS="-DWGB_FZ_PIPE=0" python cugraph-gnn_b200/build.py > /dev/null 2>&1
echo "== WGB_FZ_PIPE=0 (staged hash loops, new copy path)"
timeout 300 python profiles/overlap_probe.py c4 10 148 2>&1 | tail -1
cp /tmp/lib_shipped.so cugraph-gnn_b200/lib/libwholegraph_b200.so
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $out/r2y_bench_n1.json 2> $out/r2y_bench_n1.err
echo "== n1: $(grep -o '"value": [0-9.e+]*\|"ms_per_step": [0-9.e+]*\|"sample_renumber_ms_per_step": [0-9.e+]*\|"gather_ms_per_step": [0-9.e+]*\|"gather_alone_ms_per_step": [0-9.e+]*\|"parity_checked": [a-z]*' $out/r2y_bench_n1.json | tr '\n' ' ')"
# DRAM traffic of the sampler kernels of one call group (148 labels) and the launch list of the bench command, our kernels only (-k)
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:"fz_|mh_|rows_" --csv --log-file $out/r2y_ncu_step_traffic_c4_l148.csv python profiles/prof_step.py 3 148 c4 > $out/r2y_step.log 2>&1
tail -1 $out/r2y_step.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"fz_|mh_|rows_" -c 400 --csv --log-file $out/r2y_launches_bench.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-parity-check > $out/r2y_bench_under_ncu.log 2>&1
tail -1 $out/r2y_bench_under_ncu.log | cut -c1-120
