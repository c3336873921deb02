#!/bin/bash
# C5 with the in-bench oracle parity check (one GPU), loader-level number on C2 and C4 with the current build
out=gpurun_out
timeout 900 python bench.py --workload c5 --steps 5 --warmup 3 > $out/r2ae_bench_c5_n1.json 2> $out/r2ae_bench_c5_n1.err
echo "== c5 n1 rc=$?"; grep -i "parity" $out/r2ae_bench_c5_n1.err | cut -c1-400; cut -c1-300 $out/r2ae_bench_c5_n1.json
timeout 600 python bench.py --loader --workload c2 --steps 10 --warmup 3 > $out/r2ae_loader_c2.json 2> $out/r2ae_loader_c2.err
echo "== loader c2 rc=$?"; cut -c1-700 $out/r2ae_loader_c2.json; tail -2 $out/r2ae_loader_c2.err | cut -c1-200
timeout 600 python bench.py --loader --steps 5 --warmup 2 > $out/r2ae_loader_c4.json 2> $out/r2ae_loader_c4.err
echo "== loader c4 rc=$?"; cut -c1-700 $out/r2ae_loader_c4.json; tail -2 $out/r2ae_loader_c4.err | cut -c1-200
