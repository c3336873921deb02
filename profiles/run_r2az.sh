#!/bin/bash
out=gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 8 --steps 10 --warmup 3 > $out/r2az_bench_c4_n8.json 2> $out/r2az_bench_c4_n8.err
echo "== c4 n8 rc=$?: $(grep -o '"value": [0-9.e+]*\|"ms_per_step": [0-9.e+]*\|"gather_alone_ms_per_step": [0-9.e+]*\|"parity_checked": [a-z]*' $out/r2az_bench_c4_n8.json | tr '\n' ' ')"
