#!/bin/bash
# eight GPUs: the C5 line (heterogeneous sampler + striped P2P gather + DDP-wrapped SAGE step) and the default C4 line as the driver launches it
out=gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --workload c5 --gpus 8 --steps 10 --warmup 3 > $out/r2z_bench_c5_n8.json 2> $out/r2z_bench_c5_n8.err
echo "== c5 n8 rc=$?"; tail -2 $out/r2z_bench_c5_n8.err | cut -c1-200; cut -c1-400 $out/r2z_bench_c5_n8.json
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 8 --steps 10 --warmup 3 > $out/r2z_bench_c4_n8.json 2> $out/r2z_bench_c4_n8.err
echo "== c4 n8 rc=$?"; tail -2 $out/r2z_bench_c4_n8.err | cut -c1-200; cut -c1-400 $out/r2z_bench_c4_n8.json
