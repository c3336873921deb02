#!/bin/bash
# two GPUs: multirank tests, the default bench line as the driver launches it, the C5 line
out=gpurun_out
timeout 900 python -m pytest tests/test_gpu_multirank.py -m gpu -x -q 2>&1 | tail -3
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > $out/r2u_bench_n2.json 2> $out/r2u_bench_n2.err
echo "== c4 n2 rc=$?"; tail -3 $out/r2u_bench_n2.err | cut -c1-300; cut -c1-400 $out/r2u_bench_n2.json
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --workload c5 --gpus 2 --steps 10 --warmup 3 > $out/r2u_bench_c5_n2.json 2> $out/r2u_bench_c5_n2.err
echo "== c5 n2 rc=$?"; tail -2 $out/r2u_bench_c5_n2.err | cut -c1-300; cut -c1-300 $out/r2u_bench_c5_n2.json
