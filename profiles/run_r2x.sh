#!/bin/bash
# bulk gather with the software-pipelined index -> slot -> copy chain: parity, then two GPUs (replica / no replica) against the register kernel
out=gpurun_out
timeout 900 python -m pytest tests/test_gpu_gather_bulk.py tests/test_gpu_multirank.py -m gpu -x -q 2>&1 | tail -3
for bulk in 1 0; do
  WGB_GATHER_BULK=$bulk timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 --no-parity-check > $out/r2x_bench_n2_bulk$bulk.json 2> $out/r2x_bench_n2_bulk$bulk.err
  echo "== n2 bulk=$bulk rc=$?: $(grep -o '"value": [0-9.e+]*\|"ms_per_step": [0-9.e+]*\|"gather_ms_per_step": [0-9.e+]*\|"gather_alone_ms_per_step": [0-9.e+]*\|"frac": [0-9.e+]*' $out/r2x_bench_n2_bulk$bulk.json | tr '\n' ' ')"
done
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-parity-check > $out/r2x_bench_n1.json 2> $out/r2x_bench_n1.err
echo "== n1: $(grep -o '"value": [0-9.e+]*\|"ms_per_step": [0-9.e+]*\|"gather_ms_per_step": [0-9.e+]*\|"gather_alone_ms_per_step": [0-9.e+]*' $out/r2x_bench_n1.json | tr '\n' ' ')"
