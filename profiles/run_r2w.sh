#!/bin/bash
# two GPUs, C4, 148 labels: gather kernel x replica size
out=gpurun_out
for cfg in "1 0.1" "0 0.25" "1 0.25"; do
  set -- $cfg
  WGB_GATHER_BULK=$1 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 --hot-ratio $2 --no-parity-check > $out/r2w_bench_n2_bulk$1_hot$2.json 2> $out/r2w_bench_n2_bulk$1_hot$2.err
  echo "== bulk=$1 hot=$2 rc=$?: $(grep -o '"value": [0-9.e+]*\|"ms_per_step": [0-9.e+]*\|"gather_ms_per_step": [0-9.e+]*\|"gather_alone_ms_per_step": [0-9.e+]*' $out/r2w_bench_n2_bulk$1_hot$2.json | tr '\n' ' ')"
done
