#!/bin/bash
# two GPUs, final state: loader + gather tests on the new defaults (bulk gather for striped tables, call-group decode), multirank tests, the driver's line
out=gpurun_out
timeout 900 python -m pytest tests/test_gpu_loader.py tests/test_gpu_parity.py tests/test_gpu_gather_bulk.py tests/test_gpu_hetero.py -m gpu -x -q 2>&1 | tail -2
timeout 900 python -m pytest tests/test_gpu_multirank.py -m gpu -x -q 2>&1 | tail -2
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > $out/r2al_bench_n2.json 2> $out/r2al_bench_n2.err
echo "== c4 n2 rc=$?: $(grep -o '"value": [0-9.e+]*\|"ms_per_step": [0-9.e+]*\|"gather_ms_per_step": [0-9.e+]*\|"gather_alone_ms_per_step": [0-9.e+]*\|"parity_checked": [a-z]*' $out/r2al_bench_n2.json | tr '\n' ' ')"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 3 --warmup 1 > $out/r2al_bench_reference_n2.json 2> $out/r2al_bench_reference_n2.err
echo "== reference arm under torchrun rc=$?: $(grep -o '"cores": [0-9]*\|"value": [0-9.e+]*' $out/r2al_bench_reference_n2.json | head -3 | tr '\n' ' ')"
