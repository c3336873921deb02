#!/bin/bash
out=gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --workload c5 --gpus 2 --steps 10 --warmup 3 > $out/r2bb_bench_c5_n2.json 2> $out/r2bb_bench_c5_n2.err
echo "== c5 n2 rc=$?"; grep -i parity $out/r2bb_bench_c5_n2.err | cut -c1-200; python - <<P
import json
d=json.loads(open('$out/r2bb_bench_c5_n2.json').read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], d["stages_ms_per_step"], d["parity_checked"], d["config"].get("hot_rows_replicated_per_gpu"))
P
