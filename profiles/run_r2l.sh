#!/bin/bash
# two GPUs: multirank tests, then the bench line with the remote rows on the bulk-copy path and on the register path
out=gpurun_out
timeout 900 python -m pytest tests/test_gpu_multirank.py -m gpu -x -q 2>&1 | tail -3
for bulk in 1 0; do
  WGB_GATHER_BULK=$bulk timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 > $out/r2l_bench_n2_bulk$bulk.json 2> $out/r2l_bench_n2_bulk$bulk.err
  echo "== bulk=$bulk rc=$?"; tail -3 $out/r2l_bench_n2_bulk$bulk.err | cut -c1-300
  python - <<P
import json
try:
    d=json.loads(open('$out/r2l_bench_n2_bulk$bulk.json').read().strip().splitlines()[-1])
    print('value %.3f G ms/step %.3f gather %.3f alone %.3f | striped_only: %s' % (d['value']/1e9, d['ms_per_step'], d['stages']['gather_ms_per_step'], d['stages']['gather_alone_ms_per_step'], json.dumps(d.get('striped_only'))[:400]))
except Exception as e: print('ERR', e)
P
done
