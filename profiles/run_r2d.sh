#!/bin/bash
# fused sampler iteration: correctness (fused tests only) + in-kernel phase clock + bench line
tag=${1:-r2d}
out=gpurun_out
mkdir -p $out
timeout 600 python -m pytest tests/test_gpu_multihop_fused.py -m gpu -x -q > $out/${tag}_fused_tests.log 2>&1
echo "fused tests exit code $?" >> $out/${tag}_fused_tests.log
tail -4 $out/${tag}_fused_tests.log
WGB_MH_TIMING=1 timeout 300 python profiles/prof_step.py 15 64 c4 > $out/${tag}_stage_times.txt 2>&1
cat $out/${tag}_stage_times.txt
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-parity-check > $out/${tag}_bench_c4.json 2> $out/${tag}_bench_c4.err
python - <<'P'
import json,sys
d=json.loads(open('gpurun_out/%s_bench_c4.json' % sys.argv[1] if len(sys.argv)>1 else 'r2d').read().strip().splitlines()[-1]) if False else None
P
grep -o '"value": [0-9.e+]*\|"ms_per_step": [0-9.e+]*\|"sample_renumber_ms_per_step": [0-9.e+]*\|"gather_ms_per_step": [0-9.e+]*' $out/${tag}_bench_c4.json | head -8
