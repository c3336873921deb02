#!/bin/bash
# loader with per-call-group feature prefetch: GPU loader tests + loader-level number on C2 (prefetch on / off)
out=gpurun_out
timeout 900 python -m pytest tests/test_gpu_loader.py tests/test_gpu_temporal.py -m gpu -x -q 2>&1 | tail -2
timeout 600 python bench.py --loader --workload c2 --steps 10 --warmup 3 > $out/r2af_loader_c2.json 2> $out/r2af_loader_c2.err
echo "== loader c2 (prefetch) rc=$?"; cut -c1-500 $out/r2af_loader_c2.json
WGB_LOADER_PREFETCH=0 timeout 600 python bench.py --loader --workload c2 --steps 10 --warmup 3 > $out/r2af_loader_c2_noprefetch.json 2> $out/r2af_loader_c2_noprefetch.err
echo "== loader c2 (per-batch fetch) rc=$?"; cut -c1-500 $out/r2af_loader_c2_noprefetch.json
