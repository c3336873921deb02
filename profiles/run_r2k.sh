#!/bin/bash
out=gpurun_out
b() { # name, env...
  name=$1; shift
  env "$@" timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-parity-check $EXTRA > $out/r2k_$name.json 2> $out/r2k_$name.err
  echo "== $name: $(grep -o '"ms_per_step": [0-9.e+]*' $out/r2k_$name.json | head -1) $(grep -o '"sample_renumber_ms_per_step": [0-9.e+]*\|"gather_ms_per_step": [0-9.e+]*\|"gather_alone_ms_per_step": [0-9.e+]*' $out/r2k_$name.json | tr '\n' ' ')"
}
b t896_carve57_bulk X=1
b t896_carve0_bulk WGB_MH_CARVEOUT=0
b t896_carve0_reg WGB_MH_CARVEOUT=0 WGB_GATHER_BULK=0
b t896_carve57_reg WGB_GATHER_BULK=0
EXTRA="--gather-stream 0" b t896_inline WGB_MH_CARVEOUT=0
EXTRA=""
cp cugraph-gnn_b200/lib/libwholegraph_b200.so /tmp/lib_shipped.so
touch cugraph-gnn_b200/csrc/multihop.cu
WGB_EXTRA_NVCC_FLAGS="-DWGB_FZ_THREADS=1024 -DWGB_FZ_ILP=1" python cugraph-gnn_b200/build.py > /dev/null 2>&1
b t1024_carve0_bulk WGB_MH_CARVEOUT=0
b t1024_carve0_reg WGB_MH_CARVEOUT=0 WGB_GATHER_BULK=0
b t1024_carve0_bulk192 WGB_MH_CARVEOUT=0 WGB_GATHER_BULK_KB=192
EXTRA="--gather-stream 0" b t1024_inline WGB_MH_CARVEOUT=0
cp /tmp/lib_shipped.so cugraph-gnn_b200/lib/libwholegraph_b200.so
