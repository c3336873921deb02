#!/bin/bash
out=gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:sage_tile -s 2 -c 1 -f -o $out/r2an_sage_tile_v3 python profiles/sage_tile_bench.py 16384 1 --once > $out/r2an_sage.log 2>&1
tail -2 $out/r2an_sage.log
