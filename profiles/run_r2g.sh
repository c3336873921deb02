#!/bin/bash
out=gpurun_out
for bulk in 1 0; do
  WGB_GATHER_BULK=$bulk timeout 300 python profiles/overlap_probe.py c4 10 2>&1 | tail -1
done
