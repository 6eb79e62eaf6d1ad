#!/bin/bash
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:swb_scan -s 1 -c 1 -f -o gpurun_out/prof_scan \
  python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --nseq ${1:-2000000} $2 > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log
