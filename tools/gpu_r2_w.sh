#!/bin/bash
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:swb_scan2 -s 1 -c 1 -f -o /tmp/prof_batch \
  python bench.py --config qlen100 --batch 4 --nseq 2000000 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/r2w_ncu_batch.log 2>&1
ncu -i /tmp/prof_batch.ncu-rep --page raw --csv > gpurun_out/r2g_ncu_batch4_raw.csv 2>/dev/null
ncu -i /tmp/prof_batch.ncu-rep --page source --csv > gpurun_out/r2g_ncu_batch4_source.csv 2>/dev/null
tail -2 gpurun_out/r2w_ncu_batch.log; ls -la gpurun_out/r2g_ncu_batch4_*
