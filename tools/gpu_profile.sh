#!/bin/bash
mkdir -p gpurun_out
python tools/tune_shapes.py 2000000 375 > gpurun_out/tune_375.jsonl 2> gpurun_out/tune.err
tail -3 gpurun_out/tune.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
  python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/bench_under_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:swb_scan -s 1 -c 1 -f -o gpurun_out/prof_scan \
  python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --nseq 2000000 > gpurun_out/ncu_full.log 2>&1
tail -3 gpurun_out/ncu_full.log
cat gpurun_out/tune_375.jsonl
