#!/bin/bash
# tests + bench + launch list + full ncu capture of the scan kernel (1 GPU)
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x --timeout 600 > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active --format=csv -lms 200 > gpurun_out/clocks.csv &
SMI=$!
python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
kill $SMI
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference.json 2>> gpurun_out/bench.err; echo "ref rc=$?"
tail -3 gpurun_out/bench.err; cat gpurun_out/bench.json gpurun_out/bench_reference.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
  python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/bench_under_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:swb_scan -s 7 -c 1 -f -o gpurun_out/prof_scan \
  python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log
