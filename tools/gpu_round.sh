#!/bin/bash
# tests + bench + config sweeps + launch list + full ncu capture of the scan kernel (1 GPU)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x --timeout 900 > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()"
python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference.json 2>> gpurun_out/bench.err; echo "ref rc=$?"
tail -3 gpurun_out/bench.err; cat gpurun_out/bench.json gpurun_out/bench_reference.json
timeout 900 python tools/sweep_configs.py 5000000 ${NDNA:-50000000} 2 > gpurun_out/sweep.jsonl 2> gpurun_out/sweep.err; echo "sweep rc=$?"; cat gpurun_out/sweep.jsonl; tail -2 gpurun_out/sweep.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
  python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/bench_under_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:swb_scan -s 1 -c 1 -f -o gpurun_out/prof_scan \
  python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --nseq 1500000 > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log
