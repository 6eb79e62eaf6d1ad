#!/bin/bash
# round 2, call B: integration test, the other configs as bench lines, launch list + full ncu capture
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_integration.py -m gpu -q -x --timeout 600 > gpurun_out/r2_pytest_integration.log 2>&1
echo "integration rc=$?"; tail -8 gpurun_out/r2_pytest_integration.log
for c in qlen100 qlen1000 qlen5000; do
  timeout 900 python bench.py --config $c --steps 5 --warmup 3 > gpurun_out/r2_bench_$c.json 2> gpurun_out/r2_bench_$c.err; echo "$c rc=$?"
  cat gpurun_out/r2_bench_$c.json; tail -3 gpurun_out/r2_bench_$c.err
done
timeout 1500 python bench.py --config nt50m --steps 3 --warmup 3 > gpurun_out/r2_bench_nt50m.json 2> gpurun_out/r2_bench_nt50m.err; echo "nt50m rc=$?"
cat gpurun_out/r2_bench_nt50m.json; tail -3 gpurun_out/r2_bench_nt50m.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches.csv \
  python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/r2_bench_under_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:swb_scan -s 1 -c 1 -f -o gpurun_out/r2_prof_scan \
  python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/r2_ncu_full.log 2>&1
tail -2 gpurun_out/r2_ncu_full.log
