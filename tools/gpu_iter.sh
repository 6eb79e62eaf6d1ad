#!/bin/bash
mkdir -p gpurun_out
cd swipe_b200/csrc && cp libswipe_b200.so lib_base.so.bin && cd ../..
for lib in lib_base.so.bin lib_unroll2.so.bin lib_base.so.bin lib_unroll2.so.bin; do
  cp swipe_b200/csrc/$lib swipe_b200/csrc/libswipe_b200.so; touch swipe_b200/csrc/libswipe_b200.so
  echo "== $lib"; python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['roofline']['kernel_ms'], d['gpu_launches'])"
done
cp swipe_b200/csrc/lib_base.so.bin swipe_b200/csrc/libswipe_b200.so; touch swipe_b200/csrc/libswipe_b200.so
timeout 900 python tools/sweep_configs.py 5000000 0 2 > gpurun_out/sweep_v8.jsonl 2> gpurun_out/sweep.err; cat gpurun_out/sweep_v8.jsonl
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout 600 2>&1 | tail -2
