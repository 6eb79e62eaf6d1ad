#!/bin/bash
mkdir -p gpurun_out
python bench.py --steps ${1:-10} --warmup 3 $2 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
tail -3 gpurun_out/bench.err; cat gpurun_out/bench.json
