#!/bin/bash
mkdir -p gpurun_out
python tools/tune_shapes.py 2000000 375 16x24 1,2,3,0 > gpurun_out/tune_exp.jsonl 2> gpurun_out/tune.err
tail -3 gpurun_out/tune.err; cat gpurun_out/tune_exp.jsonl
