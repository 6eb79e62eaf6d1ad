#!/bin/bash
mkdir -p gpurun_out
V=build/variants
run() { echo "== $1 $2 $3"; SWB_LIBRARY=$1 timeout 600 python tools/tune_shapes.py 5000000 $2 $3 1 1 2>&1 | tail -3; }
run "" 1000 16x16,32x16,16x24
run $V/libswipe_b200_preslot.so 1000 16x16,32x16,16x24
run "" 5000 32x20,32x16
run $V/libswipe_b200_preslot.so 5000 32x20,32x16
