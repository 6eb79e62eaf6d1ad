#!/bin/bash
# A/B on one box: build-batch variants at 375 aa; slot-based vs shard-sized multi-pass scratch at 1000 aa
mkdir -p gpurun_out
V=build/variants
run() { echo "== $1 $2 $3"; SWB_LIBRARY=$1 timeout 600 python tools/tune_shapes.py 5000000 $2 $3 1 1 2>&1 | tail -2; }
run "" 375 16x24
run $V/libswipe_b200_batch2.so 375 16x24
run $V/libswipe_b200_batch3.so 375 16x24
run "" 375 16x24
run "" 1000 16x16,32x16
run $V/libswipe_b200_preslot.so 1000 16x16,32x16
run "" 1000 16x16
