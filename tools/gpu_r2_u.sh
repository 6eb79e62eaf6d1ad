#!/bin/bash
run() { echo "== $1"; SWB_LIBRARY=$1 timeout 600 python tools/tune_shapes.py 5000000 100 4x25 1 2 2>&1 | tail -1; }
run ""
run build/variants/libswipe_b200_g2b2.so
run build/variants/libswipe_b200_g2b4.so
