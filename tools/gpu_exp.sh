#!/bin/bash
for o in 1 2 4 8 16; do echo "oversub=$o"; SWB_OVERSUB=$o python tools/tune_shapes.py 2000000 375 16x24 1 2>&1 | tail -1; done
