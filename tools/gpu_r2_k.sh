#!/bin/bash
# launch-geometry experiments through the library's environment hooks (no rebuild), 375 aa, 5 M subjects
run() { echo "== $*"; env "$@" timeout 600 python tools/tune_shapes.py 5000000 375 16x24 1 1 2>&1 | tail -1; }
run SWB_NOTHING=1
run SWB_OVERSUB=2
run SWB_OVERSUB=4
run SWB_CTAS_PER_SM=3
run SWB_CHUNK_BYTES=536870912
run SWB_CHUNK_BYTES=2147483648
run SWB_CHUNK_BYTES=134217728
run SWB_MERGE=0
