#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x --timeout 900 > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?"; tail -5 gpurun_out/pytest_gpu.log | cut -c1-300
python - <<'PY'
import sys, time, os
sys.path.insert(0, '.')
import numpy as np
from swipe_b200 import Database, Scoring, scoring, synth
q = synth.protein_query(375)
res, off = synth.protein_db(5000000, query=q)
sc = Scoring(scoring.blosum62(), 11, 1)
for env in ("1", None, None):
    if env: os.environ["SWB_NO_STAGING"] = env
    else: os.environ.pop("SWB_NO_STAGING", None)
    t = time.perf_counter()
    db = Database(res, off)
    t1 = time.perf_counter() - t
    s = db.search(q, sc)
    print("staging" if env is None else "no staging", "open %.1f ms" % (t1 * 1e3), "open_ms", db.open_ms(), int(s.sum()))
    db.close()
PY
