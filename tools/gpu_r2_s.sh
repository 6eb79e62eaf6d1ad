#!/bin/bash
SWB_TRACE_OPEN=1 timeout 600 python tools/e2e_probe.py 375 5000000 2>&1 | grep -A14 "policy\": \"default" | head -0
SWB_TRACE_OPEN=1 timeout 600 python - <<'PY' 2>&1 | tail -30
import os, sys, time
sys.path.insert(0, '.')
import numpy as np
from swipe_b200 import Database, Scoring, HostBuffer, scoring, synth
q = synth.protein_query(375)
residues, offsets = synth.protein_db(5000000, query=q)
sc = Scoring(scoring.blosum62(), 11, 1)
pin = HostBuffer(residues.size); pin.u8[:] = residues
po = HostBuffer(8 * offsets.size); po.view(np.int64)[:] = offsets
for it in range(3):
    print("--- iteration", it, file=sys.stderr, flush=True)
    t0 = time.perf_counter()
    d = Database(pin.u8, po.view(np.int64), wait=False)
    t1 = time.perf_counter()
    d.search_hits(q, sc, 250, 1)
    t2 = time.perf_counter()
    d.close()
    t3 = time.perf_counter()
    print("open %.2f search %.2f close %.2f ms" % ((t1-t0)*1e3, (t2-t1)*1e3, (t3-t2)*1e3), file=sys.stderr, flush=True)
PY
