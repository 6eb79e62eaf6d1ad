"""Times the end-to-end step (asynchronous open + search_hits + close from pinned host buffers) for one
query length under the launch-policy switches of swb_api.cu (tuning aid).
usage: python tools/e2e_probe.py qlen [nseq]"""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from swipe_b200 import Database, Scoring, HostBuffer, scoring, synth

qlen = int(sys.argv[1])
nseq = int(sys.argv[2]) if len(sys.argv) > 2 else 5000000
q = synth.protein_query(qlen, seed=20261017 + qlen)
residues, offsets = synth.protein_db(nseq, query=synth.protein_query(375))
sc = Scoring(scoring.blosum62(), 11, 1)
pin = HostBuffer(residues.size); pin.u8[:] = residues
po = HostBuffer(8 * offsets.size); po.view(np.int64)[:] = offsets
cells = float(offsets[-1]) * qlen
for name, env in (("default", {}), ("one_stream", {"SWB_ONE_STREAM": "1"}), ("wait_resident", {"SWB_WAIT_RESIDENT": "1"}),
                  ("no_merge", {"SWB_MERGE": "0"})):
    for k in ("SWB_ONE_STREAM", "SWB_WAIT_RESIDENT", "SWB_MERGE"):
        os.environ.pop(k, None)
    os.environ.update(env)
    ts = []
    for it in range(5):
        t0 = time.perf_counter()
        d = Database(pin.u8, po.view(np.int64), wait=False)
        d.search_hits(q, sc, 250, 1)
        c = d.last_counters
        d.close()
        ts.append(time.perf_counter() - t0)
    best = min(ts[1:])
    print(json.dumps({"policy": name, "qlen": qlen, "e2e_ms": round(best * 1e3, 2), "gcups": round(cells / best * 1e-9, 1),
                      "scan_ms": round(c["scan_ms"], 2), "launches": c["kernel_launches"]}), flush=True)
