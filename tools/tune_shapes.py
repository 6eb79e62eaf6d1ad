"""Times the scan kernel for every compiled shape on one resident shard (tuning aid)."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from swipe_b200 import Database, Scoring, scoring, synth

nseq = int(sys.argv[1]) if len(sys.argv) > 1 else 2000000
qlens = [int(x) for x in sys.argv[2].split(",")] if len(sys.argv) > 2 else [375]
shapes = [(4, 25), (8, 8), (8, 13), (8, 16), (16, 12), (16, 16), (16, 20), (16, 24), (32, 12), (32, 16),
          (32, 20), (32, 24), (32, 28), (32, 32)]
if len(sys.argv) > 3:
    shapes = [tuple(int(v) for v in x.split("x")) for x in sys.argv[3].split(",")]
modes = [int(x) for x in sys.argv[4].split(",")] if len(sys.argv) > 4 else [1, 0]
geom = int(sys.argv[5]) if len(sys.argv) > 5 else 1
q0 = synth.protein_query(375)
residues, offsets = synth.protein_db(nseq, query=q0)
sc = Scoring(scoring.blosum62(), 11, 1)
ref = None
with Database(residues, offsets) as db:
    db.set_geometry(geom)
    for qlen in qlens:
        q = synth.protein_query(qlen)
        cells = float(offsets[-1]) * qlen
        ref = None
        for lane_mode in modes:
            for (G, R) in shapes:
                npass = -(-qlen // (G * R))
                if npass > 16 and qlen > 200 and geom == 1:
                    continue
                db.set_shape(G, R, lane_mode)
                best = 1e9
                for _ in range(3):
                    s = db.search(q, sc)
                    best = min(best, db.last_counters["scan_ms"])
                if ref is None:
                    ref = s.copy()
                ok = bool(np.array_equal(ref, s))
                print(json.dumps({"geom": geom, "qlen": qlen, "G": G, "R": R, "mode": lane_mode, "npass": npass,
                                  "scan_ms": round(best, 3), "gcups": round(cells / best * 1e-6, 1),
                                  "same": ok, "requeued": db.last_counters["gpu_requeued"]}), flush=True)
