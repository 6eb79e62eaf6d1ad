"""BASELINE.json configs[2] and configs[3] on one B200: the query-length sweep against the 5 M
protein shard and the nucleotide mode (1000-nt query, +1/-3, gaps 5/2, both strands = two scans
with the reverse-complemented query, query.cc:337-342).  Prints one JSON line per case; these are
parity-sized-up measurements kept under profiles/, not bench lines.

usage: python tools/sweep_configs.py [nseq_protein] [nseq_dna] [reps]"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from swipe_b200 import Database, Scoring, scoring, synth

nprot = int(sys.argv[1]) if len(sys.argv) > 1 else 5_000_000
ndna = int(sys.argv[2]) if len(sys.argv) > 2 else 50_000_000
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3


def timed(db, q, sc, reps):
    best, s = 1e18, None
    for _ in range(reps):
        s = db.search(q, sc)
        best = min(best, db.last_counters["scan_ms"] + db.last_counters["requeue_ms"])
    return best, s, db.last_counters


if nprot > 0:
    q375 = synth.protein_query(375)
    t0 = time.time()
    residues, offsets = synth.protein_db(nprot, query=q375)
    sc = Scoring(scoring.blosum62(), 11, 1)
    with Database(residues, offsets) as db:
        up, lay = db.open_ms()
        print(json.dumps({"case": "protein_open", "nseq": nprot, "residues": int(offsets[-1]),
                          "upload_and_layout_ms": round(up, 2), "layout_ms": round(lay, 2),
                          "synth_s": round(time.time() - t0, 1)}), flush=True)
        for qlen in (100, 375, 1000, 5000):
            q = q375 if qlen == 375 else synth.protein_query(qlen, seed=20261017 + qlen)
            ms, s, c = timed(db, q, sc, reps)
            cells = float(offsets[-1]) * qlen
            print(json.dumps({"case": "qlen_sweep", "qlen": qlen, "nseq": nprot, "ms": round(ms, 3),
                              "gcups": round(cells / ms * 1e-6, 1), "requeued": c["gpu_requeued"],
                              "ref_width": [c["ref_width7"], c["ref_width16"], c["ref_width63"]],
                              "checksum": int(s.sum()), "max": int(s.max())}), flush=True)
    del residues, offsets

if ndna > 0:
    import psutil
    avail = psutil.virtual_memory().available
    if avail < ndna * 200 * 4:                      # residues + generator scratch must fit comfortably
        ndna = int(avail // (200 * 4))
    t0 = time.time()
    q = synth.dna_query(1000)
    residues, offsets = synth.dna_db(ndna)
    # plant forward / reverse-complement copies of query windows in every 1000th read
    rng = np.random.default_rng(44)
    lens = offsets[1:] - offsets[:-1]
    for i in range(0, ndna, 1000):
        w = int(min(lens[i], 140))
        s0 = int(rng.integers(0, 1000 - w))
        piece = q[s0:s0 + w].copy()
        if (i // 1000) % 2:
            piece = synth.revcomp_nt(piece)
        mut = rng.random(w) < 0.05
        piece[mut] = 1 << rng.integers(0, 4, size=int(mut.sum()))
        residues[offsets[i]: offsets[i] + w] = piece
    sc = Scoring(scoring.nucleotide_matrix(1, -3), 5, 2)
    synth_s = time.time() - t0
    with Database(residues, offsets) as db:
        up, lay = db.open_ms()
        tot = 0.0
        sums = []
        for strand, qq in (("plus", q), ("minus", synth.revcomp_nt(q))):
            ms, s, c = timed(db, qq, sc, reps)
            tot += ms
            sums.append((int(s.sum()), int(s.max()), c["gpu_requeued"],
                         [c["ref_width7"], c["ref_width16"], c["ref_width63"]]))
        cells = float(offsets[-1]) * 1000 * 2
        print(json.dumps({"case": "nucleotide", "qlen": 1000, "nseq": ndna, "residues": int(offsets[-1]),
                          "strands": 2, "ms_both_strands": round(tot, 2),
                          "gcups": round(cells / tot * 1e-6, 1), "per_strand": sums,
                          "upload_and_layout_ms": round(up, 1), "synth_s": round(synth_s, 1)}), flush=True)
