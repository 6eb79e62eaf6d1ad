"""Generates tests/golden/align.json from the UNMODIFIED reference aligner (align.cc, through
oracle/_ref/libswipe_ref.so).  Run in the build container:

    python tests/golden/make_golden_align.py

For the 60 best subjects of the protein fixture (tests/golden/protein.npz, ends.npz): the
alignment the reference reports when hits_align passes search16s's end cell as a hint
(hits.cc:587-600, only when bestq > 0 and bestpos > 0) and when it has to find the end itself;
plus nucleotide pairs (+1/-3, 5/2) against the forward and the reverse-complemented subject.
Each record: [score, q_start, d_start, q_end, d_end, ops]."""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle_lib import Ref  # noqa: E402
from swipe_b200 import synth  # noqa: E402


def main():
    ref = Ref()
    p = np.load(os.path.join(HERE, "protein.npz"))
    e = np.load(os.path.join(HERE, "ends.npz"))
    q, res, off = p["query"], p["residues"], p["offsets"]
    out = {"protein": [], "nt": []}
    for name, go, ge in (("blosum62", 11, 1), ("blosum50", 10, 2)):
        ref.matrix_init(name)
        for k, s in enumerate(e["subjects"]):
            d = res[off[s]:off[s + 1]]
            rec = {"matrix": name, "go": go, "ge": ge, "subject": int(s)}
            if ref.fullsw(d, q, go, ge) == 0:
                continue
            rec["free"] = list(ref.align(q, d, go, ge))
            if name == "blosum62" and e["bestq"][k] > 0 and e["bestpos"][k] > 0:
                hint = (int(e["scores"][k]), int(e["bestq"][k]), int(e["bestpos"][k]))
                rec["hint"] = list(hint)
                rec["hinted"] = list(ref.align(q, d, go, ge, hint=hint))
            out["protein"].append(rec)
    ref.matrix_init("x", symtype=0, match=1, mismatch=-3)
    rng = np.random.default_rng(31)
    qn = synth.dna_query(400, seed=32)
    for i in range(24):
        L = int(rng.integers(60, 500))
        d = (1 << rng.integers(0, 4, size=L)).astype(np.uint8)
        w = int(rng.integers(30, min(L, 200)))
        st = int(rng.integers(0, 400 - w))
        piece = qn[st:st + w].copy()
        mut = rng.random(w) < 0.08
        piece[mut] = 1 << rng.integers(0, 4, size=int(mut.sum()))
        if i % 3 == 0 and w > 20:
            piece = np.concatenate([piece[:w // 2], piece[w // 2 + 2:]])
        if i % 2:
            piece = synth.revcomp_nt(piece)
        at = int(rng.integers(0, L - len(piece) + 1))
        d[at:at + len(piece)] = piece
        if i % 5 == 0:
            d[at + 3:at + 6] = 15
        for strand in (0, 1):
            dd = synth.revcomp_nt(d) if strand else d
            if ref.fullsw(dd, qn, 5, 2) == 0:
                continue
            out["nt"].append({"subject": d.tolist(), "strand": strand,
                              "free": list(ref.align(qn, dd, 5, 2))})
    out["nt_query_seed"] = 32
    json.dump(out, open(os.path.join(HERE, "align.json"), "w"))
    print("wrote align.json: %d protein, %d nt records" % (len(out["protein"]), len(out["nt"])))


if __name__ == "__main__":
    main()
