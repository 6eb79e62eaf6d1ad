"""Generates tests/golden/*.npz from the UNMODIFIED reference kernels (oracle/_ref, built from
/root/reference by oracle/Makefile).  Run in the build container, where /root/reference exists:

    python tests/golden/make_golden.py

What is recorded (all integers):
  matrices.npz   score_matrix_63 + SCORELIMIT_7/16 of every built-in matrix and the nt table
                 (matrices.cc:520-591)
  protein.npz    a 375-aa query against the edge-case database of tests/fixtures.py plus 400
                 planted/random subjects: scores and kept widths from search7_ssse3 -> search16 ->
                 fullsw, and from the plain SSE2 search7 path (swipe.cc:1416-1594)
  widths.npz     self hits straddling SCORELIMIT_7 and SCORELIMIT_16 (7 / 16 / 63-bit paths)
  nt.npz         1000-nt query, +1/-3, gaps 5/2, both strands
  ends.npz       search16s scores + (bestpos, bestq) for the top subjects (search16s.cc:390-405)
  asym.npz       a non-symmetric custom matrix read from a file by the reference's own parser
"""
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

import fixtures  # noqa: E402
from oracle_lib import Ref  # noqa: E402
from swipe_b200 import scoring, synth  # noqa: E402

BUILTIN = ["blosum45", "blosum50", "blosum62", "blosum80", "blosum90", "pam30", "pam70", "pam250",
           "identity_5_1"]


def main():
    ref = Ref()
    out = {}
    for name in BUILTIN:
        m, l7, l16 = ref.matrix_init(name)
        out[name] = m.astype(np.int16)
        out[name + "_limits"] = np.array([l7, l16], dtype=np.int64)
    m, l7, l16 = ref.matrix_init("x", symtype=0, match=1, mismatch=-3)
    out["nt_1_-3"] = m.astype(np.int16)
    out["nt_1_-3_limits"] = np.array([l7, l16], dtype=np.int64)
    np.savez_compressed(os.path.join(HERE, "matrices.npz"), **out)

    # ---- protein, BLOSUM62 11/1 and BLOSUM50 10/2
    q = synth.protein_query(375)
    er, eo = fixtures.edge_db(q)
    pr, po = synth.protein_db(400, query=q, seed=4242, plant_every=8, max_len=900)
    residues = np.concatenate([er, pr])
    offsets = np.concatenate([eo, po[1:] + eo[-1]])
    rec = {"query": q, "residues": residues, "offsets": offsets}
    for name, go, ge in (("blosum62", 11, 1), ("blosum50", 10, 2)):
        ref.matrix_init(name)
        s1, w1, c1 = ref.scan(residues, offsets, q, go, ge, threads=1, chunk=97, ssse3=1)
        s0, w0, c0 = ref.scan(residues, offsets, q, go, ge, threads=1, chunk=1024, ssse3=0)
        assert np.array_equal(s1, s0) and np.array_equal(w1, w0)
        rec["scores_%s_%d_%d" % (name, go, ge)] = s1.astype(np.int32)
        rec["width_%s_%d_%d" % (name, go, ge)] = w1
        rec["counts_%s_%d_%d" % (name, go, ge)] = np.array(c1, dtype=np.int64)
    np.savez_compressed(os.path.join(HERE, "protein.npz"), **rec)

    # ---- the three widths: self hits of growing length (BLOSUM62: limits 117 and 65525)
    ref.matrix_init("blosum62")
    big = synth.protein_query(13000, seed=99)
    subs = [big[:L].copy() for L in (20, 21, 22, 23, 24, 25, 30, 6000, 12300, 12400, 12500, 13000)]
    subs.append(synth.random_protein(np.random.default_rng(1), 500))
    wr, wo = fixtures.pack(subs)
    ws, ww, wc = ref.scan(wr, wo, big, 11, 1, threads=1, chunk=1024, ssse3=1)
    np.savez_compressed(os.path.join(HERE, "widths.npz"), query_seed=np.array([99]),
                        lengths=np.array([len(s) for s in subs]), scores=ws, width=ww,
                        counts=np.array(wc, dtype=np.int64))

    # ---- nucleotide, both strands
    nq = synth.dna_query(1000)
    nr, no = synth.dna_db(600, seed=77)
    rng = np.random.default_rng(78)
    lens = no[1:] - no[:-1]
    for i in range(0, 600, 20):
        w = int(min(lens[i], 140))
        s = int(rng.integers(0, 1000 - w))
        piece = nq[s:s + w].copy()
        if i % 40 == 0:
            piece = synth.revcomp_nt(piece)
        if i % 60 == 0:
            piece[10:14] = 15                                     # an N run
        nr[no[i]: no[i] + w] = piece
    ref.matrix_init("x", symtype=0, match=1, mismatch=-3)
    sp, wp, _ = ref.scan(nr, no, nq, 5, 2, threads=1, ssse3=1)
    sm, wm, _ = ref.scan(nr, no, synth.revcomp_nt(nq), 5, 2, threads=1, ssse3=1)
    np.savez_compressed(os.path.join(HERE, "nt.npz"), query=nq, residues=nr, offsets=no,
                        scores_plus=sp.astype(np.int32), scores_minus=sm.astype(np.int32),
                        width_plus=wp, width_minus=wm)

    # ---- alignment ends from search16s
    ref.matrix_init("blosum62")
    top = np.argsort(-rec["scores_blosum62_11_1"].astype(np.int64), kind="stable")[:60]
    tr, to = fixtures.pack([residues[offsets[i]:offsets[i + 1]] for i in top])
    es, ep, eq = ref.search16s(tr, to, q, 11, 1)
    np.savez_compressed(os.path.join(HERE, "ends.npz"), subjects=top, scores=es, bestpos=ep, bestq=eq)

    # ---- a non-symmetric matrix through the reference's own file parser
    m = fixtures.asym_matrix().reshape(32, 32)
    letters = scoring.SYM_AA[1:28]
    lines = ["# asymmetric test matrix", "   " + "  ".join(letters)]
    for a in range(1, 28):
        lines.append(scoring.SYM_AA[a] + " " + " ".join("%2d" % m[a, b] for b in range(1, 28)))
    text = "\n".join(lines) + "\n"
    with tempfile.NamedTemporaryFile("w", suffix=".mat", delete=False) as f:
        f.write(text)
        path = f.name
    mref, l7, l16 = ref.matrix_init(path)
    os.unlink(path)
    rng = np.random.default_rng(12)
    aq = rng.integers(1, 28, size=333).astype(np.uint8)
    asubs = [rng.integers(1, 28, size=int(rng.integers(1, 300))).astype(np.uint8) for _ in range(200)]
    asubs.append(aq.copy())
    ar, ao = fixtures.pack(asubs)
    ascore, aw, _ = ref.scan(ar, ao, aq, 7, 2, threads=1, ssse3=1)
    ascore0, _, _ = ref.scan(ar, ao, aq, 7, 2, threads=1, ssse3=0)
    assert np.array_equal(ascore, ascore0)
    np.savez_compressed(os.path.join(HERE, "asym.npz"), text=np.array(text), matrix=mref.astype(np.int16),
                        limits=np.array([l7, l16]), query=aq, residues=ar, offsets=ao,
                        scores=ascore.astype(np.int32), width=aw)
    print("golden vectors written to", HERE)


if __name__ == "__main__":
    main()
