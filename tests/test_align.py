"""swb_align (host traceback, replaces align.cc's align()) against golden alignments produced by
the unmodified reference (tests/golden/align.json) and, where oracle/_ref is present, against the
reference live on random pairs.  Bar: identical score, coordinates and op string."""
import json
import os

import numpy as np
import pytest

import oracle_lib
from swipe_b200 import Scoring, SwbError, align, scoring, synth

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def test_golden_protein_alignments():
    g = json.load(open(os.path.join(GOLD, "align.json")))
    p = np.load(os.path.join(GOLD, "protein.npz"))
    mats = np.load(os.path.join(GOLD, "matrices.npz"))
    q, res, off = p["query"], p["residues"], p["offsets"]
    assert len(g["protein"]) > 100
    hinted = 0
    for rec in g["protein"]:
        sc = Scoring(mats[rec["matrix"]].astype(np.int64), rec["go"], rec["ge"])
        d = res[off[rec["subject"]]:off[rec["subject"] + 1]]
        assert list(align(q, d, sc)) == rec["free"], rec["subject"]
        if "hint" in rec:
            hinted += 1
            assert list(align(q, d, sc, hint=rec["hint"])) == rec["hinted"], rec["subject"]
    assert hinted > 30


def test_golden_nucleotide_alignments_both_strands():
    g = json.load(open(os.path.join(GOLD, "align.json")))
    qn = synth.dna_query(400, seed=g["nt_query_seed"])
    sc = Scoring(scoring.nucleotide_matrix(1, -3), 5, 2)
    assert len(g["nt"]) > 30
    for rec in g["nt"]:
        d = np.array(rec["subject"], dtype=np.uint8)
        if rec["strand"]:
            d = synth.revcomp_nt(d)
        assert list(align(qn, d, sc)) == rec["free"]


def test_ops_cover_the_aligned_region():
    """The op string consumes exactly the aligned query / subject spans and re-scores to the
    reported score (a size-independent property, also checked on a long pair)."""
    rng = np.random.default_rng(8)
    m = scoring.blosum62()
    sc = Scoring(m, 11, 1)
    q = synth.protein_query(3000, seed=9)
    d = np.concatenate([synth.random_protein(rng, 700), q[200:1500], synth.random_protein(rng, 40),
                        q[1500:2800], synth.random_protein(rng, 300)])
    score, qs, ds, qe, de, ops = align(q, d, sc)
    import re
    i, j, total = qs, ds, 0
    for op, n in re.findall(r"([MID])(\d+)", ops):
        n = int(n)
        if op == "M":
            total += sum(int(m[(int(d[j + k]) << 5) + int(q[i + k])]) for k in range(n))
            i += n
            j += n
        elif op == "I":
            total -= 11 + n
            j += n
        else:
            total -= 11 + n
            i += n
    assert (i, j) == (qe + 1, de + 1) and total == score and score > 5000


def test_argument_and_no_alignment_errors():
    sc = Scoring(scoring.blosum62(), 11, 1)
    q = scoring.encode_protein("HEAGAWGHEE")
    with pytest.raises(SwbError) as e:
        align(q, scoring.encode_protein("PPPPPPPP"), sc)          # score 0: the reference calls fatal()
    assert e.value.status == -6
    with pytest.raises(SwbError):
        align(q, scoring.encode_protein("PAWHEAE"), sc, hint=(17, 99, 3))   # hint outside the matrix
    assert align(q, scoring.encode_protein("PAWHEAE"), sc)[0] == 17      # SURVEY 8(c) known answer


@pytest.mark.skipif(not oracle_lib.ref_available(), reason="oracle/_ref not built")
def test_random_pairs_against_live_reference():
    ref = oracle_lib.Ref()
    rng = np.random.default_rng(1)
    n = 0
    for name, go, ge in (("BLOSUM62", 11, 1), ("BLOSUM50", 10, 2), ("BLOSUM62", 0, 1), ("PAM30", 9, 1),
                         ("BLOSUM62", 5, 0)):
        m, _, _ = ref.matrix_init(name)
        sc = Scoring(m, go, ge)
        for it in range(150):
            ql, dl = int(rng.integers(1, 120)), int(rng.integers(1, 150))
            q, d = synth.random_protein(rng, ql), synth.random_protein(rng, dl)
            if it % 2 == 0 and ql > 10:
                w = q[ql // 4: ql - ql // 5].copy()
                if len(w) > 8:
                    k = int(rng.integers(2, len(w) - 2))
                    w = np.concatenate([w[:k], synth.random_protein(rng, int(rng.integers(0, 4))),
                                        w[k + int(rng.integers(0, 3)):]])
                st = int(rng.integers(0, max(1, dl - len(w))))
                d = np.concatenate([d[:st], w, d[st:]])
            if ref.fullsw(d, q, go, ge) == 0:
                continue
            a = ref.align(q, d, go, ge)
            assert align(q, d, sc) == a
            hint = (a[0], a[3], a[4])
            assert align(q, d, sc, hint=hint) == ref.align(q, d, go, ge, hint=hint)
            n += 1
    assert n > 500
