"""GPU parity against the committed golden vectors (outputs of the unmodified reference kernels,
tests/golden/make_golden.py)."""
import os

import numpy as np
import pytest

from swipe_b200 import Database, Scoring, scoring, synth

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load(name):
    return np.load(os.path.join(G, name))


@pytest.mark.parametrize("name,go,ge", [("blosum62", 11, 1), ("blosum50", 10, 2)])
def test_protein_golden(name, go, ge):
    g = load("protein.npz")
    m = load("matrices.npz")[name].astype(np.int64)
    key = "%s_%d_%d" % (name, go, ge)
    with Database(g["residues"], g["offsets"]) as db:
        for shape in ((0, 0, -1), (0, 0, 0), (32, 12, 1), (8, 8, 1)):
            db.set_shape(*shape)
            got = db.search(g["query"], Scoring(m, go, ge))
            assert np.array_equal(got, g["scores_" + key]), shape
            c = db.last_counters
            w = g["width_" + key]
            assert [c["ref_width7"], c["ref_width16"], c["ref_width63"]] == \
                [int((w == 7).sum()), int((w == 16).sum()), int((w == 63).sum())]


def test_three_widths_golden():
    g = load("widths.npz")
    big = synth.protein_query(13000, seed=int(g["query_seed"][0]))
    subs = [big[:L] for L in g["lengths"][:-1]] + [synth.random_protein(np.random.default_rng(1), 500)]
    offsets = np.zeros(len(subs) + 1, np.int64)
    np.cumsum([len(s) for s in subs], out=offsets[1:])
    with Database(np.concatenate(subs), offsets) as db:
        got = db.search(big, Scoring(scoring.blosum62(), 11, 1))
        c = db.last_counters
    assert np.array_equal(got, g["scores"])
    assert [c["ref_width7"], c["ref_width16"], c["ref_width63"]] == g["counts"].tolist()[:1] * 0 + [
        int((g["width"] == 7).sum()), int((g["width"] == 16).sum()), int((g["width"] == 63).sum())]
    assert got.max() > 65535


def test_nucleotide_golden():
    g = load("nt.npz")
    sc = Scoring(scoring.nucleotide_matrix(1, -3), 5, 2)
    with Database(g["residues"], g["offsets"]) as db:
        assert np.array_equal(db.search(g["query"], sc), g["scores_plus"])
        assert np.array_equal(db.search(synth.revcomp_nt(g["query"]), sc), g["scores_minus"])


def test_alignment_ends_golden():
    g = load("ends.npz")
    p = load("protein.npz")
    with Database(p["residues"], p["offsets"]) as db:
        s, bp, bq = db.search_end(p["query"], Scoring(scoring.blosum62(), 11, 1), g["subjects"])
    assert np.array_equal(s, g["scores"])
    assert np.array_equal(bp, g["bestpos"]) and np.array_equal(bq, g["bestq"])


def test_asymmetric_golden():
    g = load("asym.npz")
    with Database(g["residues"], g["offsets"]) as db:
        got = db.search(g["query"], Scoring(g["matrix"].astype(np.int64), 7, 2))
    assert np.array_equal(got, g["scores"])


def test_alignment_phase_golden():
    """align_chunk (swipe.cc:339-414): the end cells from the GPU (search16s's contract) are the
    hints of the host traceback; the alignments must be the ones the reference's align() gave."""
    import json
    from swipe_b200 import align
    recs = [r for r in json.load(open(os.path.join(G, "align.json")))["protein"] if "hint" in r]
    p = load("protein.npz")
    sc = Scoring(scoring.blosum62(), 11, 1)
    subjects = np.array([r["subject"] for r in recs])
    with Database(p["residues"], p["offsets"]) as db:
        s, bp, bq = db.search_end(p["query"], sc, subjects)
    for k, r in enumerate(recs):
        assert [int(s[k]), int(bq[k]), int(bp[k])] == r["hint"]
        d = p["residues"][p["offsets"][r["subject"]]:p["offsets"][r["subject"] + 1]]
        assert list(align(p["query"], d, sc, hint=(s[k], bq[k], bp[k]))) == r["hinted"]
