"""Pins the CPU oracle (oracle/sw_oracle.c) before anything else trusts it: against the known
answers obtained from the reference binary, against the golden vectors generated
from the unmodified reference kernels (tests/golden/make_golden.py) and, where oracle/_ref is
present, against those kernels live."""
import os

import numpy as np
import pytest

import fixtures
import oracle_lib
from swipe_b200 import scoring, synth

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load(name):
    return np.load(os.path.join(G, name))


def test_known_answers(oracle):
    q = scoring.encode_protein("HEAGAWGHEE")
    m = scoring.blosum62()
    table = {"PAWHEAE": 17, "HEAGAWGHEE": 62, "W": 11, "HEAGAWGHEEAAAAAAAAAAHEAGAWGHEE": 62,
             "PPPPPPPP": 0, "HEAGAWWWWWWGHEE": 46}
    for s, want in table.items():
        assert oracle.score(scoring.encode_protein(s), q, m, 11, 1) == want


def test_blosum62_and_limits_match_reference_table(oracle):
    g = load("matrices.npz")
    m = scoring.blosum62()
    assert np.array_equal(m, g["blosum62"].astype(np.int64))
    assert np.array_equal(oracle.parse_matrix(scoring.BLOSUM62_TEXT), m)
    lo, hi, l7, l16 = oracle.limits(m)
    assert (lo, hi) == (-4, 11)
    assert [l7, l16] == g["blosum62_limits"].tolist() == [117, 65525]
    assert scoring.matrix_limits(m) == (lo, hi, l7, l16)
    nt = scoring.nucleotide_matrix(1, -3)
    assert np.array_equal(nt, g["nt_1_-3"].astype(np.int64))
    assert oracle.limits(nt)[2:] == tuple(g["nt_1_-3_limits"].tolist()) == (127, 65535)


def test_matrix_parser_on_asymmetric_file(oracle):
    g = load("asym.npz")
    text = str(g["text"])
    assert np.array_equal(oracle.parse_matrix(text), g["matrix"].astype(np.int64))
    assert np.array_equal(scoring.parse_matrix(text), g["matrix"].astype(np.int64))


@pytest.mark.parametrize("name,go,ge", [("blosum62", 11, 1), ("blosum50", 10, 2)])
def test_protein_golden(oracle, name, go, ge):
    g = load("protein.npz")
    m = load("matrices.npz")[name].astype(np.int64)
    s, w, c = oracle.scan(g["residues"], g["offsets"], g["query"], m, go, ge)
    key = "%s_%d_%d" % (name, go, ge)
    assert np.array_equal(s, g["scores_" + key])
    assert np.array_equal(w, g["width_" + key])
    assert c == g["counts_" + key].tolist()


def test_golden_inputs_are_the_seeded_fixtures():
    """The fixture generators are deterministic, so the golden file needs no private inputs."""
    g = load("protein.npz")
    q = synth.protein_query(375)
    assert np.array_equal(q, g["query"])
    er, eo = fixtures.edge_db(q)
    assert np.array_equal(g["residues"][: er.size], er)


def test_three_widths_golden(oracle):
    g = load("widths.npz")
    big = synth.protein_query(13000, seed=int(g["query_seed"][0]))
    m = scoring.blosum62()
    lens = g["lengths"]
    assert set(g["width"].tolist()) == {7, 16, 63}
    for L, want, w in zip(lens[:-1], g["scores"][:-1], g["width"][:-1]):
        got = oracle.score(big[:L], big, m, 11, 1)
        assert got == want
        assert (7 if got < 117 else 16 if got < 65525 else 63) == w


def test_nucleotide_golden(oracle):
    g = load("nt.npz")
    m = scoring.nucleotide_matrix(1, -3)
    sp, wp, _ = oracle.scan(g["residues"], g["offsets"], g["query"], m, 5, 2)
    sm, wm, _ = oracle.scan(g["residues"], g["offsets"], synth.revcomp_nt(g["query"]), m, 5, 2)
    assert np.array_equal(sp, g["scores_plus"]) and np.array_equal(wp, g["width_plus"])
    assert np.array_equal(sm, g["scores_minus"]) and np.array_equal(wm, g["width_minus"])
    assert sp.max() >= 100 and sm.max() >= 100          # both strands carry planted hits


def test_alignment_ends_golden(oracle):
    g = load("ends.npz")
    p = load("protein.npz")
    m = scoring.blosum62()
    for i, s, bp, bq in zip(g["subjects"], g["scores"], g["bestpos"], g["bestq"]):
        d = p["residues"][p["offsets"][i]: p["offsets"][i + 1]]
        assert oracle.score_end(d, p["query"], m, 11, 1) == (s, bp, bq)


def test_asymmetric_golden(oracle):
    g = load("asym.npz")
    s, w, _ = oracle.scan(g["residues"], g["offsets"], g["query"], g["matrix"].astype(np.int64), 7, 2)
    assert np.array_equal(s, g["scores"]) and np.array_equal(w, g["width"])


def test_topk_rule(oracle):
    rng = np.random.default_rng(3)
    scores = rng.integers(0, 40, size=500)
    seqnos = np.arange(500)
    seq, sc, tot, obv = oracle.topk(seqnos, scores, 25, min_score=5, upper=38)
    keep = [(s, n) for n, s in zip(seqnos, scores) if 5 <= s <= 38]
    keep.sort(key=lambda t: (-t[0], -t[1]))
    assert list(zip(sc.tolist(), seq.tolist())) == keep[:25]
    assert tot == int((scores >= 5).sum()) and obv == int((scores > 38).sum())
    # arrival order does not change the kept list
    perm = rng.permutation(500)
    seq2, sc2, _, _ = oracle.topk(seqnos[perm], scores[perm], 25, min_score=5, upper=38)
    assert np.array_equal(seq, seq2) and np.array_equal(sc, sc2)


@pytest.mark.skipif(not oracle_lib.ref_available(), reason="oracle/_ref not built (no /root/reference)")
def test_oracle_against_live_reference(oracle):
    ref = oracle_lib.Ref()
    m, l7, l16 = ref.matrix_init("blosum62")
    rng = np.random.default_rng(2026)
    for qlen in (1, 5, 33, 100, 375):
        q = synth.protein_query(qlen, seed=qlen)
        residues, offsets = synth.protein_db(300, query=q, seed=qlen + 1, plant_every=6, max_len=700)
        for ssse3 in (1, 0):
            s_ref, w_ref, c_ref = ref.scan(residues, offsets, q, 11, 1, threads=2, chunk=64, ssse3=ssse3)
            s, w, c = oracle.scan(residues, offsets, q, m, 11, 1)
            assert np.array_equal(s, s_ref) and np.array_equal(w, w_ref) and c == c_ref
    for _ in range(20):
        d = synth.random_protein(rng, int(rng.integers(0, 200)))
        q = synth.random_protein(rng, int(rng.integers(1, 200)))
        assert oracle.score(d, q, m, 11, 1) == ref.fullsw(d, q, 11, 1)
