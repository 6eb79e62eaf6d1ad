"""GPU parity of the second scan-kernel geometry (swb_scan2_kernel: one warp = one pipeline stage of
32 streams, mailbox hand-off, flags in the running-maximum word) against the CPU oracle: edge cases,
query lengths around the stage / pass boundaries, every compiled shape single- and multi-pass, both
lane arithmetics, the re-queue tiers, nucleotide scoring and non-default penalties."""
import numpy as np
import pytest

import fixtures
from swipe_b200 import Database, Scoring, scoring, synth

pytestmark = pytest.mark.gpu
B62 = scoring.blosum62()
SHAPES2 = [(16, 20), (16, 21), (16, 24), (4, 25)]


def _same(db, q, sc, oracle, residues, offsets, what):
    got = db.search(q, sc)
    exp, width, _ = oracle.scan(residues, offsets, q, sc.matrix, sc.gap_open, sc.gap_extend)
    bad = np.nonzero(got != exp)[0]
    assert bad.size == 0, "%s: %d/%d scores differ, first %s got %s exp %s" % (
        what, bad.size, exp.size, bad[:8], got[bad[:8]], exp[bad[:8]])
    return got


@pytest.mark.parametrize("G,R", SHAPES2)
def test_edge_cases(oracle, G, R):
    q = synth.protein_query(375)
    residues, offsets = fixtures.edge_db(q)
    with Database(residues, offsets) as db:
        db.set_geometry(2)
        for lane_mode in (1, 0):
            db.set_shape(G, R, lane_mode)
            _same(db, q, Scoring(B62, 11, 1), oracle, residues, offsets, "edge G%d R%d mode %d" % (G, R, lane_mode))


@pytest.mark.parametrize("qlen", [1, 2, 7, 24, 25, 100, 320, 375, 384, 385, 1000, 1100])
def test_query_lengths_auto_shape(oracle, qlen):
    q = synth.protein_query(qlen, seed=300 + qlen)
    residues, offsets = synth.protein_db(1500, query=q, seed=400 + qlen, plant_every=50, max_len=1200)
    with Database(residues, offsets) as db:
        db.set_geometry(2)
        _same(db, q, Scoring(B62, 11, 1), oracle, residues, offsets, "qlen %d" % qlen)


_MP = {}


@pytest.mark.parametrize("G,R", SHAPES2)
def test_every_shape_multi_pass(oracle, G, R):
    if not _MP:
        q = synth.protein_query(1100, seed=20261017 + 1100)
        residues, offsets = synth.protein_db(500, query=q, seed=1101, plant_every=20, max_len=1400)
        _MP.update(q=q, residues=residues, offsets=offsets,
                   exp=oracle.scan(residues, offsets, q, B62, 11, 1)[0],
                   exp92=oracle.scan(residues, offsets, q, B62, 9, 2)[0])
    c = _MP
    with Database(c["residues"], c["offsets"]) as db:
        db.set_geometry(2)
        for lane_mode in (1, 0):
            db.set_shape(G, R, lane_mode)
            assert np.array_equal(db.search(c["q"], Scoring(B62, 11, 1)), c["exp"]), (G, R, lane_mode)
        db.set_shape(G, R, 1)
        assert np.array_equal(db.search(c["q"], Scoring(B62, 9, 2)), c["exp92"]), (G, R, "generic penalties")


def test_requeue_tiers(oracle):
    rng = np.random.default_rng(5)
    q = synth.protein_query(7000, seed=77)
    subs = []
    for L in (10, 20, 24, 100, 350, 380, 390, 400, 1000, 6000, 6200, 7000):
        subs.append(q[:L].copy())
        subs.append(np.concatenate([synth.random_protein(rng, 13), q[5:L], synth.random_protein(rng, 3)]))
    for _ in range(41):
        subs.append(synth.random_protein(rng, int(rng.integers(20, 500))))
    residues, offsets = fixtures.pack(subs)
    with Database(residues, offsets) as db:
        db.set_geometry(2)
        got = _same(db, q, Scoring(B62, 11, 1), oracle, residues, offsets, "requeue")
        c = db.last_counters
        assert c["gpu_middle"] > 0 and c["gpu_requeued"] - c["gpu_middle"] > 0 and got.max() > 32767
        hseq, hsc, _, _ = db.search_hits(q, Scoring(B62, 11, 1), 10, 1)
    oseq, osc, _, _ = oracle.topk(np.arange(got.size), got, 10, min_score=1)
    assert np.array_equal(hseq, oseq) and np.array_equal(hsc, osc)


def test_nucleotide_and_ambiguity_queries(oracle):
    q = synth.dna_query(1000)
    residues, offsets = synth.dna_db_planted(4000, q, seed=4, plant_every=100, ambiguity_every=3)
    m = scoring.nucleotide_matrix(1, -3)
    with Database(residues, offsets) as db:
        db.set_geometry(2)
        for qq in (q, synth.revcomp_nt(q)):
            _same(db, qq, Scoring(m, 5, 2), oracle, residues, offsets, "nt")
    # a protein query using every symbol (28 table rows: too many for the one-CTA-per-SM ring at G = 16;
    # the chooser must fall back to a shape that fits)
    q = np.concatenate([np.arange(1, 28, dtype=np.uint8), synth.protein_query(200, seed=8)])
    residues, offsets = fixtures.edge_db(q, seed=9)
    with Database(residues, offsets) as db:
        db.set_geometry(2)
        _same(db, q, Scoring(B62, 11, 1), oracle, residues, offsets, "all symbols")


def test_many_chunks_one_launch(oracle, monkeypatch):
    """Several pipeline chunks in one launch (blockIdx.y = chunk) and an asynchronous open."""
    monkeypatch.setenv("SWB_CHUNK_BYTES", "200000")
    q = synth.protein_query(375)
    residues, offsets = synth.protein_db(6000, query=q, seed=21, plant_every=40)
    exp = oracle.scan(residues, offsets, q, B62, 11, 1)[0]
    for wait in (True, False):
        with Database(residues, offsets, wait=wait) as db:
            db.set_geometry(2)
            assert np.array_equal(db.search(q, Scoring(B62, 11, 1)), exp)
