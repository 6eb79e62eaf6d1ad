"""GPU: several queries per scan (swb_search_batch / swb_search_hits_batch; the
reference searches its queries one after the other, swipe.cc:2561).  The batched results must be
exactly those of separate searches -- checked against the CPU oracle per query."""
import numpy as np
import pytest

import fixtures
from swipe_b200 import Database, Scoring, scoring, synth

pytestmark = pytest.mark.gpu
B62 = scoring.blosum62()


def _queries(lengths, seed=900):
    return [synth.protein_query(L, seed=seed + 7 * k + L) if L else np.zeros(0, np.uint8)
            for k, L in enumerate(lengths)]


@pytest.mark.parametrize("lengths", [[100, 100, 100, 100], [25, 50, 75, 100, 24, 26, 1, 99],
                                     [30, 150, 60, 25, 90, 10, 300, 400, 401, 0, 7, 380, 20],
                                     [200, 200], [5] * 20])
def test_batch_equals_separate_searches(oracle, lengths):
    qs = _queries(lengths)
    residues, offsets = synth.protein_db(2500, query=qs[0] if lengths[0] >= 8 else None, seed=77,
                                         plant_every=25, max_len=900)
    sc = Scoring(B62, 11, 1)
    with Database(residues, offsets) as db:
        got = db.search_batch(qs, sc)
    for k, q in enumerate(qs):
        exp = oracle.scan(residues, offsets, q, B62, 11, 1)[0] if q.size else np.zeros(offsets.size - 1, np.int64)
        bad = np.nonzero(got[k] != exp)[0]
        assert bad.size == 0, "query %d (len %d): %d scores differ, first %s" % (k, q.size, bad.size, bad[:5])


def test_batch_with_requeued_subjects_and_hits(oracle):
    """Planted copies push some subjects of every query beyond the first tier's exact range; the re-queue
    tiers and the device sink then run per query on top of the shared scan."""
    rng = np.random.default_rng(3)
    qs = _queries([120, 90, 110], seed=40)
    subs = []
    for i in range(900):
        if i % 9 == 0:
            q = qs[i % 3]
            subs.append(np.concatenate([synth.random_protein(rng, 5)] + [q] * 6))     # repeats: score stays modest
        elif i % 50 == 1:
            subs.append(np.concatenate([qs[0]] * 1 + [synth.random_protein(rng, 40)]))
        else:
            subs.append(synth.random_protein(rng, int(rng.integers(20, 700))))
    long_q = synth.protein_query(390, seed=41)
    qs.append(long_q)
    subs.append(np.concatenate([long_q, long_q]))
    residues, offsets = fixtures.pack(subs)
    m = scoring.blosum62()
    sc = Scoring(m, 11, 1)
    with Database(residues, offsets) as db:
        dense = db.search_batch(qs, sc)
        hits = db.search_hits_batch(qs, sc, 40, 1, 2 ** 62, seqno_base=500)
        ctr = db.last_batch_counters
    for k, q in enumerate(qs):
        exp = oracle.scan(residues, offsets, q, m, 11, 1)[0]
        assert np.array_equal(dense[k], exp), k
        oseq, osc, otot, _ = oracle.topk(np.arange(exp.size) + 500, exp, 40, min_score=1)
        assert np.array_equal(hits[k][0], oseq) and np.array_equal(hits[k][1], osc) and hits[k][2] == otot
    assert sum(c["gpu_requeued"] for c in ctr) > 0


def test_batch_nucleotide_and_generic_penalties(oracle):
    q = synth.dna_query(90)
    qs = [q, synth.revcomp_nt(q), synth.dna_query(64, seed=5)]
    residues, offsets = synth.dna_db_planted(3000, q, seed=4, plant_every=60, ambiguity_every=3)
    m = scoring.nucleotide_matrix(2, -3)
    for gaps in ((5, 2), (4, 1)):
        with Database(residues, offsets) as db:
            got = db.search_batch(qs, Scoring(m, *gaps))
        for k, qq in enumerate(qs):
            assert np.array_equal(got[k], oracle.scan(residues, offsets, qq, m, *gaps)[0]), (gaps, k)


def test_batch_int16_lanes(oracle):
    """Penalties beyond the fp16-pattern range (open + extend > 1023) put the scan on plain int16 lanes;
    batched queries must take the same build."""
    qs = _queries([80, 120, 60, 33], seed=70)
    residues, offsets = synth.protein_db(1200, query=qs[1], seed=71, plant_every=30, max_len=600)
    sc = Scoring(B62, 2000, 3)
    with Database(residues, offsets) as db:
        got = db.search_batch(qs, sc)
        assert db.last_batch_counters[0]["scan_geometry"] == 2
    for k, q in enumerate(qs):
        assert np.array_equal(got[k], oracle.scan(residues, offsets, q, B62, 2000, 3)[0]), k
