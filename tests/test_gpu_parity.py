"""GPU parity tests: the CUDA path (through the C ABI) against the CPU oracle on the same seeded
inputs.  Bar: bit-exact scores (integer work)."""
import numpy as np
import pytest

import fixtures
from swipe_b200 import Database, Scoring, scoring, synth

pytestmark = pytest.mark.gpu

B62 = scoring.blosum62()


def _check(db, q, sc, oracle, residues, offsets, what=""):
    got = db.search(q, sc)
    exp, width, counts = oracle.scan(residues, offsets, q, sc.matrix, sc.gap_open, sc.gap_extend)
    bad = np.nonzero(got != exp)[0]
    assert bad.size == 0, "%s: %d/%d scores differ, first %s got %s exp %s len %s" % (
        what, bad.size, exp.size, bad[:8], got[bad[:8]], exp[bad[:8]],
        (offsets[1:] - offsets[:-1])[bad[:8]])
    c = db.last_counters
    assert [c["ref_width7"], c["ref_width16"], c["ref_width63"]] == \
        [int((width == 7).sum()), int((width == 16).sum()), int((width == 63).sum())]
    return got, c


def test_known_answers(oracle):
    """values obtained from the reference binary."""
    q = scoring.encode_protein("HEAGAWGHEE")
    subs = ["PAWHEAE", "HEAGAWGHEE", "W", "HEAGAWGHEEAAAAAAAAAAHEAGAWGHEE", "PPPPPPPP",
            "HEAGAWWWWWWGHEE"]
    residues, offsets = fixtures.pack([scoring.encode_protein(s) for s in subs])
    with Database(residues, offsets) as db:
        got = db.search(q, Scoring(B62, 11, 1))
        assert got.tolist() == [17, 62, 11, 62, 0, 46]
        db.set_mode(2)
        assert db.search(q, Scoring(B62, 11, 1)).tolist() == [17, 62, 11, 62, 0, 46]


@pytest.mark.parametrize("shape", [(0, 0, -1), (4, 25, 1), (4, 25, 0), (8, 8, 0), (8, 13, 1), (16, 24, 0), (16, 24, 1),
                                   (32, 12, 0), (32, 12, 1), (32, 32, 1)])
def test_edge_cases_all_shapes(oracle, shape):
    q = synth.protein_query(375)
    residues, offsets = fixtures.edge_db(q)
    sc = Scoring(B62, 11, 1)
    with Database(residues, offsets) as db:
        db.set_shape(*shape)
        _check(db, q, sc, oracle, residues, offsets, "edge %s" % (shape,))


@pytest.mark.parametrize("qlen", [1, 2, 7, 100, 375, 1000, 1100])
def test_query_lengths(oracle, qlen):
    q = synth.protein_query(qlen, seed=100 + qlen)
    residues, offsets = synth.protein_db(1500, query=q, seed=200 + qlen, plant_every=50,
                                         max_len=1200)
    with Database(residues, offsets) as db:
        _check(db, q, Scoring(B62, 11, 1), oracle, residues, offsets, "qlen %d" % qlen)


def test_multi_pass_small_shape(oracle):
    """Forces 6 passes over a 375-row query with the smallest shape (64 rows per pass)."""
    q = synth.protein_query(375)
    residues, offsets = synth.protein_db(800, query=q, seed=11, plant_every=20, max_len=900)
    with Database(residues, offsets) as db:
        for lane_mode in (0, 1):
            db.set_shape(8, 8, lane_mode)
            _check(db, q, Scoring(B62, 11, 1), oracle, residues, offsets, "multipass")


def test_requeue_limits(oracle):
    """Self hits of growing length straddle the packed kernels' overflow limits (2047 - hi for
    the fp16-pattern lanes, 32767 - hi for the int16 lanes) and the reference's own limits."""
    rng = np.random.default_rng(5)
    q = synth.protein_query(7000, seed=77)
    subs = []
    for L in (10, 20, 21, 22, 23, 24, 100, 350, 380, 385, 390, 395, 400, 1000, 6000, 6100, 6200, 6300, 7000):
        subs.append(q[:L].copy())
        subs.append(np.concatenate([synth.random_protein(rng, 13), q[5:L], synth.random_protein(rng, 3)]))
    for _ in range(37):
        subs.append(synth.random_protein(rng, int(rng.integers(20, 500))))
    residues, offsets = fixtures.pack(subs)
    sc = Scoring(B62, 11, 1)
    with Database(residues, offsets) as db:
        for lane_mode, limit in ((1, 2047 - 11), (0, 32767 - 11)):
            db.set_shape(0, 0, lane_mode)
            got, c = _check(db, q, sc, oracle, residues, offsets, "requeue mode %d" % lane_mode)
            assert c["gpu_requeued"] == int((got >= limit).sum())
            assert c["gpu_narrow"] + c["gpu_requeued"] == got.size
            if lane_mode == 1:       # the 16-bit tier takes what fits int16 lanes, the wide kernel the rest
                assert c["gpu_middle"] == int(((got >= limit) & (got < 32767 - 11)).sum()) > 0
                assert c["gpu_requeued"] - c["gpu_middle"] == int((got >= 32767 - 11).sum()) > 0
            else:
                assert c["gpu_middle"] == 0
        assert got.max() > 32767
        # the same through a list search (positions of the re-queue are list positions)
        db.set_shape(0, 0, -1)
        sel = np.arange(len(subs) - 1, -1, -3)
        exp, _, _ = oracle.scan(residues, offsets, q, B62, 11, 1)
        assert np.array_equal(db.search_list(q, sc, sel), exp[sel])
        assert db.last_counters["gpu_middle"] > 0


def test_many_requeued_subjects(oracle):
    """A database where a fifth of the subjects score above the 11-bit range (long self copies):
    the 16-bit tier handles a re-queue list of hundreds of subjects in one launch."""
    rng = np.random.default_rng(15)
    q = synth.protein_query(900, seed=78)
    subs = []
    for i in range(1500):
        if i % 5 == 0:
            a = int(rng.integers(0, 200))
            piece = q[a:a + int(rng.integers(450, 700))].copy()
            mut = rng.random(piece.size) < 0.05
            piece[mut] = synth.random_protein(rng, int(mut.sum()))
            subs.append(np.concatenate([synth.random_protein(rng, int(rng.integers(0, 30))), piece]))
        else:
            subs.append(synth.random_protein(rng, int(rng.integers(20, 600))))
    residues, offsets = fixtures.pack(subs)
    sc = Scoring(B62, 11, 1)
    with Database(residues, offsets) as db:
        got, c = _check(db, q, sc, oracle, residues, offsets, "many requeued")
        assert c["gpu_requeued"] >= 250 and c["gpu_middle"] == c["gpu_requeued"]


def test_wide_only_matches(oracle):
    q = synth.protein_query(200, seed=9)
    residues, offsets = fixtures.edge_db(q, seed=8)
    with Database(residues, offsets) as db:
        db.set_mode(2)
        _check(db, q, Scoring(B62, 11, 1), oracle, residues, offsets, "wide")


@pytest.mark.parametrize("gaps", [(11, 1), (10, 2), (0, 1), (5, 0), (0, 0), (40, 3)])
def test_gap_penalties(oracle, gaps):
    q = synth.protein_query(120, seed=31)
    residues, offsets = synth.protein_db(600, query=q, seed=32, plant_every=10, max_len=600)
    with Database(residues, offsets) as db:
        _check(db, q, Scoring(B62, *gaps), oracle, residues, offsets, "gaps %s" % (gaps,))


def test_asymmetric_matrix_and_all_symbols(oracle):
    m = fixtures.asym_matrix()
    rng = np.random.default_rng(12)
    q = rng.integers(0, 28, size=333).astype(np.uint8)              # every code incl. '-' and '*'
    subs = [rng.integers(0, 28, size=int(rng.integers(1, 300))).astype(np.uint8) for _ in range(301)]
    subs.append(q.copy())
    residues, offsets = fixtures.pack(subs)
    with Database(residues, offsets) as db:
        for lane_mode in (0, 1):
            db.set_shape(0, 0, lane_mode)
            _check(db, q, Scoring(m, 7, 2), oracle, residues, offsets, "asym")


def test_all_32_query_symbols(oracle):
    """A query using every one of the 32 symbol codes (the sound alphabet of -p 5 uses 31): the packed
    kernel's per-block tables hold up to 32 symbol rows, in every shape."""
    rng = np.random.default_rng(2)
    q = np.concatenate([np.arange(32), rng.integers(0, 32, size=300)]).astype(np.uint8)
    subs = [rng.integers(0, 32, size=int(rng.integers(1, 400))).astype(np.uint8) for _ in range(300)]
    subs.append(q[10:200].copy())
    residues, offsets = fixtures.pack(subs)
    sound = np.full((32, 32), -1, dtype=np.int64)
    for a in range(1, 32):
        sound[a, a] = 5
    with Database(residues, offsets) as db:
        for shape in ((0, 0, -1), (8, 13, 1), (16, 24, 0), (32, 12, 1)):
            db.set_shape(*shape)
            _, c = _check(db, q, Scoring(fixtures.asym_matrix(), 3, 1), oracle, residues, offsets, "32 symbols %s" % (shape,))
            assert c["gpu_narrow"] + c["gpu_middle"] == len(subs)
        db.set_shape(0, 0, -1)
        _check(db, q, Scoring(sound.reshape(-1), 15, 5), oracle, residues, offsets, "sound identity")


def test_nucleotide_both_strands(oracle):
    """Config 4 at parity size: +1/-3, gaps 5/2, both strands = two scans with the
    reverse-complemented query (query.cc:337-342)."""
    q = synth.dna_query(1000)
    residues, offsets = synth.dna_db(3000, seed=4)
    rng = np.random.default_rng(6)
    # plant forward and reverse-complement copies of query windows, with an N run
    lens = offsets[1:] - offsets[:-1]
    for i in range(0, 3000, 100):
        w = int(min(lens[i], 120))
        s = int(rng.integers(0, 1000 - w))
        piece = q[s:s + w].copy()
        if i % 200 == 0:
            piece = synth.revcomp_nt(piece)
        if i % 300 == 0:
            piece[10:14] = 15
        residues[offsets[i]: offsets[i] + w] = piece
    m = scoring.nucleotide_matrix(1, -3)
    with Database(residues, offsets) as db:
        for strand_q in (q, synth.revcomp_nt(q)):
            _check(db, strand_q, Scoring(m, 5, 2), oracle, residues, offsets, "nt")


def test_trailing_separator_layout(oracle):
    """offsets into a raw .psq: NUL before the first and after every sequence (database.cc:1246)."""
    q = synth.protein_query(90, seed=41)
    subs_res, subs_off = synth.protein_db(300, query=q, seed=42, plant_every=10, max_len=300)
    lens = subs_off[1:] - subs_off[:-1]
    psq = [np.zeros(1, dtype=np.uint8)]
    off = [1]
    for i in range(300):
        psq.append(subs_res[subs_off[i]:subs_off[i + 1]])
        psq.append(np.zeros(1, dtype=np.uint8))
        off.append(off[-1] + int(lens[i]) + 1)
    psq = np.concatenate(psq)
    off = np.array(off, dtype=np.int64)
    with Database(psq, off, trailing=1) as db:
        got = db.search(q, Scoring(B62, 11, 1))
    exp, _, _ = oracle.scan(subs_res, subs_off, q, B62, 11, 1)
    assert np.array_equal(got, exp)


def test_search_list_and_end(oracle):
    q = synth.protein_query(150, seed=51)
    residues, offsets = synth.protein_db(400, query=q, seed=52, plant_every=7, max_len=500)
    sc = Scoring(B62, 11, 1)
    exp, _, _ = oracle.scan(residues, offsets, q, B62, 11, 1)
    rng = np.random.default_rng(53)
    with Database(residues, offsets) as db:
        for n in (0, 1, 5, 77, 400):
            sel = rng.permutation(400)[:n]
            got = db.search_list(q, sc, sel)
            assert np.array_equal(got, exp[sel]), "list of %d" % n
        sel = np.argsort(-exp)[:40]
        s, bp, bq = db.search_end(q, sc, sel)
        assert np.array_equal(s, exp[sel])
        for k, i in enumerate(sel):
            es, ed, eq = oracle.score_end(residues[offsets[i]:offsets[i + 1]], q, B62, 11, 1)
            assert (s[k], bp[k], bq[k]) == (es, ed, eq)


def test_empty_inputs(oracle):
    q = synth.protein_query(50, seed=61)
    residues, offsets = fixtures.pack([np.zeros(0, dtype=np.uint8)] * 3)
    sc = Scoring(B62, 11, 1)
    with Database(residues, offsets) as db:
        assert db.search(q, sc).tolist() == [0, 0, 0]
    with Database(np.zeros(0, dtype=np.uint8), np.zeros(1, dtype=np.int64)) as db:
        assert db.search(q, sc).size == 0
    residues, offsets = synth.protein_db(10, seed=62)
    with Database(residues, offsets) as db:
        assert db.search(np.zeros(0, dtype=np.uint8), sc).tolist() == [0] * 10


def test_full_size_properties():
    """BASELINE config-2-shaped run at a size the oracle cannot cover: size-independent checks.
    (a) a second search returns identical scores; (b) scoring a list of the same subjects in a
    different order gives the same per-subject scores (layout independence); (c) the wide kernel
    agrees on a random sample; (d) planted self copies score the query's self score."""
    q = synth.protein_query(375)
    residues, offsets = synth.protein_db(200000, query=q, seed=71)
    sc = Scoring(B62, 11, 1)
    with Database(residues, offsets) as db:
        a = db.search(q, sc)
        b = db.search(q, sc)
        assert np.array_equal(a, b)
        rng = np.random.default_rng(72)
        sel = rng.permutation(200000)[:30000]
        assert np.array_equal(db.search_list(q, sc, sel), a[sel])
        db.set_mode(2)
        sel2 = sel[:3000]
        assert np.array_equal(db.search_list(q, sc, sel2), a[sel2])
        assert a.min() >= 0


def test_chunked_async_open(oracle, monkeypatch):
    """The shard is uploaded, re-laid-out and scanned in pipeline chunks; force many small chunks
    and the asynchronous open and check scores, list searches and counters are unchanged."""
    q = synth.protein_query(200, seed=81)
    residues, offsets = synth.protein_db(3000, query=q, seed=82, plant_every=11, max_len=800)
    sc = Scoring(B62, 11, 1)
    exp, _, _ = oracle.scan(residues, offsets, q, B62, 11, 1)
    for chunk, merge in (("50000", "1"), ("50000", "0"), ("1", "1"), ("300000000", "1")):
        monkeypatch.setenv("SWB_CHUNK_BYTES", chunk)
        monkeypatch.setenv("SWB_MERGE", merge)     # 1: one launch over all resident chunks
        for wait in (True, False):
            with Database(residues, offsets, wait=wait) as db:
                got = db.search(q, sc)
                assert np.array_equal(got, exp), (chunk, wait)
                assert db.last_counters["gpu_narrow"] + db.last_counters["gpu_requeued"] == 3000
                sel = np.arange(0, 3000, 7)
                assert np.array_equal(db.search_list(q, sc, sel), exp[sel])
                db.set_shape(8, 8, 1)                     # multi-pass over chunks
                assert np.array_equal(db.search(q, sc), exp)


def test_random_scoring_systems(oracle):
    """Seeded fuzz: random (asymmetric) tables over random symbol subsets, random gap penalties, query
    lengths and subject sets, through whatever shape / lane mode / cascade tier the library picks."""
    rng = np.random.default_rng(20261017)
    for it in range(24):
        ncodes = int(rng.integers(2, 32))
        codes = rng.permutation(32)[:ncodes]
        m = np.full((32, 32), -1, dtype=np.int64)
        lo, hi = int(rng.integers(-12, 0)), int(rng.integers(1, 16))
        m[np.ix_(codes, codes)] = rng.integers(lo, hi + 1, size=(ncodes, ncodes))
        if it % 3 == 0:
            for c in codes:
                m[c, c] = hi
        go, ge = int(rng.integers(0, 25)), int(rng.integers(0, 6))
        if go + ge == 0:
            ge = 1
        qlen = int(rng.choice([1, 3, 17, 64, 100, 333, 700]))
        q = rng.choice(codes, size=qlen).astype(np.uint8)
        subs = []
        for _ in range(150):
            L = int(rng.integers(0, 500))
            s = rng.choice(codes, size=L).astype(np.uint8)
            if L > 20 and qlen > 8 and rng.random() < 0.2:
                w = int(min(L, qlen, rng.integers(8, 300)))
                a = int(rng.integers(0, qlen - w + 1))
                s[:w] = q[a:a + w]
            subs.append(s)
        subs.append(rng.integers(0, 32, size=77).astype(np.uint8))          # symbols outside the table
        residues, offsets = fixtures.pack(subs)
        with Database(residues, offsets) as db:
            _check(db, q, Scoring(m.reshape(-1), go, ge), oracle, residues, offsets, "fuzz %d" % it)


@pytest.mark.parametrize("keep,min_score,upper", [(100, 1, 2 ** 62), (10, 40, 2 ** 62), (250, 0, 90),
                                                   (5000, 1, 2 ** 62), (0, 30, 60), (64, 5000, 2 ** 62)])
def test_device_sink_equals_hits_enter(oracle, keep, min_score, upper):
    """swb_search_hits: the admission rule of hits_enter (hits.cc:163-222) applied on the device
    returns the list the oracle's sink holds after seeing every score, ties included."""
    q = synth.protein_query(200, seed=61)
    residues, offsets = synth.protein_db(3000, query=q, seed=62, plant_every=40, max_len=500)
    sc = Scoring(B62, 11, 1)
    exp, _, _ = oracle.scan(residues, offsets, q, B62, 11, 1)
    oseq, osc, otot, oobv = oracle.topk(np.arange(exp.size) + 1000, exp, keep, min_score=min_score, upper=upper)
    with Database(residues, offsets) as db:
        seq, s, tot, obv = db.search_hits(q, sc, keep, min_score, upper, seqno_base=1000)
        c = db.last_counters
    assert np.array_equal(seq, oseq) and np.array_equal(s, osc)
    assert (tot, obv) == (otot, oobv)
    assert c["ref_width7"] + c["ref_width16"] + c["ref_width63"] == exp.size


def test_device_sink_with_requeued_scores_and_ties(oracle):
    """Scores above the first histogram's range (re-queued self hits) and a database that is one
    subject repeated (every score tied): the cut bin and the seqno tie rule must still hold."""
    q = synth.protein_query(6000, seed=63)
    rng = np.random.default_rng(64)
    subs = [q[:L].copy() for L in (300, 800, 900, 5000, 6000)] + [synth.random_protein(rng, 77)] * 300
    residues, offsets = fixtures.pack(subs)
    sc = Scoring(B62, 11, 1)
    exp, _, _ = oracle.scan(residues, offsets, q, B62, 11, 1)
    for keep in (3, 5, 50, 1000):
        oseq, osc, otot, _ = oracle.topk(np.arange(exp.size), exp, keep, min_score=1)
        with Database(residues, offsets) as db:
            seq, s, tot, _ = db.search_hits(q, sc, keep, 1)
        assert np.array_equal(seq, oseq) and np.array_equal(s, osc) and tot == otot


def test_list_with_repeated_subjects(oracle):
    """A list may name a subject several times (search7.cc:894-895 only promises scores[k] <->
    seqnos[k]); the list layout must size itself for the repeats."""
    q = synth.protein_query(150, seed=71)
    residues, offsets = synth.protein_db(40, query=q, seed=72, plant_every=5, min_len=25, max_len=3000)
    lens = offsets[1:] - offsets[:-1]
    longest = int(np.argmax(lens))
    sel = np.array([longest] * 300 + [0, 1, longest, 2] * 20)
    exp, _, _ = oracle.scan(residues, offsets, q, B62, 11, 1)
    with Database(residues, offsets) as db:
        got = db.search_list(q, Scoring(B62, 11, 1), sel)
    assert np.array_equal(got, exp[sel])


def test_query_5000_auto_shape(oracle):
    """BASELINE configs[2], longest query: the automatically chosen multi-pass shape (G32 x R20,
    8 passes at 5000 rows) against the oracle, planted copies re-queueing to the 16-bit tier."""
    q = synth.protein_query(5000, seed=20261017 + 5000)
    residues, offsets = synth.protein_db(700, query=q, seed=5001, plant_every=25, max_len=1500)
    with Database(residues, offsets) as db:
        got, c = _check(db, q, Scoring(B62, 11, 1), oracle, residues, offsets, "qlen 5000")
        assert c["gpu_requeued"] > 0          # some planted copies leave the 11-bit range
        hseq, hsc, _, _ = db.search_hits(q, Scoring(B62, 11, 1), 50, 1)
    oseq, osc, _, _ = oracle.topk(np.arange(got.size), got, 50, min_score=1)
    assert np.array_equal(hseq, oseq) and np.array_equal(hsc, osc)


_MP = {}


def _mp_case(oracle):
    if not _MP:
        q = synth.protein_query(1100, seed=20261017 + 1100)
        residues, offsets = synth.protein_db(500, query=q, seed=1101, plant_every=20, max_len=1400)
        exp, _, _ = oracle.scan(residues, offsets, q, B62, 11, 1)
        _MP.update(q=q, residues=residues, offsets=offsets, exp=exp)
    return _MP


SHAPES = [(4, 25), (8, 8), (8, 13), (8, 16), (16, 12), (16, 16), (16, 20), (16, 24), (32, 12), (32, 16),
          (32, 20), (32, 24), (32, 28), (32, 32)]


@pytest.mark.parametrize("G,R", SHAPES)
def test_every_shape_multi_pass(oracle, G, R):
    """Every compiled scan shape in multi-pass mode (1100 rows > G * R for all of them), both lane
    arithmetics, default penalties (immediate builds) and non-default ones (generic builds)."""
    case = _mp_case(oracle)
    with Database(case["residues"], case["offsets"]) as db:
        for lane_mode in (1, 0):
            db.set_shape(G, R, lane_mode)
            got = db.search(case["q"], Scoring(B62, 11, 1))
            assert np.array_equal(got, case["exp"]), "G%d R%d mode %d" % (G, R, lane_mode)
        db.set_shape(G, R, 1)
        got = db.search(case["q"], Scoring(B62, 9, 2))
        exp, _, _ = oracle.scan(case["residues"], case["offsets"], case["q"], B62, 9, 2)
        assert np.array_equal(got, exp), "G%d R%d generic penalties" % (G, R)


def test_nucleotide_one_million_reads_vs_reference():
    """BASELINE configs[3] at 1/50 of its size: 1 M reads (200 M nt), 1000-nt query, +1/-3, gaps 5/2,
    both strands, against the UNMODIFIED reference kernels (oracle/_ref: search7 -> search16 ->
    fullsw) where they are built, else the C oracle."""
    import os
    import oracle_lib
    q = synth.dna_query(1000)
    residues, offsets = synth.dna_db_planted(1_000_000, q, seed=31, plant_every=500, ambiguity_every=7)
    m = scoring.nucleotide_matrix(1, -3)
    threads = os.cpu_count() or 1
    if oracle_lib.ref_available():
        ref = oracle_lib.Ref()
        ref.matrix_init("x", symtype=0, match=1, mismatch=-3)

        def cpu(qq):
            return ref.scan(residues, offsets, qq, 5, 2, threads=threads, chunk=4096, ssse3=1)[0]
    else:
        orc = oracle_lib.Oracle()

        def cpu(qq):
            return orc.scan(residues, offsets, qq, m, 5, 2, threads=threads)[0]
    with Database(residues, offsets) as db:
        for name, qq in (("plus", q), ("minus", synth.revcomp_nt(q))):
            got = db.search(qq, Scoring(m, 5, 2))
            exp = cpu(qq)
            bad = np.nonzero(got != exp)[0]
            assert bad.size == 0, "%s strand: %d scores differ, first %s" % (name, bad.size, bad[:5])
            assert got.max() > 100                      # the planted copies are found


@pytest.mark.parametrize("qlen", [1, 40, 256, 257, 600, 1024, 1025, 2500])
def test_end_cells_every_query_length(oracle, qlen):
    """swb_search_end (search16s's contract, search16s.cc:390-405) through the warp-per-subject kernel:
    every strip width (8 / 16 / 32 rows per lane), queries beyond 1024 rows in passes, and subjects that
    repeat a piece of the query so that the maximum is reached in several cells (the first column, then the
    smallest row, must be reported)."""
    rng = np.random.default_rng(qlen)
    q = synth.protein_query(qlen, seed=700 + qlen)
    subs = []
    for i in range(60):
        L = int(rng.integers(1, 1500))
        s = synth.random_protein(rng, L)
        if i % 3 == 0 and qlen >= 8:
            w = int(rng.integers(4, min(qlen, 200) + 1))
            a = int(rng.integers(0, qlen - w + 1))
            piece = q[a:a + w]
            s = np.concatenate([s[:L // 3], piece, s[L // 3: L // 2], piece, s[L // 2:]])    # the same best score twice
        subs.append(s)
    subs.append(q.copy())
    subs.append(np.concatenate([q, q]))
    residues, offsets = fixtures.pack(subs)
    sc = Scoring(B62, 11, 1)
    sel = np.arange(len(subs))
    with Database(residues, offsets) as db:
        s, bp, bq = db.search_end(q, sc, sel)
    for k in sel:
        es, ed, eq = oracle.score_end(residues[offsets[k]:offsets[k + 1]], q, B62, 11, 1)
        assert (s[k], bp[k], bq[k]) == (es, ed, eq), "subject %d (len %d)" % (k, offsets[k + 1] - offsets[k])


@pytest.mark.parametrize("geometry", [1, 2])
def test_multi_pass_scratch_sized_for_resident_ctas(oracle, monkeypatch, geometry):
    """Shards whose full pass-boundary scratch (32 B per block) would not fit use one region per resident
    CTA, claimed when the CTA starts and released when it ends (ScanParams::slot_flags).  Forced here with a
    zero budget, over several chunks so that regions are reused by later CTAs."""
    monkeypatch.setenv("SWB_BND_BUDGET_MB", "0")
    monkeypatch.setenv("SWB_CHUNK_BYTES", "150000")
    q = synth.protein_query(1100, seed=20261017 + 1100)
    residues, offsets = synth.protein_db(4000, query=q, seed=1102, plant_every=40, max_len=1400)
    exp = oracle.scan(residues, offsets, q, B62, 11, 1)[0]
    with Database(residues, offsets) as db:
        db.set_geometry(geometry)
        got = db.search(q, Scoring(B62, 11, 1))
        assert db.last_counters["scan_passes"] > 1
    assert np.array_equal(got, exp)


def test_device_sink_wide_cells(oracle):
    """A scoring system that needs 64-bit cells (penalties beyond 2^30) cannot use the 32-bit score half
    of a candidate key: swb_search_hits then reads the scores back and runs the host sink inside the
    library -- same list."""
    q = synth.protein_query(60, seed=81)
    residues, offsets = synth.protein_db(300, query=q, seed=82, plant_every=10, max_len=200)
    sc = Scoring(B62, 2 ** 31, 2 ** 31)
    exp = oracle.scan(residues, offsets, q, B62, 2 ** 31, 2 ** 31)[0]
    oseq, osc, otot, oobv = oracle.topk(np.arange(exp.size), exp, 20, min_score=1)
    with Database(residues, offsets) as db:
        seq, s, tot, obv = db.search_hits(q, sc, 20, 1)
    assert np.array_equal(seq, oseq) and np.array_equal(s, osc) and (tot, obv) == (otot, oobv)


def test_device_sink_subject_filter(oracle):
    """swb_db_set_filter: subjects whose bit is clear are scored but never enter the hit list nor the
    totalhits / obvious counts (db_check_inclusion before the kernels, swipe.cc:1373-1376) -- the list
    equals the oracle's sink over the included subjects only; dense scores ignore the filter; NULL
    removes it.  Narrow cells (device sink) and 64-bit cells (host sink inside the library)."""
    q = synth.protein_query(120, seed=91)
    residues, offsets = synth.protein_db(2500, query=q, seed=92, plant_every=30, max_len=400)
    rng = np.random.default_rng(93)
    include = rng.random(2500) < 0.4
    include[::30] = rng.random(include[::30].size) < 0.5          # planted subjects on both sides
    idx = np.flatnonzero(include)
    for open_, ext in ((11, 1), (2 ** 31, 2 ** 31)):
        sc = Scoring(B62, open_, ext)
        exp = oracle.scan(residues, offsets, q, B62, open_, ext)[0]
        want = oracle.topk(idx, exp[idx], 25, min_score=1, upper=200)
        everything = oracle.topk(np.arange(exp.size), exp, 25, min_score=1, upper=200)
        with Database(residues, offsets) as db:
            db.set_filter(include)
            seq, s, tot, obv = db.search_hits(q, sc, 25, 1, 200)
            assert np.array_equal(seq, want[0]) and np.array_equal(s, want[1]) and (tot, obv) == want[2:]
            assert db.last_counters["ref_width7"] + db.last_counters["ref_width16"] \
                + db.last_counters["ref_width63"] == exp.size
            assert np.array_equal(db.search(q, sc), exp)
            if open_ == 11:
                got = db.search_hits_batch([q, q[:50]], sc, 25, 1, 200)
                assert np.array_equal(got[0][0], want[0]) and np.array_equal(got[0][1], want[1])
            db.set_filter(None)
            seq, s, tot, obv = db.search_hits(q, sc, 25, 1, 200)
            assert np.array_equal(seq, everything[0]) and (tot, obv) == everything[2:]
