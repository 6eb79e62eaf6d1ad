"""The library's BLAST-v4 reader (swb_blastdb_*, replacing database.cc's db_open / db_getsequence)
against databases written by tests/blastdb.py -- the same bytes the unmodified reference CLI read
when tests/golden/cli_*.npz was generated -- and the oracle against those CLI scores."""
import os

import numpy as np
import pytest

import blastdb
import fixtures
from swipe_b200 import BlastDB, SwbError, scoring, synth

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _protein(tmp_path):
    g = np.load(os.path.join(GOLD, "cli_protein.npz"))
    q = synth.protein_query(375)
    assert np.array_equal(q, g["query"])
    subs = fixtures.blast_protein_subjects(q)
    return g, q, subs


def _nt(tmp_path):
    g = np.load(os.path.join(GOLD, "cli_nt.npz"))
    q = synth.dna_query(600, seed=77)
    assert np.array_equal(q, g["query"])
    return g, q, fixtures.blast_nt_subjects(q)


def test_protein_reader(tmp_path):
    g, q, subs = _protein(tmp_path)
    base = str(tmp_path / "p")
    blastdb.write_protein(base, subs, title="my title")
    with BlastDB(base) as db:
        assert (db.nseq, db.volumes, db.title) == (len(subs), 1, "my title")
        assert db.symbols == sum(len(s) for s in subs) and db.longest == max(len(s) for s in subs)
        for i in (0, 1, 17, len(subs) - 1):
            assert db.seqlen(i) == len(subs[i])
            assert np.array_equal(db.sequence(i), subs[i])
            assert b"s%d" % i in db.header(i)
            assert db.included(i)
        with pytest.raises(SwbError):
            db.seqlen(len(subs))


def test_alias_volumes_and_mask(tmp_path):
    g, q, subs = _protein(tmp_path)
    cut = int(g["volume_cut"][0])
    blastdb.write_protein(str(tmp_path / "v0"), subs[:cut])
    blastdb.write_protein(str(tmp_path / "v1"), subs[cut:])
    (tmp_path / "pa.pal").write_text("TITLE two volumes\nDBLIST v0 \"v1\"\n")
    with BlastDB(str(tmp_path / "pa")) as db:
        assert (db.nseq, db.volumes, db.title) == (len(subs), 2, "two volumes")
        for i in (0, cut - 1, cut, len(subs) - 1):
            assert np.array_equal(db.sequence(i), subs[i])
    # membership mask: OIDLIST names a bitmap, MSB first, after a 4-byte header (database.cc:687-706)
    keep = np.zeros(cut, dtype=bool)
    keep[::3] = True
    bits = np.packbits(keep)
    (tmp_path / "m.msk").write_bytes(b"\0\0\0\0" + bits.tobytes())
    (tmp_path / "pm.pal").write_text("TITLE masked\nDBLIST v0\nOIDLIST m.msk\nMEMB_BIT 1\nMAXOID %d\n" % (cut - 1))
    with BlastDB(str(tmp_path / "pm")) as db:
        assert [db.included(i) for i in range(cut)] == keep.tolist()


def test_nucleotide_reader_decodes_like_db_getsequence(tmp_path):
    g, q, subs = _nt(tmp_path)
    for big in (False, True):
        base = str(tmp_path / ("n%d" % big))
        blastdb.write_nucleotide(base, subs, big_table=big)
        with BlastDB(base, nucleotide=True) as db:
            assert db.nseq == len(subs) and db.longest == max(len(s) for s in subs)
            for i, s in enumerate(subs):
                assert db.seqlen(i) == len(s)
                assert np.array_equal(db.sequence(i), s), i
                assert np.array_equal(db.sequence(i, strand=1), synth.revcomp_nt(s)), i


def test_open_errors(tmp_path):
    with pytest.raises(SwbError) as e:
        BlastDB(str(tmp_path / "missing"))
    assert e.value.status == -7 and "Unable to open file" in str(e.value)
    g, q, subs = _protein(tmp_path)
    blastdb.write_protein(str(tmp_path / "p"), subs[:5])
    with pytest.raises(SwbError):
        BlastDB(str(tmp_path / "p"), nucleotide=True)          # no .nin
    raw = (tmp_path / "p.pin").read_bytes()
    (tmp_path / "p.pin").write_bytes(b"\0\0\0\5" + raw[4:])
    with pytest.raises(SwbError) as e:
        BlastDB(str(tmp_path / "p"))
    assert "version" in str(e.value)
    (tmp_path / "p.pin").write_bytes(raw[:40])
    with pytest.raises(SwbError):
        BlastDB(str(tmp_path / "p"))


def test_oracle_matches_reference_cli(oracle):
    """Pins the oracle (and the fixtures) to the reference program end to end: database files ->
    db_getsequence -> search7/16/63 -> hits_enter order."""
    g, q, subs = _protein(None)
    res, off = fixtures.pack(subs)
    mats = np.load(os.path.join(GOLD, "matrices.npz"))
    for name, go, ge in (("blosum62", 11, 1), ("blosum50", 10, 2)):
        m = mats[name].astype(np.int64)
        exp, _, _ = oracle.scan(res, off, q, m, go, ge)
        assert np.array_equal(exp, g["scores_%s_%d_%d" % (name, go, ge)])
        order = np.lexsort((-np.arange(exp.size), -exp))         # score desc, seqno desc
        assert np.array_equal(order, g["order_%s_%d_%d" % (name, go, ge)])
    gn, qn, nsubs = _nt(None)
    res, off = fixtures.pack(nsubs)
    m = scoring.nucleotide_matrix(1, -3)
    a, _, _ = oracle.scan(res, off, qn, m, 5, 2)
    b, _, _ = oracle.scan(res, off, synth.revcomp_nt(qn), m, 5, 2)
    assert np.array_equal(np.sort(np.stack([a, b], 1), 1), gn["strand_scores_sorted"])


def test_vectorised_writers_equal_the_per_sequence_ones(tmp_path):
    """bench.py's reference arm writes its sample with the vectorised writers: they must produce the very
    files the per-sequence writers (the ones pinned against the reference program) produce."""
    q = synth.protein_query(120)
    res, off = synth.protein_db(700, query=q, seed=5, plant_every=50, max_len=400)
    subs = [res[off[i]:off[i + 1]] for i in range(700)]
    a, b = str(tmp_path / "slow"), str(tmp_path / "fast")
    s1 = blastdb.write_protein(a, subs)
    s2 = blastdb.write_protein_fast(b, res, off)
    assert np.array_equal(s1, s2)
    assert open(a + ".psq", "rb").read() == open(b + ".psq", "rb").read()
    with BlastDB(b) as db:
        assert db.nseq == 700 and db.symbols == int(off[-1])
        assert np.array_equal(db.sequence(123), subs[123])
    qn = synth.dna_query(300)
    r, o = synth.dna_db_planted(900, qn, seed=6, plant_every=40, ambiguity_every=2)
    nsubs = [r[o[i]:o[i + 1]] for i in range(900)]
    a, b = str(tmp_path / "nslow"), str(tmp_path / "nfast")
    s1, a1 = blastdb.write_nucleotide(a, nsubs)
    s2, a2 = blastdb.write_nucleotide_fast(b, r, o, chunk=250)
    assert np.array_equal(s1, s2) and np.array_equal(a1, a2)
    assert open(a + ".nsq", "rb").read() == open(b + ".nsq", "rb").read()
    with BlastDB(b, nucleotide=True) as db:
        for i in (0, 40, 80, 899):
            assert np.array_equal(db.sequence(i), nsubs[i])
