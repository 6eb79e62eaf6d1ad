"""GPU: shards uploaded straight from BLAST database files (swb_db_open_blast: .psq bytes copied
as they lie, .nsq unpacked on the device) must give the scores the unmodified reference program
printed for the same files (tests/golden/cli_*.npz, made by tests/golden/make_golden_cli.py)."""
import os

import numpy as np
import pytest

import blastdb
import fixtures
from swipe_b200 import BlastDB, Scoring, scoring, synth

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")


def test_protein_files_match_reference_cli(tmp_path, monkeypatch):
    g = np.load(os.path.join(GOLD, "cli_protein.npz"))
    mats = np.load(os.path.join(GOLD, "matrices.npz"))
    q = g["query"]
    subs = fixtures.blast_protein_subjects(q)
    n = len(subs)
    cut = int(g["volume_cut"][0])
    blastdb.write_protein(str(tmp_path / "p"), subs)
    blastdb.write_protein(str(tmp_path / "v0"), subs[:cut])
    blastdb.write_protein(str(tmp_path / "v1"), subs[cut:])
    (tmp_path / "pa.pal").write_text("TITLE two volumes\nDBLIST v0 v1\n")
    for base in ("p", "pa"):
        for chunk in ("300000000", "20000"):
            monkeypatch.setenv("SWB_CHUNK_BYTES", chunk)
            with BlastDB(str(tmp_path / base)) as bdb:
                for wait in (True, False):
                    with bdb.upload(wait=wait) as db:
                        assert db.info() == {"nseq": n, "residues": bdb.symbols, "longest": bdb.longest}
                        for name, go, ge in (("blosum62", 11, 1), ("blosum50", 10, 2)):
                            got = db.search(q, Scoring(mats[name].astype(np.int64), go, ge))
                            assert np.array_equal(got, g["scores_%s_%d_%d" % (name, go, ge)]), (base, chunk, wait)
                # a shard that starts and ends inside different volumes
                first, count = cut - 7, 40
                with bdb.upload(first=first, count=count) as db:
                    got = db.search(q, Scoring(mats["blosum62"].astype(np.int64), 11, 1))
                    assert np.array_equal(got, g["scores_blosum62_11_1"][first:first + count])
                    db.set_mode(2)                     # the wide kernel reads the raw residue buffer
                    got = db.search(q, Scoring(mats["blosum62"].astype(np.int64), 11, 1))
                    assert np.array_equal(got, g["scores_blosum62_11_1"][first:first + count])


def test_nucleotide_files_match_reference_cli(tmp_path, monkeypatch):
    g = np.load(os.path.join(GOLD, "cli_nt.npz"))
    q = g["query"]
    subs = fixtures.blast_nt_subjects(q)
    sc = Scoring(scoring.nucleotide_matrix(1, -3), 5, 2)
    for big in (False, True):
        base = str(tmp_path / ("n%d" % big))
        blastdb.write_nucleotide(base, subs, big_table=big)
        for chunk in ("300000000", "3000"):
            monkeypatch.setenv("SWB_CHUNK_BYTES", chunk)
            with BlastDB(base, nucleotide=True) as bdb:
                for wait in (True, False):
                    with bdb.upload(wait=wait) as db:
                        a = db.search(q, sc)
                        b = db.search(synth.revcomp_nt(q), sc)
                        got = np.sort(np.stack([a, b], 1), 1)
                        assert np.array_equal(got, g["strand_scores_sorted"]), (big, chunk, wait)
                        db.set_mode(2)                 # wide kernel: reads the device-decoded residues
                        a2 = db.search(q, sc)
                        assert np.array_equal(a2, a)
                with bdb.upload(first=5, count=50) as db:
                    a = db.search(q, sc)
                    b = db.search(synth.revcomp_nt(q), sc)
                    assert np.array_equal(np.sort(np.stack([a, b], 1), 1), g["strand_scores_sorted"][5:55])


def test_device_decode_equals_host_decode(tmp_path, oracle):
    """Larger random nt database with long ambiguity runs: GPU scores of the device-decoded shard
    against the oracle run on the reader's host decode (db_getsequence semantics)."""
    rng = np.random.default_rng(99)
    q = synth.dna_query(200, seed=5)
    subs = []
    for i in range(2000):
        L = int(rng.integers(1, 900))
        s = (1 << rng.integers(0, 4, size=L)).astype(np.uint8)
        if i % 5 == 0 and L > 100:
            a = int(rng.integers(0, L - 90))
            s[a:a + int(rng.integers(1, 90))] = int(rng.choice([15, 5, 10]))
        if i % 50 == 0 and L > 120:
            s[:100] = q[50:150]
        subs.append(s)
    base = str(tmp_path / "big")
    blastdb.write_nucleotide(base, subs, big_table=True)
    sc = Scoring(scoring.nucleotide_matrix(2, -3), 5, 2)
    with BlastDB(base, nucleotide=True) as bdb:
        host = [bdb.sequence(i) for i in range(len(subs))]
        res, off = fixtures.pack(host)
        exp, _, _ = oracle.scan(res, off, q, sc.matrix, 5, 2)
        with bdb.upload() as db:
            assert np.array_equal(db.search(q, sc), exp)
