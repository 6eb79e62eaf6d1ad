"""Generates tests/golden/cli_*.npz by running the UNMODIFIED reference command-line program
(oracle/_ref/swipe, built from /root/reference by oracle/Makefile) on BLAST databases written by
tests/blastdb.py.  Run in the build container:

    python tests/golden/make_golden_cli.py

`swipe -m 7 -v <all> -b 0 -e 1e30 -c 0` prints <track>seqno</track> and the raw <score> of every
subject (hits.cc:1673-1691); nucleotide searches print one hit per strand.  Recorded: the query,
and per subject the score (protein)
or the sorted pair of strand scores (nucleotide), plus the order of the hit list (score desc,
seqno desc: hits.cc:188-191).
"""
import os
import re
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

import blastdb  # noqa: E402
import fixtures  # noqa: E402
from swipe_b200 import synth  # noqa: E402

SWIPE = os.path.join(os.path.dirname(os.path.dirname(HERE)), "oracle", "_ref", "swipe")


def run(args):
    r = subprocess.run([SWIPE] + args, capture_output=True, text=True, check=True)
    return [(int(t), int(s)) for t, s in
            re.findall(r"<track>(\d+)</track>.*?<score>(\d+)</score>", r.stdout, re.S)]


def main():
    tmp = tempfile.mkdtemp()
    # ---- protein: one volume, and the same subjects split over two volumes behind a .pal
    q = synth.protein_query(375)
    subs = fixtures.blast_protein_subjects(q)
    n = len(subs)
    blastdb.write_protein(os.path.join(tmp, "p"), subs)
    blastdb.write_fasta(os.path.join(tmp, "q.fa"), q)
    rec = {"query": q}
    for name, go, ge in (("BLOSUM62", 11, 1), ("BLOSUM50", 10, 2)):
        hits = run(["-d", os.path.join(tmp, "p"), "-i", os.path.join(tmp, "q.fa"), "-M", name,
                    "-G", str(go), "-E", str(ge), "-m", "7", "-v", str(n), "-b", "0", "-e", "1e30",
                    "-c", "0", "-a", "2"])
        assert len(hits) == n
        sc = np.full(n, -1, dtype=np.int64)
        for t, s in hits:
            sc[t] = s
        rec["scores_%s_%d_%d" % (name.lower(), go, ge)] = sc
        rec["order_%s_%d_%d" % (name.lower(), go, ge)] = np.array([t for t, _ in hits])
    cut = n // 3
    blastdb.write_protein(os.path.join(tmp, "v0"), subs[:cut])
    blastdb.write_protein(os.path.join(tmp, "v1"), subs[cut:])
    open(os.path.join(tmp, "pa.pal"), "w").write("TITLE two volumes\nDBLIST v0 \"v1\"\n")
    hits = run(["-d", os.path.join(tmp, "pa"), "-i", os.path.join(tmp, "q.fa"), "-m", "7", "-v", str(n),
                "-b", "0", "-e", "1e30", "-c", "0", "-a", "1"])
    sc = np.full(n, -1, dtype=np.int64)
    for t, s in hits:
        sc[t] = s
    assert np.array_equal(sc, rec["scores_blosum62_11_1"])        # global seqnos run through the volumes
    rec["volume_cut"] = np.array([cut])
    np.savez_compressed(os.path.join(HERE, "cli_protein.npz"), **rec)

    # ---- nucleotide: old (32-bit) and new (64-bit) ambiguity tables, both strands
    qn = synth.dna_query(600, seed=77)
    nsubs = fixtures.blast_nt_subjects(qn)
    m = len(nsubs)
    blastdb.write_fasta(os.path.join(tmp, "qn.fa"), qn, protein=False)
    rec = {"query": qn}
    for big in (0, 1):
        base = os.path.join(tmp, "n%d" % big)
        blastdb.write_nucleotide(base, nsubs, big_table=bool(big))
        hits = run(["-d", base, "-i", os.path.join(tmp, "qn.fa"), "-p", "0", "-m", "7", "-v", str(2 * m),
                    "-b", "0", "-e", "1e30", "-c", "0", "-a", "1"])
        assert len(hits) == 2 * m
        pairs = [[] for _ in range(m)]
        for t, s in hits:
            pairs[t].append(s)
        arr = np.array([sorted(p) for p in pairs], dtype=np.int64)
        if big == 0:
            rec["strand_scores_sorted"] = arr
        else:
            assert np.array_equal(arr, rec["strand_scores_sorted"])
    np.savez_compressed(os.path.join(HERE, "cli_nt.npz"), **rec)
    print("wrote cli_protein.npz (%d subjects), cli_nt.npz (%d subjects)" % (n, m))


if __name__ == "__main__":
    main()
