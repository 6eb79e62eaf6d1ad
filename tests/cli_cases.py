"""TEST INFRASTRUCTURE: the command-line parity cases.  Builds the BLAST-file fixtures (seeded, so
the reference program and ours read byte-identical files) and lists the command lines whose output
is compared: tests/golden/cli_out/<name>.txt holds what the UNMODIFIED reference program printed
(tests/golden/make_golden_cli_out.py), tests/test_gpu_cli.py runs swipe-b200 on the same files."""
import os
import re

import numpy as np

import blastdb
import fixtures
from swipe_b200 import scoring, synth

CODON = {"A": "GCT", "R": "CGT", "N": "AAT", "D": "GAT", "C": "TGT", "Q": "CAA", "E": "GAA", "G": "GGT",
         "H": "CAT", "I": "ATT", "L": "CTT", "K": "AAA", "M": "ATG", "F": "TTT", "P": "CCT", "S": "TCT",
         "T": "ACT", "W": "TGG", "Y": "TAT", "V": "GTT", "B": "GAT", "Z": "GAA", "X": "GCT", "U": "TGT",
         "O": "AAA", "J": "CTT", "*": "TAA", "-": "GCT"}
NT = {"A": 1, "C": 2, "G": 4, "T": 8}


def back_translate(codes):
    text = "".join(CODON[scoring.SYM_AA[int(c)]] for c in codes)
    return np.array([NT[c] for c in text], dtype=np.uint8)


def random_nt(rng, n):
    return (1 << rng.integers(0, 4, size=n)).astype(np.uint8)


def build(tmp):
    """Writes every fixture under tmp; returns nothing (names are fixed)."""
    q = synth.protein_query(375)
    subs = fixtures.blast_protein_subjects(q)
    cut = len(subs) // 3
    blastdb.write_protein(os.path.join(tmp, "p"), subs)
    blastdb.write_protein(os.path.join(tmp, "v0"), subs[:cut])
    blastdb.write_protein(os.path.join(tmp, "v1"), subs[cut:])
    open(os.path.join(tmp, "pa.pal"), "w").write("TITLE two volumes\nDBLIST v0 v1\n")
    blastdb.write_fasta(os.path.join(tmp, "q.fa"), q, name="query one protein")
    # two queries in one file
    q2 = np.concatenate([subs[30][:60], synth.protein_query(40, seed=5), subs[31][10:90]])
    with open(os.path.join(tmp, "q2.fa"), "w") as f:
        for name, codes in (("first", q[:200]), ("second some text", q2)):
            f.write(">%s\n%s\n" % (name, "".join(scoring.SYM_AA[int(c)] for c in codes)))
    # nucleotide database and query
    qn = synth.dna_query(600, seed=77)
    nsubs = fixtures.blast_nt_subjects(qn)
    blastdb.write_nucleotide(os.path.join(tmp, "n"), nsubs)
    blastdb.write_fasta(os.path.join(tmp, "qn.fa"), qn, protein=False, name="ntquery")
    # translated searches: a nucleotide query that encodes pieces of protein subjects (blastx),
    # and a nucleotide database whose sequences encode pieces of the protein query (tblastn/tblastx)
    rng = np.random.default_rng(2026)
    piece = np.concatenate([random_nt(rng, 31), back_translate(subs[25][20:150]), random_nt(rng, 17),
                            synth.revcomp_nt(back_translate(subs[24][5:120])), random_nt(rng, 8)])
    blastdb.write_fasta(os.path.join(tmp, "qx.fa"), piece, protein=False, name="ntquery coding")
    tsubs = []
    for i in range(60):
        L = int(rng.integers(30, 700))
        s = random_nt(rng, L)
        if i % 3 == 0:
            a = int(rng.integers(0, 250))
            frag = back_translate(q[a:a + int(rng.integers(20, 110))])
            if i % 2:
                frag = synth.revcomp_nt(frag)
            off = int(rng.integers(0, 3))
            s = np.concatenate([random_nt(rng, off + 3 * int(rng.integers(0, 20))), frag, s[:40]])
        if i % 10 == 0 and len(s) > 30:
            s[7:12] = 15
        tsubs.append(s)
    tsubs.append(random_nt(rng, 2))
    tsubs.append(random_nt(rng, 4))
    blastdb.write_nucleotide(os.path.join(tmp, "t"), tsubs)
    # deflines with gi numbers, taxids and membership bits; some sequences carry two deflines
    nx = 45
    headers = []
    for i in range(nx):
        tax = (9606, 10090, 562)[i % 3]
        d = [{"title": "protein %d of organism %d" % (i, tax), "gi": 1000 + i, "lcl": "x%d" % i, "taxid": tax,
              "memb": 1 if i % 4 == 0 else 0}]
        if i % 5 == 0:
            d.append({"title": "identical protein, other source", "gi": 5000 + i, "lcl": "y%d" % i,
                      "taxid": 7227, "memb": 1 if i % 4 == 0 else 0})
        headers.append(d)
    blastdb.write_protein(os.path.join(tmp, "px"), subs[:nx], title="annotated db", headers=headers)
    open(os.path.join(tmp, "tax.txt"), "w").write("10090\n7227\n")
    keep = np.array([i % 4 == 0 for i in range(nx)])
    open(os.path.join(tmp, "mx.msk"), "wb").write(b"\0\0\0\0" + np.packbits(keep).tobytes())
    open(os.path.join(tmp, "pm.pal"), "w").write(
        "TITLE masked subset\nDBLIST px\nOIDLIST mx.msk\nMEMB_BIT 1\nNSEQ %d\nLENGTH %d\nMAXOID %d\n" % (
            int(keep.sum()), int(sum(len(subs[i]) for i in range(nx) if keep[i])), nx - 1))
    # -p 5: the "sound" alphabet (A-Z, a-e = codes 1..31) over a protein-type database
    srng = np.random.default_rng(555)
    sq = srng.integers(1, 32, size=120).astype(np.uint8)
    ssubs = []
    for i in range(40):
        L = int(srng.integers(5, 200))
        s_ = srng.integers(1, 32, size=L).astype(np.uint8)
        if i % 4 == 0 and L > 30:
            w = int(min(L, 60))
            a = int(srng.integers(0, 120 - w))
            s_[:w] = sq[a:a + w]
            s_[w // 2] = (s_[w // 2] % 31) + 1
        ssubs.append(s_)
    blastdb.write_protein(os.path.join(tmp, "snd"), ssubs, title="sound db")
    sound = "-ABCDEFGHIJKLMNOPQRSTUVWXYZabcde"
    with open(os.path.join(tmp, "qs.fa"), "w") as f:
        f.write(">soundquery\n%s\n" % "".join(sound[int(c)] for c in sq))
    # a custom matrix without statistics
    m = fixtures.asym_matrix().reshape(32, 32)
    letters = scoring.SYM_AA[1:28]
    lines = ["# asymmetric test matrix", "   " + "  ".join(letters)]
    for a in range(1, 28):
        lines.append(scoring.SYM_AA[a] + " " + " ".join("%2d" % m[a, b] for b in range(1, 28)))
    open(os.path.join(tmp, "asym.mat"), "w").write("\n".join(lines) + "\n")


CASES = {
    "protein_plain": "-d p -i q.fa -v 20 -b 8",
    "protein_tsv": "-d p -i q.fa -m 8 -b 40",
    "protein_tsv_comments": "-d p -i q.fa -m 9 -b 12 -v 5",
    "protein_xml": "-d p -i q.fa -m 7 -v 15 -b 5",
    "protein_blosum50": "-d p -i q.fa -M BLOSUM50 -G 10 -E 2 -m 8 -b 25",
    "protein_pam30_defaults": "-d p -i q.fa -M PAM30 -v 10 -b 3",
    "protein_volumes": "-d pa -i q.fa -m 8 -b 30",
    "protein_two_queries": "-d p -i q2.fa -v 8 -b 4",
    "protein_thresholds": "-d p -i q.fa -e 1e-5 -m 8",
    "protein_minscore_nolimit": "-d p -i q.fa -e 1e30 -c 30 -u 600 -m 7 -b 0 -v 1000",
    "protein_custom_matrix": "-d p -i q.fa -M asym.mat -G 7 -E 2 -v 12 -b 4",
    "protein_effdbsize": "-d p -i q.fa -z 5000000 -m 8 -b 10",
    "protein_min_evalue": "-d p -i q.fa -k 1e-100 -e 1e30 -m 8 -b 12",
    "protein_long_options": "--db=p --query=q.fa --matrix=BLOSUM80 --gapopen=10 --gapextend=1 --outfmt=8 --num_alignments=9 --evalue=1e-3",
    "annotated_plain": "-d px -i q.fa -e 1e30 -v 12 -b 3 -I -H",
    "annotated_tsv": "-d px -i q.fa -e 1e30 -m 8 -b 45",
    "taxid_list": "-d px -i q.fa -e 1e30 -x tax.txt -m 8 -b 45",
    "taxid_list_plain": "-d px -i q.fa -e 1e30 -x tax.txt -v 30 -b 2 -H",
    "masked_alias": "-d pm -i q.fa -e 1e30 -v 30 -b 2",
    "masked_alias_tsv": "-d pm -i q.fa -m 8 -b 45",
    "dump_protein": "-d px -N 1",
    "dump_protein_split": "-d px -N 2",
    "dump_nt": "-d n -p 0 -N 1",
    "paralign_protein": "-d px -i q2.fa -m 99 -e 1e30 -v 6 -b 2",
    "paralign_nt": "-d n -i qn.fa -p 0 -m 99 -v 5 -b 2",
    "paralign_tblastx": "-d t -i qx.fa -p 4 -m 99 -v 4 -b 2 -e 100",
    "paralign_taxid_masked": "-d pm -i q.fa -m 99 -x tax.txt -e 1e30 -v 5 -b 1",
    "sound_plain": "-d snd -i qs.fa -p 5 -v 10 -b 3",
    "sound_xml": "-d snd -i qs.fa -p sound -m 7 -v 12 -b 4 -G 8 -E 2",
    "nt_plain": "-d n -i qn.fa -p 0 -v 20 -b 10",
    "nt_tsv": "-d n -i qn.fa -p 0 -m 8 -b 60",
    "nt_plus_only": "-d n -i qn.fa -p 0 -S 1 -m 8 -b 30",
    "nt_minus_only": "-d n -i qn.fa -p 0 -S 2 -m 8 -b 30",
    "nt_2_3": "-d n -i qn.fa -p 0 -r 2 -q -3 -G 5 -E 2 -m 8 -b 20",
    "blastx_plain": "-d p -i qx.fa -p 2 -v 10 -b 6",
    "blastx_tsv": "-d p -i qx.fa -p 2 -m 8 -b 20 -e 1000",
    "tblastn_plain": "-d t -i q.fa -p 3 -v 12 -b 8",
    "tblastn_tsv": "-d t -i q.fa -p 3 -m 8 -b 40 -e 1000",
    "tblastx_tsv": "-d t -i qx.fa -p 4 -m 8 -b 30 -e 100",
    "tblastx_plain": "-d t -i qx.fa -p 4 -v 6 -b 3 -Q 4 -D 11",
}

# lines that legitimately differ: program banner, wall-clock lines, the thread count's meaning
DROP = re.compile(r"^(SWIPE|Reference:|with inter-sequence|Score-only Smith-Waterman|SWIPE:|# SWIPE|"
                  r"Search started:|Search completed:|Elapsed:|Speed:|Threads:|"
                  r"\s*<(searchStarted|searchCompleted|searchElapsedTime|searchSpeed|programVersion|threads)>)")


def normalise(text):
    lines = [ln.rstrip() for ln in text.splitlines() if not DROP.match(ln)]
    while lines and not lines[0]:
        lines.pop(0)
    return "\n".join(lines) + "\n"
