"""Seeded inputs shared by the CPU and GPU tests."""
import numpy as np

from swipe_b200 import scoring, synth


def pack(subjects):
    """list of uint8 arrays -> (residues, offsets)"""
    offsets = np.zeros(len(subjects) + 1, dtype=np.int64)
    if subjects:
        np.cumsum([len(s) for s in subjects], out=offsets[1:])
    residues = np.concatenate([np.asarray(s, dtype=np.uint8) for s in subjects] +
                              [np.zeros(0, dtype=np.uint8)])
    return residues, offsets


def edge_db(query, seed=7):
    """The edge cases SURVEY.md section 4 lists: lengths 0/1/3/4/5/15/16/17, every protein code
    including B Z X U * O J and '-', exact / partial copies of the query, an odd subject count."""
    rng = np.random.default_rng(seed)
    q = np.asarray(query, dtype=np.uint8)
    subs = []
    for L in (0, 1, 3, 4, 5, 15, 16, 17, 0, 2, 31, 32, 33, 63, 64, 65, 127, 128, 129):
        subs.append(synth.random_protein(rng, L))
    subs.append(np.arange(0, 28, dtype=np.uint8))                 # every code once
    subs.append(np.arange(27, -1, -1, dtype=np.uint8))
    subs.append(scoring.encode_protein("BZXU*OJ-BZXU*OJ"))
    subs.append(q.copy())                                          # self hit
    subs.append(q[: len(q) // 2].copy())
    subs.append(np.concatenate([synth.random_protein(rng, 40), q[len(q) // 3:], synth.random_protein(rng, 9)]))
    subs.append(np.concatenate([q[:20], synth.random_protein(rng, 5), q[20:]]) if len(q) > 20 else q.copy())
    for _ in range(60):
        subs.append(synth.random_protein(rng, int(rng.integers(1, 400))))
    subs.append(np.zeros(0, dtype=np.uint8))
    subs.append(synth.random_protein(rng, 7))                      # odd count
    if len(subs) % 2 == 0:
        subs.append(synth.random_protein(rng, 11))
    return pack(subs)


def asym_matrix(seed=3):
    """A non-symmetric custom table (CHANGES:28-30) over codes 1..27, -1 elsewhere."""
    rng = np.random.default_rng(seed)
    m = np.full((32, 32), -1, dtype=np.int64)
    m[1:28, 1:28] = rng.integers(-6, 4, size=(27, 27))
    for a in range(1, 28):
        m[a, a] = int(rng.integers(3, 13))
    return m.reshape(-1)


def blast_protein_subjects(query, seed=11):
    """Subjects of the BLAST-file fixtures (two volumes' worth): edge cases + planted copies.
    No zero-length subjects (makeblastdb never writes them)."""
    rng = np.random.default_rng(seed)
    res, off = edge_db(query, seed=seed)
    subs = [res[off[i]:off[i + 1]] for i in range(off.size - 1)]
    subs = [s if s.size else synth.random_protein(rng, 2) for s in subs]
    pr, po = synth.protein_db(150, query=query, seed=seed + 1, plant_every=9, max_len=700)
    subs += [pr[po[i]:po[i + 1]] for i in range(150)]
    return subs


def blast_nt_subjects(query, seed=12):
    """Nucleotide subjects with every feature the .nsq format has: lengths of every residue class
    mod 4, N runs longer than one ambiguity entry, single IUPAC codes, planted forward and
    reverse-complement copies of query windows."""
    rng = np.random.default_rng(seed)
    qlen = len(query)
    subs = []
    for i in range(120):
        L = int(rng.integers(1, 500)) if i >= 8 else i + 1
        s = (1 << rng.integers(0, 4, size=L)).astype(np.uint8)
        if i % 3 == 0 and L > 50:
            w = min(L, 150)
            st = int(rng.integers(0, qlen - w))
            p = np.asarray(query[st:st + w]).copy()
            if i % 2 == 0:
                p = synth.revcomp_nt(p)
            s[:w] = p
        if i % 4 == 0 and L > 40:
            a = int(rng.integers(0, L - 35))
            s[a:a + int(rng.integers(1, 35))] = 15
        if i % 7 == 0 and L > 10:
            s[int(rng.integers(0, L))] = int(rng.choice([5, 10, 3, 12, 6, 9, 14, 13, 11, 7]))
        subs.append(s)
    return subs
