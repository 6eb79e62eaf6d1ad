"""Deterministic synthetic inputs of the shape BASELINE.json / SURVEY.md 8(d) names: a
background-frequency protein query, a log-normal-length protein database with planted mutated
copies of query windows, and a uniform ACGT read set.  Used by tests and bench.py only."""
import numpy as np

from .scoring import SYM_AA, map_aa

# Robinson-Robinson-like background frequencies
AA_FREQ = {"A": .078, "R": .051, "N": .045, "D": .054, "C": .019, "Q": .043, "E": .063,
           "G": .074, "H": .022, "I": .051, "L": .090, "K": .057, "M": .022, "F": .039,
           "P": .052, "S": .071, "T": .058, "W": .013, "Y": .032, "V": .064}


def _aa_lut():
    letters = sorted(AA_FREQ)
    p = np.array([AA_FREQ[c] for c in letters], dtype=np.float64)
    p /= p.sum()
    edges = np.floor(np.cumsum(p) * 65536.0 + 0.5).astype(np.int64)
    lut = np.zeros(65536, dtype=np.uint8)
    lo = 0
    for c, hi in zip(letters, edges):
        lut[lo:hi] = map_aa(c)
        lo = hi
    lut[lo:] = map_aa(letters[-1])
    return lut


def random_protein(rng, n):
    """n residues i.i.d. from the background frequencies, as NCBIstdaa codes."""
    return _aa_lut()[rng.integers(0, 65536, size=n, dtype=np.uint16)]


def protein_query(qlen, seed=20261017):
    return random_protein(np.random.default_rng(seed), qlen)


def protein_db(nseq, query=None, seed=20261018, plant_every=1000, mu=5.65, sigma=0.65,
               min_len=25, max_len=5000):
    """(residues uint8[total], offsets int64[nseq+1]); lengths clamp(round(exp(N(mu, sigma^2)))).
    Every plant_every-th subject carries a mutated copy of a random query window (identity
    30-95 %, occasional indels) so that a small fraction re-queues to the wider kernels."""
    rng = np.random.default_rng(seed)
    lens = np.clip(np.rint(np.exp(rng.normal(mu, sigma, size=nseq))), min_len, max_len).astype(np.int64)
    offsets = np.zeros(nseq + 1, dtype=np.int64)
    np.cumsum(lens, out=offsets[1:])
    total = int(offsets[-1])
    residues = np.empty(total, dtype=np.uint8)
    chunk = 1 << 26
    lut = _aa_lut()
    for s in range(0, total, chunk):
        e = min(total, s + chunk)
        residues[s:e] = lut[rng.integers(0, 65536, size=e - s, dtype=np.uint16)]
    if query is not None and plant_every and len(query) >= 8:
        q = np.asarray(query, dtype=np.uint8)
        for i in range(plant_every - 1, nseq, plant_every):
            L = int(lens[i])
            w = int(rng.integers(8, min(len(q), L) + 1))
            qs = int(rng.integers(0, len(q) - w + 1))
            piece = q[qs:qs + w].copy()
            ident = rng.uniform(0.30, 0.95)
            mut = rng.random(w) > ident
            piece[mut] = random_protein(rng, int(mut.sum()))
            if w > 20 and rng.random() < 0.3:           # an occasional deletion
                cut = int(rng.integers(1, 6))
                at = int(rng.integers(5, w - cut - 5))
                piece = np.concatenate([piece[:at], piece[at + cut:]])
            ds = int(rng.integers(0, L - len(piece) + 1))
            residues[offsets[i] + ds: offsets[i] + ds + len(piece)] = piece
    return residues, offsets


def dna_query(qlen=1000, seed=20261019):
    """One-hot nt codes A=1 C=2 G=4 T=8 (database.cc:915-921)."""
    rng = np.random.default_rng(seed)
    return (1 << rng.integers(0, 4, size=qlen)).astype(np.uint8)


def dna_db(nseq, seed=20261020, min_len=150, max_len=250):
    rng = np.random.default_rng(seed)
    lens = rng.integers(min_len, max_len + 1, size=nseq).astype(np.int64)
    offsets = np.zeros(nseq + 1, dtype=np.int64)
    np.cumsum(lens, out=offsets[1:])
    total = int(offsets[-1])
    residues = np.empty(total, dtype=np.uint8)
    chunk = 1 << 26
    lut = np.array([1, 2, 4, 8], dtype=np.uint8)
    for s in range(0, total, chunk):
        e = min(total, s + chunk)
        residues[s:e] = lut[rng.integers(0, 4, size=e - s, dtype=np.uint8)]
    return residues, offsets


def dna_db_planted(nseq, query, seed=20261020, plant_every=1000, ambiguity_every=0):
    """dna_db with a mutated (5 %) copy of a query window in every plant_every-th read, forward and
    reverse-complemented in turn, and -- every ambiguity_every-th planted read -- a run of N (code 15)
    inside it (the DNA workload of BASELINE configs[3])."""
    residues, offsets = dna_db(nseq, seed=seed)
    q = np.asarray(query, dtype=np.uint8)
    rng = np.random.default_rng(seed + 24)
    lens = offsets[1:] - offsets[:-1]
    for n, i in enumerate(range(0, nseq, plant_every)):
        w = int(min(lens[i], 140, q.size))
        s0 = int(rng.integers(0, q.size - w + 1))
        piece = q[s0:s0 + w].copy()
        if n % 2:
            piece = revcomp_nt(piece)
        mut = rng.random(w) < 0.05
        piece[mut] = 1 << rng.integers(0, 4, size=int(mut.sum()))
        if ambiguity_every and n % ambiguity_every == 0 and w > 20:
            piece[10:14] = 15
        residues[offsets[i]: offsets[i] + w] = piece
    return residues, offsets


def revcomp_nt(codes):
    """Reverse complement of one-hot/ambiguity nt codes: bit-reverse the 4-bit code
    (query.cc:112 ntcompl) and reverse the order."""
    c = np.asarray(codes, dtype=np.uint8)
    rev = ((c & 1) << 3) | ((c & 2) << 1) | ((c & 4) >> 1) | ((c & 8) >> 3)
    return rev[::-1].copy()
