"""Host-side scoring inputs for the scan: symbol alphabets, score tables and their limits.

Mirrors what the reference prepares before it calls its kernels:
  symbol codes .......... query.cc:175-179 (sym_ncbi_aa, sym_ncbi_nt16), query.cc:51-109 (maps)
  32x32 table, -1 fill .. matrices.cc:520-538
  matrix text format .... matrices.cc:437-517
  hi/lo and the limits .. matrices.cc:560-577
The tables are plain inputs of the C ABI (swb_scoring.matrix); nothing here computes scores.
"""
import numpy as np

SYM_AA = "-ABCDEFGHIKLMNPQRSTVWXYZU*OJ"      # NCBIstdaa, code = index (query.cc:178)
SYM_NT16 = "-ACMGRSVTWYHKDBN"                # 4-bit one-hot nucleotide codes (query.cc:179)
DIM = 32

# BLOSUM62 as distributed with current NCBI BLAST (public-domain data, includes J), checked against the reference's table in
# tests/test_oracle_pins.py via tests/golden/matrices.json.
BLOSUM62_TEXT = """\
   A  R  N  D  C  Q  E  G  H  I  L  K  M  F  P  S  T  W  Y  V  B  J  Z  X  *
A  4 -1 -2 -2  0 -1 -1  0 -2 -1 -1 -1 -1 -2 -1  1  0 -3 -2  0 -2 -1 -1 -1 -4
R -1  5  0 -2 -3  1  0 -2  0 -3 -2  2 -1 -3 -2 -1 -1 -3 -2 -3 -1 -2  0 -1 -4
N -2  0  6  1 -3  0  0  0  1 -3 -3  0 -2 -3 -2  1  0 -4 -2 -3  4 -3  0 -1 -4
D -2 -2  1  6 -3  0  2 -1 -1 -3 -4 -1 -3 -3 -1  0 -1 -4 -3 -3  4 -3  1 -1 -4
C  0 -3 -3 -3  9 -3 -4 -3 -3 -1 -1 -3 -1 -2 -3 -1 -1 -2 -2 -1 -3 -1 -3 -1 -4
Q -1  1  0  0 -3  5  2 -2  0 -3 -2  1  0 -3 -1  0 -1 -2 -1 -2  0 -2  4 -1 -4
E -1  0  0  2 -4  2  5 -2  0 -3 -3  1 -2 -3 -1  0 -1 -3 -2 -2  1 -3  4 -1 -4
G  0 -2  0 -1 -3 -2 -2  6 -2 -4 -4 -2 -3 -3 -2  0 -2 -2 -3 -3 -1 -4 -2 -1 -4
H -2  0  1 -1 -3  0  0 -2  8 -3 -3 -1 -2 -1 -2 -1 -2 -2  2 -3  0 -3  0 -1 -4
I -1 -3 -3 -3 -1 -3 -3 -4 -3  4  2 -3  1  0 -3 -2 -1 -3 -1  3 -3  3 -3 -1 -4
L -1 -2 -3 -4 -1 -2 -3 -4 -3  2  4 -2  2  0 -3 -2 -1 -2 -1  1 -4  3 -3 -1 -4
K -1  2  0 -1 -3  1  1 -2 -1 -3 -2  5 -1 -3 -1  0 -1 -3 -2 -2  0 -3  1 -1 -4
M -1 -1 -2 -3 -1  0 -2 -3 -2  1  2 -1  5  0 -2 -1 -1 -1 -1  1 -3  2 -1 -1 -4
F -2 -3 -3 -3 -2 -3 -3 -3 -1  0  0 -3  0  6 -4 -2 -2  1  3 -1 -3  0 -3 -1 -4
P -1 -2 -2 -1 -3 -1 -1 -2 -2 -3 -3 -1 -2 -4  7 -1 -1 -4 -3 -2 -2 -3 -1 -1 -4
S  1 -1  1  0 -1  0  0  0 -1 -2 -2  0 -1 -2 -1  4  1 -3 -2 -2  0 -2  0 -1 -4
T  0 -1  0 -1 -1 -1 -1 -2 -2 -1 -1 -1 -1 -2 -1  1  5 -2 -2  0 -1 -1 -1 -1 -4
W -3 -3 -4 -4 -2 -2 -3 -2 -2 -3 -2 -3 -1  1 -4 -3 -2 11  2 -3 -4 -2 -2 -1 -4
Y -2 -2 -2 -3 -2 -1 -2 -3  2 -1 -1 -2 -1  3 -3 -2 -2  2  7 -1 -3 -1 -2 -1 -4
V  0 -3 -3 -3 -1 -2 -2 -3 -3  3  1 -2  1 -1 -2 -2  0 -3 -1  4 -3  2 -2 -1 -4
B -2 -1  4  4 -3  0  1 -1  0 -3 -4  0 -3 -3 -2  0 -1 -4 -3 -3  4 -3  0 -1 -4
J -1 -2 -3 -3 -1 -2 -3 -4 -3  3  3 -3  2  0 -3 -2 -1 -2 -1  2 -3  3 -3 -1 -4
Z -1  0  0  1 -3  4  4 -2  0 -3 -3  1 -1 -3 -1  0 -1 -2 -2 -2  0 -3  4 -1 -4
X -1 -1 -1 -1 -1 -1 -1 -1 -1 -1 -1 -1 -1 -1 -1 -1 -1 -1 -1 -1 -1 -1 -1 -1 -4
* -4 -4 -4 -4 -4 -4 -4 -4 -4 -4 -4 -4 -4 -4 -4 -4 -4 -4 -4 -4 -4 -4 -4 -4  1
"""


def map_aa(ch):
    """Character -> NCBIstdaa code, or -1 when the reference drops it (query.cc:51-69)."""
    if ch == "-":
        return 0
    if ch == "*":
        return 25
    i = SYM_AA.find(ch.upper(), 1)
    return i if i > 0 and ch != "*" else -1


def map_nt16(ch):
    ch = ch.upper()
    if ch == "U":
        ch = "T"
    i = SYM_NT16.find(ch, 1)
    return i if i > 0 else -1


def encode_protein(text):
    codes = [map_aa(c) for c in text]
    return np.array([c for c in codes if c >= 0], dtype=np.uint8)


def encode_nucleotide(text):
    codes = [map_nt16(c) for c in text]
    return np.array([c for c in codes if c >= 0], dtype=np.uint8)


def parse_matrix(text):
    """NCBI matrix text -> int64[32*32] indexed [(row_code << 5) + column_code]; undefined
    pairs score -1 (matrices.cc:531).  The reference reads the database residue as the row and
    the query residue as the column (search63.cc:52-58)."""
    m = np.full(DIM * DIM, -1, dtype=np.int64)
    order = []
    for line in text.splitlines():
        if not line or line[0] == "#":
            continue
        if line[0] in " \t":
            order = [map_aa(tok) for tok in line.split()]
            continue
        a = map_aa(line[0])
        vals = line[1:].split()
        for b, v in zip(order, vals):
            if a >= 0 and b >= 0:
                m[(a << 5) + b] = int(v)
    return m


def blosum62():
    return parse_matrix(BLOSUM62_TEXT)


def nucleotide_matrix(match=1, mismatch=-3):
    """Codes 1..15: equal -> match, different -> mismatch; everything else -1
    (matrices.cc:533-538)."""
    m = np.full(DIM * DIM, -1, dtype=np.int64)
    for a in range(1, 16):
        for b in range(1, 16):
            m[(a << 5) + b] = match if a == b else mismatch
    return m


def matrix_limits(m):
    """(lo, hi, SCORELIMIT_7, SCORELIMIT_16) as matrices.cc:560-577."""
    lo, hi = int(min(100, m.min())), int(max(-100, m.max()))
    return lo, hi, 128 - hi, 65536 - hi
