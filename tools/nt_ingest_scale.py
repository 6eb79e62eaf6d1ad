"""Scale check of the nucleotide ingest: a packed .nsq database of N reads (vectorised writer, every
read 200 nt, a sprinkling of ambiguity runs) uploaded with swb_db_open_blast (2-bit -> codes on the
device) must score exactly like the same reads uploaded as plain symbol bytes; also reports the open
times and, for the six-frame translated upload, compares with host-translated subjects.
usage: python tools/nt_ingest_scale.py [nreads]"""
import os
import struct
import sys
import tempfile
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from swipe_b200 import BlastDB, Database, Scoring, scoring, synth

n = int(sys.argv[1]) if len(sys.argv) > 1 else 5_000_000
L = 200
rng = np.random.default_rng(7)
two = rng.integers(0, 4, size=(n, L), dtype=np.uint8)
codes = (1 << two).astype(np.uint8)
# ambiguity: every 1000th read gets an N run of 12 at position 40 (one old-format table entry)
amb_reads = np.arange(0, n, 1000)
codes[amb_reads, 40:52] = 15
two[amb_reads, 40:52] = 0
packed = ((two[:, 0::4] << 6) | (two[:, 1::4] << 4) | (two[:, 2::4] << 2) | two[:, 3::4]).astype(np.uint8)
rec = np.zeros((n, L // 4 + 1), dtype=np.uint8)
rec[:, : L // 4] = packed                                  # last byte: 0 remaining bases, count 0
has_amb = np.zeros(n, dtype=bool)
has_amb[amb_reads] = True
entry = struct.pack(">II", 1, (15 << 28) | (11 << 24) | 40)
reclen = np.where(has_amb, rec.shape[1] + len(entry), rec.shape[1]).astype(np.int64)
soff = np.concatenate([[1], 1 + np.cumsum(reclen)]).astype(np.int64)
aoff = soff[:-1] + rec.shape[1]
nsq = np.zeros(int(soff[-1]), dtype=np.uint8)
idx = (soff[:-1, None] + np.arange(rec.shape[1])[None, :]).reshape(-1)
nsq[idx] = rec.reshape(-1)
e = np.frombuffer(entry, dtype=np.uint8)
for k in range(len(entry)):
    nsq[aoff[amb_reads] + k] = e[k]
tmp = tempfile.mkdtemp()
nsq.tofile(os.path.join(tmp, "big.nsq"))
open(os.path.join(tmp, "big.nhr"), "wb").close()
t = b"synthetic reads"
d = b"Oct 17, 2026  5:00 AM"
buf = struct.pack(">II", 4, 0) + struct.pack(">I", len(t)) + t + struct.pack(">I", len(d)) + d
buf += bytes(-len(buf) % 4)
buf += struct.pack(">I", n) + struct.pack("<Q", n * L) + struct.pack(">I", L)
tabs = np.zeros(n + 1, dtype=">u4").tobytes() + soff.astype(">u4").tobytes() + np.concatenate([aoff, [soff[-1]]]).astype(">u4").tobytes()
open(os.path.join(tmp, "big.nin"), "wb").write(buf + tabs)
print("written: %d reads, %.2f GB packed" % (n, nsq.size / 1e9), flush=True)

q = synth.dna_query(300, seed=3)
sc = Scoring(scoring.nucleotide_matrix(1, -3), 5, 2)
residues = codes.reshape(-1)
offsets = np.arange(n + 1, dtype=np.int64) * L
t0 = time.perf_counter()
with Database(residues, offsets) as db:
    raw_open = time.perf_counter() - t0
    a = db.search(q, sc)
with BlastDB(os.path.join(tmp, "big"), nucleotide=True) as bdb:
    assert bdb.nseq == n and bdb.symbols == n * L
    t0 = time.perf_counter()
    with bdb.upload() as db:
        nsq_open = time.perf_counter() - t0
        b = db.search(q, sc)
        up, lay = db.open_ms()
    print("raw bytes open %.0f ms, packed open %.0f ms (device: upload+decode+layout %.1f ms); scores equal: %s, max %d"
          % (raw_open * 1e3, nsq_open * 1e3, up, bool(np.array_equal(a, b)), int(a.max())), flush=True)
    # six-frame translated upload of the first 200k reads against host translation
    import ctypes as C
    m = 200_000
    lib = bdb._lib
    table = np.zeros(4096, dtype=np.uint8)
    lib.swb_translate_table(1, table.ctypes.data)
    lib.swb_translate.restype = C.c_int64
    lib.swb_translate.argtypes = [C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
    prot = []
    out = np.zeros(L // 3 + 1, dtype=np.uint8)
    for s in range(0, m, 997):                              # a sample of reads, all six frames
        for strand in (0, 1):
            for frame in (0, 1, 2):
                k = lib.swb_translate(codes[s].ctypes.data, L, strand, frame, table.ctypes.data, out.ctypes.data)
                prot.append((s * 6 + strand * 3 + frame, out[:k].copy()))
    lib.swb_db_open_blast_translated.argtypes = [C.c_int, C.c_void_p, C.c_int64, C.c_int64, C.c_void_p, C.c_int,
                                                 C.c_void_p, C.POINTER(C.c_void_p)]
    h = C.c_void_p()
    assert lib.swb_db_open_blast_translated(0, bdb._h, 0, m, table.ctypes.data, 0, None, C.byref(h)) == 0
    pq = synth.protein_query(120)
    psc = Scoring(scoring.blosum62(), 11, 1)
    scores = np.zeros(6 * m, dtype=np.int64)
    s_ = psc._c()
    assert lib.swb_search(h, pq.ctypes.data, pq.size, C.byref(s_), scores.ctypes.data, None) == 0
    lib.swb_db_close(h)
    sub = np.concatenate([p for _, p in prot])
    off = np.concatenate([[0], np.cumsum([len(p) for _, p in prot])]).astype(np.int64)
    with Database(sub, off) as db:
        ref = db.search(pq, psc)
    got = scores[[i for i, _ in prot]]
    print("translated upload: %d sampled frames equal host translation: %s" % (len(prot), bool(np.array_equal(got, ref))), flush=True)
