"""Small end-to-end run for compute-sanitizer (memcheck / racecheck / synccheck): every kernel family
once -- chunked open, hybrid scan (single and multi-pass), the int16 middle tier, the wide kernel, the
end-cell search, nucleotide decode and six-frame translation -- checked against nothing here (the
parity tests do that); the sanitizer's report is the result."""
import os
import sys
import tempfile

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import blastdb
from swipe_b200 import BlastDB, Database, Scoring, scoring, synth

os.environ["SWB_CHUNK_BYTES"] = "20000"
q = synth.protein_query(500)
res, off = synth.protein_db(400, query=q, seed=3, plant_every=7, max_len=500)
# a self copy scores above the 11-bit range: the int16 middle tier runs; mode 2 below runs the wide kernel
res = np.concatenate([res, q])
off = np.concatenate([off, [off[-1] + q.size]])
sc = Scoring(scoring.blosum62(), 11, 1)
with Database(res, off) as db:
    a = db.search(q, sc)
    db.set_shape(8, 8, 1)
    b = db.search(q, sc)                      # multi-pass
    assert np.array_equal(a, b)
    db.set_shape(0, 0, -1)
    assert db.last_counters["gpu_middle"] >= 1
    s, bp, bq = db.search_end(q, sc, np.arange(0, 400, 40))
    db.set_mode(2)
    w = db.search_list(q, sc, np.arange(380, 401))
    assert np.array_equal(w, a[380:401])
    print("protein ok", int(a.max()), db.last_counters)
with Database(res, off, wait=False) as db:
    assert np.array_equal(db.search(q, sc), a)
# round 2: the second kernel geometry (single pass, multi-pass with the cp.async feed, per-CTA scratch regions),
# the device sink, batched queries, the warp-per-subject end-cell kernel in passes
with Database(res, off) as db:
    db.set_geometry(2)
    assert np.array_equal(db.search(q[:375], sc), Database(res, off).search(q[:375], sc))
    assert np.array_equal(db.search(q, sc), a)          # 500 rows: two passes of G16
    db.set_shape(4, 25, 1)
    assert np.array_equal(db.search(q, sc), a)          # five passes of the four-stage CTA
    db.set_shape(0, 0, -1)
    os.environ["SWB_BND_BUDGET_MB"] = "0"
    assert np.array_equal(db.search(q, sc), a)          # regions claimed / released per CTA
    db.set_geometry(1)
    db.set_shape(8, 8, 1)
    assert np.array_equal(db.search(q, sc), a)
    db.set_shape(0, 0, -1)
    del os.environ["SWB_BND_BUDGET_MB"]
    db.set_geometry(0)
    seq, hs, tot, obv = db.search_hits(q, sc, 25, 1)
    order = np.lexsort((-np.arange(a.size), -a))[:25]
    assert np.array_equal(seq, order) and np.array_equal(hs, a[order])
    qs = [q[:90], q[100:160], q[200:330], q[:25]]
    got = db.search_batch(qs, sc)
    for qq, g in zip(qs, got):
        assert np.array_equal(g, db.search(qq, sc))
    long_q = synth.protein_query(1300, seed=9)
    s1, bp1, bq1 = db.search_end(long_q, sc, np.arange(0, 400, 25))     # two passes of 1024 rows
    print("round-2 paths ok", int(hs[0]), int(s1.max()))
tmp = tempfile.mkdtemp()
qn = synth.dna_query(200, seed=5)
rng = np.random.default_rng(1)
subs = [(1 << rng.integers(0, 4, size=int(rng.integers(1, 300)))).astype(np.uint8) for _ in range(100)]
subs[3][5:20] = 15
blastdb.write_nucleotide(os.path.join(tmp, "n"), subs)
with BlastDB(os.path.join(tmp, "n"), nucleotide=True) as bdb:
    with bdb.upload() as db:
        db.search(qn, Scoring(scoring.nucleotide_matrix(1, -3), 5, 2))
    import ctypes as C
    lib = db._lib
    table = np.zeros(4096, dtype=np.uint8)
    lib.swb_translate_table(1, table.ctypes.data)
    h = C.c_void_p()
    lib.swb_db_open_blast_translated.argtypes = [C.c_int, C.c_void_p, C.c_int64, C.c_int64, C.c_void_p, C.c_int,
                                                 C.c_void_p, C.POINTER(C.c_void_p)]
    assert lib.swb_db_open_blast_translated(0, bdb._h, 0, -1, table.ctypes.data, 0, None, C.byref(h)) == 0
    lib.swb_db_close(h)
print("all ok")
