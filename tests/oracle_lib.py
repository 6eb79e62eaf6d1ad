"""TEST INFRASTRUCTURE: ctypes access to the parity checkers under oracle/.

  Oracle  -> oracle/liborc.so        (our plain-C restatement, oracle/sw_oracle.c)
  Ref     -> oracle/_ref/libswipe_ref.so (the UNMODIFIED reference kernels + oracle/ref_harness.cc)

Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may import this.
"""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORC_PATH = os.path.join(ROOT, "oracle", "liborc.so")
REF_PATH = os.path.join(ROOT, "oracle", "_ref", "libswipe_ref.so")

_p = C.c_void_p
_l = C.c_long


def _ensure_built():
    if not os.path.exists(ORC_PATH) or (os.path.isdir("/root/reference") and not os.path.exists(REF_PATH)):
        subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "all"], check=True)


class Oracle:
    def __init__(self):
        _ensure_built()
        lib = C.CDLL(ORC_PATH)
        lib.orc_sw_score.restype = _l
        lib.orc_sw_score.argtypes = [_p, _l, _p, _l, _p, _l, _l, _p]
        lib.orc_sw_score_end.restype = _l
        lib.orc_sw_score_end.argtypes = [_p, _l, _p, _l, _p, _l, _l, _p, _p, _p]
        lib.orc_scan.restype = None
        lib.orc_scan.argtypes = [_p, _p, _l, _p, _l, _p, _l, _l, C.c_int, _p, _p, _p]
        lib.orc_topk.restype = _l
        lib.orc_topk.argtypes = [_p, _p, _l, _l, _l, _l, _p, _p, _p, _p]
        lib.orc_matrix_parse.restype = C.c_int
        lib.orc_matrix_parse.argtypes = [_p, C.c_char_p]
        lib.orc_matrix_nt.argtypes = [_p, _l, _l]
        lib.orc_matrix_limits.argtypes = [_p, _p, _p, _p, _p]
        lib.orc_map_aa.argtypes = [C.c_int]
        lib.orc_map_nt16.argtypes = [C.c_int]
        self.lib = lib

    def score(self, d, q, m, gap_open, gap_extend):
        d = np.ascontiguousarray(d, dtype=np.uint8)
        q = np.ascontiguousarray(q, dtype=np.uint8)
        m = np.ascontiguousarray(m, dtype=np.int64)
        he = np.zeros(2 * max(q.size, 1), dtype=np.int64)
        return int(self.lib.orc_sw_score(d.ctypes.data, d.size, q.ctypes.data, q.size,
                                         m.ctypes.data, gap_open + gap_extend, gap_extend,
                                         he.ctypes.data))

    def score_end(self, d, q, m, gap_open, gap_extend):
        d = np.ascontiguousarray(d, dtype=np.uint8)
        q = np.ascontiguousarray(q, dtype=np.uint8)
        m = np.ascontiguousarray(m, dtype=np.int64)
        he = np.zeros(2 * max(q.size, 1), dtype=np.int64)
        eq, ed = _l(), _l()
        s = self.lib.orc_sw_score_end(d.ctypes.data, d.size, q.ctypes.data, q.size, m.ctypes.data,
                                      gap_open + gap_extend, gap_extend, he.ctypes.data,
                                      C.byref(eq), C.byref(ed))
        return int(s), int(ed.value), int(eq.value)

    def scan(self, residues, offsets, q, m, gap_open, gap_extend, threads=None):
        residues = np.ascontiguousarray(residues, dtype=np.uint8)
        offsets = np.ascontiguousarray(offsets, dtype=np.int64)
        q = np.ascontiguousarray(q, dtype=np.uint8)
        m = np.ascontiguousarray(m, dtype=np.int64)
        n = offsets.size - 1
        scores = np.zeros(max(n, 1), dtype=np.int64)
        width = np.zeros(max(n, 1), dtype=np.uint8)
        counts = (_l * 3)()
        self.lib.orc_scan(residues.ctypes.data, offsets.ctypes.data, n, q.ctypes.data, q.size,
                          m.ctypes.data, gap_open + gap_extend, gap_extend,
                          threads or os.cpu_count() or 1, scores.ctypes.data, width.ctypes.data,
                          counts)
        return scores[:n], width[:n], [int(x) for x in counts]

    def topk(self, seqnos, scores, keep, min_score=0, upper=2 ** 62):
        seqnos = np.ascontiguousarray(seqnos, dtype=np.int64)
        scores = np.ascontiguousarray(scores, dtype=np.int64)
        oseq = np.zeros(max(keep, 1), dtype=np.int64)
        osc = np.zeros(max(keep, 1), dtype=np.int64)
        tot, obv = _l(), _l()
        k = self.lib.orc_topk(seqnos.ctypes.data, scores.ctypes.data, scores.size, keep, min_score,
                              upper, oseq.ctypes.data, osc.ctypes.data, C.byref(tot), C.byref(obv))
        return oseq[:k].copy(), osc[:k].copy(), int(tot.value), int(obv.value)

    def parse_matrix(self, text):
        m = np.zeros(1024, dtype=np.int64)
        rc = self.lib.orc_matrix_parse(m.ctypes.data, text.encode())
        if rc != 0:
            raise ValueError("matrix text rejected")
        return m

    def limits(self, m):
        m = np.ascontiguousarray(m, dtype=np.int64)
        v = [_l() for _ in range(4)]
        self.lib.orc_matrix_limits(m.ctypes.data, *[C.byref(x) for x in v])
        return tuple(int(x.value) for x in v)


def ref_available():
    _ensure_built()
    return os.path.exists(REF_PATH)


class Ref:
    """The reference's own kernels (search7/search7_ssse3/search16/search16s/fullsw) driven in
    memory.  The library keeps global state (score matrices), so use one instance per process."""

    def __init__(self):
        _ensure_built()
        lib = C.CDLL(REF_PATH)
        lib.ref_matrix_init.argtypes = [C.c_char_p, _l, _l, _l]
        lib.ref_matrix_get.argtypes = [_p, _p, _p]
        lib.ref_fullsw.restype = _l
        lib.ref_fullsw.argtypes = [_p, _l, _p, _l, _l, _l]
        lib.ref_scan.argtypes = [_p, _p, _l, _p, _l, _l, _l, C.c_int, _l, C.c_int, _p, _p, _p]
        lib.ref_search16.argtypes = [_p, _p, _l, _p, _l, _l, _l, _p, _p]
        lib.ref_search16s.argtypes = [_p, _p, _l, _p, _l, _l, _l, _p, _p, _p]
        if hasattr(lib, "ref_align"):
            lib.ref_align.restype = _l
            lib.ref_align.argtypes = [_p, _l, _p, _l, _l, _l, _p, _p, C.c_char_p, _l]
        self.lib = lib

    def align(self, q, d, gap_open, gap_extend, hint=None):
        """The reference's align() (align.cc:469-519) with the matrix of the last matrix_init:
        (score, q_start, d_start, q_end, d_end, ops)."""
        q = np.ascontiguousarray(q, dtype=np.uint8)
        d = np.ascontiguousarray(d, dtype=np.uint8)
        coords = np.zeros(4, dtype=np.int64)
        score = C.c_int64(0)
        if hint is not None:
            score.value, coords[2], coords[3] = int(hint[0]), int(hint[1]), int(hint[2])
        cap = 16 * (q.size + d.size) + 64
        buf = C.create_string_buffer(cap)
        self.lib.ref_align(q.ctypes.data, q.size, d.ctypes.data, d.size, gap_open, gap_extend,
                           coords.ctypes.data, C.byref(score), buf, cap)
        return (int(score.value), int(coords[0]), int(coords[1]), int(coords[2]), int(coords[3]),
                buf.value.decode())

    def matrix_init(self, name="BLOSUM62", symtype=1, match=1, mismatch=-3):
        self.lib.ref_matrix_init(name.encode(), symtype, match, mismatch)
        m = np.zeros(1024, dtype=np.int64)
        l7, l16 = C.c_int64(), C.c_int64()
        self.lib.ref_matrix_get(m.ctypes.data, C.byref(l7), C.byref(l16))
        return m, int(l7.value), int(l16.value)

    def fullsw(self, d, q, gap_open, gap_extend):
        d = np.ascontiguousarray(d, dtype=np.uint8)
        q = np.ascontiguousarray(q, dtype=np.uint8)
        return int(self.lib.ref_fullsw(d.ctypes.data, d.size, q.ctypes.data, q.size,
                                       gap_open + gap_extend, gap_extend))

    def scan(self, residues, offsets, q, gap_open, gap_extend, threads=1, chunk=1024, ssse3=1):
        residues = np.ascontiguousarray(residues, dtype=np.uint8)
        # the reference kernels read up to 3 bytes beyond a subject when they pad a 4-column
        # block, so keep slack behind the last one
        residues = np.concatenate([residues, np.zeros(64, dtype=np.uint8)])
        offsets = np.ascontiguousarray(offsets, dtype=np.int64)
        q = np.ascontiguousarray(q, dtype=np.uint8)
        n = offsets.size - 1
        scores = np.zeros(max(n, 1), dtype=np.int64)
        width = np.zeros(max(n, 1), dtype=np.uint8)
        counts = (_l * 3)()
        self.lib.ref_scan(residues.ctypes.data, offsets.ctypes.data, n, q.ctypes.data, q.size,
                          gap_open, gap_extend, threads, chunk, ssse3, scores.ctypes.data,
                          width.ctypes.data, counts)
        return scores[:n], width[:n], [int(x) for x in counts]

    def search16s(self, residues, offsets, q, gap_open, gap_extend):
        residues = np.concatenate([np.ascontiguousarray(residues, dtype=np.uint8),
                                   np.zeros(64, dtype=np.uint8)])
        offsets = np.ascontiguousarray(offsets, dtype=np.int64)
        q = np.ascontiguousarray(q, dtype=np.uint8)
        n = offsets.size - 1
        sc = np.zeros(max(n, 1), dtype=np.int64)
        bp = np.zeros(max(n, 1), dtype=np.int64)
        bq = np.zeros(max(n, 1), dtype=np.int64)
        self.lib.ref_search16s(residues.ctypes.data, offsets.ctypes.data, n, q.ctypes.data, q.size,
                               gap_open, gap_extend, sc.ctypes.data, bp.ctypes.data, bq.ctypes.data)
        return sc[:n], bp[:n], bq[:n]
