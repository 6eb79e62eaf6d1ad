"""ctypes binding of include/swipe_b200.h (the C ABI is the product; this is plumbing)."""
import ctypes as C
import os

import numpy as np

from . import build as _build

_LIB = None

STATUS = {0: "SWB_OK", -1: "SWB_ERR_ARG", -2: "SWB_ERR_NO_DEVICE", -3: "SWB_ERR_CUDA",
          -4: "SWB_ERR_NOMEM", -5: "SWB_ERR_RANGE", -6: "SWB_ERR_INTERNAL", -7: "SWB_ERR_IO"}

# every symbol include/swipe_b200.h declares (tests check the library exports all of them)
EXPORTS = ["swb_abi_version", "swb_align", "swb_alu_peak", "swb_blastdb_close", "swb_blastdb_date",
           "swb_blastdb_error", "swb_blastdb_header", "swb_blastdb_included", "swb_blastdb_info",
           "swb_blastdb_masked_info", "swb_blastdb_open", "swb_blastdb_seqlen", "swb_blastdb_sequence",
           "swb_blastdb_title", "swb_db_close", "swb_db_info", "swb_db_open",
           "swb_db_open_async", "swb_db_open_blast", "swb_db_open_blast_translated", "swb_db_open_ms", "swb_db_set_filter",
           "swb_db_wait", "swb_defline_text", "swb_device_count", "swb_gencode_name",
           "swb_host_alloc", "swb_host_free", "swb_last_cuda_error", "swb_matrix_builtin",
           "swb_matrix_limits", "swb_matrix_nucleotide", "swb_matrix_parse", "swb_matrix_read",
           "swb_matrix_read_sound", "swb_query_parse", "swb_revcomp", "swb_search",
           "swb_search_batch", "swb_search_end", "swb_search_hits", "swb_search_hits_batch", "swb_search_list", "swb_set_cache_limit",
           "swb_hits_merge", "swb_set_geometry", "swb_set_mode", "swb_set_shape",
           "swb_stats_bits", "swb_stats_default_gaps", "swb_stats_evalue", "swb_stats_init",
           "swb_stats_length_adjustment", "swb_stats_params", "swb_stats_params_nt", "swb_strerror",
           "swb_topk_merge", "swb_translate", "swb_translate_table", "swb_trim"]


class SwbError(RuntimeError):
    def __init__(self, status, detail=""):
        self.status = status
        msg = "%s (%d)" % (STATUS.get(status, "?"), status)
        if detail:
            msg += ": " + detail
        super().__init__(msg)


class _Scoring(C.Structure):
    _fields_ = [("matrix", C.POINTER(C.c_int64)), ("gap_open_extend", C.c_int64),
                ("gap_extend", C.c_int64)]


class Counters(C.Structure):
    _fields_ = [("subjects", C.c_int64), ("cells", C.c_int64), ("ref_width7", C.c_int64),
                ("ref_width16", C.c_int64), ("ref_width63", C.c_int64), ("gpu_narrow", C.c_int64),
                ("gpu_requeued", C.c_int64), ("gpu_middle", C.c_int64), ("kernel_launches", C.c_int64),
                ("scan_ms", C.c_double), ("requeue_ms", C.c_double), ("scan_geometry", C.c_int64),
                ("scan_G", C.c_int64), ("scan_R", C.c_int64), ("scan_passes", C.c_int64)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


def load_library():
    """Loads (building first if stale and nvcc is present) the CUDA library.  Raises when it is
    missing: there is deliberately no fallback implementation."""
    global _LIB
    if _LIB is not None:
        return _LIB
    path = os.environ.get("SWB_LIBRARY") or _build.build_lib()       # (an experiment build, tools/build_variant.sh)
    if not os.path.exists(path):
        raise RuntimeError("swipe_b200: %s is not built (run python -m swipe_b200.build)" % path)
    lib = C.CDLL(path)
    p64 = C.POINTER(C.c_int64)
    pu8 = C.POINTER(C.c_uint8)
    lib.swb_abi_version.restype = C.c_int
    lib.swb_strerror.restype = C.c_char_p
    lib.swb_strerror.argtypes = [C.c_int]
    lib.swb_last_cuda_error.restype = C.c_char_p
    lib.swb_device_count.argtypes = [C.POINTER(C.c_int)]
    lib.swb_host_alloc.argtypes = [C.POINTER(C.c_void_p), C.c_int64]
    lib.swb_host_free.argtypes = [C.c_void_p]
    lib.swb_db_open.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_void_p,
                                C.POINTER(C.c_void_p)]
    lib.swb_db_open_async.argtypes = lib.swb_db_open.argtypes
    lib.swb_db_wait.argtypes = [C.c_void_p]
    lib.swb_trim.restype = C.c_int
    lib.swb_db_close.argtypes = [C.c_void_p]
    lib.swb_db_info.argtypes = [C.c_void_p, p64, p64, p64]
    lib.swb_search.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.POINTER(_Scoring), C.c_void_p,
                               C.POINTER(Counters)]
    lib.swb_search_list.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.POINTER(_Scoring),
                                    C.c_void_p, C.c_int64, C.c_void_p, C.POINTER(Counters)]
    lib.swb_search_end.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.POINTER(_Scoring),
                                   C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.swb_search_hits.restype = C.c_int
    lib.swb_search_hits.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.POINTER(_Scoring), C.c_int64,
                                    C.c_int64, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p, p64, p64,
                                    p64, C.POINTER(Counters)]
    lib.swb_search_batch.restype = C.c_int
    lib.swb_search_batch.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_void_p), p64, C.POINTER(_Scoring),
                                     C.POINTER(C.c_void_p), C.POINTER(Counters)]
    lib.swb_search_hits_batch.restype = C.c_int
    lib.swb_search_hits_batch.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_void_p), p64, C.POINTER(_Scoring),
                                          C.c_int64, C.c_int64, C.c_int64, C.c_int64, C.POINTER(C.c_void_p),
                                          C.POINTER(C.c_void_p), p64, p64, p64, C.POINTER(Counters)]
    lib.swb_hits_merge.restype = C.c_int64
    lib.swb_hits_merge.argtypes = [C.c_int, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), p64, C.c_int64,
                                   C.c_void_p, C.c_void_p]
    lib.swb_set_cache_limit.restype = C.c_int
    lib.swb_db_set_filter.restype = C.c_int
    lib.swb_db_set_filter.argtypes = [C.c_void_p, C.c_void_p]
    lib.swb_alu_peak.restype = C.c_int
    lib.swb_alu_peak.argtypes = [C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double)]
    lib.swb_set_cache_limit.argtypes = [C.c_int64]
    lib.swb_topk_merge.restype = C.c_int64
    lib.swb_topk_merge.argtypes = [C.c_int, C.POINTER(C.c_void_p), p64, p64, C.c_int64, C.c_int64,
                                   C.c_int64, C.c_void_p, C.c_void_p, p64, p64]
    lib.swb_set_mode.argtypes = [C.c_void_p, C.c_int]
    lib.swb_db_open_ms.argtypes = [C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_double)]
    lib.swb_set_shape.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int]
    lib.swb_set_geometry.argtypes = [C.c_void_p, C.c_int]
    lib.swb_set_geometry.restype = C.c_int
    for name in ("swb_device_count", "swb_host_alloc", "swb_host_free", "swb_db_open",
                 "swb_db_close", "swb_db_info", "swb_search", "swb_search_list", "swb_search_end",
                 "swb_set_mode", "swb_db_open_ms", "swb_db_set_filter", "swb_set_shape", "swb_db_open_async", "swb_db_wait"):
        getattr(lib, name).restype = C.c_int
    lib.swb_blastdb_open.argtypes = [C.c_char_p, C.c_int, C.POINTER(C.c_void_p)]
    lib.swb_blastdb_close.argtypes = [C.c_void_p]
    lib.swb_blastdb_error.restype = C.c_char_p
    lib.swb_blastdb_info.argtypes = [C.c_void_p, p64, p64, p64, C.POINTER(C.c_int)]
    lib.swb_blastdb_title.restype = C.c_char_p
    lib.swb_blastdb_title.argtypes = [C.c_void_p]
    lib.swb_blastdb_date.restype = C.c_char_p
    lib.swb_blastdb_date.argtypes = [C.c_void_p]
    lib.swb_blastdb_seqlen.restype = C.c_int64
    lib.swb_blastdb_seqlen.argtypes = [C.c_void_p, C.c_int64]
    lib.swb_blastdb_sequence.argtypes = [C.c_void_p, C.c_int64, C.c_int, C.c_void_p, C.c_int64, p64]
    lib.swb_blastdb_header.argtypes = [C.c_void_p, C.c_int64, C.POINTER(C.c_void_p), p64]
    lib.swb_blastdb_included.argtypes = [C.c_void_p, C.c_int64]
    lib.swb_db_open_blast.argtypes = [C.c_int, C.c_void_p, C.c_int64, C.c_int64, C.c_int, C.c_void_p,
                                      C.POINTER(C.c_void_p)]
    lib.swb_align.restype = C.c_int
    lib.swb_align.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64,
                              C.c_int64, p64, p64, p64, p64, p64, C.c_char_p, C.c_int64, p64]
    for name in ("swb_blastdb_open", "swb_blastdb_close", "swb_blastdb_info", "swb_blastdb_sequence",
                 "swb_blastdb_header", "swb_blastdb_included", "swb_db_open_blast"):
        getattr(lib, name).restype = C.c_int
    _LIB = lib
    return lib


def _check(rc):
    if rc != 0:
        lib = load_library()
        detail = lib.swb_strerror(rc).decode()
        cuda = lib.swb_last_cuda_error().decode()
        if rc in (-2, -3, -4) and cuda:
            detail += " [" + cuda + "]"
        if rc == -7:
            detail += " [" + lib.swb_blastdb_error().decode() + "]"
        raise SwbError(rc, detail)


def device_count():
    lib = load_library()
    n = C.c_int(0)
    _check(lib.swb_device_count(C.byref(n)))
    return n.value


class HostBuffer:
    """Pinned host memory from swb_host_alloc, viewed as a numpy array."""

    def __init__(self, nbytes):
        lib = load_library()
        self._ptr = C.c_void_p()
        _check(lib.swb_host_alloc(C.byref(self._ptr), int(nbytes)))
        self.nbytes = int(nbytes)
        buf = (C.c_uint8 * max(self.nbytes, 1)).from_address(self._ptr.value)
        self.u8 = np.frombuffer(buf, dtype=np.uint8, count=self.nbytes)

    def view(self, dtype):
        return self.u8.view(dtype)

    def free(self):
        if self._ptr is not None and self._ptr.value:
            load_library().swb_host_free(self._ptr)
            self._ptr = None
            self.u8 = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class Scoring:
    """matrix: int64[1024] indexed [(db_symbol << 5) + query_symbol] (score_matrix_63 layout);
    gap_open / gap_extend as given on the reference's command line (-G / -E)."""

    def __init__(self, matrix, gap_open, gap_extend):
        self.matrix = np.ascontiguousarray(matrix, dtype=np.int64).reshape(-1)
        if self.matrix.size != 1024:
            raise ValueError("matrix must hold 32 x 32 scores")
        self.gap_open = int(gap_open)
        self.gap_extend = int(gap_extend)

    def _c(self):
        s = _Scoring()
        s.matrix = self.matrix.ctypes.data_as(C.POINTER(C.c_int64))
        s.gap_open_extend = self.gap_open + self.gap_extend     # swipe.cc:1126
        s.gap_extend = self.gap_extend
        return s


def _u8(a):
    return np.ascontiguousarray(a, dtype=np.uint8)


class BlastDB:
    """A BLAST version-4 database opened by the library's own reader (swb_blastdb)."""

    def __init__(self, basename, nucleotide=False):
        lib = load_library()
        self._lib = lib
        self._h = C.c_void_p()
        _check(lib.swb_blastdb_open(os.fsencode(basename), int(bool(nucleotide)), C.byref(self._h)))
        a, b, c, v = C.c_int64(), C.c_int64(), C.c_int64(), C.c_int()
        _check(lib.swb_blastdb_info(self._h, C.byref(a), C.byref(b), C.byref(c), C.byref(v)))
        self.nseq, self.symbols, self.longest, self.volumes = a.value, b.value, c.value, v.value
        self.nucleotide = bool(nucleotide)
        self.title = lib.swb_blastdb_title(self._h).decode()
        self.date = lib.swb_blastdb_date(self._h).decode()

    def close(self):
        if self._h is not None and self._h.value:
            self._lib.swb_blastdb_close(self._h)
            self._h = None

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def seqlen(self, seqno):
        n = self._lib.swb_blastdb_seqlen(self._h, int(seqno))
        if n < 0:
            _check(int(n))
        return int(n)

    def sequence(self, seqno, strand=0):
        n = self.seqlen(seqno)
        buf = np.empty(max(n, 1), dtype=np.uint8)
        got = C.c_int64()
        _check(self._lib.swb_blastdb_sequence(self._h, int(seqno), int(strand), buf.ctypes.data, n,
                                              C.byref(got)))
        return buf[:got.value].copy()

    def header(self, seqno):
        ptr, n = C.c_void_p(), C.c_int64()
        _check(self._lib.swb_blastdb_header(self._h, int(seqno), C.byref(ptr), C.byref(n)))
        return C.string_at(ptr.value, n.value) if n.value else b""

    def included(self, seqno):
        return bool(self._lib.swb_blastdb_included(self._h, int(seqno)))

    def upload(self, device=0, first=0, count=-1, stream=None, wait=True):
        """The shard [first, first+count) resident on one GPU (swb_db_open_blast)."""
        return Database._from_blast(self, device, first, count, stream, wait)


class Database:
    """One database shard resident on one GPU (swb_db)."""

    @classmethod
    def _from_blast(cls, bdb, device, first, count, stream, wait):
        self = cls.__new__(cls)
        self._lib = bdb._lib
        self._bdb = bdb                       # keeps the mappings alive for an asynchronous open
        self._h = C.c_void_p()
        _check(self._lib.swb_db_open_blast(int(device), bdb._h, int(first), int(count), int(not wait),
                                           C.c_void_p(stream or 0), C.byref(self._h)))
        self.nseq = self.info()["nseq"]
        self.last_counters = None
        return self

    def __init__(self, residues, offsets, device=0, trailing=0, stream=None, wait=True):
        lib = load_library()
        self._lib = lib
        self._residues = _u8(residues)
        self._offsets = np.ascontiguousarray(offsets, dtype=np.int64)
        self.nseq = int(self._offsets.size - 1)
        self._h = C.c_void_p()
        opener = lib.swb_db_open if wait else lib.swb_db_open_async
        _check(opener(int(device), self._residues.ctypes.data, self._offsets.ctypes.data,
                      self.nseq, int(trailing), C.c_void_p(stream or 0), C.byref(self._h)))
        self.last_counters = None

    def close(self):
        if self._h is not None and self._h.value:
            self._lib.swb_db_close(self._h)
            self._h = None

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def wait(self):
        _check(self._lib.swb_db_wait(self._h))

    def info(self):
        a, b, c = C.c_int64(), C.c_int64(), C.c_int64()
        _check(self._lib.swb_db_info(self._h, C.byref(a), C.byref(b), C.byref(c)))
        return {"nseq": a.value, "residues": b.value, "longest": c.value}

    def open_ms(self):
        a, b = C.c_double(), C.c_double()
        _check(self._lib.swb_db_open_ms(self._h, C.byref(a), C.byref(b)))
        return a.value, b.value

    def set_mode(self, mode):
        _check(self._lib.swb_set_mode(self._h, int(mode)))

    def set_filter(self, include=None):
        """swb_db_set_filter: include = boolean array over the subjects (None removes the filter)."""
        if include is None:
            _check(self._lib.swb_db_set_filter(self._h, None))
            return
        bits = np.packbits(np.asarray(include, dtype=bool), bitorder="little")
        _check(self._lib.swb_db_set_filter(self._h, bits.ctypes.data))

    def set_geometry(self, geometry=0):
        _check(self._lib.swb_set_geometry(self._h, int(geometry)))

    def set_shape(self, G=0, R=0, lane_mode=-1):
        _check(self._lib.swb_set_shape(self._h, int(G), int(R), int(lane_mode)))

    def search(self, query, scoring, out=None):
        q = _u8(query)
        scores = out if out is not None else np.empty(self.nseq, dtype=np.int64)
        ctr = Counters()
        sc = scoring._c()
        _check(self._lib.swb_search(self._h, q.ctypes.data, q.size, C.byref(sc),
                                    scores.ctypes.data, C.byref(ctr)))
        self.last_counters = ctr.as_dict()
        return scores

    def search_hits(self, query, scoring, keep, min_score=1, upper_score=2 ** 62, seqno_base=0):
        """swb_search_hits: the best `keep` admissible subjects, selected on the device, in the
        sink's order.  Returns (seqnos, scores, totalhits, obvious)."""
        q = _u8(query)
        keep = int(keep)
        if getattr(self, "_hit_keep", -1) < keep:
            self._hit_seq = np.empty(max(keep, 1), dtype=np.int64)
            self._hit_sc = np.empty(max(keep, 1), dtype=np.int64)
            self._hit_keep = keep
        k, tot, obv = C.c_int64(), C.c_int64(), C.c_int64()
        ctr = Counters()
        sc = scoring._c()
        _check(self._lib.swb_search_hits(self._h, q.ctypes.data, q.size, C.byref(sc), int(seqno_base),
                                         keep, int(min_score), int(upper_score),
                                         self._hit_seq.ctypes.data, self._hit_sc.ctypes.data,
                                         C.byref(k), C.byref(tot), C.byref(obv), C.byref(ctr)))
        self.last_counters = ctr.as_dict()
        return self._hit_seq[:k.value].copy(), self._hit_sc[:k.value].copy(), tot.value, obv.value

    def search_batch(self, queries, scoring):
        """swb_search_batch: list of score arrays, one per query (queries sharing scans where they fit)."""
        qs = [_u8(q) for q in queries]
        n = len(qs)
        out = [np.empty(self.nseq, dtype=np.int64) for _ in qs]
        qp = (C.c_void_p * max(n, 1))(*[q.ctypes.data for q in qs])
        ql = (C.c_int64 * max(n, 1))(*[q.size for q in qs])
        op = (C.c_void_p * max(n, 1))(*[o.ctypes.data for o in out])
        ctr = (Counters * max(n, 1))()
        sc = scoring._c()
        _check(self._lib.swb_search_batch(self._h, n, qp, ql, C.byref(sc), op, ctr))
        self.last_batch_counters = [ctr[k].as_dict() for k in range(n)]
        return out

    def search_hits_batch(self, queries, scoring, keep, min_score=1, upper_score=2 ** 62, seqno_base=0):
        """swb_search_hits_batch: [(seqnos, scores, totalhits, obvious)] per query."""
        qs = [_u8(q) for q in queries]
        n = len(qs)
        keep = int(keep)
        seqs = [np.empty(max(keep, 1), dtype=np.int64) for _ in qs]
        scs = [np.empty(max(keep, 1), dtype=np.int64) for _ in qs]
        qp = (C.c_void_p * max(n, 1))(*[q.ctypes.data for q in qs])
        ql = (C.c_int64 * max(n, 1))(*[q.size for q in qs])
        p1 = (C.c_void_p * max(n, 1))(*[a.ctypes.data for a in seqs])
        p2 = (C.c_void_p * max(n, 1))(*[a.ctypes.data for a in scs])
        nh = (C.c_int64 * max(n, 1))()
        tot = (C.c_int64 * max(n, 1))()
        obv = (C.c_int64 * max(n, 1))()
        ctr = (Counters * max(n, 1))()
        sc = scoring._c()
        _check(self._lib.swb_search_hits_batch(self._h, n, qp, ql, C.byref(sc), int(seqno_base), keep,
                                               int(min_score), int(upper_score), p1, p2, nh, tot, obv, ctr))
        self.last_batch_counters = [ctr[k].as_dict() for k in range(n)]
        return [(seqs[k][:nh[k]].copy(), scs[k][:nh[k]].copy(), tot[k], obv[k]) for k in range(n)]

    def search_list(self, query, scoring, seqnos):
        """seqnos: plain sequence numbers; coded (seqno << 3) for the ABI as the reference does."""
        q = _u8(query)
        coded = np.ascontiguousarray(np.asarray(seqnos, dtype=np.int64) << 3)
        scores = np.empty(coded.size, dtype=np.int64)
        ctr = Counters()
        sc = scoring._c()
        _check(self._lib.swb_search_list(self._h, q.ctypes.data, q.size, C.byref(sc),
                                         coded.ctypes.data, coded.size, scores.ctypes.data,
                                         C.byref(ctr)))
        self.last_counters = ctr.as_dict()
        return scores

    def search_end(self, query, scoring, seqnos):
        q = _u8(query)
        coded = np.ascontiguousarray(np.asarray(seqnos, dtype=np.int64) << 3)
        scores = np.empty(coded.size, dtype=np.int64)
        bestpos = np.empty(coded.size, dtype=np.int64)
        bestq = np.empty(coded.size, dtype=np.int64)
        sc = scoring._c()
        _check(self._lib.swb_search_end(self._h, q.ctypes.data, q.size, C.byref(sc),
                                        coded.ctypes.data, coded.size, scores.ctypes.data,
                                        bestpos.ctypes.data, bestq.ctypes.data))
        return scores, bestpos, bestq


def align(query, subject, scoring, hint=None):
    """swb_align: (score, q_start, d_start, q_end, d_end, ops).  hint = (score, q_end, d_end) from
    search_end (hits.cc:589-600) or None to let the aligner find the end cell itself."""
    lib = load_library()
    q, d = _u8(query), _u8(subject)
    vals = [C.c_int64(0) for _ in range(5)]                  # q_start d_start q_end d_end score
    if hint is not None:
        vals[4].value, vals[2].value, vals[3].value = int(hint[0]), int(hint[1]), int(hint[2])
    cap = 16 * (q.size + d.size) + 64
    buf = C.create_string_buffer(cap)
    n = C.c_int64()
    _check(lib.swb_align(q.ctypes.data, q.size, d.ctypes.data, d.size, scoring.matrix.ctypes.data,
                         scoring.gap_open, scoring.gap_extend, C.byref(vals[0]), C.byref(vals[1]),
                         C.byref(vals[2]), C.byref(vals[3]), C.byref(vals[4]), buf, cap, C.byref(n)))
    return (vals[4].value, vals[0].value, vals[1].value, vals[2].value, vals[3].value,
            buf.value.decode())


def topk_merge(score_arrays, seqno_bases, keep, min_score=0, upper_score=2 ** 62):
    """hits_enter's rule (hits.cc:163-222) over the score arrays of one or more shards."""
    lib = load_library()
    arrs = [np.ascontiguousarray(a, dtype=np.int64) for a in score_arrays]
    n = len(arrs)
    ptrs = (C.c_void_p * max(n, 1))(*[a.ctypes.data for a in arrs])
    ns = (C.c_int64 * max(n, 1))(*[a.size for a in arrs])
    bases = (C.c_int64 * max(n, 1))(*[int(b) for b in seqno_bases])
    out_seq = np.empty(max(keep, 1), dtype=np.int64)
    out_sc = np.empty(max(keep, 1), dtype=np.int64)
    tot, obv = C.c_int64(), C.c_int64()
    k = lib.swb_topk_merge(n, ptrs, ns, bases, int(keep), int(min_score), int(upper_score),
                           out_seq.ctypes.data, out_sc.ctypes.data, C.byref(tot), C.byref(obv))
    if k < 0:
        _check(int(k))
    return out_seq[:k].copy(), out_sc[:k].copy(), tot.value, obv.value


def hits_merge(lists, keep):
    """swb_hits_merge over [(seqnos, scores), ...], each already in the sink's order."""
    lib = load_library()
    seqs = [np.ascontiguousarray(a, dtype=np.int64) for a, _ in lists]
    scs = [np.ascontiguousarray(b, dtype=np.int64) for _, b in lists]
    n = len(seqs)
    p1 = (C.c_void_p * max(n, 1))(*[a.ctypes.data for a in seqs])
    p2 = (C.c_void_p * max(n, 1))(*[a.ctypes.data for a in scs])
    ns = (C.c_int64 * max(n, 1))(*[a.size for a in seqs])
    out_seq = np.empty(max(keep, 1), dtype=np.int64)
    out_sc = np.empty(max(keep, 1), dtype=np.int64)
    k = lib.swb_hits_merge(n, p1, p2, ns, int(keep), out_seq.ctypes.data, out_sc.ctypes.data)
    if k < 0:
        _check(int(k))
    return out_seq[:k].copy(), out_sc[:k].copy()


def set_cache_limit(nbytes):
    _check(load_library().swb_set_cache_limit(int(nbytes)))


def alu_peak(device=0):
    """swb_alu_peak: (DPX warp instructions per clock per SM, SM clock in MHz during the measurement)."""
    a, b = C.c_double(), C.c_double()
    _check(load_library().swb_alu_peak(int(device), C.byref(a), C.byref(b)))
    return a.value, b.value
