"""Generates tests/golden/host.json and host_tables.npz from the UNMODIFIED reference objects
(stats.cc / NCBI blastkar tables, query.cc) through oracle/_ref/libswipe_ref.so:

  stats        Karlin-Altschul parameters for every built-in matrix x gap penalties 0..32 / 0..4 that is
               tabulated (stats.cc:163-245), the nucleotide sets (stats.cc:44-161), preferred gaps
  lenadj       BlastComputeLengthAdjustment over a grid of query / database sizes
  codon tables translate_createtable for every assigned genetic code (query.cc:366-444)

Run in the build container:  python tests/golden/make_golden_host.py"""
import ctypes as C
import json
import math
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import oracle_lib  # noqa: E402

MATS = ["BLOSUM45", "BLOSUM50", "BLOSUM62", "BLOSUM80", "BLOSUM90", "PAM30", "PAM70", "PAM250"]
NT = [(1, -5), (1, -4), (2, -7), (1, -3), (2, -5), (1, -2), (2, -3), (3, -4), (4, -5), (1, -1), (3, -2), (5, -4), (1, -6)]


def main():
    lib = oracle_lib.Ref().lib
    lib.ref_stats_params.restype = C.c_long
    lib.ref_stats_params.argtypes = [C.c_char_p, C.c_long, C.c_long, C.POINTER(C.c_double)]
    lib.ref_stats_params_nt.restype = C.c_long
    lib.ref_stats_params_nt.argtypes = [C.c_long, C.c_long, C.c_long, C.c_long, C.POINTER(C.c_double)]
    lib.ref_stats_prefs.restype = C.c_long
    lib.ref_stats_prefs.argtypes = [C.c_char_p, C.POINTER(C.c_long), C.POINTER(C.c_long)]
    lib.ref_length_adjustment.restype = C.c_long
    lib.ref_length_adjustment.argtypes = [C.c_double, C.c_double, C.c_double, C.c_double, C.c_long, C.c_long, C.c_long]
    lib.ref_translate_table.argtypes = [C.c_long, C.c_void_p]
    out = {"protein": [], "nt": [], "prefs": {}, "lenadj": []}
    p = (C.c_double * 5)()
    for m in MATS + ["blosum62", "NOSUCH"]:
        for go in list(range(0, 33)) + [32767]:
            for ge in list(range(0, 5)) + [32767]:
                if lib.ref_stats_params(m.encode(), go, ge, p):
                    out["protein"].append([m, go, ge, list(p)])
        a, b = C.c_long(), C.c_long()
        if lib.ref_stats_prefs(m.encode(), C.byref(a), C.byref(b)):
            out["prefs"][m] = [a.value, b.value]
    for r, q in NT:
        for go in range(0, 30):
            for ge in range(0, 12):
                if lib.ref_stats_params_nt(r, q, go, ge, p):
                    out["nt"].append([r, q, go, ge, list(p)])
    for lam, K, alpha, beta in ((0.267, 0.041, 1.9, -30.0), (0.3176, 0.134, 0.7916, -3.2), (1.374, 0.711, 1.05, 0.0),
                                (0.206, 0.010, 4.0, -87.0)):
        for qlen in (1, 10, 60, 375, 1000, 5000, 100000):
            for dblen, nseq in ((100, 1), (15429, 89), (1755796755, 5000000), (10000160084, 50000000), (500, 400)):
                adj = lib.ref_length_adjustment(K, math.log(K), alpha / lam, beta, qlen, dblen, nseq)
                out["lenadj"].append([lam, K, alpha, beta, qlen, dblen, nseq, adj])
    json.dump(out, open(os.path.join(HERE, "host.json"), "w"))
    tabs = {}
    buf = np.zeros(4096, dtype=np.uint8)
    for g in (1, 2, 3, 4, 5, 6, 9, 10, 11, 12, 13, 14, 15, 16, 21, 22, 23):
        lib.ref_translate_table(g, buf.ctypes.data)
        tabs["code%d" % g] = buf.copy()
    np.savez_compressed(os.path.join(HERE, "host_tables.npz"), **tabs)
    print("protein %d, nt %d, lenadj %d, tables %d" % (len(out["protein"]), len(out["nt"]), len(out["lenadj"]), len(tabs)))


if __name__ == "__main__":
    main()
