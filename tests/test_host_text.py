"""Host modules of the library against golden data from the reference (tests/golden/host.json,
host_tables.npz, matrices.npz; generating scripts beside them): substitution tables, Karlin-Altschul
parameters and length adjustment, codon tables, query parsing, deflines."""
import ctypes as C
import json
import math
import os

import numpy as np
import pytest

import blastdb
from swipe_b200 import load_library, scoring, synth

GOLD = os.path.join(os.path.dirname(__file__), "golden")
P64 = C.POINTER(C.c_int64)


@pytest.fixture(scope="module")
def lib():
    lib = load_library()
    lib.swb_matrix_builtin.argtypes = [C.c_char_p, C.c_void_p]
    lib.swb_matrix_parse.argtypes = [C.c_char_p, C.c_void_p]
    lib.swb_matrix_read.argtypes = [C.c_char_p, C.c_void_p]
    lib.swb_matrix_nucleotide.argtypes = [C.c_int64, C.c_int64, C.c_void_p]
    lib.swb_matrix_limits.argtypes = [C.c_void_p, P64, P64, P64, P64]
    lib.swb_stats_params.argtypes = [C.c_char_p, C.c_int64, C.c_int64, C.POINTER(C.c_double)]
    lib.swb_stats_params_nt.argtypes = [C.c_int64, C.c_int64, C.c_int64, C.c_int64, C.POINTER(C.c_double)]
    lib.swb_stats_default_gaps.argtypes = [C.c_char_p, P64, P64]
    lib.swb_stats_length_adjustment.restype = C.c_int64
    lib.swb_stats_length_adjustment.argtypes = [C.c_double, C.c_double, C.c_double, C.c_int64, C.c_int64, C.c_int64]
    lib.swb_translate_table.argtypes = [C.c_int, C.c_void_p]
    lib.swb_translate.restype = C.c_int64
    lib.swb_translate.argtypes = [C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
    lib.swb_revcomp.argtypes = [C.c_void_p, C.c_int64, C.c_void_p]
    lib.swb_query_parse.restype = C.c_int64
    lib.swb_query_parse.argtypes = [C.c_char_p, C.c_int64, C.c_int, C.c_void_p, C.c_int64, P64, C.c_char_p, C.c_int64]
    lib.swb_defline_text.restype = C.c_int64
    lib.swb_defline_text.argtypes = [C.c_char_p, C.c_int64, C.c_int, C.c_int, C.c_int64, C.c_void_p, C.c_int64,
                                     C.c_char_p, C.c_int64, P64]
    lib.swb_gencode_name.restype = C.c_char_p
    return lib


def test_builtin_matrices_and_limits(lib):
    g = np.load(os.path.join(GOLD, "matrices.npz"))
    m = np.zeros(1024, dtype=np.int64)
    for name in ("blosum45", "blosum50", "blosum62", "blosum80", "blosum90", "pam30", "pam70", "pam250", "identity_5_1"):
        assert lib.swb_matrix_builtin(name.upper().encode(), m.ctypes.data) == 0
        assert np.array_equal(m, g[name].astype(np.int64)), name
        lo, hi, l7, l16 = (C.c_int64() for _ in range(4))
        lib.swb_matrix_limits(m.ctypes.data, lo, hi, l7, l16)
        assert [l7.value, l16.value] == g[name + "_limits"].tolist()
    assert lib.swb_matrix_builtin(b"nosuch", m.ctypes.data) == -1
    lib.swb_matrix_nucleotide(1, -3, m.ctypes.data)
    assert np.array_equal(m, g["nt_1_-3"].astype(np.int64))


def test_matrix_file_parser(lib, tmp_path):
    g = np.load(os.path.join(GOLD, "asym.npz"))
    m = np.zeros(1024, dtype=np.int64)
    assert lib.swb_matrix_parse(str(g["text"]).encode(), m.ctypes.data) == 0
    assert np.array_equal(m, g["matrix"].astype(np.int64))                 # the reference's own file parser
    path = tmp_path / "m.mat"
    path.write_text(str(g["text"]))
    assert lib.swb_matrix_read(str(path).encode(), m.ctypes.data) == 0
    assert np.array_equal(m, g["matrix"].astype(np.int64))
    assert lib.swb_matrix_read(str(tmp_path / "missing").encode(), m.ctypes.data) == -7
    assert lib.swb_matrix_parse(b"   A  R\nA  4  x\n", m.ctypes.data) == -1


def test_karlin_altschul_tables(lib):
    g = json.load(open(os.path.join(GOLD, "host.json")))
    p = (C.c_double * 5)()
    seen = set()
    for name, go, ge, exp in g["protein"]:
        assert lib.swb_stats_params(name.encode(), go, ge, p) == 1
        assert list(p) == exp, (name, go, ge)
        seen.add((name, go, ge))
    assert len(seen) > 90
    for name in ("BLOSUM62", "PAM30", "NOSUCH"):
        for go, ge in ((11, 1), (5, 5), (0, 0), (32767, 32767)):
            if (name, go, ge) not in seen:
                assert lib.swb_stats_params(name.encode(), go, ge, p) == 0
    for r, q, go, ge, exp in g["nt"]:
        assert lib.swb_stats_params_nt(r, q, go, ge, p) == 1
        assert list(p) == exp, (r, q, go, ge)
    assert lib.swb_stats_params_nt(1, -6, 5, 2, p) == 0 and lib.swb_stats_params_nt(2, -3, 1, 1, p) == 0
    a, b = C.c_int64(), C.c_int64()
    for name, (go, ge) in g["prefs"].items():
        assert lib.swb_stats_default_gaps(name.encode(), a, b) == 1 and (a.value, b.value) == (go, ge)
    assert lib.swb_stats_default_gaps(b"NOSUCH", a, b) == 0


def test_length_adjustment(lib):
    g = json.load(open(os.path.join(GOLD, "host.json")))
    for lam, K, alpha, beta, qlen, dblen, nseq, exp in g["lenadj"]:
        got = lib.swb_stats_length_adjustment(K, alpha / lam, beta, qlen, dblen, nseq)
        assert got == exp, (lam, qlen, dblen, nseq)


def test_codon_tables_and_translation(lib):
    g = np.load(os.path.join(GOLD, "host_tables.npz"))
    t = np.zeros(4096, dtype=np.uint8)
    for code in range(1, 24):
        rc = lib.swb_translate_table(code, t.ctypes.data)
        if "code%d" % code in g:
            assert rc == 0 and np.array_equal(t, g["code%d" % code]), code
            assert lib.swb_gencode_name(code)
        else:
            assert rc == -1 and lib.swb_gencode_name(code) is None
    lib.swb_translate_table(1, t.ctypes.data)
    nt = scoring.encode_nucleotide("ATGGCNTAARAYTGA")                     # M A * (RAY = D/N -> B) *
    out = np.zeros(8, dtype=np.uint8)
    n = lib.swb_translate(nt.ctypes.data, nt.size, 0, 0, t.ctypes.data, out.ctypes.data)
    assert "".join(scoring.SYM_AA[c] for c in out[:n]) == "MA*B*"
    n = lib.swb_translate(nt.ctypes.data, nt.size, 0, 2, t.ctypes.data, out.ctypes.data)
    assert n == 4
    rc_nt = synth.revcomp_nt(nt)
    a = np.zeros(8, dtype=np.uint8)
    b = np.zeros(8, dtype=np.uint8)
    for frame in range(3):
        na = lib.swb_translate(nt.ctypes.data, nt.size, 1, frame, t.ctypes.data, a.ctypes.data)
        nb = lib.swb_translate(rc_nt.ctypes.data, rc_nt.size, 0, frame, t.ctypes.data, b.ctypes.data)
        assert na == nb and np.array_equal(a[:na], b[:nb])                  # strand 1 = frames of the reverse complement
    r = np.zeros(nt.size, dtype=np.uint8)
    lib.swb_revcomp(nt.ctypes.data, nt.size, r.ctypes.data)
    assert np.array_equal(r, rc_nt)


def _parse(lib, text, nucleotide):
    seq = np.zeros(len(text) + 1, dtype=np.uint8)
    n = C.c_int64()
    descr = C.create_string_buffer(len(text) + 2)
    used = lib.swb_query_parse(text, len(text), nucleotide, seq.ctypes.data, seq.size, n, descr, len(text) + 2)
    return used, seq[:n.value].copy(), descr.value.decode()


def test_query_parsing(lib):
    text = b">first query\nHEAG awg\nhe1e*-\n>second\nPAW\n"
    used, seq, d = _parse(lib, text, 0)
    assert d == "first query" and np.array_equal(seq, scoring.encode_protein("HEAGAWGHEE*-"))
    used2, seq2, d2 = _parse(lib, text[used:], 0)
    assert d2 == "second" and np.array_equal(seq2, scoring.encode_protein("PAW")) and used + used2 == len(text)
    assert _parse(lib, b"", 0)[0] == 0
    used, seq, d = _parse(lib, b"ACGTUNRY-x\nacgt", 1)                         # no header line; '-' and 'x' dropped
    assert d == "" and seq.tolist() == [1, 2, 4, 8, 8, 15, 5, 10, 1, 2, 4, 8]


def test_deflines(lib):
    def text(data, gis=0, taxid=0, memb=0, keep=None):
        buf = C.create_string_buffer(4096)
        need = C.c_int64()
        bitmap = None
        if keep is not None:
            bitmap = np.zeros(2048, dtype=np.uint8)
            for t in keep:
                bitmap[t // 8] |= 1 << (t & 7)
        n = lib.swb_defline_text(data, len(data), gis, taxid, memb, bitmap.ctypes.data if keep is not None else None,
                                 2048 if keep is not None else 0, buf, 4096, need)
        return n, buf.value.decode()
    assert text(blastdb._defline("s22", "subject 22")) == (1, "lcl|s22 subject 22")
    # hand-encoded: title, seqids { gi 12345, ref { accession NP_000001, version 2 } }, taxid 9606
    vs = lambda s: bytes([0x1A, len(s)]) + s.encode()                        # noqa: E731
    wrap = lambda tag, body: bytes([tag, 0x80]) + body + b"\0\0"            # noqa: E731
    gi = wrap(0xAB, bytes([0x02, 0x02, 0x30, 0x39]))
    ref = wrap(0xA9, wrap(0x30, wrap(0xA1, vs("NP_000001")) + wrap(0xA3, bytes([0x02, 0x01, 0x02]))))
    gnl = wrap(0xAA, wrap(0x30, wrap(0xA0, vs("mydb")) + wrap(0xA1, wrap(0xA0, bytes([0x02, 0x01, 0x07])))))
    d1 = wrap(0x30, wrap(0xA0, vs("some protein")) + wrap(0xA1, wrap(0x30, gi + ref)) +
              wrap(0xA2, bytes([0x02, 0x02, 0x25, 0x86])))
    d2 = wrap(0x30, wrap(0xA0, vs("another")) + wrap(0xA1, wrap(0x30, gnl)))
    data = wrap(0x30, d1 + d2)
    assert text(data) == (2, "ref|NP_000001.2| some protein\ngnl|mydb|7 another")
    assert text(data, gis=1)[1].startswith("gi|12345|ref|NP_000001.2| some protein")
    assert text(data, taxid=1)[1].startswith("ref|NP_000001.2||taxid|9606 some protein")
    assert text(data, keep=[9606]) == (1, "ref|NP_000001.2| some protein")     # -x list: only that taxid's defline
    assert text(data, keep=[1, 2])[0] == 0 and text(data, memb=2)[0] == 0
    assert text(b"\x31\x80\0\0")[0] == -7 and text(data[:20])[0] == -7      # not a defline set / truncated
