"""TEST INFRASTRUCTURE: writes BLAST version-4 databases (the format the reference reads,
database.cc:515-608, :1237-1401; asnparse.cc for the .phr deflines) so that the unmodified
reference CLI (oracle/_ref/swipe) and our reader/ingest see the same bytes.  makeblastdb is not
available offline; format facts are the ones listed at the end of SURVEY.md section 8."""
import struct

import numpy as np


def _defline(seq_id, title):
    """Minimal Blast-def-line-set, BER with indefinite lengths: parses as 'lcl|id title'."""
    t = title.encode()
    i = seq_id.encode()
    assert len(t) < 128 and len(i) < 128
    return (bytes([0x30, 0x80, 0x30, 0x80, 0xA0, 0x80, 0x1A, len(t)]) + t + bytes([0, 0]) +
            bytes([0xA1, 0x80, 0x30, 0x80, 0xA0, 0x80, 0xA1, 0x80, 0x1A, len(i)]) + i +
            bytes(12))


def _int(v):
    n = max(1, (int(v).bit_length() + 8) // 8)
    return bytes([0x02, n]) + int(v).to_bytes(n, "big")


def _wrap(tag, body):
    return bytes([tag, 0x80]) + body + b"\0\0"


def defline_set(deflines):
    """deflines: list of dicts {title, lcl, gi, taxid, memb}: one Blast-def-line each (asnparse.cc:753-887)."""
    out = b""
    for d in deflines:
        t = d["title"].encode()
        body = _wrap(0xA0, bytes([0x1A, len(t)]) + t)
        ids = b""
        if d.get("gi") is not None:
            ids += _wrap(0xAB, _int(d["gi"]))
        if d.get("lcl") is not None:
            i = d["lcl"].encode()
            ids += _wrap(0xA0, _wrap(0xA1, bytes([0x1A, len(i)]) + i))
        body += _wrap(0xA1, _wrap(0x30, ids))
        if d.get("taxid"):
            body += _wrap(0xA2, _int(d["taxid"]))
        if d.get("memb"):
            body += _wrap(0xA3, _wrap(0x30, _int(d["memb"])))
        out += _wrap(0x30, body)
    return _wrap(0x30, out)


def _index(path, protein, title, date, nseq, residues, longest, tables):
    t = title.encode()
    d = date.encode()
    buf = struct.pack(">II", 4, 1 if protein else 0)
    buf += struct.pack(">I", len(t)) + t + struct.pack(">I", len(d)) + d
    buf += bytes(-len(buf) % 4)                      # database.cc:587-592
    buf += struct.pack(">I", nseq) + struct.pack("<Q", residues) + struct.pack(">I", longest)
    for tab in tables:
        buf += np.asarray(tab, dtype=">u4").tobytes()
    with open(path, "wb") as f:
        f.write(buf)


def write_protein(basename, subjects, title="synthetic protein db", date="Oct 17, 2026  5:00 AM",
                  ids=None, headers=None):
    """subjects: uint8 arrays of NCBIstdaa codes 1..27.  Writes basename.pin/.psq/.phr.
    headers: optional list of defline_set() inputs, one per subject."""
    n = len(subjects)
    ids = ids or ["s%d" % i for i in range(n)]
    hdr, hoff = b"", [0]
    for i in range(n):
        hdr += defline_set(headers[i]) if headers else _defline(ids[i], "subject %d" % i)
        hoff.append(len(hdr))
    sq = bytearray(b"\0")
    soff = [1]
    for s in subjects:
        sq += np.asarray(s, dtype=np.uint8).tobytes() + b"\0"
        soff.append(len(sq))
    lens = [len(s) for s in subjects]
    _index(basename + ".pin", True, title, date, n, int(sum(lens)), max(lens + [0]), [hoff, soff])
    open(basename + ".psq", "wb").write(bytes(sq))
    open(basename + ".phr", "wb").write(hdr)
    return np.asarray(soff, dtype=np.int64)


_TWOBIT = np.zeros(16, dtype=np.uint8)
_TWOBIT[[1, 2, 4, 8]] = [0, 1, 2, 3]


def pack_nt(codes, big_table=False):
    """4-bit codes -> the .nsq record: 4 bases per byte MSB first, the last byte carries the
    remaining len%4 bases and their count in its low two bits (database.cc:1260-1261), then the
    ambiguity table (database.cc:1288-1322)."""
    c = np.asarray(codes, dtype=np.uint8)
    L = c.size
    two = _TWOBIT[c & 15]
    nfull = L // 4
    out = bytearray()
    if nfull:
        q = two[:4 * nfull].reshape(-1, 4)
        out += ((q[:, 0] << 6) | (q[:, 1] << 4) | (q[:, 2] << 2) | q[:, 3]).astype(np.uint8).tobytes()
    last = 0
    for k in range(L % 4):
        last |= int(two[4 * nfull + k]) << (6 - 2 * k)
    out.append(last | (L % 4))
    amb = ~np.isin(c, [1, 2, 4, 8])
    entries = []
    i = 0
    maxrun = 4096 if big_table else 16
    while i < L:
        if amb[i]:
            j = i
            while j < L and amb[j] and c[j] == c[i] and j - i < maxrun:
                j += 1
            entries.append((int(c[i]), j - i, i))
            i = j
        else:
            i += 1
    tail = b""
    if entries:
        if big_table:
            tail = struct.pack(">I", 0x80000000 | (2 * len(entries)))
            for code, run, off in entries:
                tail += struct.pack(">Q", (code << 60) | ((run - 1) << 48) | off)
        else:
            tail = struct.pack(">I", len(entries))
            for code, run, off in entries:
                assert off < (1 << 24)
                tail += struct.pack(">I", (code << 28) | ((run - 1) << 24) | off)
    return bytes(out), tail


def write_nucleotide(basename, subjects, title="synthetic nt db", date="Oct 17, 2026  5:00 AM",
                     ids=None, big_table=False):
    """subjects: uint8 arrays of 4-bit nt codes (A=1 C=2 G=4 T=8, others ambiguity).
    Writes basename.nin/.nsq/.nhr; returns (seq_offsets[n+1], amb_offsets[n+1])."""
    n = len(subjects)
    ids = ids or ["s%d" % i for i in range(n)]
    hdr, hoff = b"", [0]
    for i in range(n):
        hdr += _defline(ids[i], "subject %d" % i)
        hoff.append(len(hdr))
    sq = bytearray(b"\0")
    soff, aoff = [], []
    for s in subjects:
        soff.append(len(sq))
        packed, tail = pack_nt(s, big_table)
        sq += packed
        aoff.append(len(sq))
        sq += tail
    soff.append(len(sq))
    aoff.append(len(sq))
    lens = [len(s) for s in subjects]
    _index(basename + ".nin", False, title, date, n, int(sum(lens)), max(lens + [0]),
           [hoff, soff, aoff])
    open(basename + ".nsq", "wb").write(bytes(sq))
    open(basename + ".nhr", "wb").write(hdr)
    return np.asarray(soff, dtype=np.int64), np.asarray(aoff, dtype=np.int64)


AA = "-ABCDEFGHIKLMNPQRSTVWXYZU*OJ"
NT = {1: "A", 2: "C", 4: "G", 8: "T", 15: "N", 5: "R", 10: "Y", 3: "M", 12: "K", 6: "S", 9: "W",
      14: "B", 13: "D", 11: "H", 7: "V"}


def write_fasta(path, codes, protein=True, name="query"):
    c = np.asarray(codes, dtype=np.uint8)
    text = "".join(AA[x] for x in c) if protein else "".join(NT[int(x)] for x in c)
    with open(path, "w") as f:
        f.write(">%s\n" % name)
        for i in range(0, len(text), 60):
            f.write(text[i:i + 60] + "\n")


# ---- vectorised writers for large synthetic databases (bench.py's reference arm, scale tests) -------
def _fixed_headers(n):
    """n deflines of identical size ('lcl|s%08d subject'): (bytes, offsets[n+1])."""
    rec = np.frombuffer(_defline("s00000000", "subject"), dtype=np.uint8)
    at = bytes(rec).index(b"s00000000") + 1
    hdr = np.tile(rec, (n, 1))
    idx = np.arange(n, dtype=np.int64)
    for d in range(8):
        hdr[:, at + 7 - d] = 48 + (idx // 10 ** d) % 10
    return hdr.reshape(-1), np.arange(n + 1, dtype=np.int64) * rec.size


def write_protein_fast(basename, residues, offsets, title="synthetic protein db",
                       date="Oct 17, 2026  5:00 AM"):
    """Like write_protein for (residues, offsets) arrays of millions of subjects."""
    offsets = np.asarray(offsets, dtype=np.int64)
    n = offsets.size - 1
    total = int(offsets[-1] - offsets[0])
    res = np.asarray(residues, dtype=np.uint8)[offsets[0]: offsets[-1]]
    psq = np.zeros(total + n + 1, dtype=np.uint8)
    keep = np.ones(total + n + 1, dtype=bool)
    keep[(offsets - offsets[0]) + np.arange(n + 1)] = False        # the NUL before / after every subject
    psq[keep] = res
    soff = (offsets - offsets[0]) + np.arange(n + 1) + 1
    hdr, hoff = _fixed_headers(n)
    lens = offsets[1:] - offsets[:-1]
    _index(basename + ".pin", True, title, date, n, total, int(lens.max()) if n else 0, [hoff, soff])
    psq.tofile(basename + ".psq")
    hdr.tofile(basename + ".phr")
    return soff


def write_nucleotide_fast(basename, residues, offsets, title="synthetic nt db",
                          date="Oct 17, 2026  5:00 AM", chunk=500_000):
    """Like write_nucleotide for millions of short reads: 2-bit packing vectorised over chunks of
    reads; the few reads holding ambiguity codes get their table through pack_nt."""
    offsets = np.asarray(offsets, dtype=np.int64)
    n = offsets.size - 1
    res = np.asarray(residues, dtype=np.uint8)
    lens = offsets[1:] - offsets[:-1]
    plen = lens // 4 + 1
    pieces = [np.zeros(1, dtype=np.uint8)]
    tails = {}
    for c0 in range(0, n, chunk):
        c1 = min(n, c0 + chunk)
        L = lens[c0:c1]
        width = int(-(-int(L.max()) // 4) * 4 + 4)
        seg = res[offsets[c0]: offsets[c1]]
        amb = ~np.isin(seg, [1, 2, 4, 8])
        if amb.any():
            rid = np.unique(np.searchsorted(offsets[c0: c1 + 1], np.nonzero(amb)[0] + offsets[c0],
                                            side="right") - 1) + c0
            for r in rid:
                tails[int(r)] = pack_nt(res[offsets[r]: offsets[r + 1]])[1]
        grid = np.zeros((c1 - c0, width), dtype=np.uint8)
        grid[np.arange(width)[None, :] < L[:, None]] = _TWOBIT[seg & 15]
        by = (grid[:, 0::4] << 6) | (grid[:, 1::4] << 4) | (grid[:, 2::4] << 2) | grid[:, 3::4]
        by[np.arange(c1 - c0), L // 4] |= (L % 4).astype(np.uint8)
        pieces.append(by[np.arange(width // 4)[None, :] < plen[c0:c1, None]])
    sq = np.concatenate(pieces)
    soff = np.ones(n + 1, dtype=np.int64)
    soff[1:] += np.cumsum(plen)
    aoff = soff[1:].copy()                                  # no table: the next record follows the packed bases
    if tails:
        rid = np.array(sorted(tails), dtype=np.int64)
        tl = np.array([len(tails[int(r)]) for r in rid], dtype=np.int64)
        sq = np.insert(sq, np.repeat(soff[rid + 1], tl),
                       np.frombuffer(b"".join(tails[int(r)] for r in rid), dtype=np.uint8))
        shift = np.zeros(n + 1, dtype=np.int64)
        shift[rid + 1] = tl
        shift = np.cumsum(shift)
        aoff = soff[1:] + shift[:-1]
        soff = soff + shift
    aoff_full = np.concatenate([aoff, soff[-1:]])
    hdr, hoff = _fixed_headers(n)
    _index(basename + ".nin", False, title, date, n, int(lens.sum()), int(lens.max()) if n else 0,
           [hoff, soff, aoff_full])
    sq.tofile(basename + ".nsq")
    hdr.tofile(basename + ".nhr")
    return soff, aoff_full
