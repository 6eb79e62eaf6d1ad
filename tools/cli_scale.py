"""Scale check of the command-line front end: writes a synthetic BLAST-v4 protein database of N
sequences (vectorised writer), runs swipe-b200 on it with 1..G GPUs and compares the reported hits with
the library's own top-K over the same residues.  usage: python tools/cli_scale.py [nseq] [gpus]"""
import os
import re
import struct
import subprocess
import sys
import tempfile
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import blastdb
from swipe_b200 import Database, Scoring, build, scoring, synth, topk_merge

nseq = int(sys.argv[1]) if len(sys.argv) > 1 else 2_000_000
gpus = [int(x) for x in sys.argv[2].split(",")] if len(sys.argv) > 2 else [1]
tmp = tempfile.mkdtemp()
q = synth.protein_query(375)
t0 = time.time()
res, off = synth.protein_db(nseq, query=q)
lens = (off[1:] - off[:-1]).astype(np.int64)
# .psq: NUL, then every sequence followed by NUL
psq = np.zeros(int(off[-1]) + nseq + 1, dtype=np.uint8)
starts = off[:-1] + np.arange(nseq) + 1
idx = np.arange(int(off[-1]), dtype=np.int64) + np.repeat(np.arange(nseq, dtype=np.int64) + 1, lens)
psq[idx] = res
psq.tofile(os.path.join(tmp, "big.psq"))
soff = np.concatenate([starts, [psq.size]]).astype(">u4")
one = blastdb._defline("s0", "x")
hdr = bytearray()
hoff = np.zeros(nseq + 1, dtype=np.int64)
for i in range(nseq):
    d = blastdb._defline("s%d" % i, "synthetic")
    hdr += d
    hoff[i + 1] = len(hdr)
open(os.path.join(tmp, "big.phr"), "wb").write(bytes(hdr))
t = b"synthetic protein db"
d = b"Oct 17, 2026  5:00 AM"
buf = struct.pack(">II", 4, 1) + struct.pack(">I", len(t)) + t + struct.pack(">I", len(d)) + d
buf += bytes(-len(buf) % 4)
buf += struct.pack(">I", nseq) + struct.pack("<Q", int(off[-1])) + struct.pack(">I", int(lens.max()))
open(os.path.join(tmp, "big.pin"), "wb").write(buf + hoff.astype(">u4").tobytes() + soff.tobytes())
blastdb.write_fasta(os.path.join(tmp, "q.fa"), q)
print("database written: %d sequences, %d residues, %.1f s" % (nseq, int(off[-1]), time.time() - t0), flush=True)

with Database(res, off) as db:
    s = db.search(q, Scoring(scoring.blosum62(), 11, 1))
exe = build.build_cli()
for g in gpus:
    t0 = time.time()
    r = subprocess.run([exe, "-d", "big", "-i", "q.fa", "-m", "8", "-b", "50", "-v", "50", "-a", str(g)], cwd=tmp,
                       capture_output=True, text=True)
    wall = time.time() - t0
    assert r.returncode == 0, r.stderr
    rows = [ln.split("\t") for ln in r.stdout.strip().splitlines()]
    ids = [int(x[1][5:]) for x in rows]
    # expected: hits above the E-value threshold in (score desc, seqno desc) order
    exp_seq, exp_sc, _, _ = topk_merge([s], [0], 50, min_score=1)
    n = len(ids)
    ok = ids == exp_seq[:n].tolist()
    r0 = subprocess.run([exe, "-d", "big", "-i", "q.fa", "-v", "5", "-b", "0", "-a", str(g)], cwd=tmp, capture_output=True, text=True)
    speed = re.search(r"Speed:\s+([0-9.]+) GCUPS", r0.stdout)
    elapsed = re.search(r"Elapsed:\s+([0-9.]+)s", r0.stdout)
    print("gpus %d: wall %.2f s, %d hits reported, order matches library top-K: %s, program's own search phase: %s s / %s GCUPS" % (
        g, wall, n, ok, elapsed.group(1) if elapsed else "?", speed.group(1) if speed else "?"), flush=True)
