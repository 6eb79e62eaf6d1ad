"""CPU-side checks: the C-ABI library loads and exports what include/swipe_b200.h declares, the
host-only entry points behave, and the product path refuses to run without a GPU."""
import ctypes
import os
import re

import numpy as np
import pytest

import swipe_b200
from swipe_b200 import api, scoring, shard, synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    lib = swipe_b200.load_library()
    header = open(os.path.join(ROOT, "include", "swipe_b200.h")).read()
    declared = set(re.findall(r"\b(swb_[a-z0-9_]+)\s*\(", header))
    assert declared == set(api.EXPORTS)
    for name in declared:
        assert getattr(lib, name) is not None
    assert lib.swb_abi_version() == 2
    assert lib.swb_strerror(0) == b"ok"
    assert b"no CPU path" in lib.swb_strerror(-2)


def test_product_has_no_cpu_fallback():
    """Without a device the compute entry points fail loudly; with one this test is moot."""
    try:
        n = swipe_b200.device_count()
    except swipe_b200.SwbError as e:
        assert e.status == -2
        with pytest.raises(swipe_b200.SwbError):
            swipe_b200.Database(np.zeros(4, np.uint8), np.array([0, 4], np.int64))
        return
    assert n >= 1


def test_product_never_imports_the_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "swipe_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cc", ".cpp", ".inc", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "oracle_lib" not in text and "liborc" not in text and "libswipe_ref" not in text, f


def test_argument_checks_do_not_need_a_gpu():
    lib = swipe_b200.load_library()
    h = ctypes.c_void_p()
    off = np.array([0, 4], np.int64)
    assert lib.swb_db_open(0, None, off.ctypes.data, -1, 0, None, ctypes.byref(h)) == -1
    bad = np.array([4, 0], np.int64)
    res = np.zeros(4, np.uint8)
    assert lib.swb_db_open(0, res.ctypes.data, bad.ctypes.data, 1, 0, None, ctypes.byref(h)) == -1
    assert lib.swb_db_close(None) == 0
    assert lib.swb_search(None, None, 0, None, None, None) == -1


def test_topk_merge_matches_oracle_rule(oracle):
    rng = np.random.default_rng(11)
    shards = [rng.integers(0, 60, size=n) for n in (1000, 1, 0, 777)]
    bases = [0, 1000, 1001, 1001]
    seq, sc, tot, obv = swipe_b200.topk_merge(shards, bases, keep=50, min_score=10, upper_score=57)
    allsc = np.concatenate(shards)
    allseq = np.arange(allsc.size)
    oseq, osc, otot, oobv = oracle.topk(allseq, allsc, 50, min_score=10, upper=57)
    assert np.array_equal(seq, oseq) and np.array_equal(sc, osc) and (tot, obv) == (otot, oobv)
    # the same list whatever the shard split (what makes 1/2/4/8-GPU output identical)
    seq1, sc1, _, _ = swipe_b200.topk_merge([allsc], [0], keep=50, min_score=10, upper_score=57)
    assert np.array_equal(seq, seq1) and np.array_equal(sc, sc1)
    # ties: higher sequence number first (hits.cc:188-191)
    seq2, sc2, _, _ = swipe_b200.topk_merge([np.array([5, 5, 5, 9])], [0], keep=3)
    assert seq2.tolist() == [3, 2, 1] and sc2.tolist() == [9, 5, 5]
    assert swipe_b200.topk_merge([np.array([1, 2])], [0], keep=0)[0].size == 0


def test_shard_helpers():
    assert shard.shard_bounds(10, 3) == [(0, 4), (4, 7), (7, 10)]
    assert shard.shard_bounds(2, 4) == [(0, 1), (1, 2), (2, 2), (2, 2)]
    s, c = shard.merge_topk([(np.array([1, 2]), np.array([7, 9])), (np.array([9]), np.array([7]))], 2)
    assert s.tolist() == [2, 9] and c.tolist() == [9, 7]


def test_symbol_maps_and_encoders():
    assert scoring.SYM_AA.index("A") == 1 and scoring.SYM_AA.index("J") == 27
    assert scoring.encode_protein("a-*xZ?1").tolist() == [1, 0, 25, 21, 23]
    assert scoring.encode_nucleotide("ACGTUN-x").tolist() == [1, 2, 4, 8, 8, 15]
    assert synth.revcomp_nt(np.array([1, 2, 4, 8, 15, 3], np.uint8)).tolist() == [12, 15, 1, 2, 4, 8]


def test_synthetic_inputs_are_deterministic_and_shaped():
    q = synth.protein_query(375)
    assert q.size == 375 and q.min() >= 1 and q.max() <= 22
    r1, o1 = synth.protein_db(3000, query=q, seed=5)
    r2, o2 = synth.protein_db(3000, query=q, seed=5)
    assert np.array_equal(r1, r2) and np.array_equal(o1, o2)
    lens = o1[1:] - o1[:-1]
    assert lens.min() >= 25 and lens.max() <= 5000 and 250 < lens.mean() < 450
    d, o = synth.dna_db(100)
    assert set(np.unique(d).tolist()) <= {1, 2, 4, 8} and (o[1:] - o[:-1]).min() >= 150


def test_hits_merge_equals_the_dense_sink(oracle):
    """swb_hits_merge over per-shard lists (each the shard's own top-K in the sink's order) gives
    the list hits_enter would hold after seeing every subject (hits.cc:163-222)."""
    rng = np.random.default_rng(21)
    allsc = rng.integers(0, 40, size=5000)
    cuts = [0, 1200, 1200, 3100, 5000]
    lists = []
    for lo, hi in zip(cuts[:-1], cuts[1:]):
        s, v, _, _ = swipe_b200.topk_merge([allsc[lo:hi]], [lo], keep=64, min_score=5, upper_score=38)
        lists.append((s, v))
    seq, sc = swipe_b200.hits_merge(lists, 64)
    oseq, osc, _, _ = oracle.topk(np.arange(allsc.size), allsc, 64, min_score=5, upper=38)
    assert np.array_equal(seq, oseq) and np.array_equal(sc, osc)
    assert swipe_b200.hits_merge([], 10)[0].size == 0
    assert swipe_b200.hits_merge(lists, 0)[0].size == 0
    short, _ = swipe_b200.hits_merge([(np.array([7, 3]), np.array([9, 9]))], 10)
    assert short.tolist() == [7, 3]


def test_batch_and_sink_argument_checks_do_not_need_a_gpu():
    lib = swipe_b200.load_library()
    n = ctypes.c_int64()
    assert lib.swb_search_hits(None, None, 0, None, 0, 10, 1, 100, None, None, ctypes.byref(n), None, None, None) == -1
    assert lib.swb_search_batch(None, 1, None, None, None, None, None) == -1
    assert lib.swb_search_hits_batch(None, -1, None, None, None, 0, 1, 0, 0, None, None, None, None, None, None) == -1
    assert lib.swb_set_cache_limit(-5) == -1 and lib.swb_set_cache_limit(8 << 30) == 0
    assert lib.swb_set_geometry(None, 1) == -1
    assert lib.swb_db_set_filter(None, None) == -1
    assert lib.swb_alu_peak(0, None, None) == -1


def test_residue_balanced_shard_cuts():
    """shard_cuts: contiguous ranges covering every sequence once, each within one sequence of an equal
    share of the residues -- the cut swipe-b200 -a N and bench.py --gpus N make."""
    rng = np.random.default_rng(3)
    lens = rng.integers(1, 5000, size=20000)
    off = np.concatenate([[0], np.cumsum(lens)])
    for world in (1, 2, 3, 4, 8):
        cuts = shard.shard_cuts(off, world)
        assert cuts[0][0] == 0 and cuts[-1][1] == lens.size
        assert all(cuts[r][1] == cuts[r + 1][0] for r in range(world - 1))
        share = off[-1] / world
        for a, b in cuts:
            assert abs((off[b] - off[a]) - share) <= lens.max()
    # more ranks than sequences: empty tail shards, nothing lost
    tiny = shard.shard_cuts(np.array([0, 10, 30]), 4)
    assert tiny[0][0] == 0 and tiny[-1][1] == 2 and sum(b - a for a, b in tiny) == 2
    ex = shard.HitExchange(10, 2)                       # no process group: the single-rank path merges locally
    seq, sc = ex([(np.array([5, 1]), np.array([9, 3])), (np.array([7]), np.array([9]))])
    assert seq.tolist() == [7, 5, 1] and sc.tolist() == [9, 9, 3]
