"""The N > 1 host path on CPU: two gloo ranks each own a shard, take the local top-K through the
C ABI's swb_topk_merge and exchange K pairs; the merged list must equal the single-shard list and
the oracle's hits_enter restatement.  (Scores come from the CPU oracle here: test infrastructure.)"""
import os
import socket
import sys

import numpy as np
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch.distributed as dist
    from oracle_lib import Oracle
    from swipe_b200 import scoring, shard, synth
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    q = synth.protein_query(60, seed=8)
    residues, offsets = synth.protein_db(501, query=q, seed=9, plant_every=5, max_len=200)
    lo, hi = shard.shard_bounds(501, world)[rank]
    m = scoring.blosum62()
    scores, _, _ = Oracle().scan(residues[offsets[lo]:offsets[hi]], offsets[lo:hi + 1] - offsets[lo],
                                 q, m, 11, 1, threads=1)
    seq, sc, _, _ = shard.local_topk(scores, lo, keep=40, min_score=20)
    gseq, gsc = shard.gather_topk(seq, sc, 40)
    np.savez(os.path.join(out_dir, "rank%d.npz" % rank), seq=gseq, sc=gsc)
    # the bench's N > 1 path: ONE database cut by residue count, per-rank hit lists (two "strands" here),
    # gathered and merged on rank 0 with swb_hits_merge
    a, b = shard.shard_cuts(offsets, world)[rank]
    sc2, _, _ = Oracle().scan(residues[offsets[a]:offsets[b]], offsets[a:b + 1] - offsets[a], q, m, 11, 1, threads=1)
    lists = []
    for strand in range(2):
        s_, v_, _, _ = shard.local_topk(sc2 + strand, a, keep=25, min_score=20)
        lists.append((s_ * 2 + strand, v_))
    merged = shard.HitExchange(25, 2)(lists)
    if rank == 0:
        np.savez(os.path.join(out_dir, "exchange.npz"), seq=merged[0], sc=merged[1], cut=np.array([a, b]))
    else:
        assert merged is None
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_topk_merge(tmp_path, oracle):
    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    from swipe_b200 import scoring, synth
    q = synth.protein_query(60, seed=8)
    residues, offsets = synth.protein_db(501, query=q, seed=9, plant_every=5, max_len=200)
    scores, _, _ = oracle.scan(residues, offsets, q, scoring.blosum62(), 11, 1)
    want_seq, want_sc, _, _ = oracle.topk(np.arange(501), scores, 40, min_score=20)
    for r in range(2):
        got = np.load(os.path.join(str(tmp_path), "rank%d.npz" % r))
        assert np.array_equal(got["seq"], want_seq) and np.array_equal(got["sc"], want_sc)
    # HitExchange over residue-balanced shards and two lists per rank = the sink over everything
    ex = np.load(os.path.join(str(tmp_path), "exchange.npz"))
    allseq = np.concatenate([np.arange(501) * 2, np.arange(501) * 2 + 1])
    allsc = np.concatenate([scores, scores + 1])
    want2_seq, want2_sc, _, _ = oracle.topk(allseq, allsc, 25, min_score=20)
    # (strand 1 was filtered at min_score 20 AFTER the +1, like strand 0: same rule on both)
    assert np.array_equal(ex["seq"], want2_seq) and np.array_equal(ex["sc"], want2_sc)
    half = int(offsets[-1]) // 2
    assert abs(int(offsets[ex["cut"][1]]) - half) <= int((offsets[1:] - offsets[:-1]).max())
