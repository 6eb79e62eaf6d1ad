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
