"""Multi-GPU plumbing: the database shards by sequence, each rank scans its own shard, and the
only exchange is the K best (seqno, score) pairs per rank, merged with the reference's hits_enter
ordering (hits.cc:163-222; the MPI build does the same through its master, swipe.cc:1957-1974).
No data-path collective."""
import numpy as np

from .api import topk_merge


def shard_bounds(nseq, world):
    """Contiguous sequence ranges [lo, hi) per rank, sizes differing by at most one."""
    base, extra = divmod(int(nseq), int(world))
    bounds, lo = [], 0
    for r in range(world):
        hi = lo + base + (1 if r < extra else 0)
        bounds.append((lo, hi))
        lo = hi
    return bounds


def local_topk(scores, seqno_base, keep, min_score=1, upper_score=2 ** 62):
    return topk_merge([scores], [seqno_base], keep, min_score, upper_score)


def merge_topk(lists, keep):
    """lists: iterable of (seqnos, scores) already filtered by the thresholds; returns the global
    top `keep` ordered by score descending, then sequence number descending."""
    seq = np.concatenate([np.asarray(a, dtype=np.int64) for a, _ in lists] + [np.zeros(0, np.int64)])
    sc = np.concatenate([np.asarray(b, dtype=np.int64) for _, b in lists] + [np.zeros(0, np.int64)])
    order = np.lexsort((-seq, -sc))[:keep]
    return seq[order], sc[order]


def gather_topk(local_seq, local_sc, keep, group=None, device=None):
    """All ranks contribute their local top-K; every rank returns the merged global top-K.  Uses
    torch.distributed (NCCL on GPUs, gloo on CPU) for the K x 2 int64 exchange."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    mine = torch.full((keep, 2), -1, dtype=torch.int64, device=device)
    n = len(local_seq)
    if n:
        mine[:n, 0] = torch.as_tensor(np.asarray(local_seq, dtype=np.int64), device=device)
        mine[:n, 1] = torch.as_tensor(np.asarray(local_sc, dtype=np.int64), device=device)
    gathered = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(gathered, mine, group=group)
    allhits = torch.cat(gathered).cpu().numpy()
    allhits = allhits[allhits[:, 0] >= 0]
    return merge_topk([(allhits[:, 0], allhits[:, 1])], keep)
