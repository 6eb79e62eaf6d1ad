"""Multi-GPU plumbing: ONE database is cut by residue count into one shard per GPU, each rank scans its
shard and selects its hits on the device (swb_search_hits), and the only exchange is the K best
(seqno, score) pairs per rank, merged on the host with the reference's hits_enter ordering
(swb_hits_merge; hits.cc:163-222 -- the reference's MPI build does the same through its master,
swipe.cc:1957-1974).  No data-path collective."""
import numpy as np

from .api import topk_merge, hits_merge


def shard_bounds(nseq, world):
    """Contiguous sequence ranges [lo, hi) per rank, sizes differing by at most one."""
    base, extra = divmod(int(nseq), int(world))
    bounds, lo = [], 0
    for r in range(world):
        hi = lo + base + (1 if r < extra else 0)
        bounds.append((lo, hi))
        lo = hi
    return bounds


def shard_cuts(offsets, world):
    """Sequence ranges [lo, hi) per rank holding equal shares of the RESIDUES -- the cut swipe-b200 -a N
    makes (swipe_main.cpp), so that every GPU gets the same number of DP cells."""
    offsets = np.asarray(offsets)
    total = int(offsets[-1])
    cuts = [0]
    for r in range(1, world):
        cuts.append(max(cuts[-1], int(np.searchsorted(offsets, total * r // world, side="left"))))
    cuts.append(int(offsets.size - 1))
    return [(cuts[r], max(cuts[r + 1], cuts[r])) for r in range(world)]


def local_topk(scores, seqno_base, keep, min_score=1, upper_score=2 ** 62):
    return topk_merge([scores], [seqno_base], keep, min_score, upper_score)


def merge_topk(lists, keep):
    """lists: iterable of (seqnos, scores) already filtered by the thresholds; returns the global
    top `keep` ordered by score descending, then sequence number descending."""
    seq = np.concatenate([np.asarray(a, dtype=np.int64) for a, _ in lists] + [np.zeros(0, np.int64)])
    sc = np.concatenate([np.asarray(b, dtype=np.int64) for _, b in lists] + [np.zeros(0, np.int64)])
    order = np.lexsort((-seq, -sc))[:keep]
    return seq[order], sc[order]


class HitExchange:
    """Gathers every rank's hit lists (each in the sink's order, at most `keep` entries, `nlists` per
    rank -- one per query strand) and merges them on rank 0 with swb_hits_merge.  Buffers are
    allocated once; NCCL needs device tensors (pinned staging on both sides), gloo takes CPU tensors."""

    def __init__(self, keep, nlists=1, device=None, stream=None):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.keep, self.nlists = int(keep), int(nlists)
        self.world = dist.get_world_size() if dist.is_initialized() else 1
        self.rank = dist.get_rank() if dist.is_initialized() else 0
        self.device, self.stream = device, stream
        shape = (self.keep * self.nlists, 2)
        self.mine = torch.empty(shape, dtype=torch.int64)
        self.all = torch.empty((self.world,) + shape, dtype=torch.int64)
        if device is not None:
            self.mine, self.all = self.mine.pin_memory(), self.all.pin_memory()
            self.mine_dev = torch.empty(shape, dtype=torch.int64, device=device)
            self.all_dev = [torch.empty(shape, dtype=torch.int64, device=device) for _ in range(self.world)]
        self.mine.fill_(-1)
        self.mine_np, self.all_np = self.mine.numpy(), self.all.numpy()

    def __call__(self, lists):
        """lists: this rank's [(seqnos, scores)] * nlists.  Returns the merged (seqnos, scores) on rank 0,
        None elsewhere."""
        torch, dist = self.torch, self.dist
        if self.world == 1:
            return hits_merge(lists, self.keep) if len(lists) > 1 else lists[0]
        mine = self.mine_np                              # numpy view of the (pinned) staging tensor
        for k, (seq, sc) in enumerate(lists):
            n = len(seq)
            base = k * self.keep
            mine[base: base + n, 0] = seq
            mine[base: base + n, 1] = sc
            if n < self.keep:
                mine[base + n, 0] = -1                   # end marker: a list is sorted, so the first -1 ends it
        if self.device is not None:
            ctx = torch.cuda.stream(self.stream) if self.stream is not None else torch.cuda.stream(torch.cuda.current_stream())
            with ctx:
                self.mine_dev.copy_(self.mine, non_blocking=True)
                dist.all_gather(self.all_dev, self.mine_dev)
                if self.rank == 0:
                    for r in range(self.world):
                        self.all[r].copy_(self.all_dev[r], non_blocking=True)
            (self.stream or torch.cuda.current_stream()).synchronize()
        else:
            parts = [torch.empty_like(self.mine) for _ in range(self.world)]
            dist.all_gather(parts, self.mine)
            for r in range(self.world):
                self.all[r].copy_(parts[r])
        if self.rank != 0:
            return None
        g = self.all_np
        parts = []
        for r in range(self.world):
            for k in range(self.nlists):
                blk = g[r, k * self.keep: (k + 1) * self.keep]
                neg = np.flatnonzero(blk[:, 0] < 0)
                n = int(neg[0]) if neg.size else self.keep
                parts.append((blk[:n, 0], blk[:n, 1]))
        return hits_merge(parts, self.keep)


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
