#!/usr/bin/env python
"""bench.py -- GCUPS of the score-only Smith-Waterman database scan on N B200s.

Workloads (BASELINE.json `configs`), selected with --config:

  protein375 (default)  configs[1] at N=1: 375-aa query, BLOSUM62 11/1, 5,000,000-sequence synthetic
                        protein database (log-normal lengths, ~1.76 G residues, 0.1 % planted homologs).
                        At N>1 it is configs[4]: THE SAME database cut by residue count into N shards,
                        one per GPU; every step each rank scans its shard and selects its hits on the
                        device (swb_search_hits), the K (seqno, score) pairs per rank are gathered and
                        rank 0 merges them with swb_hits_merge INSIDE the timed region ("scaling":
                        "strong").  The merged list must equal the 1-GPU list (`topk_identical`).
                        A secondary `weak` block times a full 5 M shard per GPU.
  qlen100 / qlen1000 / qlen5000   configs[2]: the same database, other query lengths (N=1).
  nt50m                 configs[3]: 1000-nt query, +1/-3, gaps 5/2, both strands (two scans, the second
                        with the reverse-complemented query, query.cc:337-342) against 50 M reads (N=1).

  value : GCUPS = 1e-9 * residues * qlen [* 2 strands] / s (swipe.cc:1744-1775), database resident in
          HBM, K steps timed with CUDA events on the handle's stream, max over ranks.
  e2e   : the same through the C ABI from pinned HOST buffers: every step opens the shard (host ->
          device copy + device re-layout, overlapped with the scan), searches, and reads the hit
          list back; with N>1 the gather and merge are inside as well.
  roofline : the scan is bound by the issue rate of the packed 16x2 DPX instructions, which
          swb_alu_peak measures in this run; HBM is reported as a secondary block.
  --impl reference : the UNMODIFIED reference program (oracle/_ref/swipe, built from /root/reference)
          on all host cores over a bounded sample of the same workload, its own "Speed:" line.
"""
import argparse
import json
import os
import re
import shutil
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

TOPK = 250                                  # the reference keeps max(-v, -b) = 250 hits by default (hits.cc:287)

CONFIGS = {
    "protein375": dict(kind="protein", qlen=375, nseq=5_000_000, gap_open=11, gap_extend=1,
                       workload="375-aa query vs 5M-seq synthetic protein DB, BLOSUM62 11/1 (BASELINE configs[1]; "
                                "sharded over the GPUs at N>1 = configs[4])"),
    "qlen100": dict(kind="protein", qlen=100, nseq=5_000_000, gap_open=11, gap_extend=1,
                    workload="100-aa query vs 5M-seq synthetic protein DB, BLOSUM62 11/1 (BASELINE configs[2])"),
    "qlen1000": dict(kind="protein", qlen=1000, nseq=5_000_000, gap_open=11, gap_extend=1,
                     workload="1000-aa query vs 5M-seq synthetic protein DB, BLOSUM62 11/1 (BASELINE configs[2])"),
    "qlen5000": dict(kind="protein", qlen=5000, nseq=5_000_000, gap_open=11, gap_extend=1,
                     workload="5000-aa query vs 5M-seq synthetic protein DB, BLOSUM62 11/1 (BASELINE configs[2])"),
    "nt50m": dict(kind="nt", qlen=1000, nseq=50_000_000, gap_open=5, gap_extend=2,
                  workload="1000-nt query, +1/-3, gaps 5/2, both strands vs 50M-read synthetic DNA DB "
                           "(BASELINE configs[3])"),
}


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="protein375", choices=sorted(CONFIGS))
    ap.add_argument("--nseq", type=int, default=0, help="override the database size (subjects)")
    ap.add_argument("--qlen", type=int, default=0, help="override the query length")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-weak", action="store_true")
    ap.add_argument("--no-product-path", action="store_true")
    ap.add_argument("--shape", default="", help="G,R,lane_mode test hook")
    ap.add_argument("--batch", type=int, default=1,
                    help="distinct queries per step, searched with swb_search_hits_batch (protein configs, N=1)")
    return ap.parse_args()


class Workload:
    """The synthetic inputs of one config: queries (one per strand), database, scoring."""

    def __init__(self, cfg, nseq, qlen, batch=1):
        from swipe_b200 import scoring, synth
        self.kind = cfg["kind"]
        self.gap_open, self.gap_extend = cfg["gap_open"], cfg["gap_extend"]
        if self.kind == "protein":
            q = synth.protein_query(qlen) if qlen == 375 else synth.protein_query(qlen, seed=20261017 + qlen)
            self.queries = [q] + [synth.protein_query(qlen, seed=20261017 + qlen + 31 * k) for k in range(1, batch)]
            # planted homologs always derive from the 375-aa query so every config scans the same database
            self.residues, self.offsets = synth.protein_db(nseq, query=synth.protein_query(375))
            self.matrix = scoring.blosum62()
            self.matrix_name = "BLOSUM62"
        else:
            q = synth.dna_query(qlen)
            self.queries = [q, synth.revcomp_nt(q)]
            self.residues, self.offsets = synth.dna_db_planted(nseq, q, ambiguity_every=10)
            self.matrix = scoring.nucleotide_matrix(1, -3)
            self.matrix_name = "+1/-3"
        self.nseq = int(self.offsets.size - 1)
        self.qlen = int(qlen)
        self.total_res = int(self.offsets[-1])
        self.cells = float(self.total_res) * self.qlen * len(self.queries)


class ClockSampler(threading.Thread):
    """Samples SM clocks and throttle reasons while the timed region runs (NVML, else nvidia-smi)."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.samples = []
        self.reasons = set()
        self.max_mhz = None
        self._halt = threading.Event()

    def run(self):
        try:
            self._run_nvml()
        except Exception:
            self._run_smi()

    def _run_nvml(self):
        import pynvml
        pynvml.nvmlInit()
        try:
            h = pynvml.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM))
            bits = {"hw_slowdown": pynvml.nvmlClocksThrottleReasonHwSlowdown,
                    "hw_thermal_slowdown": pynvml.nvmlClocksThrottleReasonHwThermalSlowdown,
                    "sw_thermal_slowdown": pynvml.nvmlClocksThrottleReasonSwThermalSlowdown,
                    "sw_power_cap": pynvml.nvmlClocksThrottleReasonSwPowerCap}
            while not self._halt.is_set():
                self.samples.append(float(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)))
                r = pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                for name, bit in bits.items():
                    if r & bit:
                        self.reasons.add(name)
                self._halt.wait(0.02)
        finally:
            pynvml.nvmlShutdown()

    def _run_smi(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        while not self._halt.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q,
                                      "--format=csv,noheader,nounits"], capture_output=True,
                                     text=True, timeout=5).stdout.strip().split(",")
                self.samples.append(float(out[0]))
                self.max_mhz = float(out[1])
                for name, v in zip(names, out[2:]):
                    if v.strip().lower() == "active":
                        self.reasons.add(name)
            except Exception:
                pass
            self._halt.wait(0.1)

    def stop(self):
        self._halt.set()
        self.join(timeout=10)
        med = float(np.median(self.samples)) if self.samples else None
        return {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(self.samples)}


# ------------------------------------------------------------------------------------------------
# CPU legs: the only places that touch oracle/ (test infrastructure used as the reported baseline)

def harness_scan(w, budget_s, threads):
    """The reference's own search7/search16/fullsw cascade (oracle/_ref, driven in memory by
    oracle/ref_harness.cc) on a bounded prefix of the database, all strands.  Returns
    (gcups, kind, description, [scores per strand])."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib
    kind = "reference" if oracle_lib.ref_available() else "port"
    if kind == "reference":
        ref = oracle_lib.Ref()
        if w.kind == "protein":
            ref.matrix_init("BLOSUM62")
        else:
            ref.matrix_init("x", symtype=0, match=1, mismatch=-3)

        def scan(n, q):
            return ref.scan(w.residues[: w.offsets[n]], w.offsets[: n + 1], q, w.gap_open, w.gap_extend,
                            threads=threads, chunk=1024, ssse3=1)[0]
    else:
        orc = oracle_lib.Oracle()

        def scan(n, q):
            return orc.scan(w.residues[: w.offsets[n]], w.offsets[: n + 1], q, w.matrix, w.gap_open,
                            w.gap_extend, threads=threads)[0]

    n0 = min(w.nseq, 20000)
    scan(n0, w.queries[0])                     # cold: page faults, thread start-up
    t0 = time.perf_counter()
    scan(n0, w.queries[0])
    rate = float(w.offsets[n0]) * w.qlen / max(time.perf_counter() - t0, 1e-6)       # cells / s
    per_seq = w.qlen * (w.offsets[n0] / n0) * len(w.queries)
    n = int(min(w.nseq, max(n0, budget_s * rate / per_seq)))
    t0 = time.perf_counter()
    outs = [np.asarray(scan(n, q)[:n]) for q in w.queries]
    dt = time.perf_counter() - t0
    cells = float(w.offsets[n]) * w.qlen * len(w.queries)
    return cells / dt * 1e-9, kind, "first %d subjects (%d residues) of the database, %d strand(s), %.1f s" % (
        n, int(w.offsets[n]), len(w.queries), dt), outs


def write_blast_sample(w, n, basename):
    """The first n subjects as a BLAST v4 database the unmodified reference program reads."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import blastdb
    if w.kind == "protein":
        blastdb.write_protein_fast(basename, w.residues[: w.offsets[n]], w.offsets[: n + 1])
    else:
        blastdb.write_nucleotide_fast(basename, w.residues[: w.offsets[n]], w.offsets[: n + 1])
    blastdb.write_fasta(basename + ".query.fa", w.queries[0], protein=w.kind == "protein")


def reference_cli(w, basename, threads):
    """One run of oracle/_ref/swipe -a threads -v 10 -b 0 (alignment phase idle):
    (its own "Speed:" GCUPS, its "Elapsed:" seconds, external wall seconds)."""
    exe = os.path.join(ROOT, "oracle", "_ref", "swipe")
    cmd = [exe, "-d", basename, "-i", basename + ".query.fa", "-a", str(threads), "-v", "10", "-b", "0"]
    if w.kind == "nt":
        cmd += ["-p", "0", "-r", "1", "-q", "-3"]
    t0 = time.perf_counter()
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=1800).stdout
    wall = time.perf_counter() - t0
    m = re.search(r"Speed:\s+([0-9.]+) GCUPS", out)
    e = re.search(r"Elapsed:\s+([0-9.]+)s", out)
    if not m:
        raise RuntimeError("reference program printed no Speed line:\n" + out[-2000:])
    return float(m.group(1)), float(e.group(1)) if e else 0.0, wall


def bind_to_gpu_numa_node(index):
    """Multi-rank runs: keep this rank's host threads and its pinned buffers on the NUMA node the GPU
    hangs off, so that simultaneous uploads do not cross sockets.  Best effort."""
    try:
        import torch
        p = torch.cuda.get_device_properties(index)
        bdf = "%04x:%02x:%02x.0" % (p.pci_domain_id, p.pci_bus_id, p.pci_device_id)
        node = int(open("/sys/bus/pci/devices/%s/numa_node" % bdf).read())
        if node < 0:
            return
        cpus = set()
        for part in open("/sys/devices/system/node/node%d/cpulist" % node).read().strip().split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
    except Exception:
        pass


def base_config(args, cfg, w_nseq, qlen, world):
    return {"workload": cfg["workload"], "name": args.config, "qlen": qlen, "nseq": w_nseq,
            "strands": 2 if cfg["kind"] == "nt" else 1,
            "gap_open": cfg["gap_open"], "gap_extend": cfg["gap_extend"],
            "matrix": "BLOSUM62" if cfg["kind"] == "protein" else "+1/-3",
            "sharding": "one database cut by residue count into %d shard(s), no data-path collective; "
                        "K (seqno, score) pairs per rank gathered and merged on the host" % world,
            "keep": TOPK,
            "l2": "inputs (%.1f GB database) larger than L2" % (w_nseq * (350 if cfg["kind"] == "protein" else 200) / 1e9),
            "lanes": "two int16 lanes per 32-bit register: DPX s16x2 max / add-max, adds as fp16x2 on integer "
                     "bit patterns (exact to 2047), re-queue to plain int16 lanes and then 32/64-bit cells"}


# ------------------------------------------------------------------------------------------------
def run_reference(args, cfg, nseq, qlen):
    """The reference arm: the unmodified program's own number on this box's host cores."""
    cores = os.cpu_count() or 1
    sample = min(nseq, 2_000_000 if cfg["kind"] == "protein" else 8_000_000)
    if qlen >= 1000 and cfg["kind"] == "protein":
        sample = min(sample, 1_000_000)
    w = Workload(cfg, sample, qlen)
    config = base_config(args, cfg, nseq, qlen, 1)
    exe = os.path.join(ROOT, "oracle", "_ref", "swipe")
    line = {"metric": "GCUPS", "unit": "GCUPS", "impl": "reference", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "higher_is_better": True,
            "scaling": "strong" if args.gpus > 1 else "weak", "vs_baseline": None, "dtype": "int8",
            "data": "synthetic", "config": config}
    if os.path.exists(exe):
        tmp = tempfile.mkdtemp(prefix="swb_ref_")
        try:
            base = os.path.join(tmp, "db")
            write_blast_sample(w, w.nseq, base)
            speeds, elapsed, walls = [], [], []
            for it in range(args.warmup + args.steps):
                s, e, wl = reference_cli(w, base, cores)
                if it >= args.warmup:
                    speeds.append(s); elapsed.append(e); walls.append(wl)
            # one thread, on a tenth of the sample (per-core figure)
            n1 = max(1000, w.nseq // 10)
            base1 = os.path.join(tmp, "db1")
            w1 = w
            if n1 < w.nseq:
                write_blast_sample(w, n1, base1)
            else:
                base1 = base
            s1, _, _ = reference_cli(w1, base1, 1)
        finally:
            shutil.rmtree(tmp, ignore_errors=True)
        v = float(np.median(speeds))
        cells = w.cells
        # "Speed" is cells / the program's own Elapsed (10 ms ticks): the step time is its search phase
        ms = float(np.median(elapsed)) * 1e3
        desc = ("oracle/_ref/swipe -a %d -v 10 -b 0 on the first %d subjects (%d residues) written as a BLAST v4 "
                "database; median of %d runs of its own Speed line" % (cores, w.nseq, w.total_res, len(speeds)))
        line.update({"value": v, "ms_per_step": ms,
                     "cpu_baseline": {"value": v, "unit": "GCUPS", "cores": cores, "kind": "reference", "sample": desc,
                                      "best": float(np.max(speeds)), "median": v, "runs": [round(x, 2) for x in speeds],
                                      "one_thread_gcups": s1, "wall_s_median": float(np.median(walls)),
                                      "elapsed_s_median": float(np.median(elapsed)), "cells_per_step": cells}})
    else:
        vals = []
        desc = kind = ""
        t0 = time.perf_counter()
        for it in range(args.warmup + args.steps):
            g, kind, desc, _ = harness_scan(w, 6.0, cores)
            if it >= args.warmup:
                vals.append(g)
        v = float(np.median(vals))
        line.update({"value": v, "ms_per_step": (time.perf_counter() - t0) * 1e3 / max(1, args.warmup + args.steps),
                     "cpu_baseline": {"value": v, "unit": "GCUPS", "cores": cores, "kind": kind, "sample": desc}})
    line["e2e"] = {"value": line["value"], "unit": "GCUPS", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    print(json.dumps(line))
    return 0


# ------------------------------------------------------------------------------------------------
def main():
    args = parse_args()
    cfg = dict(CONFIGS[args.config])
    nseq = args.nseq or cfg["nseq"]
    qlen = args.qlen or cfg["qlen"]
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    cores = os.cpu_count() or 1

    if args.impl == "reference":
        return run_reference(args, cfg, nseq, qlen) if rank == 0 else 0

    import torch
    import torch.distributed as dist
    from swipe_b200 import (Database, Scoring, HostBuffer, topk_merge, hits_merge, set_cache_limit, alu_peak)
    from swipe_b200.shard import shard_cuts, HitExchange

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device - the scan has no CPU path")
    torch.cuda.set_device(local_rank)
    if world > 1:
        bind_to_gpu_numa_node(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        cpu_group = dist.new_group(backend="gloo")       # host-side waits that must not occupy the GPUs

    if cfg["kind"] == "nt":
        import psutil
        avail = psutil.virtual_memory().available
        if avail < nseq * 200 * 3.2:                    # codes + pinned copy + generator scratch
            nseq = int(avail // (200 * 3.2))
    # the e2e loop re-opens the shard every step: let the library keep its device buffers (shard, layout,
    # pass-boundary scratch) between handles instead of returning gigabytes to the driver at every close
    set_cache_limit(96 << 30)
    w = Workload(cfg, nseq, qlen, args.batch if cfg["kind"] == "protein" and world == 1 else 1)
    config = base_config(args, cfg, w.nseq, qlen, world)
    batched = len(w.queries) > 1 and cfg["kind"] == "protein"
    if batched:
        config["batch"] = "%d distinct %d-aa queries per step through swb_search_hits_batch (one shared scan where they fit)" % (
            len(w.queries), qlen)
    sc = Scoring(w.matrix, w.gap_open, w.gap_extend)
    nq = len(w.queries)

    # this rank's shard of the one database
    lo, hi = shard_cuts(w.offsets, world)[rank]
    sh_off = (w.offsets[lo: hi + 1] - w.offsets[lo]).astype(np.int64)
    sh_nseq = hi - lo
    sh_res = int(sh_off[-1])
    pin_res = HostBuffer(max(sh_res, 1))
    pin_res.u8[:sh_res] = w.residues[w.offsets[lo]: w.offsets[hi]]
    pin_off = HostBuffer(8 * (sh_nseq + 1))
    pin_off.view(np.int64)[:] = sh_off
    if cfg["kind"] == "nt" and world == 1:
        w.residues = None                               # the pinned copy is the database from here on

    stream = torch.cuda.Stream()
    shape = [int(x) for x in args.shape.split(",")] if args.shape else None
    exchange = HitExchange(TOPK, nq, device=torch.device("cuda", local_rank) if world > 1 else None, stream=stream)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    split = {"search": 0.0, "gather_and_merge": 0.0}

    def step(db, record=None):
        """One pass of the hot path over this rank's shard: scan(s) + device sink, then the exchange of
        the per-rank hit lists and the merge on rank 0.  Returns rank 0's merged (seqnos, scores)."""
        t0 = time.perf_counter()
        lists = []
        if batched:
            res = db.search_hits_batch(w.queries, sc, TOPK, 1, 2 ** 62, seqno_base=lo)
            for s, (seq, scv, tot, obv) in enumerate(res):
                lists.append((seq * nq + s, scv))
            if record is not None:
                for c in db.last_batch_counters:
                    record["launches"] += c["kernel_launches"]
                    record["scan_ms"].append(c["scan_ms"])
                    record["requeue_ms"].append(c["requeue_ms"])
                    record["counters"] = c
        for s, q in enumerate(w.queries if not batched else []):
            seq, scv, tot, obv = db.search_hits(q, sc, TOPK, 1, 2 ** 62, seqno_base=lo)
            lists.append((seq * nq + s, scv))            # strand in the low bit keeps (seqno, strand) unique
            if record is not None:
                c = db.last_counters
                record["launches"] += c["kernel_launches"]
                record["scan_ms"].append(c["scan_ms"])
                record["requeue_ms"].append(c["requeue_ms"])
                record["counters"] = c
        t1 = time.perf_counter()
        t2 = t1
        merged = exchange(lists)                         # all_gather of K pairs per rank + swb_hits_merge on rank 0
        t3 = time.perf_counter()
        split["search"] += t1 - t0
        split["gather_and_merge"] += t3 - t2
        return merged

    # ---- the 1-GPU answer the sharded run must reproduce (rank 0, outside the timed region) ----------
    single = None
    dense_checksum = None
    if rank == 0 and world > 1:
        with Database(w.residues, w.offsets, device=local_rank) as full:
            ls = []
            for s, q in enumerate(w.queries):
                seq, scv, _, _ = full.search_hits(q, sc, TOPK, 1, 2 ** 62)
                ls.append((seq * nq + s, scv))
            single = hits_merge(ls, TOPK)

    # ---- resident-database timing ------------------------------------------------------------------
    db = Database(pin_res.u8[:sh_res], pin_off.view(np.int64), device=local_rank, stream=stream.cuda_stream)
    if shape:
        db.set_shape(*shape)
    for _ in range(args.warmup):
        step(db)
    rec = {"launches": 0, "scan_ms": [], "requeue_ms": [], "counters": None}
    for k in split:
        split[k] = 0.0
    sampler = ClockSampler(local_rank)
    barrier()
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    merged = None
    for _ in range(args.steps):
        merged = step(db, rec)
    e1.record(stream)
    barrier()
    clocks = sampler.stop()
    my_ms = e0.elapsed_time(e1)
    res_split = {k: v * 1e3 / args.steps for k, v in split.items()}

    # ---- parity of the sink at full size: dense scores of the last query + host sink (rank-local) ----
    pin_scores = HostBuffer(8 * max(sh_nseq, 1))
    dense = pin_scores.view(np.int64)[:sh_nseq]
    local_lists = []
    for s, q in enumerate(w.queries):
        db.search(q, sc, out=dense)
        seq, scv, _, _ = topk_merge([dense], [lo], TOPK, min_score=1)
        local_lists.append((seq * nq + s, scv))
        hseq, hsc, _, _ = db.search_hits(q, sc, TOPK, 1, 2 ** 62, seqno_base=lo)
        if not (np.array_equal(hseq, seq) and np.array_equal(hsc, scv)):
            raise SystemExit("bench.py: device sink differs from the host sink on rank %d" % rank)
        if s == 0:
            dense_checksum = int(dense.sum())
    topk_identical = None
    if rank == 0:
        if world > 1:
            topk_identical = bool(np.array_equal(merged[0], single[0]) and np.array_equal(merged[1], single[1]))
        else:
            ref_list = hits_merge(local_lists, TOPK)
            topk_identical = bool(np.array_equal(merged[0], ref_list[0]) and np.array_equal(merged[1], ref_list[1]))
    # the alignment phase's device part (search16s's contract, swipe.cc:381-393): exact score + end cell
    # of the K best hits of the first query, as align_chunk asks for them
    end_cell = None
    if rank == 0 and local_lists and len(local_lists[0][0]):
        top = (local_lists[0][0] // nq) - lo
        db.search_end(w.queries[0], sc, top)             # warm-up (scratch allocation)
        t0 = time.perf_counter()
        es, ep, eq = db.search_end(w.queries[0], sc, top)
        end_cell = {"what": "swb_search_end over the %d best hits of this shard (search16s: score, first column "
                            "reaching it, smallest row in it)" % top.size,
                    "subjects": int(top.size), "ms": (time.perf_counter() - t0) * 1e3,
                    "scores_equal_scan": bool(np.array_equal(es, local_lists[0][1]))}
    dense_first_strand = None
    if world == 1 and not args.no_cpu_baseline:
        db.search(w.queries[0], sc, out=dense)
        dense_first_strand = dense.copy() if nq > 1 else dense
    db.close()

    # ---- end to end from host buffers ------------------------------------------------------------------
    e2e_ms = None
    e2e_split = None
    h2d = sh_res + 8 * (sh_nseq + 1) + nq * (qlen + 8 * 1024 + 2256 + 2 * 1024)
    d2h = nq * (16 * TOPK + 64)
    if not args.no_e2e:
        host = {"open_enqueue": 0.0, "search_gather_merge": 0.0, "close": 0.0}

        def e2e_step():
            t0 = time.perf_counter()
            d = Database(pin_res.u8[:sh_res], pin_off.view(np.int64), device=local_rank,
                         stream=stream.cuda_stream, wait=False)   # upload / re-layout / scan overlap
            if shape:
                d.set_shape(*shape)
            t1 = time.perf_counter()
            step(d)
            t2 = time.perf_counter()
            d.close()
            t3 = time.perf_counter()
            host["open_enqueue"] += t1 - t0
            host["search_gather_merge"] += t2 - t1
            host["close"] += t3 - t2
        for _ in range(max(1, min(args.warmup, 2))):
            e2e_step()
        for k in host:
            host[k] = 0.0
        barrier()
        f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        f0.record(stream)
        for _ in range(args.steps):
            e2e_step()
        f1.record(stream)
        barrier()
        e2e_ms = f0.elapsed_time(f1)
        e2e_split = {k: round(v * 1e3 / args.steps, 3) for k, v in host.items()}

    # ---- weak scaling, secondary: every GPU scans a full copy of the database -------------------------
    weak_ms = None
    weak_steps = max(1, min(args.steps, 5))
    if world > 1 and not args.no_weak and w.residues is not None:
        with Database(w.residues, w.offsets, device=local_rank, stream=stream.cuda_stream) as full:
            full.search_hits(w.queries[0], sc, TOPK, 1, 2 ** 62)
            barrier()
            g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            g0.record(stream)
            for _ in range(weak_steps):
                for q in w.queries:
                    full.search_hits(q, sc, TOPK, 1, 2 ** 62)
            g1.record(stream)
            barrier()
            weak_ms = g0.elapsed_time(g1)

    # ---- the product's own multi-GPU path, secondary: ONE process, one host thread + handle per GPU ----
    # (what swipe-b200 -a N does, swipe_main.cpp); rank 0 drives all N devices while the other ranks wait.
    product = None
    if world > 1:
        barrier()
        # (the other ranks wait on the gloo group below: an NCCL barrier would spin a kernel on their GPUs,
        # which rank 0 is about to use from its own context)
        if rank == 0 and not args.no_product_path and torch.cuda.device_count() >= world:
            try:
                cuts = shard_cuts(w.offsets, world)
                dbs = []
                for r, (a, b) in enumerate(cuts):
                    dbs.append((a, Database(w.residues[w.offsets[a]: w.offsets[b]],
                                            (w.offsets[a: b + 1] - w.offsets[a]).astype(np.int64), device=r)))
                out = [None] * world

                def worker(r):
                    a, d = dbs[r]
                    ls = []
                    for s, q in enumerate(w.queries):
                        seq, scv, _, _ = d.search_hits(q, sc, TOPK, 1, 2 ** 62, seqno_base=a)
                        ls.append((seq * nq + s, scv))
                    out[r] = ls

                def product_step():
                    th = [threading.Thread(target=worker, args=(r,)) for r in range(world)]
                    for t in th:
                        t.start()
                    for t in th:
                        t.join()
                    return hits_merge([l for ls in out for l in ls], TOPK)
                for _ in range(2):
                    product_step()
                t0 = time.perf_counter()
                for _ in range(args.steps):
                    pm = product_step()
                dt = time.perf_counter() - t0
                for _, d in dbs:
                    d.close()
                torch.cuda.set_device(local_rank)            # the handles switched this thread's device
                product = {"value": w.cells * args.steps / dt * 1e-9, "unit": "GCUPS", "ms_per_step": dt * 1e3 / args.steps,
                           "timing": "host wall clock around N threads (each search ends in a stream synchronize)",
                           "topk_identical": bool(np.array_equal(pm[0], single[0]) and np.array_equal(pm[1], single[1])),
                           "what": "one process, one host thread + swb_db handle per GPU, swb_search_hits + swb_hits_merge"}
            except Exception as e:               # e.g. GPUs in exclusive-process mode: the other ranks own them
                product = {"unavailable": "%s: %s" % (type(e).__name__, e)}
                torch.cuda.set_device(local_rank)
        dist.barrier(group=cpu_group)
        barrier()

    # ---- reduce over ranks --------------------------------------------------------------------------------
    if world > 1:
        t = torch.tensor([my_ms, e2e_ms or 0.0, weak_ms or 0.0, float(rec["launches"]), float(np.mean(rec["scan_ms"])),
                          res_split["search"], float(sh_res)], device="cuda", dtype=torch.float64)
        tmax = t.clone()
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        tsum = t.clone()
        dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
        ms_all, e2e_all, weak_all = float(tmax[0]), float(tmax[1]), float(tmax[2])
        launches_all = int(tsum[3])
        scan_max, search_max = float(tmax[4]), float(tmax[5])
        h2d_all, d2h_all = int(tsum[6]) + world * (h2d - sh_res), d2h * world
    else:
        ms_all, e2e_all, weak_all, launches_all = my_ms, e2e_ms, None, rec["launches"]
        scan_max, search_max = float(np.mean(rec["scan_ms"])), res_split["search"]
        h2d_all, d2h_all = h2d, d2h

    if rank == 0:
        value = w.cells * args.steps / (ms_all * 1e-3) * 1e-9
        scan_avg = float(np.mean(rec["scan_ms"]))            # per scan launch (one strand), this rank
        my_cells_per_scan = float(sh_res) * qlen
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        # the binding roofline: issue rate of the DPX instructions, measured now on this device
        dpx_rate, ub_mhz = alu_peak(local_rank)
        sm_mhz = clocks["sm_mhz"] or ub_mhz or 1965.0
        sms = torch.cuda.get_device_properties(local_rank).multi_processor_count
        dpx_per_cell_pair = 3.5                              # 1 VIMNMX3 + 2 VIADDMNMX + 1/2 VIMNMX3 (running maximum)
        peak_cells_clk_sm = dpx_rate * 64.0 / dpx_per_cell_pair
        peak_gcups = peak_cells_clk_sm * sms * sm_mhz * 1e6 * 1e-9
        kernel_gcups = my_cells_per_scan / (scan_avg * 1e-3) * 1e-9
        cells_clk_sm = my_cells_per_scan / (scan_avg * 1e-3) / (sms * sm_mhz * 1e6)
        hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
        alg_bytes = sh_res + 8 * sh_nseq                     # 1 B per residue + 8 B per subject
        hbm_achieved = alg_bytes / (scan_avg * 1e-3) * 1e-9
        traffic = None
        traffic_src = None
        try:
            cap = json.load(open(os.path.join(ROOT, "profiles", "r2_ncu_traffic.json")))
            for c in cap["captures"]:
                if c["config"] == args.config and c["subjects"] == sh_nseq and c["residues"] == sh_res:
                    traffic = c["dram_bytes_read"] + c["dram_bytes_write"]
                    traffic_src = c["source"]
        except Exception:
            pass
        line = {
            "metric": "GCUPS", "value": value, "unit": "GCUPS", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_all / args.steps, "higher_is_better": True,
            "scaling": "strong" if world > 1 else "weak", "vs_baseline": None, "dtype": "int16",
            "data": "synthetic", "config": config, "clocks": clocks,
            "gpu_launches": launches_all,
            "roofline": {"bound": "int-alu-issue", "achieved": kernel_gcups, "peak": peak_gcups, "unit": "GCUPS",
                         "frac": kernel_gcups / peak_gcups,
                         "kernel": "%s<G=%d,R=%d,hybrid fp16-pattern/DPX lanes>, %d pass(es)" % (
                             "swb_scan2_kernel" if rec["counters"]["scan_geometry"] == 2 else "swb_scan_kernel",
                             rec["counters"]["scan_G"], rec["counters"]["scan_R"], rec["counters"]["scan_passes"]),
                         "kernel_ms": scan_avg,
                         "cells_per_launch": my_cells_per_scan, "cells_per_clk_per_sm": cells_clk_sm,
                         "peak_cells_per_clk_per_sm": peak_cells_clk_sm,
                         "dpx_warp_instr_per_clk_per_sm_measured": dpx_rate, "dpx_per_cell_pair": dpx_per_cell_pair,
                         "lane_ops_per_cell": 9, "sm_mhz": sm_mhz, "ubench_sm_mhz": ub_mhz, "sms": sms,
                         "peak_source": "swb_alu_peak in this run (issue rate of VIADDMNMX/VIMNMX3.S16x2) x 64 cells / 3.5 "
                                        "DPX per cell pair x SMs x median SM clock of the timed region",
                         "nominal_8bit_x4_gcups": 64 * 4 / 9.0 * sms * sm_mhz * 1e-3,
                         "frac_of_nominal_8bit_x4": kernel_gcups / (64 * 4 / 9.0 * sms * sm_mhz * 1e-3),
                         "traffic": traffic, "traffic_source": traffic_src,
                         "hbm": {"bound": "hbm", "achieved": hbm_achieved, "peak": hbm_peak, "unit": "GB/s",
                                 "frac": hbm_achieved / hbm_peak, "algorithmic_bytes": alg_bytes,
                                 "peak_source": "MEASURED_PEAKS.json" if peaks else "fallback",
                                 "note": "not the binding bound (SURVEY 8d): 1 byte per residue per scan"}},
            "split_ms": {"scan_kernel_max_over_ranks": scan_max * nq, "search_call_max_over_ranks": search_max,
                         "sink_and_d2h": max(0.0, search_max - scan_max * nq),
                         "gather_and_merge": res_split["gather_and_merge"], "step": ms_all / args.steps},
            "counters": {k: rec["counters"][k] for k in ("ref_width7", "ref_width16", "ref_width63", "gpu_narrow",
                                                         "gpu_requeued", "gpu_middle", "kernel_launches")},
            "requeue_ms": float(np.mean(rec["requeue_ms"])),
            "topk_identical": topk_identical, "checksum": dense_checksum,
            "top_hit": [int(merged[0][0]) // nq, int(merged[1][0])] if len(merged[0]) else None,
        }
        sp = line["split_ms"]
        parts = {"scan": sp["scan_kernel_max_over_ranks"], "sink_and_d2h": sp["sink_and_d2h"],
                 "gather_and_merge": sp["gather_and_merge"]}
        line["split_ms"]["limiter"] = max(parts, key=parts.get)
        if e2e_all:
            line["e2e"] = {"value": w.cells * args.steps / (e2e_all * 1e-3) * 1e-9, "unit": "GCUPS",
                           "h2d_bytes_per_step": int(h2d_all), "d2h_bytes_per_step": int(d2h_all),
                           "ms_per_step": e2e_all / args.steps, "host_ms": e2e_split}
        if weak_all:
            line["weak"] = {"value": w.cells * world * weak_steps / (weak_all * 1e-3) * 1e-9, "unit": "GCUPS",
                            "ms_per_step": weak_all / weak_steps, "steps": weak_steps,
                            "what": "every GPU scans its own full copy of the database (per-GPU work fixed)"}
        if end_cell:
            line["alignment_phase"] = end_cell
        if product:
            line["product_path"] = product
        if world == 1 and not args.no_cpu_baseline:
            budget = 12.0
            g, kind, desc, outs = harness_scan(w if w.residues is not None else _host_view(w, pin_res, sh_res),
                                               budget, cores)
            line["cpu_baseline"] = {"value": g, "unit": "GCUPS", "cores": cores, "kind": kind, "sample": desc}
            # the CPU leg scored the same subjects: a bit-exact parity check at the bench's own size
            k = int(outs[0].size)
            line["cpu_baseline"]["scores_compared"] = k * 1
            line["cpu_baseline"]["scores_equal"] = bool(np.array_equal(outs[0], dense_first_strand[:k]))
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def _host_view(w, pin_res, sh_res):
    """nt50m drops the pageable copy of the database; the CPU leg reads the pinned one."""
    w.residues = pin_res.u8[:sh_res]
    return w


if __name__ == "__main__":
    sys.exit(main())
