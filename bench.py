#!/usr/bin/env python
"""bench.py -- GCUPS of the score-only Smith-Waterman database scan on N B200s.

Workload (BASELINE.json configs[1]): a 375-aa background-frequency query, BLOSUM62, gap 11/1,
against a synthetic 5,000,000-sequence protein database (log-normal lengths, ~1.75 G residues,
0.1 % planted homologs) PER GPU.  With N GPUs every rank holds its own 5 M-sequence shard (weak
scaling: a 5 M x N database sharded by sequence), there is no data-path collective, and rank 0
merges the per-shard top-K lists with the reference's hits_enter rule.

  value : GCUPS = 1e-9 * residues * qlen / s (swipe.cc:1744-1775) with the shard resident in HBM;
          K swb_search calls timed with CUDA events on the handle's stream, max over ranks.
  e2e   : the same metric through the C ABI from HOST buffers: every step opens the shard
          (pinned host -> device copy + device re-layout), searches, reads the scores back and
          takes the local top-K.
  --impl reference : the UNMODIFIED reference kernels (oracle/_ref, built from /root/reference)
          on all host cores over a bounded sample of the same workload.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

QLEN = 375
NSEQ = 5_000_000
GAP_OPEN, GAP_EXTEND = 11, 1
TOPK = 100                                  # the reference keeps max(-v, -b) = 250 by default; any K works


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--nseq", type=int, default=NSEQ, help="subjects per GPU shard")
    ap.add_argument("--qlen", type=int, default=QLEN)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--shape", default="", help="G,R,lane_mode test hook")
    return ap.parse_args()


def make_workload(nseq, qlen, rank):
    from swipe_b200 import synth
    q = synth.protein_query(qlen)
    residues, offsets = synth.protein_db(nseq, query=q, seed=20261018 + 1000 * rank)
    return q, residues, offsets


class ClockSampler(threading.Thread):
    """Samples SM clocks and throttle reasons with nvidia-smi while the timed region runs."""

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
        """NVML directly (nvidia-ml-py): a sample every 20 ms instead of one per nvidia-smi start-up."""
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


def run_reference_cpu(q, residues, offsets, budget_s, threads):
    """Times the reference's own search7/search16/fullsw cascade (oracle/_ref) on a bounded
    prefix of the shard.  Returns (gcups, description)."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib
    kind = "reference" if oracle_lib.ref_available() else "port"
    nseq = offsets.size - 1

    if kind == "reference":
        ref = oracle_lib.Ref()
        ref.matrix_init("BLOSUM62")

        def scan(n):
            return ref.scan(residues[: offsets[n]], offsets[: n + 1], q, GAP_OPEN, GAP_EXTEND,
                            threads=threads, chunk=1024, ssse3=1)
    else:
        from swipe_b200 import scoring
        orc = oracle_lib.Oracle()
        m = scoring.blosum62()

        def scan(n):
            return orc.scan(residues[: offsets[n]], offsets[: n + 1], q, m, GAP_OPEN, GAP_EXTEND,
                            threads=threads)

    n0 = min(nseq, 20000)
    scan(n0)                                   # cold: page faults, thread start-up
    t0 = time.perf_counter()
    scan(n0)
    t1 = time.perf_counter()
    rate = float(offsets[n0]) * q.size / max(t1 - t0, 1e-6)          # cells / s, cold
    n = int(min(nseq, max(n0, budget_s * rate / (q.size * (offsets[n0] / n0)))))
    t0 = time.perf_counter()
    out = scan(n)
    t1 = time.perf_counter()
    cells = float(offsets[n]) * q.size
    run_reference_cpu.last_scores = np.asarray(out[0][:n])      # kept for the bench's full-size parity check
    return cells / (t1 - t0) * 1e-9, kind, "first %d subjects (%d residues) of the shard, %.1f s" % (
        n, int(offsets[n]), t1 - t0)


def bind_to_gpu_numa_node(index):
    """Multi-rank runs: keep this rank's host threads and its pinned buffers on the NUMA node the GPU
    hangs off, so that eight simultaneous uploads do not cross sockets.  Best effort."""
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


def main():
    args = parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    cores = os.cpu_count() or 1
    config = {"workload": "375-aa query vs 5M-seq synthetic protein DB per GPU, BLOSUM62 11/1 "
                          "(BASELINE configs[1])",
              "qlen": args.qlen, "nseq_per_gpu": args.nseq, "gap_open": GAP_OPEN,
              "gap_extend": GAP_EXTEND, "matrix": "BLOSUM62", "sharding": "by sequence, no collective",
              "l2": "inputs (1.8 GB/shard) larger than L2",
              "lanes": "two int16 lanes per 32-bit register: DPX s16x2 max / add-max, adds as fp16x2 on integer "
                       "bit patterns (exact to 2047), re-queue to plain int16 lanes and then 32/64-bit cells"}

    # ------------------------------------------------------------------ reference arm
    if args.impl == "reference":
        if rank != 0:
            return 0
        q, residues, offsets = make_workload(min(args.nseq, 1_000_000), args.qlen, 0)
        vals = []
        desc = kind = ""
        budget = 8.0
        for it in range(args.warmup + args.steps):
            g, kind, desc = run_reference_cpu(q, residues, offsets, budget, cores)
            if it >= args.warmup:
                vals.append(g)
            if it == 0 and args.warmup + args.steps > 12:
                budget = 4.0
        v = float(np.mean(vals)) if vals else 0.0
        # one "step" of the metric's workload (a 5 M-sequence shard) at the measured rate
        ms = (1.75e9 * args.qlen / (v * 1e9)) * 1e3 if v > 0 else 0.0
        line = {"metric": "GCUPS", "value": v, "unit": "GCUPS", "impl": "reference",
                "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "int8",
                "data": "synthetic", "config": config,
                "cpu_baseline": {"value": v, "unit": "GCUPS", "cores": cores, "kind": kind,
                                 "sample": desc},
                "e2e": {"value": v, "unit": "GCUPS", "h2d_bytes_per_step": 0,
                        "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return 0

    # ------------------------------------------------------------------ B200 arm
    import torch
    import torch.distributed as dist
    from swipe_b200 import Database, Scoring, HostBuffer, scoring, topk_merge

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device - the scan has no CPU path")
    torch.cuda.set_device(local_rank)
    if world > 1:
        bind_to_gpu_numa_node(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    q, residues, offsets = make_workload(args.nseq, args.qlen, rank)
    nseq = offsets.size - 1
    total_res = int(offsets[-1])
    cells = float(total_res) * args.qlen
    sc = Scoring(scoring.blosum62(), GAP_OPEN, GAP_EXTEND)

    # pinned host copies: what a caller that mmaps the .psq would hand over
    pin_res = HostBuffer(total_res)
    pin_res.u8[:] = residues
    pin_off = HostBuffer(8 * (nseq + 1))
    pin_off.view(np.int64)[:] = offsets
    pin_scores = HostBuffer(8 * nseq)
    scores = pin_scores.view(np.int64)

    stream = torch.cuda.Stream()
    shape = [int(x) for x in args.shape.split(",")] if args.shape else None

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- resident-database timing ------------------------------------------------------------
    db = Database(pin_res.u8, pin_off.view(np.int64), device=local_rank, stream=stream.cuda_stream)
    if shape:
        db.set_shape(*shape)
    for _ in range(args.warmup):
        db.search(q, sc, out=scores)
    sampler = ClockSampler(local_rank)
    barrier()
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    launches = 0
    scan_ms = []
    requeue_ms = []
    e0.record(stream)
    for _ in range(args.steps):
        db.search(q, sc, out=scores)
        c = db.last_counters
        launches += c["kernel_launches"]
        scan_ms.append(c["scan_ms"])
        requeue_ms.append(c["requeue_ms"])
    e1.record(stream)
    barrier()
    clocks = sampler.stop()
    my_ms = e0.elapsed_time(e1)
    counters = db.last_counters
    checksum = int(scores.sum())
    top_seq, top_sc, _, _ = topk_merge([scores], [rank * nseq], TOPK, min_score=1)
    db.close()

    # ---- end to end from host buffers ---------------------------------------------------------
    e2e_ms = None
    h2d = total_res + 8 * (nseq + 1) + args.qlen + 8 * 1024 + 33 * 32 * 2 + 2 * 1024
    d2h = 8 * nseq + 64
    if not args.no_e2e:
        split = np.zeros(4)

        def e2e_step():
            t0 = time.perf_counter()
            d = Database(pin_res.u8, pin_off.view(np.int64), device=local_rank,
                         stream=stream.cuda_stream, wait=False)   # upload / re-layout / scan overlap
            if shape:
                d.set_shape(*shape)
            t1 = time.perf_counter()
            d.search(q, sc, out=scores)
            t2 = time.perf_counter()
            k = d.last_counters["kernel_launches"]
            d.close()
            t3 = time.perf_counter()
            topk_merge([scores], [rank * nseq], TOPK, min_score=1)
            t4 = time.perf_counter()
            split[:] += (t1 - t0, t2 - t1, t3 - t2, t4 - t3)
            return k
        for _ in range(max(1, min(args.warmup, 2))):
            e2e_step()
        barrier()
        f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        split[:] = 0
        f0.record(stream)
        for _ in range(args.steps):
            e2e_step()
        f1.record(stream)
        barrier()
        e2e_ms = f0.elapsed_time(f1)

    # ---- reduce over ranks ----------------------------------------------------------------------
    if world > 1:
        t = torch.tensor([my_ms, e2e_ms or 0.0, cells, float(launches)], device="cuda",
                         dtype=torch.float64)
        tmax = t.clone()
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        tsum = t.clone()
        dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
        ms_all, e2e_all = float(tmax[0]), float(tmax[1])
        cells_all, launches_all = float(tsum[2]), int(tsum[3])
        # host-side top-K merge of the shards (hits_enter rule): gather K (seqno, score) per rank
        mine = torch.full((TOPK, 2), -1, dtype=torch.int64, device="cuda")
        mine[: top_seq.size, 0] = torch.from_numpy(top_seq).cuda()
        mine[: top_seq.size, 1] = torch.from_numpy(top_sc).cuda()
        gathered = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(gathered, mine)
        if rank == 0:
            allhits = torch.cat(gathered).cpu().numpy()
            allhits = allhits[allhits[:, 0] >= 0]
            order = np.lexsort((-allhits[:, 0], -allhits[:, 1]))[:TOPK]
            top_seq, top_sc = allhits[order, 0], allhits[order, 1]
    else:
        ms_all, e2e_all, cells_all, launches_all = my_ms, e2e_ms, cells, launches

    if rank == 0:
        value = cells_all * args.steps / (ms_all * 1e-3) * 1e-9
        scan_avg = float(np.mean(scan_ms))
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
        alg_bytes = total_res + 8 * nseq                 # 1 B per residue + 8 B per subject
        achieved = alg_bytes / (scan_avg * 1e-3) * 1e-9
        # DRAM traffic of the scan kernel from the committed ncu --set full capture, scaled from the
        # captured shard to this one by algorithmic bytes (the kernel streams every block once)
        traffic = None
        try:
            cap = json.load(open(os.path.join(ROOT, "profiles", "r1_ncu_traffic.json")))
            cap_alg = cap["residues"] + 8 * cap["subjects"]
            traffic = (cap["dram_bytes_read"] + cap["dram_bytes_write"]) * (alg_bytes / cap_alg)
        except Exception:
            pass
        sm_mhz = clocks["sm_mhz"] or 1965.0
        cells_per_clk_sm = cells / (scan_avg * 1e-3) / (148 * sm_mhz * 1e6)
        line = {
            "metric": "GCUPS", "value": value, "unit": "GCUPS", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_all / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "int16",
            "data": "synthetic", "config": config, "clocks": clocks,
            "gpu_launches": launches_all,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s",
                         "frac": achieved / hbm_peak, "traffic": traffic,
                         "traffic_source": "profiles/r1_ncu_traffic.json (ncu capture at 1.5M subjects, scaled by algorithmic bytes)",
                         "algorithmic_bytes": alg_bytes,
                         "peak_source": "MEASURED_PEAKS.json" if peaks else "fallback",
                         "kernel": "swb_scan_kernel", "kernel_ms": scan_avg,
                         "note": "integer-issue bound, not HBM bound (SURVEY 8d): see alu"},
            "alu": {"kernel_gcups": cells / (scan_avg * 1e-3) * 1e-9,
                    "cells_per_clk_per_sm": cells_per_clk_sm,
                    "lane_ops_per_cell": 9,
                    "peak_cells_per_clk_per_sm_ubench_int16": 25.6,
                    "peak_cells_per_clk_per_sm_ubench_hybrid": 29.9,
                    "frac_of_ubench_hybrid": cells_per_clk_sm / 29.9,
                    # 3.5 DPX ops per cell pair at one warp instruction per 2 clk on the 16-lane ALU
                    # pipe = 7 clk per 64 cells per SM sub-partition
                    "alu_pipe_bound_cells_per_clk_per_sm": 4 * 64 / 7.0,
                    "frac_of_alu_pipe_bound": cells_per_clk_sm / (4 * 64 / 7.0),
                    "nominal_8bit_tcups": 8.27,
                    "frac_of_nominal_8bit": cells / (scan_avg * 1e-3) * 1e-12 / 8.27},
            "counters": {k: counters[k] for k in ("ref_width7", "ref_width16", "ref_width63",
                                                   "gpu_narrow", "gpu_requeued", "kernel_launches")},
            "requeue_ms": float(np.mean(requeue_ms)),
            "checksum": checksum, "top_hit": [int(top_seq[0]), int(top_sc[0])] if len(top_seq) else None,
        }
        if e2e_all:
            line["e2e"] = {"value": cells_all * args.steps / (e2e_all * 1e-3) * 1e-9, "unit": "GCUPS",
                           "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                           "ms_per_step": e2e_all / args.steps,
                           "host_ms": dict(zip(["open_enqueue", "search", "close", "topk"],
                                               [round(float(x) * 1e3 / args.steps, 2) for x in split]))}
        if world == 1 and not args.no_cpu_baseline:
            g, kind, desc = run_reference_cpu(q, residues, offsets, 12.0, cores)
            line["cpu_baseline"] = {"value": g, "unit": "GCUPS", "cores": cores, "kind": kind,
                                    "sample": desc}
            # the CPU leg scored the same subjects: a bit-exact parity check at the bench's full size
            ref_scores = getattr(run_reference_cpu, "last_scores", None)
            if ref_scores is not None:
                k = int(ref_scores.size)
                line["cpu_baseline"]["scores_compared"] = k
                line["cpu_baseline"]["scores_equal"] = bool(np.array_equal(ref_scores, scores[:k]))
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
