"""Copies the call-G outputs from gpurun_out/ into profiles/ (bench lines, ncu summaries, launch list) and
prints the table DESIGN.md section 7 quotes."""
import collections
import csv
import json
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = os.path.join(ROOT, "gpurun_out")
P = os.path.join(ROOT, "profiles")


def line(path):
    ls = [l for l in open(path) if l.startswith("{")]
    return json.loads(ls[-1]) if ls else None


rows = []
for cfg in ("", "_qlen100", "_qlen1000", "_qlen5000", "_nt50m", "_qlen100_batch4"):
    src = os.path.join(G, "r2g_bench%s.json" % cfg)
    if not os.path.exists(src):
        continue
    d = line(src)
    if d is None:
        continue
    json.dump(d, open(os.path.join(P, "r2_bench%s_final.json" % cfg), "w"))
    ref = None
    rsrc = os.path.join(G, "r2g_bench_reference%s.json" % cfg)
    if os.path.exists(rsrc):
        ref = line(rsrc)
        if ref:
            json.dump(ref, open(os.path.join(P, "r2_bench_reference%s_final.json" % cfg), "w"))
    cb = d.get("cpu_baseline", {})
    rows.append((d["config"]["name"] + ("+batch" if "batch" in d["config"] else ""), d["roofline"]["kernel"],
                 d["roofline"]["achieved"], d["value"], d["roofline"]["frac"], d["e2e"]["value"],
                 ref["value"] if ref else None, ref["cpu_baseline"].get("one_thread_gcups") if ref else None,
                 cb.get("value"), cb.get("scores_equal"), cb.get("scores_compared"), d.get("topk_identical"),
                 d.get("alignment_phase", {}).get("ms")))
print("| config | scan kernel | kernel GCUPS | value GCUPS | frac | e2e GCUPS | reference program GCUPS (1 thread) | "
      "reference kernels GCUPS | scores_equal (subjects) | topk | end cells ms |")
print("|---|---|---|---|---|---|---|---|---|---|---|")
for r in rows:
    print("| %s | %s | %.0f | %.0f | %.3f | %.0f | %s (%s) | %s | %s (%s) | %s | %s |" % (
        r[0], r[1], r[2], r[3], r[4], r[5], "%.0f" % r[6] if r[6] else "-", "%.1f" % r[7] if r[7] else "-",
        "%.0f" % r[8] if r[8] else "-", r[9], r[10], r[11], "%.2f" % r[12] if r[12] else "-"))

for name in ("375", "qlen100", "qlen1000", "nt"):
    raw = os.path.join(G, "r2g_ncu_%s_raw.csv" % name)
    src = os.path.join(G, "r2g_ncu_%s_source.csv" % name)
    if os.path.exists(raw) and os.path.getsize(raw) > 1000:
        out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_summary.py"), raw],
                             capture_output=True, text=True).stdout
        open(os.path.join(P, "r2_ncu_scan_final_%s_raw_summary.txt" % name), "w").write(out)
    if os.path.exists(src) and os.path.getsize(src) > 1000:
        out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_source_summary.py"), src, "15"],
                             capture_output=True, text=True).stdout
        open(os.path.join(P, "r2_ncu_scan_final_%s_source_summary.txt" % name), "w").write(out)

lc = os.path.join(G, "r2g_launches.csv")
if os.path.exists(lc):
    shutil.copy(lc, os.path.join(P, "r2_launches_final.csv"))
    rows = list(csv.reader(open(lc)))
    for i, r in enumerate(rows):
        if "Kernel Name" in r:
            hdr, start = r, i + 2
            break
    ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
    tot, cnt = collections.Counter(), collections.Counter()
    for r in rows[start:]:
        if len(r) <= vi:
            continue
        try:
            v = float(r[vi].replace(",", ""))
        except ValueError:
            continue
        k = r[ki].split("(")[0][:70]
        tot[k] += v
        cnt[k] += 1
    s = sum(tot.values())
    out = ["# ncu --metrics gpu__time_duration.sum --clock-control none of `python bench.py --steps 2 --warmup 1 "
           "--no-cpu-baseline --no-e2e` (round 2, final)", "# per kernel: launches, total ns, share of device time"]
    for k, v in tot.most_common():
        out.append("%-72s %5d %14.0f %6.2f%%" % (k, cnt[k], v, 100 * v / s))
    open(os.path.join(P, "r2_launches_final_summary.txt"), "w").write("\n".join(out) + "\n")
    print("\n".join(out[:8]))
if os.path.exists(os.path.join(G, "r2g_pytest_gpu.log")):
    shutil.copy(os.path.join(G, "r2g_pytest_gpu.log"), os.path.join(P, "r2_pytest_gpu_final.txt"))
