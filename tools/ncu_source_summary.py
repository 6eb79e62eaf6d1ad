"""Aggregates an `ncu --page source --csv` dump: executed warp instructions and stall samples per opcode,
and the hottest instructions."""
import csv, sys, collections, re
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
ops = collections.Counter(); samp = collections.Counter(); total = 0; tot_s = 0
hot = []
for r in rows[2:]:
    if len(r) < len(hdr): continue
    src = r[ix["Source"]].strip()
    m = re.match(r"(@!?U?P\d+\s+)?([A-Z0-9_.]+)", src)
    op = m.group(2) if m else src
    n = int(r[ix["Instructions Executed"]] or 0); s = int(r[ix["# Samples"]] or 0)
    ops[op] += n; samp[op] += s; total += n; tot_s += s
    hot.append((s, n, src[:70]))
print("total warp instructions", total, "samples", tot_s)
for op, n in ops.most_common(28):
    print("  %-28s %14d  %5.1f%%   samples %5.1f%%" % (op, n, 100.0 * n / total, 100.0 * samp[op] / max(tot_s, 1)))
if len(sys.argv) > 2:
    print("hottest by samples:")
    for s, n, src in sorted(hot, reverse=True)[:int(sys.argv[2])]:
        print("  %7d %12d  %s" % (s, n, src))
