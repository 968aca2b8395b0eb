"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list into a per-kernel table of the last complete
training step (between two adam_kernel launches):  python scripts/launch_summary.py profiles/r2_launches_tf32.csv"""
import collections
import csv
import re
import sys

rows, hdr = [], None
for r in csv.reader(open(sys.argv[1])):
    if hdr is None:
        if "Kernel Name" in r:
            hdr = r
        continue
    rows.append(r)
ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
idx = [i for i, r in enumerate(rows) if "adam_kernel" in r[ki]]
a, b = idx[-2], idx[-1]
agg = collections.OrderedDict()
tot = 0.0
for r in rows[a + 1:b + 1]:
    name = re.sub(r"\(.*", "", r[ki]).replace("void ", "").replace("pu::", "")
    t = float(r[vi].replace(",", ""))
    t = t / 1000.0 if r[ui].startswith("ns") or r[ui] == "nsecond" else t
    agg.setdefault(name, [0, 0.0])
    agg[name][0] += 1
    agg[name][1] += t
    tot += t
print("last complete step: %d launches, %.1f us (serialised, cold caches)\n" % (b - a, tot))
print("| kernel | launches | total us | share | avg us |\n|---|---:|---:|---:|---:|")
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("| `%s` | %d | %.1f | %.1f %% | %.1f |" % (k[:70], n, t, 100 * t / tot, t / n))
