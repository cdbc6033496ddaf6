#!/usr/bin/env python3
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel count, total time, share."""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
H = rows[hdr]
ki, vi = H.index("Kernel Name"), H.index("Metric Value")
agg = collections.defaultdict(lambda: [0, 0.0])
for r in rows[hdr + 1:]:
    if len(r) <= vi:
        continue
    name = r[ki].split("(")[0]
    agg[name][0] += 1
    agg[name][1] += float(r[vi].replace(",", ""))
tot = sum(v[1] for v in agg.values())
print("%-70s %6s %12s %7s" % ("kernel", "n", "total_ms", "share"))
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("%-70s %6d %12.3f %6.1f%%" % (k[:70], v[0], v[1] / 1e6, 100 * v[1] / tot))
print("%-70s %6d %12.3f" % ("TOTAL", sum(v[0] for v in agg.values()), tot / 1e6))
