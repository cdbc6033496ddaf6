#!/usr/bin/env python3
"""Per CUDA source line summary of `ncu -i X.ncu-rep --page source --csv --print-source cuda,sass -k regex:K -c 1`:
stall samples and executed warp instructions per line (top N)."""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
cur_file = ""
agg = {}
for r in rows:
    if len(r) >= 2 and r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
        continue
    if len(r) > 8 and r[0].isdigit() and r[2] == "-":
        key = (cur_file, int(r[0]))
        a = agg.setdefault(key, [0, 0, r[1].strip()])
        a[0] += int(r[4] or 0); a[1] += int(r[7] or 0)
tot_s = sum(a[0] for a in agg.values()); tot_i = sum(a[1] for a in agg.values())
print("total samples %d, warp instructions %d" % (tot_s, tot_i))
for (f, ln), a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print("%-16s %5d  samples %5.1f%%  inst %5.1f%%  %s" % (f[:16], ln, 100.0 * a[0] / max(tot_s, 1), 100.0 * a[1] / max(tot_i, 1), a[2][:110]))
