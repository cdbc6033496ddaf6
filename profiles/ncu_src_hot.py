#!/usr/bin/env python3
"""Hot SASS instructions of an `ncu --page source --csv` export: stall samples by opcode class and top lines."""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
body = [r for r in rows[2:] if len(r) > 5 and r[0].startswith("0x")]
tot = sum(int(r[2]) for r in body)
print("total samples", tot, "instructions", len(body))
byop = collections.Counter()
for r in body:
    op = r[1].split()
    op = [o for o in op if not o.startswith("@")][0].split(".")[0]
    byop[op] += int(r[2])
print("by opcode:", ", ".join("%s=%.1f%%" % (k, 100 * v / tot) for k, v in byop.most_common(12)))
for i, r in sorted(enumerate(body), key=lambda x: -int(x[1][2]))[:top]:
    print("%5d %6.2f%%  exec=%-9s %s" % (i, 100 * int(r[2]) / tot, r[5], r[1].strip()[:90]))
