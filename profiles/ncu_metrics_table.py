#!/usr/bin/env python3
"""Per-launch table from an `ncu --metrics ... --csv` log with several metrics per launch (long format).
usage: ncu_metrics_table.py launches.csv [first_id last_id]   -> one line per launch: kernel, grid, time, dram read/write, sm%, dram%, regs, warp instr"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1], errors="ignore")))
hdr = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
H = rows[hdr]
ix = {h: i for i, h in enumerate(H)}
L = {}
for r in rows[hdr + 1:]:
    if len(r) <= ix["Metric Value"]:
        continue
    i = int(r[0])
    d = L.setdefault(i, {"name": r[ix["Kernel Name"]], "grid": r[ix["Grid Size"]], "block": r[ix["Block Size"]]})
    v = r[ix["Metric Value"]].replace(",", "")
    try:
        v = float(v)
    except ValueError:
        pass
    unit = r[ix["Metric Unit"]]
    if unit == "Kbyte": v *= 1e3
    elif unit == "Mbyte": v *= 1e6
    elif unit == "Gbyte": v *= 1e9
    elif unit in ("usecond", "us"): v *= 1e3
    elif unit in ("msecond", "ms"): v *= 1e6
    elif unit in ("second", "s"): v *= 1e9
    d[r[ix["Metric Name"]]] = v
lo, hi = (int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else (0, 1 << 30)
print("%4s %-58s %10s %9s %9s %9s %6s %6s %5s %10s" % ("id", "kernel", "grid", "ms", "rd_GB", "wr_GB", "sm%", "dram%", "regs", "winst_M"))
tot = 0.0
for i in sorted(L):
    if i < lo or i > hi:
        continue
    d = L[i]
    name = d["name"].split("(")[0].replace("cdgsort::", "").replace("cdgkmer::", "").replace("cdgalign::", "").replace("cdgscan::", "").replace("cdgexp::", "")
    if "<" in d["name"].split("(")[0]:
        pass
    t = d.get("gpu__time_duration.sum", 0.0) / 1e6
    tot += t
    print("%4d %-58s %10s %9.3f %9.3f %9.3f %6.1f %6.1f %5d %10.2f" % (
        i, name[:58], d["grid"].split(",")[0].strip("("), t, d.get("dram__bytes_read.sum", 0) / 1e9, d.get("dram__bytes_write.sum", 0) / 1e9,
        d.get("sm__throughput.avg.pct_of_peak_sustained_elapsed", 0), d.get("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", 0),
        int(d.get("launch__registers_per_thread", 0)), d.get("smsp__inst_executed.sum", 0) / 1e6))
print("total ms %.3f" % tot)
