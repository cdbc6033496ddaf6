#!/usr/bin/env python3
"""profiles/traffic.json from an `ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,... --csv` log of
`bench.py` and the bench line of the same command run without ncu (for the record count of a sort-1 pass).
usage: ncu_traffic_summary.py metrics.csv bench_line.json > traffic.json
Per kernel family: launches per captured run, mean duration, mean dram read+write bytes per launch; for the sort-1 pass
(onesweep_kernel<256,16,1,...> with the large grid) also dram bytes over the algorithmic 2*16*records."""
import csv
import json
import re
import sys

rows = list(csv.reader(open(sys.argv[1], errors="ignore")))
hdr = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
ix = {h: i for i, h in enumerate(rows[hdr])}
L = {}
for r in rows[hdr + 1:]:
    if len(r) <= ix["Metric Value"]:
        continue
    d = L.setdefault(int(r[0]), {"name": r[ix["Kernel Name"]], "grid": int(r[ix["Grid Size"]].strip("()").split(",")[0])})
    v = float(r[ix["Metric Value"]].replace(",", ""))
    v *= {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "usecond": 1e3, "us": 1e3, "msecond": 1e6, "ms": 1e6, "second": 1e9}.get(r[ix["Metric Unit"]], 1.0)
    d[r[ix["Metric Name"]]] = v
bench = json.loads([l for l in open(sys.argv[2]) if l.startswith("{")][-1])
records = int(re.search(r"(\d+) records", bench["roofline"]["kernel"]).group(1))
fam = {}
for d in L.values():
    name = re.sub(r"\(.*", "", d["name"]).replace("void ", "").replace("<unnamed>::", "")
    if "onesweep_kernel<256, 16, 1" in name or "onesweep_kernel<(int)256, (int)16, (bool)1" in name:
        name = "onesweep_pairs" if d["grid"] * 4096 >= 0.9 * records else "onesweep_pairs(small)"
    elif "onesweep_kernel" in name:
        name = "onesweep_keys"
    f = fam.setdefault(name, {"launches": 0, "ns": 0.0, "dram": 0.0})
    f["launches"] += 1; f["ns"] += d.get("gpu__time_duration.sum", 0.0)
    f["dram"] += d.get("dram__bytes_read.sum", 0.0) + d.get("dram__bytes_write.sum", 0.0)
out = {"source": sys.argv[1].split("/")[-1] + " (ncu --clock-control none, per-launch dram__bytes_read.sum + dram__bytes_write.sum; workload: " +
       bench["config"]["workload"] + ")", "sort1_records": records}
for name, f in sorted(fam.items(), key=lambda kv: -kv[1]["ns"]):
    if f["ns"] < 5e4 * f["launches"] and not name.startswith("onesweep"):
        continue
    out[name] = {"launches": f["launches"], "ms_per_launch_under_ncu": f["ns"] / f["launches"] / 1e6, "dram_bytes_per_launch": f["dram"] / f["launches"]}
p = out.get("onesweep_pairs")
if p:
    p["algorithmic_bytes_per_launch"] = 2 * 16 * records
    p["dram_over_algorithmic"] = p["dram_bytes_per_launch"] / (2 * 16 * records)
json.dump(out, sys.stdout, indent=1)
print()
