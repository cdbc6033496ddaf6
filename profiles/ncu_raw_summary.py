#!/usr/bin/env python3
"""Per-kernel summary of an `ncu --page raw --csv` export: duration, DRAM bytes, throughput %, occupancy, registers, top stalls."""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
H = rows[0]
def col(name):
    return H.index(name) if name in H else None
base = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__inst_executed.sum", "launch__grid_size", "launch__block_size", "sm__maximum_warps_per_active_cycle_pct"]
stall = [h for h in H if h.startswith("smsp__average_warp") and "issue_stalled" in h and h.endswith("_per_warp_active.pct") is False and "ratio" in h]
if not stall:
    stall = [h for h in H if "issue_stalled" in h and h.endswith(".ratio")]
idx = [col(b) for b in base]
units = rows[1]
for r in rows[2:]:
    name = r[idx[0]].split("(")[0][:48]
    g = lambda i: r[idx[i]] if idx[i] is not None else "-"
    print("%-48s dur=%s%s rd=%s%s wr=%s%s dram%%=%s occ%%=%s regs=%s sm%%=%s l1%%=%s l2%%=%s inst=%s grid=%s blk=%s" % (
        name, g(1), units[idx[1]], g(2), units[idx[2]] if idx[2] is not None else "", g(3), units[idx[3]] if idx[3] is not None else "",
        g(4), g(5), g(6), g(7), g(8), g(9), g(10), g(11), g(12)))
    st = sorted(((float(r[H.index(s)].replace(",", "")) if r[H.index(s)] not in ("", "n/a") else 0.0, s) for s in stall), reverse=True)[:4]
    print("      top stalls: " + "; ".join("%s=%.2f" % (s.replace("smsp__average_warps_issue_stalled_", "").replace("smsp__average_warp_latency_issue_stalled_", "").replace("_per_issue_active.ratio", "").replace(".ratio", ""), v) for v, s in st))
