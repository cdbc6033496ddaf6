#!/usr/bin/env python3
"""profiles/r02_sass_summary.txt + r02_sass_hot_kernels.txt.gz from `cuobjdump -sass carpedeam_b200/libcarpedeam_gpu.so` (runs without a GPU).
Per hot kernel: static instruction count and counts of the memory / collective / fp64 / TMA mnemonics."""
import collections
import gzip
import os
import re
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(os.path.dirname(HERE), "carpedeam_b200", "libcarpedeam_gpu.so")
HOT = ("onesweep_kernelILi256ELi16ELb1E", "onesweep_kernelILi512ELi16ELb0ELb0Ej", "hist_kernel", "extract_group_kernelILi8ELi8E", "group_kernelILb0E", "reduce_kernelILb0E",
       "pair_end_count_kernel", "id_fixup_kernel", "rescore_kernel", "corr_prep_kernel", "correct_fast_kernel", "pair_loglik_kernelILi1E", "ext_lines_kernel",
       "merge_lines_kernel", "merge_score_kernel", "hamming_kernel", "cycle_emit_kernel", "cycle_match_kernel", "cycle_band_kernel", "hash_hist_kernel", "compact_records_kernel")
KEEP = re.compile(r"^(LDG|STG|LDS|STS|ATOM|RED|SHFL|VOTE|BAR|WARPSYNC|POPC|BREV|MUFU|D(ADD|MUL|FMA)|UBLKCP|UTMA|SYNCS|CREDUX|MATCH|LDGSTS)")

out = subprocess.run(["cuobjdump", "-sass", SO], capture_output=True, text=True, check=True).stdout
funcs, cur = collections.OrderedDict(), None
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1) if any(h in m.group(1) for h in HOT) else None
        if cur:
            funcs[cur] = []
        continue
    if cur and re.match(r"\s+/\*[0-9a-f]{4}\*/", line):
        funcs[cur].append(line)
with open(os.path.join(HERE, "r02_sass_summary.txt"), "w") as f:
    f.write("SASS of the hot kernels (cuobjdump -sass carpedeam_b200/libcarpedeam_gpu.so, sm_100a; profiles/sass_summary.py): static instruction counts of the "
            "memory / collective / fp64 mnemonics.\nFull listings: r02_sass_hot_kernels.txt.gz\n\n")
    for name, lines in funcs.items():
        ops = collections.Counter()
        for l in lines:
            t = re.sub(r"^\s+/\*[0-9a-f]+\*/\s+", "", l).split()
            op = t[1] if t and t[0].startswith("@") and len(t) > 1 else (t[0] if t else "")
            op = op.rstrip(";")
            if KEEP.match(op):
                ops[op] += 1
        f.write("%s\n   %d instructions; %s\n" % (name, len(lines), ", ".join("%s %d" % kv for kv in sorted(ops.items()))))
with gzip.GzipFile(os.path.join(HERE, "r02_sass_hot_kernels.txt.gz"), "wb", mtime=0) as f:
    for name, lines in funcs.items():
        f.write(("Function : %s\n" % name).encode() + "\n".join(lines).encode() + b"\n\n")
print(len(funcs), "kernels")
