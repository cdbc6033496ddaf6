#!/usr/bin/env python3
"""One context beyond 2^30 k-mer records (64-bit tile states in the sort): 40 M reads -> 1.38e9 records.  Checks the
size-independent properties of the result (hit list strictly sorted, representatives longest, self alignments, only base bytes
rewritten) and prints the stage times next to a 30 M-read run that stays on the 32-bit path."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import carpedeam_b200 as cd  # noqa: E402
from bench import synth_reads_torch  # noqa: E402

dev = torch.device("cuda", 0)
ctx = cd.Context(0)
ctx.set_damage(cd.load_damage(os.path.join(ROOT, "tests", "golden", "dhigh")))
for n in (30_000_000, 40_000_000):
    s = synth_reads_torch(n, 5 * n, 20260102, dev)
    ctx.upload(s["blob"], s["off"], s["len"], None)
    hits, alns, out, changed = ctx.iteration(fetch_hits=True)
    st = ctx.stats()
    rep, tgt = hits["rep"].astype(np.int64), hits["tgt"].astype(np.int64)
    pair = rep * (1 << 32) + tgt
    L = s["len"].astype(np.int64)
    lo = alns["line_off"].astype(np.int64)
    ok = dict(sorted=bool(np.all(np.diff(pair) > 0)), rep_longest=bool(np.all(L[rep] >= L[tgt])), no_self=bool(np.all(rep != tgt)),
              self_kept=bool(np.all(alns["keep"][lo[:-1]] == 1) and np.all(alns["score"][lo[:-1]] == 2 * L)),
              only_bases=bool(set(np.unique(out[np.nonzero(out != s["blob"])[0]]).tolist()) <= set(b"ACGT")),
              changed=int(changed) == int((out != s["blob"]).sum()))
    print("n=%d records=%d grouped=%d hits=%d changed=%d passes=%d/%d  ms: extract %.1f sort1 %.1f group %.1f sort2 %.1f reduce %.1f rescore %.1f correct %.1f  checks %s" % (
        n, st["n_records"], st["n_grouped"], st["n_hits"], changed, st["sort1_passes"], st["sort2_passes"], st["ms_extract"], st["ms_sort1"], st["ms_group"],
        st["ms_sort2"], st["ms_reduce"], st["ms_rescore"], st["ms_correct"], ok), flush=True)
    assert all(ok.values()), ok
    del s, hits, alns, out
ctx.close()
