#!/usr/bin/env python3
"""TEST INFRASTRUCTURE — regenerates the known answers of cyclecheck: tests/golden/synth_cycle.gz (sequence set) and
tests/golden/cycle_kat.txt (`key <TAB> split diagonal`, 0 = not reported).

Runs the UNMODIFIED reference binary (oracle/_ref/carpedeam_ref cyclecheck --chop-cycle 1: the reported entries are cut at the split diagonal,
so the entry length IS the diagonal) on a seeded synthetic set: linear contigs, circular genomes sequenced past their start (exact and mutated
overlaps of 3-60 %), tandem repeats, low-complexity sequences, N runs, sequences shorter than three k-mers."""
import os
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from make_golden import canonical, load_db  # noqa: E402
from make_merge_kat import gz_write  # noqa: E402
from support import GOLD, write_db  # noqa: E402


def synth_cycle_set(seed=20260120):
    rng = np.random.default_rng(seed)
    acgt = np.frombuffer(b"ACGT", np.uint8)
    rnd = lambda L: acgt[rng.integers(0, 4, L)]
    seqs = []
    for _ in range(40):                                   # linear
        seqs.append(rnd(int(rng.integers(60, 6000))))
    for _ in range(60):                                   # circular, read past the start
        C = int(rng.integers(300, 5000))
        g = rnd(C)
        ov = int(C * rng.uniform(0.03, 0.6))
        tail = g[:ov].copy()
        if rng.random() < 0.5:                            # mutated copy
            m = rng.random(ov) < rng.choice([0.005, 0.02, 0.08])
            tail[m] = acgt[rng.integers(0, 4, int(m.sum()))]
        s = np.concatenate([g, tail])
        if rng.random() < 0.2:
            s[rng.integers(0, len(s), 4)] = ord("N")
        seqs.append(s)
    for _ in range(10):                                   # tandem repeats / low complexity
        unit = rnd(int(rng.integers(1, 120)))
        seqs.append(np.tile(unit, int(rng.integers(3, 40)))[:4000])
    seqs.append(np.full(500, ord("A"), np.uint8))
    seqs.append(np.full(90, ord("N"), np.uint8))
    for L in (1, 21, 22, 23, 65, 66, 67, 68):             # around the length limits of the k-mer parts
        seqs.append(rnd(L))
    g = rnd(30000); seqs.append(np.concatenate([g, g[:9000]]))   # a long one
    return [s.tobytes() for s in seqs]


def main():
    binary = os.path.join(HERE, "_ref", "carpedeam_ref")
    if not os.path.exists(binary):
        sys.exit("build oracle/_ref first: python oracle/build_ref.py")
    work = tempfile.mkdtemp(prefix="cdcyc")
    seqs = synth_cycle_set()
    db = os.path.join(work, "seq")
    write_db(db, list(range(len(seqs))), [s + b"\n" for s in seqs], [1] * len(seqs), 1)
    subprocess.run([binary, "cyclecheck", db, db + "_cycle", "--chop-cycle", "1", "--threads", "4", "-v", "1"], check=True)
    out = load_db(db + "_cycle")
    gz_write(os.path.join(GOLD, "synth_cycle.gz"), canonical(load_db(db)))
    with open(os.path.join(GOLD, "cycle_kat.txt"), "w") as f:
        for k in range(len(seqs)):
            f.write("%d\t%d\n" % (k, len(out[k][0].rstrip(b"\n")) if k in out else 0))
    print(len(seqs), "sequences,", len(out), "reported as circular")


if __name__ == "__main__":
    main()
