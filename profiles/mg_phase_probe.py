#!/usr/bin/env python3
"""Single-GPU probe of the multi-GPU phases: rank 0 of an emulated world runs cdg_mg_extract on `world` x the single-GPU
read set, then sorts/groups its own partitioned records as if they had arrived through the all-to-all (ownership is wrong,
record counts and kernel work per rank are right).  Run under `ncu --metrics gpu__time_duration.sum` for the per-kernel
times of the per-rank work that grows with the world size (length order, identity hashes, scans over all sequences)."""
import argparse
import ctypes as C
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import carpedeam_b200 as cd  # noqa: E402
from bench import synth_reads_torch  # noqa: E402
from carpedeam_b200.multigpu import ShardedIteration  # noqa: E402


class FakeComm:
    def __init__(self, world, dev):
        self.world, self.rank, self.device, self.torch = world, 0, dev, torch


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--world", type=int, default=8)
    ap.add_argument("--reads", type=int, default=10_000_000)
    a = ap.parse_args()
    dev = torch.device("cuda", 0)
    s = synth_reads_torch(a.reads * a.world, 50_000_000 * a.world, 20260102, dev)
    ctx = cd.Context(0)
    ctx.upload(s["blob"], s["off"], s["len"], None)
    ctx.set_damage(cd.load_damage(os.path.join(ROOT, "tests", "golden", "dhigh")))
    sh = ShardedIteration(ctx, FakeComm(a.world, dev))
    L = sh.L
    kp = cd.kmer_params()
    for it in range(2):
        t = {}
        def lap(name, t0):
            torch.cuda.synchronize(); t[name] = (time.perf_counter() - t0) * 1e3
        t0 = time.perf_counter()
        sc = (C.c_uint64 * a.world)(); isc = (C.c_uint64 * a.world)()
        sk, sv, ik, iv = C.c_void_p(), C.c_void_p(), C.c_void_p(), C.c_void_p()
        ctx._ck(L.cdg_mg_extract(ctx._h, C.byref(kp), a.world, 0, 1, sc, C.byref(sk), C.byref(sv), isc, C.byref(ik), C.byref(iv)), "cdg_mg_extract")
        lap("extract+partition", t0)
        total, id_total = sum(sc), sum(isc)
        rk, rv, irk, irv = C.c_void_p(), C.c_void_p(), C.c_void_p(), C.c_void_p()
        ctx._ck(L.cdg_mg_recv_records(ctx._h, total, id_total, C.byref(rk), C.byref(rv), C.byref(irk), C.byref(irv)), "cdg_mg_recv_records")
        sh._t(rk.value, total).copy_(sh._t(sk.value, total)); sh._t(rv.value, total).copy_(sh._t(sv.value, total))
        sh._t(irk.value, id_total).copy_(sh._t(ik.value, id_total)); sh._t(irv.value, id_total).copy_(sh._t(iv.value, id_total))
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        mk = C.c_uint64()
        ctx._ck(L.cdg_mg_sort(ctx._h, C.byref(kp), C.byref(mk)), "cdg_mg_sort")
        lap("identity+sort1", t0)
        t0 = time.perf_counter()
        sc2 = (C.c_uint64 * a.world)(); s2 = C.c_void_p()
        ctx._ck(L.cdg_mg_group(ctx._h, C.byref(kp), 1, 1, sc2, C.byref(s2)), "cdg_mg_group")
        lap("group+route", t0)
        print("iteration %d: records %d  " % (it, total) + "  ".join("%s %.1f ms" % kv for kv in t.items()), flush=True)
    ctx.close()


if __name__ == "__main__":
    main()
