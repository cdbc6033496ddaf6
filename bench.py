#!/usr/bin/env python3
"""bench.py — Mreads/s of one reads-only iteration (kmermatch + rescore + correction) on synthetic damaged reads.

Workload = BASELINE.json configs[1]: N merged reads, length U{35..70}, one i.i.d. uniform
genome at ~10.5x coverage, dhigh damage profile, numpy default_rng(20260102).  A "step" is one pass of the hot path
over that read set.

  value : reads / device time of the fused iteration with the packed read set already resident in HBM
  e2e   : same, through the C ABI with HOST buffers: pinned ASCII reads -> H2D + pack -> iteration -> alignment SoA and
          corrected reads back in pinned host memory, all inside the timed region
  roofline : dominant kernel (LSD radix pass `onesweep_kernel`), algorithmic bytes / measured launch time vs the
          measured HBM copy bandwidth in MEASURED_PEAKS.json
  cpu_baseline / --impl reference : the UNMODIFIED reference binary (oracle/_ref/carpedeam_ref) running its three stage
          commands with all host threads on a bounded sample of the same workload (same coverage, fewer reads)
"""
import argparse
import ctypes as C
import json
import os
import shutil
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

DHIGH5 = np.array([.329405, .221745, .187678, .161196, .144011])
DHIGH3 = np.array([.32891, .223405, .188599, .164419, .146352])   # 3' positions L-1, L-2, ... L-5
ACGT = np.frombuffer(b"ACGT", np.uint8)
COMP = np.zeros(256, np.uint8)
for _a, _b in zip(b"ACGT", b"TGCA"):
    COMP[_a] = _b


def synth_reads_fast(n, genome_len, seed, lmin=35, lmax=70, interior=0.01, sub_err=0.001, chunk=1_000_000):
    """SURVEY §8(d) config-2 generator, vectorised.  Returns the MMseqs2 data-file layout (`BASES\\n\\0` per entry)."""
    rng = np.random.default_rng(seed)
    genome = ACGT[rng.integers(0, 4, genome_len, dtype=np.uint8)]
    lens = rng.integers(lmin, lmax + 1, n).astype(np.uint32)
    off = np.zeros(n, np.uint64)
    off[1:] = np.cumsum(lens[:-1].astype(np.uint64) + 2)
    total = int(lens.sum())
    blob = np.zeros(total + 2 * n, np.uint8)
    j = np.arange(lmax, dtype=np.int64)[None, :]
    for c0 in range(0, n, chunk):
        m = min(chunk, n - c0)
        L = lens[c0:c0 + m].astype(np.int64)[:, None]
        start = rng.integers(0, genome_len - lmax, m)[:, None]
        strand = (rng.random(m) < 0.5)[:, None]
        valid = j < L
        idx = np.where(strand, start + L - 1 - j, start + j)
        idx = np.where(valid, idx, 0)
        r = genome[idx]
        r = np.where(strand, COMP[r], r)
        p_ct = np.full((m, lmax), interior); p_ct[:, :5] = DHIGH5
        d3 = L - 1 - j                                  # distance from the 3' end
        p_ga = np.where((d3 >= 0) & (d3 < 5), DHIGH3[np.clip(d3, 0, 4)], interior)
        u = rng.random((m, lmax))
        r = np.where((r == ord("C")) & (u < p_ct), ord("T"), r)
        r = np.where((r == ord("G")) & (u < p_ga), ord("A"), r).astype(np.uint8)
        e = rng.random((m, lmax)) < sub_err
        r = np.where(e, ACGT[rng.integers(0, 4, (m, lmax), dtype=np.uint8)], r).astype(np.uint8)
        flat = r[valid]
        row = np.repeat(np.arange(m, dtype=np.int64), L[:, 0])
        pos = np.arange(len(flat), dtype=np.int64) + 2 * row + int(off[c0])
        blob[pos] = flat
        ends = (off[c0:c0 + m] + lens[c0:c0 + m]).astype(np.int64)
        blob[ends] = ord("\n")
    return dict(blob=blob, off=off, len=lens, n=n, total=total)


def synth_reads_torch(n, genome_len, seed, device, lmin=35, lmax=70, interior=0.01, sub_err=0.001, chunk=4_000_000):
    """Same read model generated on the GPU (torch.Generator, identical on every rank) — used for the multi-GPU runs where
    the numpy generator would take minutes; returns host numpy arrays in the MMseqs2 data-file layout."""
    import torch
    g = torch.Generator(device=device); g.manual_seed(seed)
    acgt = torch.tensor(list(b"ACGT"), dtype=torch.uint8, device=device)
    comp = torch.zeros(256, dtype=torch.uint8, device=device)
    for a, b in zip(b"ACGT", b"TGCA"):
        comp[a] = b
    genome = acgt[torch.randint(0, 4, (genome_len,), generator=g, device=device)]
    lens = torch.randint(lmin, lmax + 1, (n,), generator=g, device=device, dtype=torch.int64)
    off = torch.zeros(n, dtype=torch.int64, device=device)
    off[1:] = torch.cumsum(lens[:-1] + 2, 0)
    total = int(lens.sum().item())
    blob = torch.zeros(total + 2 * n, dtype=torch.uint8, device=device)
    j = torch.arange(lmax, device=device)[None, :]
    d5 = torch.tensor(DHIGH5, device=device, dtype=torch.float32); d3 = torch.tensor(DHIGH3, device=device, dtype=torch.float32)
    for c0 in range(0, n, chunk):
        m = min(chunk, n - c0)
        L = lens[c0:c0 + m][:, None]
        start = torch.randint(0, genome_len - lmax, (m, 1), generator=g, device=device)
        strand = torch.rand((m, 1), generator=g, device=device) < 0.5
        valid = j < L
        idx = torch.where(strand, start + L - 1 - j, start + j)
        idx = torch.where(valid, idx, torch.zeros_like(idx))
        r = genome[idx]
        r = torch.where(strand, comp[r.long()], r)
        p_ct = torch.full((m, lmax), interior, device=device); p_ct[:, :5] = d5
        dd = L - 1 - j
        p_ga = torch.where((dd >= 0) & (dd < 5), d3[dd.clamp(0, 4)], torch.full_like(p_ct, interior))
        u = torch.rand((m, lmax), generator=g, device=device)
        r = torch.where((r == ord("C")) & (u < p_ct), torch.full_like(r, ord("T")), r)
        r = torch.where((r == ord("G")) & (u < p_ga), torch.full_like(r, ord("A")), r)
        e = torch.rand((m, lmax), generator=g, device=device) < sub_err
        r = torch.where(e, acgt[torch.randint(0, 4, (m, lmax), generator=g, device=device)], r)
        flat = r[valid]
        row = torch.repeat_interleave(torch.arange(m, device=device), lens[c0:c0 + m])
        pos = torch.arange(flat.numel(), device=device) + 2 * row + off[c0]
        blob[pos] = flat
        blob[off[c0:c0 + m] + lens[c0:c0 + m]] = ord("\n")
        del r, u, e, idx, flat, row, pos, p_ct, p_ga
    out = dict(blob=blob.cpu().numpy(), off=off.cpu().numpy().astype(np.uint64), len=lens.cpu().numpy().astype(np.uint32), n=n, total=total)
    del blob, genome
    torch.cuda.empty_cache()
    return out


# ---------------------------------------------------------------------------------------------------------------
def clocks_sampler(stop, out, dev):
    q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
    while not stop.is_set():
        try:
            r = subprocess.run(["nvidia-smi", "-i", str(dev), "--query-gpu=" + q, "--format=csv,noheader,nounits"],
                               capture_output=True, text=True, timeout=5)
            if r.returncode == 0 and r.stdout.strip():
                out.append([x.strip() for x in r.stdout.strip().split(",")])
        except Exception:
            pass
        stop.wait(0.2)


def clocks_summary(samples):
    if not samples:
        return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
    sm = sorted(int(s[0]) for s in samples if s[0].isdigit())
    reasons = set()
    names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
    for s in samples:
        for nm, v in zip(names, s[2:6]):
            if v.lower().startswith("active"):
                reasons.add(nm)
    return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": int(samples[0][1]) if samples[0][1].isdigit() else None,
            "reasons": sorted(reasons)}


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def alg_bytes(st, k):
    """SURVEY §8(d) pass-aware algorithmic bytes of one iteration, from the run's own counters."""
    S, N, R1, R2, H, A = st["n_bases"], st["n_seqs"], st["n_records"], st["n_grouped"], st["n_hits"], st["n_alns"]
    P1, P2 = st["sort1_passes"], st["sort2_passes"]
    Lbar = S / max(N, 1)
    return (S / 4 + 12 * N) + 16 * R1 + 16 * R1 * (1 + 2 * P1) + 16 * (R1 + R2) + 16 * R2 * (1 + 2 * P2) + 16 * R2 + 8 * H + \
        H * (8 + 2 * Lbar / 4 + 32) + A * (32 + Lbar / 4) + S / 4 * 2


# dram__bytes_read.sum + dram__bytes_write.sum of onesweep_kernel<256,16,true> over its algorithmic 2*16*records bytes, measured with
# ncu --set full (profiles/r01j_onesweep_full_raw_2Mreads.csv: 1.088 GB read + 1.044 GB written per launch of 16358 tiles x 4096
# records = 2.144 GB algorithmic)
ONESWEEP_TRAFFIC_RATIO = 0.995


# ---------------------------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the unmodified reference binary on a bounded sample
REF_BIN = os.path.join(ROOT, "oracle", "_ref", "carpedeam_ref")


def ref_available():
    if not os.path.exists(REF_BIN):
        return False
    try:
        flags = open("/proc/cpuinfo").read()
        return " avx2 " in flags or "avx2" in flags.split()
    except Exception:
        return True


def write_fasta(s, path):
    blob, off, lens = s["blob"], s["off"], s["len"]
    with open(path, "wb") as f:
        parts = []
        for i in range(s["n"]):
            o = int(off[i])
            parts.append(b">r%d\n" % i + blob[o:o + int(lens[i])].tobytes() + b"\n")
            if len(parts) >= 100000:
                f.write(b"".join(parts)); parts = []
        f.write(b"".join(parts))


class RefRunner:
    """createdb once, then time kmermatcher + rescorediagonal + ancient_correction (default argv)."""

    def __init__(self, sample, threads):
        self.n = sample["n"]; self.threads = threads
        self.dir = tempfile.mkdtemp(prefix="cdref")
        fa = os.path.join(self.dir, "reads.fa")
        write_fasta(sample, fa)
        self.db = os.path.join(self.dir, "db")
        self._run(["createdb", fa, self.db, "--dbtype", "2", "--shuffle", "0"])
        self.damage = os.path.join(ROOT, "tests", "golden", "dhigh")

    def _run(self, args):
        r = subprocess.run([REF_BIN] + args, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("reference stage failed: %s\n%s" % (" ".join(args), r.stdout[-2000:] + r.stderr[-2000:]))

    def step(self):
        d = self.dir
        for f in os.listdir(d):
            if f.startswith(("pref", "aln", "corr")):
                os.remove(os.path.join(d, f))
        th = str(self.threads)
        t0 = time.perf_counter()
        self._run(["kmermatcher", self.db, d + "/pref", "--min-seq-id", "0.9", "--kmer-per-seq", "200", "--kmer-per-seq-scale", "0.2",
                   "--cov-mode", "1", "-k", "20", "-c", "0", "--max-seq-len", "200000", "--hash-shift", "67", "--include-only-extendable", "0",
                   "--ignore-multi-kmer", "1", "--threads", th, "-v", "1"])
        t1 = time.perf_counter()
        self._run(["rescorediagonal", self.db, self.db, d + "/pref", d + "/aln", "--rescore-mode", "3", "-e", "0.001", "-c", "0", "-a", "0",
                   "--cov-mode", "1", "--min-seq-id", "0.9", "--min-aln-len", "0", "--seq-id-mode", "0", "--sort-results", "0", "--threads", th,
                   "-v", "1"])
        t2 = time.perf_counter()
        self._run(["ancient_correction", self.db, d + "/aln", d + "/corr", "--min-seq-id", "0.9", "--max-seq-len", "200000", "--threads", th,
                   "-v", "1", "--ancient-damage", self.damage])
        t3 = time.perf_counter()
        return t3 - t0, (t1 - t0, t2 - t1, t3 - t2)

    def close(self):
        shutil.rmtree(self.dir, ignore_errors=True)


def reference_arm(args, sample_reads):
    cores = os.cpu_count() or 1
    genome = max(200_000, int(sample_reads * 5))   # 52.5 b * n / genome = 10.5x like the full workload
    sample = synth_reads_fast(sample_reads, genome, 20260102)
    sample_desc = "%d reads (35-70 bp) over a %d bp genome: same 10.5x coverage and dhigh damage as the full workload" % (sample_reads, genome)
    if ref_available():
        rr = RefRunner(sample, cores)
        try:
            for _ in range(args.warmup):
                rr.step()
            times = []; stages = None
            for _ in range(args.steps):
                t, stages = rr.step()
                times.append(t)
        finally:
            rr.close()
        kind = "reference"
        note = "kmermatcher %.2fs rescorediagonal %.2fs ancient_correction %.2fs (last step)" % stages
    else:   # oracle port (scalar, 1 thread)
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        from support import Oracle, SeqSet
        small = min(sample_reads, 100_000)
        sample = synth_reads_fast(small, max(200_000, small * 5), 20260102)
        sample_desc = "%d reads over a %d bp genome (oracle port, oracle/_ref unavailable on this host)" % (small, max(200_000, small * 5))
        o = Oracle()
        s = SeqSet.__new__(SeqSet)
        s.n = sample["n"]; s.blob = sample["blob"]; s.off = sample["off"]; s.len = sample["len"]; s.total = sample["total"]
        s.ext = np.zeros(s.n, np.uint8); s.keys = np.arange(s.n, dtype=np.uint32)
        times = []
        for _ in range(max(1, args.steps)):
            t0 = time.perf_counter(); o.kmermatch(s); times.append(time.perf_counter() - t0)
        cores = 1; kind = "port"; note = "kmermatch stage only (oracle port)"
    sec = float(np.mean(times))
    val = sample["n"] / sec / 1e6
    return dict(value=val, sec=sec, cores=cores, kind=kind, sample=sample_desc, note=note, n=sample["n"])


# ---------------------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--reads", type=int, default=10_000_000)
    ap.add_argument("--genome", type=int, default=50_000_000)
    ap.add_argument("--cpu-sample", type=int, default=1_000_000)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--exchange", default="p2p", choices=["p2p", "nccl"],
                    help="multi-GPU record exchange: sender-side partition kernel writing into peer memory over NVLink (p2p) or NCCL all-to-all")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
    workload = "synthetic %d merged reads, 35-70 bp, dhigh damage, %d bp genome (BASELINE configs[1])" % (args.reads, args.genome)

    if args.impl == "reference":
        if rank != 0:
            return 0
        r = reference_arm(args, args.cpu_sample)
        line = {"impl": "reference", "metric": "Mreads/s kmermatch+rescore+correction", "value": r["value"], "unit": "Mreads/s",
                "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": r["sec"] * 1e3, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "u8/int64 (+f64/f80 likelihoods)", "data": "synthetic",
                "config": {"workload": workload, "sample": r["sample"]},
                "cpu_baseline": {"value": r["value"], "unit": "Mreads/s", "cores": r["cores"], "kind": r["kind"], "sample": r["sample"], "note": r["note"]},
                "e2e": {"value": r["value"], "unit": "Mreads/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return 0

    import torch
    import carpedeam_b200 as cd
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: a CUDA device is required (no CPU fallback)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        import torch.distributed as dist
        from carpedeam_b200.multigpu import ShardedIteration, TorchComm
        dist.init_process_group("nccl", device_id=dev)
    # ---- workload: identical on every rank; weak scaling = world x the single-GPU read set (same coverage) ----
    n_total = args.reads * world
    genome = args.genome * world
    if world == 1:
        s = synth_reads_fast(n_total, genome, 20260102)
        data_note = "synthetic (numpy default_rng(20260102))"
    else:
        s = synth_reads_torch(n_total, genome, 20260102, dev)
        data_note = "synthetic (same read model, torch.Generator seed 20260102 on the GPU, %d x the single-GPU set)" % world
    workload = "synthetic %d merged reads, 35-70 bp, dhigh damage, %d bp genome (BASELINE configs[1]%s)" % (
        n_total, genome, "" if world == 1 else " x %d, hash-sharded" % world)
    n = s["n"]
    pin = lambda a: torch.from_numpy(a).pin_memory()
    h_blob, h_off, h_len = pin(s["blob"]), pin(s["off"].view(np.int64)), pin(s["len"].view(np.int32))
    h_out = torch.empty(s["blob"].nbytes, dtype=torch.uint8).pin_memory()
    ctx = cd.Context(local)
    ctx.set_damage(cd.load_damage(os.path.join(ROOT, "tests", "golden", "dhigh")))
    ip = cd.IterParams(cd.kmer_params(), cd.RescoreParams(0.9, 0.001, 0), cd.CorrParams(0.9, 0.99))
    stream = torch.cuda.ExternalStream(ctx.stream)
    sharded = ShardedIteration(ctx, TorchComm(dev), exchange=args.exchange) if world > 1 else None

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def upload():
        ctx.upload_raw(h_blob.data_ptr(), h_blob.numel(), h_off.data_ptr(), h_len.data_ptr(), 0, n)

    def step_resident():
        if world == 1:
            ctx.iteration_raw(ip, None, None, 0)
        else:
            last["phase_ms"] = sharded.run()["phase_ms"]

    alns = cd.Alns()
    last = {}

    def step_e2e():
        if world == 1:   # asynchronous upload: the iteration packs and extracts piece c while piece c+1 is still on the bus
            ctx.upload_raw_async(h_blob.data_ptr(), h_blob.numel(), h_off.data_ptr(), h_len.data_ptr(), 0, n)
        else:   # every rank copies 1/world of the host arrays, NVLink all-gather, pack from device memory
            last["h2d"] = sharded.upload_sharded(h_blob, h_off, h_len)
        if world == 1:
            ctx.iteration_raw(ip, None, alns, h_out.data_ptr())
            last["M"] = alns.n_lines; last["d2h_seq"] = h_blob.numel()
        else:
            t_up = time.perf_counter()
            r = sharded.run(fetch_alns="raw", out_range=h_out.data_ptr())
            if os.environ.get("CDG_MG_TRACE") is not None and rank == 0:
                print("e2e run (incl. exports) %.1f ms  phases %s" % ((time.perf_counter() - t_up) * 1e3, {k: round(v, 1) for k, v in r["phase_ms"].items()}),
                      file=sys.stderr, flush=True)
            last["M"] = int(r["alns_raw"].n_lines); last["d2h_seq"] = r["range_bytes"]; last["sent"] = r["records_sent"]; last["sent2"] = r["grouped_sent"]

    # ---- resident-input timing ----
    upload()
    for _ in range(args.warmup):
        step_resident()
    samples = []; stop = threading.Event()
    th = threading.Thread(target=clocks_sampler, args=(stop, samples, local)); th.start()
    l0 = ctx.stats()["kernel_launches"]
    barrier()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(args.steps):
        step_resident()
    e1.record(stream)
    barrier()
    ms = e0.elapsed_time(e1)
    launches = ctx.stats()["kernel_launches"] - l0
    st = ctx.stats()
    # ---- end to end through the C ABI with host buffers ----
    step_e2e()
    barrier()
    t0 = time.perf_counter()
    f0 = torch.cuda.Event(enable_timing=True); f1 = torch.cuda.Event(enable_timing=True)
    f0.record(stream)
    for _ in range(args.steps):
        step_e2e()
    f1.record(stream)
    barrier()
    ms_e2e = f0.elapsed_time(f1)
    wall_e2e = (time.perf_counter() - t0) * 1e3
    stop.set(); th.join()
    M = last["M"]
    h2d = int(h_blob.numel() + n * 8 + n * 4) if world == 1 else int(last["h2d"])   # per rank
    d2h = int(last["d2h_seq"] + (st["n_seqs"] // world + 1) * 8 + M * (4 + 1 + 4 + 4 + 8 + 4 + 16))
    if world > 1:
        t = torch.tensor([ms, ms_e2e, wall_e2e], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, ms_e2e, wall_e2e = t.tolist()
    ms_step = ms / args.steps
    e2e_ms_step = max(ms_e2e, wall_e2e) / args.steps
    value = n / (ms_step * 1e-3) / 1e6
    e2e_val = n / (e2e_ms_step * 1e-3) / 1e6
    peak, peak_src = measured_peak()
    line = {"metric": "Mreads/s kmermatch+rescore+correction", "value": value, "unit": "Mreads/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u8/int64 (+f64 likelihoods)", "data": data_note,
            "config": {"workload": workload, "l2": "inputs larger than L2 (k-mer records of one pass: GBs)",
                       "parallelism": "single GPU" if world == 1 else "k-mer records owned by hash range, %d ranks, %s" % (
                           world, "partition kernel writes into peer memory over NVLink (CUDA IPC)" if (sharded.exchange if sharded else args.exchange) == "p2p" else "NCCL all-to-all")},
            "e2e": {"value": e2e_val, "unit": "Mreads/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "ms_per_step": e2e_ms_step},
            "gpu_launches": int(launches), "clocks": clocks_summary(samples)}
    if world == 1:
        # roofline of the dominant kernel: one LSD pass of sort 1 reads and writes every 16-byte record once; launch times
        # from CUDA events recorded around every onesweep launch on the library's stream (last timed step)
        launches1 = max(1, st["n_sort1_pass_launches"])
        per_launch_bytes = 2 * 16 * st["sort1_records"]
        per_launch_ms = st["ms_sort1_passes"] / launches1
        ach = per_launch_bytes / (per_launch_ms * 1e-3) / 1e9 if per_launch_ms > 0 else 0.0
        whole = alg_bytes(st, 20) / (ms_step * 1e-3) / 1e9
        line["roofline"] = {"bound": "hbm", "kernel": "cdgsort::onesweep_kernel<256,16,true> (sort 1, %d launches per step, %.3f ms each, %d records x 32 B)" % (
                                launches1, per_launch_ms, st["sort1_records"]),
                            "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak, "traffic": ONESWEEP_TRAFFIC_RATIO * per_launch_bytes,
                            "traffic_note": "dram read+write per launch = %.3f x algorithmic bytes, ratio from the ncu --set full capture in profiles/ (2 M-read run)" % ONESWEEP_TRAFFIC_RATIO,
                            "peak_source": peak_src, "whole_iteration_alg_GBps": whole, "whole_iteration_frac": whole / peak}
        line["stage_ms"] = {k: st[k] for k in st if k.startswith("ms_")}
        line["counters"] = {k: st[k] for k in ("n_seqs", "n_bases", "n_records", "n_grouped", "n_hits", "n_alns", "sort1_passes", "sort2_passes")}
    else:
        # same dominant kernel per rank: the sort-1 passes over the records this rank received (rank 0's launches of the last step)
        launches1 = max(1, st["n_sort1_pass_launches"])
        per_launch_bytes = 2 * 16 * st["sort1_records"]
        per_launch_ms = st["ms_sort1_passes"] / launches1
        ach = per_launch_bytes / (per_launch_ms * 1e-3) / 1e9 if per_launch_ms > 0 else None
        line["roofline"] = {"bound": "hbm", "kernel": "cdgsort::onesweep_kernel<256,16,true> on rank 0 (sort 1 of the received records, %d launches per step, %.3f ms each, "
                                      "%d records x 32 B)" % (launches1, per_launch_ms, st["sort1_records"]),
                            "achieved": ach, "peak": peak, "unit": "GB/s", "frac": (ach / peak) if ach else None,
                            "traffic": ONESWEEP_TRAFFIC_RATIO * per_launch_bytes if ach else None, "peak_source": peak_src}
        line["phase_ms_rank0"] = {k: round(v, 2) for k, v in last.get("phase_ms", {}).items()}
        line["exchange"] = {"records_sent_per_rank": last.get("sent"), "grouped_sent_per_rank": last.get("sent2"),
                            "nvlink_bytes_per_rank_per_step": int(16 * (last.get("sent") or 0) + 8 * (last.get("sent2") or 0))}
    if rank == 0 and not args.no_cpu_baseline and world == 1:
        try:
            r = reference_arm(argparse.Namespace(steps=1, warmup=0), args.cpu_sample)
            line["cpu_baseline"] = {"value": r["value"], "unit": "Mreads/s", "cores": r["cores"], "kind": r["kind"], "sample": r["sample"], "note": r["note"]}
        except Exception as ex:   # the baseline is reported, never required for the GPU numbers
            line["cpu_baseline"] = {"value": None, "unit": "Mreads/s", "cores": os.cpu_count(), "kind": "reference", "sample": "failed: %s" % str(ex)[:200]}
    if rank == 0:
        print(json.dumps(line))
    if sharded is not None:
        sharded.release()
    ctx.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
