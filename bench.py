#!/usr/bin/env python3
"""bench.py — Mreads/s of one reads-only iteration (kmermatch + rescore + correction) on synthetic damaged reads.

N = 1: BASELINE.json configs[1]: 10 M merged reads, length U{35..70}, one i.i.d. uniform genome at
~10.5x coverage, dhigh damage profile, numpy default_rng(20260102).
N > 1: BASELINE.json configs[2] weak-scaled: N x 12.5 M reads, log-normal length (mean ~45, 30-150),
25 N genomes of 0.5-5 Mbp at log-normal abundance (mixed coverage), same damage, seed 20260103 — at N = 8 exactly config 3
(10^8 reads, 200 genomes).  Reads are block-partitioned over the ranks (every rank generates and holds only its block on the
host), k-mer records are hash-sharded.  The config-2 x N workload of round 1 is measured too and reported as `config2_weak`.
A "step" is one pass of the hot path over the read set.

  e2e   : THE HEADLINE (SURVEY §8(d) metric): through the C ABI with HOST buffers — 2-bit packed reads in pinned host memory
          -> cdg_upload_packed -> iteration -> compact alignment SoA (cdg_alns_compact) + corrected reads as patches
          (cdg_patches) back in pinned host memory, all copies inside the timed region
  value : same iteration with the packed read set already resident in HBM (explains e2e; also printed as value_resident)
  roofline : dominant kernel (LSD radix pass `onesweep_kernel`), algorithmic bytes / launch time measured live with CUDA events
          on the library's stream vs the measured HBM copy bandwidth in MEASURED_PEAKS.json; `traffic` = dram read+write bytes
          per launch from the committed ncu capture named in `traffic_source`
  parity_check : before timing, the path being timed (single GPU or sharded, full and compact outputs) runs the reference's
          config-1 reads (k=20) and a contig iteration (k=22, include-only-extendable) and must reproduce the golden DBs written
          by the unmodified reference (tests/golden) byte for byte; a mismatch exits non-zero
  cpu_baseline / --impl reference : the UNMODIFIED reference binary (oracle/_ref/carpedeam_ref) running its three stage
          commands on the box's host threads on a bounded sample of the same workload (same coverage, fewer reads), plus a
          --threads 1 leg and (N = 1, reference arm) one pass over the full 10 M-read set
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


def synth_block_torch(model, n_block, world, rank, device, chunk=2_000_000):
    """This rank's block of `n_block` reads of a `world`-block set, generated with torch on `device` (GPU, or CPU for the
    reference arm's sample).  model "config2": one uniform genome of 50 Mbp x world, length U{35..70}, seed 20260102;
    model "config3": 25 x world genomes of 0.5-5 Mbp (log-uniform), reads drawn with p ~ LogNormal(0,1) abundance x size
    (coverage ~1x-100x), length clip(round(LogNormal(ln 44, 0.2)), 30, 150), seed 20260103.  The genomes come from
    the shared seed (identical on every rank), the reads from seed + 1 + rank.  dhigh damage on the read strand, 0.01 interior
    C->T / G->A, 0.001 substitution errors.  Returns host numpy arrays in the MMseqs2 data-file layout."""
    import torch
    seed = 20260102 if model == "config2" else 20260103
    g = torch.Generator(device=device); g.manual_seed(seed)
    acgt = torch.tensor(list(b"ACGT"), dtype=torch.uint8, device=device)
    comp = torch.zeros(256, dtype=torch.uint8, device=device)
    for a, b in zip(b"ACGT", b"TGCA"):
        comp[a] = b
    if model == "config2":
        sizes = torch.tensor([int(50_000_000 * world)], dtype=torch.int64, device=device)
        ab = torch.ones(1, device=device, dtype=torch.float64)
        lmin, lmax = 35, 70
    else:
        ng = max(1, int(round(25 * world)))
        u = torch.rand(ng, generator=g, device=device, dtype=torch.float64)
        sizes = torch.exp(np.log(5e5) + u * (np.log(5e6) - np.log(5e5))).to(torch.int64)
        ab = torch.exp(torch.randn(ng, generator=g, device=device, dtype=torch.float64)) * sizes.to(torch.float64)
        lmin, lmax = 30, 150
    gstart = torch.zeros(len(sizes) + 1, dtype=torch.int64, device=device)
    gstart[1:] = torch.cumsum(sizes, 0)
    genome = acgt[torch.randint(0, 4, (int(gstart[-1].item()),), generator=g, device=device)]
    g.manual_seed(seed + 1 + rank)
    if model == "config2":
        lens = torch.randint(lmin, lmax + 1, (n_block,), generator=g, device=device, dtype=torch.int64)
    else:
        lens = torch.round(torch.exp(np.log(44.0) + 0.2 * torch.randn(n_block, generator=g, device=device, dtype=torch.float64))).clamp(lmin, lmax).to(torch.int64)
    off = torch.zeros(n_block, dtype=torch.int64, device=device)
    off[1:] = torch.cumsum(lens[:-1] + 2, 0)
    total = int(lens.sum().item())
    blob = torch.zeros(total + 2 * n_block, dtype=torch.uint8, device=device)
    j = torch.arange(lmax, device=device)[None, :]
    d5 = torch.tensor(DHIGH5, device=device, dtype=torch.float32); d3 = torch.tensor(DHIGH3, device=device, dtype=torch.float32)
    interior, sub_err = 0.01, 0.001
    for c0 in range(0, n_block, chunk):
        m = min(chunk, n_block - c0)
        L = lens[c0:c0 + m][:, None]
        which = torch.multinomial(ab / ab.sum(), m, replacement=True, generator=g)
        span = (sizes[which] - lmax).to(torch.float64)
        start = (gstart[which] + (torch.rand(m, generator=g, device=device, dtype=torch.float64) * span).to(torch.int64))[:, None]
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
        del r, u, e, idx, flat, row, pos, p_ct, p_ga, dd, valid
    out = dict(blob=blob.cpu().numpy(), off=off.cpu().numpy().astype(np.uint64), len=lens.cpu().numpy().astype(np.uint32), n=n_block, total=total,
               genome_bp=int(gstart[-1].item()), n_genomes=len(sizes))
    del blob, genome
    if device != "cpu" and getattr(device, "type", "cuda") != "cpu":
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




def survey_passes(n_seqs, k):
    """P1 = ceil(2k/8), P2 = ceil((2*ceil(log2 N) + 16)/8) — the pass counts the ALGORITHMIC bytes are charged with,
    whatever the implementation does (ours runs fewer: 8-byte grouped keys, 57 bits -> 7 passes at 10 M reads)."""
    id_bits = max(1, int(np.ceil(np.log2(max(2, n_seqs)))))
    return (2 * k + 7) // 8, (2 * id_bits + 16 + 7) // 8


def traffic_from_profile():
    """dram read+write bytes per launch of the sort-1 pass over its algorithmic bytes, from the committed ncu capture
    (profiles/traffic.json, written by profiles/ncu_traffic_summary.py from the named ncu csv of the benchmarked workload)."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    try:
        t = json.load(open(p))
        return float(t["onesweep_pairs"]["dram_over_algorithmic"]), "profiles/traffic.json <- " + t["source"]
    except Exception:
        return None, "no committed ncu capture found (profiles/traffic.json)"


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


def cpu_model():
    try:
        for line in open("/proc/cpuinfo"):
            if line.startswith("model name"):
                return line.split(":", 1)[1].strip()
    except Exception:
        pass
    return "unknown"


def write_fasta(s, path, chunk=1_000_000):
    """>r<9 digits>\\nBASES\\n per read, built with numpy (the DB layout already holds BASES\\n)"""
    blob, off, lens, n = s["blob"], s["off"].astype(np.int64), s["len"].astype(np.int64), s["n"]
    with open(path, "wb") as f:
        for c0 in range(0, n, chunk):
            m = min(chunk, n - c0)
            L = lens[c0:c0 + m]
            rec_off = np.zeros(m + 1, np.int64)
            rec_off[1:] = np.cumsum(L + 12)
            out = np.empty(int(rec_off[-1]), np.uint8)
            ids = np.arange(c0, c0 + m, dtype=np.int64)
            hdr = np.empty((m, 11), np.uint8)
            hdr[:, 0] = ord(">")
            for d in range(9):
                hdr[:, 9 - d] = ord("0") + (ids // 10 ** d) % 10
            hdr[:, 10] = ord("\n")
            out[(rec_off[:-1, None] + np.arange(11)[None, :]).ravel()] = hdr.ravel()
            b0 = int(off[c0]); b1 = int(off[c0 + m - 1] + L[-1] + 2)
            seg = blob[b0:b1]
            keep = seg != 0                                         # BASES + "\n", the NUL of every entry dropped
            row = np.repeat(np.arange(m, dtype=np.int64), L + 1)
            src = np.nonzero(keep)[0]
            out[src - 2 * row + (rec_off[:-1] - (off[c0:c0 + m] - b0) + 2 * np.arange(m) + 11)[row]] = seg[src]
            out.tofile(f)


class RefRunner:
    """createdb once, then time kmermatcher + rescorediagonal + ancient_correction (default argv)."""

    def __init__(self, sample):
        self.n = sample["n"]
        self.dir = tempfile.mkdtemp(prefix="cdref")
        fa = os.path.join(self.dir, "reads.fa")
        write_fasta(sample, fa)
        self.db = os.path.join(self.dir, "db")
        self._run(["createdb", fa, self.db, "--dbtype", "2", "--shuffle", "0"])
        os.remove(fa)
        self.damage = os.path.join(ROOT, "tests", "golden", "dhigh")

    def _run(self, args):
        r = subprocess.run([REF_BIN] + args, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("reference stage failed: %s\n%s" % (" ".join(args), r.stdout[-2000:] + r.stderr[-2000:]))

    def step(self, threads):
        d = self.dir
        for f in os.listdir(d):
            if f.startswith(("pref", "aln", "corr")):
                os.remove(os.path.join(d, f))
        th = str(threads)
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


def sample_of(world, n_reads):
    """bounded sample of the N-GPU workload: same read model and coverage, fewer reads"""
    if world == 1:
        genome = max(200_000, int(n_reads * 5))   # 52.5 b * n / genome = 10.5x like the full workload
        return synth_reads_fast(n_reads, genome, 20260102), \
            "%d reads (35-70 bp) over a %d bp genome: same 10.5x coverage and dhigh damage as the full workload" % (n_reads, genome)
    # config-3 model scaled down: genomes in proportion to the reads (1e8 reads : 200 genomes), so coverage and mixture stay
    frac = n_reads / 12.5e6             # "world" such that 25 * world genomes match the read count
    s = synth_block_torch("config3", n_reads, max(frac, 0.04), 0, "cpu", chunk=250_000)
    return s, "%d reads of the config-3 model (log-normal length, %d genomes, %d bp, mixed coverage, dhigh damage)" % (n_reads, s["n_genomes"], s["genome_bp"])


def reference_arm(steps, warmup, world, sample_reads, threads1_reads=0, full=None):
    """-> dict for the JSON line.  steps x the three stage commands on the sample with all host threads; optionally a --threads 1
    leg on a smaller sample and one pass over a full-size set (`full` = read-set dict)."""
    cores = os.cpu_count() or 1
    sample, sample_desc = sample_of(world, sample_reads)
    extra = {}
    if ref_available():
        rr = RefRunner(sample)
        try:
            for _ in range(warmup):
                rr.step(cores)
            times = []; stages = None
            for _ in range(steps):
                t, stages = rr.step(cores)
                times.append(t)
        finally:
            rr.close()
        kind = "reference"
        note = "kmermatcher %.2fs rescorediagonal %.2fs ancient_correction %.2fs (last step); %s; binary built -march=x86-64-v3 (it is compiled in the " \
               "build container and must run on any host of the pool, so not -march=native)" % (stages + (cpu_model(),))
        if threads1_reads:
            s1, d1 = sample_of(world, threads1_reads)
            r1 = RefRunner(s1)
            try:
                t, st1 = r1.step(1)
            finally:
                r1.close()
            extra["threads1"] = {"value": s1["n"] / t / 1e6, "unit": "Mreads/s", "cores": 1, "sample": d1,
                                 "note": "kmermatcher %.2fs rescorediagonal %.2fs ancient_correction %.2fs" % st1}
        if full is not None:
            rf = RefRunner(full)
            try:
                t, stf = rf.step(cores)
            finally:
                rf.close()
            extra["full_config"] = {"value": full["n"] / t / 1e6, "unit": "Mreads/s", "cores": cores, "reads": full["n"], "sec": t, "steps": 1,
                                    "note": "one pass over the whole benchmarked set (cold): kmermatcher %.2fs rescorediagonal %.2fs ancient_correction %.2fs" % stf}
    else:   # oracle port (scalar, 1 thread)
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        from support import Oracle, SeqSet
        small = min(sample_reads, 100_000)
        sample, sample_desc = sample_of(1, small)
        sample_desc += " (oracle port, oracle/_ref unavailable on this host)"
        o = Oracle()
        s = SeqSet.__new__(SeqSet)
        s.n = sample["n"]; s.blob = sample["blob"]; s.off = sample["off"]; s.len = sample["len"]; s.total = sample["total"]
        s.ext = np.zeros(s.n, np.uint8); s.keys = np.arange(s.n, dtype=np.uint32)
        times = []
        for _ in range(max(1, steps)):
            t0 = time.perf_counter(); o.kmermatch(s); times.append(time.perf_counter() - t0)
        cores = 1; kind = "port"; note = "kmermatch stage only (oracle port)"
    sec = float(np.mean(times))
    val = sample["n"] / sec / 1e6
    return dict(value=val, sec=sec, cores=cores, kind=kind, sample=sample_desc, note=note, n=sample["n"], extra=extra)


def cpu_baseline_obj(r):
    o = {"value": r["value"], "unit": "Mreads/s", "cores": r["cores"], "kind": r["kind"], "sample": r["sample"], "note": r["note"]}
    o.update(r["extra"])
    return o


# ---------------------------------------------------------------------------------------------------------------
def workload_text(world, reads_per_gpu):
    if world == 1:
        return "synthetic %d merged reads, 35-70 bp, dhigh damage, %d bp genome (BASELINE configs[1])" % (reads_per_gpu, reads_per_gpu * 5)
    return "synthetic %d reads, log-normal length (mean ~45 bp, 30-150), %d genomes of 0.5-5 Mbp at mixed coverage, dhigh damage, seed 20260103 " \
           "(BASELINE configs[2]%s), reads block-partitioned over %d GPUs, k-mer records hash-sharded" % (
               reads_per_gpu * world, 25 * world, "" if world == 8 and reads_per_gpu == 12_500_000 else " model, weak-scaled", world)


def config_obj(world, reads_per_gpu, exchange):
    return {"workload": workload_text(world, reads_per_gpu), "l2": "inputs larger than L2 (k-mer records of one pass: GBs)",
            "parallelism": "single GPU" if world == 1 else "k-mer records owned by hash range, %d ranks, %s" % (
                world, "partition kernel writes into peer memory over NVLink (CUDA IPC)" if exchange == "p2p" else "NCCL all-to-all")}


# ---------------------------------------------------------------------------------------------------------------
def golden_parity_check(cd, ctx, sharded, dist, rank, world):
    """The path about to be timed on the reference's own data: config-1 reads (k=20) and a contig iteration (k=22,
    include-only-extendable) against the golden DBs the unmodified reference wrote (tests/golden).  Full-fidelity outputs
    (pref hits, alignment lines, corrected bytes) and the compact outputs the e2e step uses (expanded on the host)."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from support import SeqSet, format_aln, format_pref, load_golden
    cases = [("reads k=20", "nucl_reads", "pref_0", "aln_0", "correction_0", cd.kmer_params()),
             ("contigs k=22 ext", "assembly_contigs_8_noneCycle", "pref_asm_9", "aln_asm_9", "correction_9", cd.kmer_params(k=22, include_only_extendable=1))]
    report, ok = [], True

    def corrected_equal(s, out, cpay):
        return all(out[int(s.off[i]):int(s.off[i]) + int(s.len[i])].tobytes() + b"\n" == p for i, p in enumerate(cpay))

    for name, seqdb, pref, aln, corr, kp in cases:
        s = SeqSet.from_golden(seqdb)
        pkeys, ppay, _ = load_golden(pref); akeys, apay, _ = load_golden(aln); _, cpay, _ = load_golden(corr)
        ip = cd.IterParams(kp, cd.RescoreParams(0.9, 0.001, 0), cd.CorrParams(0.9, 0.99))
        try:
            packed = cd.pack_seqs(s.blob, s.off, s.len)
        except cd.CdgError:
            packed = None                                          # letters beyond ACGTN: the byte upload is the only input form
        res = {}
        if world == 1:
            ctx.upload(s.blob, s.off, s.len, s.ext)
            hits, alns, out, _ = ctx.iteration(kmer=kp, fetch_hits=True)
            mine = format_pref(s, hits)
            res["pref"] = all(mine[k] == p for k, p in zip(pkeys, ppay))
            mine = format_aln(s, alns["line_off"], alns["tgt"], alns)
            res["aln"] = all(mine[k] == p for k, p in zip(akeys, apay))
            res["corrected"] = corrected_equal(s, out, cpay)
            if packed is not None:
                ctx.upload_packed(packed[0], packed[1], s.len, s.ext)
            ca, pa = cd.AlnsCompact(), cd.Patches()
            ctx.iteration_compact(ip, ca, pa)
            ex = cd.expand_alns(ca, s.len)
            mine = format_aln(s, ex["line_off"], ex["tgt"], ex)
            res["aln_compact"] = all(mine[k] == p for k, p in zip(akeys, apay))
            res["corrected_patches"] = corrected_equal(s, cd.apply_patches(s.blob, pa), cpay)
        else:
            ctx.upload(s.blob, s.off, s.len, s.ext)
            out_range = np.zeros(len(s.blob), np.uint8)
            r = sharded.run(kmer=kp, fetch_hits=True, fetch_alns=True, out_range=out_range)
            parts = [None] * world
            dist.all_gather_object(parts, dict(hits=r["hits"], alns=r["alns"], off=r["range_offset"], nbytes=r["range_bytes"],
                                               out=out_range[:r["range_bytes"]].copy()))
            r = sharded.run(kmer=kp, fetch_alns="compact")
            ex = cd.expand_alns(r["alns_compact"], s.len)
            pent = np.ctypeslib.as_array(r["patches"].entry, shape=(max(1, int(r["patches"].n)),))[:int(r["patches"].n)].copy()
            cparts = [None] * world
            dist.all_gather_object(cparts, dict(ex={k: v.copy() for k, v in ex.items()}, patches=pent))
            if rank == 0:
                hits = {f: np.concatenate([p["hits"][f] for p in parts]) for f in ("rep", "tgt", "score", "diag")}
                mine = format_pref(s, hits)
                res["pref"] = all(mine[k] == p for k, p in zip(pkeys, ppay))

                def cat_alns(items):
                    lo, base = [], 0
                    for a in items:
                        lo.append(a["line_off"][:-1].astype(np.int64) + base); base += int(a["line_off"][-1])
                    o = {f: np.concatenate([a[f] for a in items]) for f in items[0] if f != "line_off"}
                    o["line_off"] = np.concatenate(lo + [np.array([base], np.int64)])
                    return o
                al = cat_alns([p["alns"] for p in parts])
                mine = format_aln(s, al["line_off"], al["tgt"], al)
                res["aln"] = all(mine[k] == p for k, p in zip(akeys, apay))
                merged = s.blob.copy()
                for p in parts:
                    merged[p["off"]:p["off"] + p["nbytes"]] = p["out"]
                res["corrected"] = corrected_equal(s, merged, cpay)
                al = cat_alns([p["ex"] for p in cparts])
                mine = format_aln(s, al["line_off"], al["tgt"], al)
                res["aln_compact"] = all(mine[k] == p for k, p in zip(akeys, apay))
                patched = s.blob.copy()
                for p in cparts:
                    e = p["patches"]
                    patched[(e >> np.uint64(8)).astype(np.int64)] = (e & np.uint64(0xFF)).astype(np.uint8)
                res["corrected_patches"] = corrected_equal(s, patched, cpay)
        if rank == 0:
            good = all(res.values())
            ok = ok and good
            report.append({"case": name, "input": "packed + bytes" if packed is not None else "bytes", "ok": good, **res})
    return ok, report


# ---------------------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--reads", type=int, default=0, help="reads per GPU (default: 10 M at N = 1, 12.5 M per GPU at N > 1)")
    ap.add_argument("--cpu-sample", type=int, default=1_000_000)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-secondary", action="store_true", help="N > 1: skip the config-2 x N workload (`config2_weak`)")
    ap.add_argument("--no-parity", action="store_true")
    ap.add_argument("--exchange", default="p2p", choices=["p2p", "nccl"],
                    help="multi-GPU record exchange: sender-side partition kernel writing into peer memory over NVLink (p2p) or NCCL all-to-all")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
    per_gpu = args.reads or (10_000_000 if world == 1 else 12_500_000)
    metric = "Mreads/s kmermatch+rescore+correction"

    if args.impl == "reference":
        if rank != 0:
            return 0
        full = synth_reads_fast(per_gpu, per_gpu * 5, 20260102) if (world == 1 and ref_available()) else None
        r = reference_arm(args.steps, args.warmup, world, args.cpu_sample, threads1_reads=200_000, full=full)
        line = {"impl": "reference", "metric": metric, "value": r["value"], "unit": "Mreads/s",
                "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": r["sec"] * 1e3, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "u8/int64 (+f64/f80 likelihoods)", "data": "synthetic",
                "config": config_obj(world, per_gpu, args.exchange),
                "cpu_baseline": cpu_baseline_obj(r),
                "e2e": {"value": r["value"], "unit": "Mreads/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return 0

    import torch
    import carpedeam_b200 as cd
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: a CUDA device is required (no CPU fallback)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        from carpedeam_b200.multigpu import ShardedIteration, TorchComm
        dist.init_process_group("nccl", device_id=dev)
    ctx = cd.Context(local)
    ctx.set_damage(cd.load_damage(os.path.join(ROOT, "tests", "golden", "dhigh")))
    ip = cd.IterParams(cd.kmer_params(), cd.RescoreParams(0.9, 0.001, 0), cd.CorrParams(0.9, 0.99))
    stream = torch.cuda.ExternalStream(ctx.stream)
    sharded = ShardedIteration(ctx, TorchComm(dev), exchange=args.exchange) if world > 1 else None

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- parity first: the path that is about to be timed, on the reference's own data ----
    parity = None
    if not args.no_parity:
        ok, report = golden_parity_check(cd, ctx, sharded, dist, rank, world)
        if world > 1:
            flag = torch.tensor([1 if ok else 0], device=dev)
            dist.broadcast(flag, 0)
            ok = bool(flag.item())
        parity = {"cases": report, "ok": ok, "mode": "single GPU" if world == 1 else "sharded, %s exchange, %d ranks" % (sharded.exchange, world),
                  "against": "tests/golden (DBs written by the unmodified reference)"}
        if not ok:
            if rank == 0:
                print(json.dumps({"metric": metric, "parity_check": parity, "error": "parity check failed, nothing timed"}))
            return 1

    def measure(model, reads_per_gpu, steps, warmup, sample_clocks):
        """one workload: generate this rank's block, pack it on the host, time the resident and the end-to-end step"""
        if world == 1 and model == "config2":
            s = synth_reads_fast(reads_per_gpu, reads_per_gpu * 5, 20260102)
            data_note = "synthetic (numpy default_rng(20260102))"
        else:
            s = synth_block_torch(model, reads_per_gpu, world, rank, dev)
            data_note = "synthetic (%s read model, torch.Generator on the GPU: genomes from the shared seed, reads of block r from seed + 1 + r)" % model
        words, _, _ = cd.pack_seqs(s["blob"], s["off"], s["len"], with_mask=False)        # host-side createdb product: 2-bit words, no N in this set
        n_local = s["n"]; n = n_local * world
        h_words = torch.from_numpy(words.view(np.int64)).pin_memory()
        h_lens = torch.from_numpy(s["len"].view(np.int32)).pin_memory()
        del s, words
        alns, patches = cd.AlnsCompact(), cd.Patches()
        last = {}

        def upload():
            if world == 1:
                ctx.upload_packed_raw(h_words.data_ptr(), 0, h_lens.data_ptr(), 0, n, h_words.numel())
                return h_words.numel() * 8 + h_lens.numel() * 4
            return sharded.upload_packed_blocks(h_words, None, h_lens)

        def step_resident():
            if world == 1:
                ctx.iteration_compact(ip, None, None)
            else:
                last["phase_ms"] = sharded.run()["phase_ms"]

        def step_e2e():
            last["h2d"] = upload()
            if world == 1:
                ctx.iteration_compact(ip, alns, patches)
                a, p = alns, patches
            else:
                r = sharded.run(fetch_alns="compact")
                a, p = r["alns_compact"], r["patches"]
                last["sent"] = r["records_sent"]; last["sent2"] = r["grouped_sent"]; last["phase_ms_e2e"] = r["phase_ms"]
            last["d2h"] = int(a.n_queries * 4 + a.n_alns * (4 + 5 * a.width) + p.n * 8)
            last["n_alns"] = int(a.n_alns); last["n_patches"] = int(p.n)

        # ---- resident-input timing ----
        upload()
        for _ in range(warmup):
            step_resident()
        samples = []; stop = threading.Event(); th = None
        if sample_clocks:
            th = threading.Thread(target=clocks_sampler, args=(stop, samples, local)); th.start()
        l0 = ctx.stats()["kernel_launches"]
        barrier()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(steps):
            step_resident()
        e1.record(stream)
        barrier()
        ms = e0.elapsed_time(e1)
        launches = ctx.stats()["kernel_launches"] - l0
        st = ctx.stats()
        # ---- end to end through the C ABI with host buffers ----
        for _ in range(min(2, warmup)):
            step_e2e()
        barrier()
        t0 = time.perf_counter()
        f0 = torch.cuda.Event(enable_timing=True); f1 = torch.cuda.Event(enable_timing=True)
        f0.record(stream)
        for _ in range(steps):
            step_e2e()
        f1.record(stream)
        barrier()
        ms_e2e = f0.elapsed_time(f1)
        wall_e2e = (time.perf_counter() - t0) * 1e3
        if th is not None:
            stop.set(); th.join()
        h2d, d2h = last["h2d"], last["d2h"]
        if world > 1:
            t = torch.tensor([ms, ms_e2e, wall_e2e], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms, ms_e2e, wall_e2e = t.tolist()
            t = torch.tensor([h2d, d2h], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
            h2d, d2h = [int(x) for x in t.tolist()]
        ms_step = ms / steps
        e2e_ms_step = max(ms_e2e, wall_e2e) / steps
        return dict(n=n, ms_step=ms_step, e2e_ms_step=e2e_ms_step, value=n / (ms_step * 1e-3) / 1e6, e2e=n / (e2e_ms_step * 1e-3) / 1e6,
                    h2d=h2d, d2h=d2h, launches=int(launches), st=st, last=last, samples=samples, data_note=data_note)

    primary_model = "config2" if world == 1 else "config3"
    m = measure(primary_model, per_gpu, args.steps, args.warmup, True)
    st, last = m["st"], m["last"]
    peak, peak_src = measured_peak()
    line = {"metric": metric, "value": m["value"], "unit": "Mreads/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": m["ms_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u8/int64 (+f64 likelihoods)", "data": m["data_note"],
            "config": config_obj(world, per_gpu, sharded.exchange if sharded else args.exchange),
            "e2e": {"value": m["e2e"], "unit": "Mreads/s", "h2d_bytes_per_step": m["h2d"], "d2h_bytes_per_step": m["d2h"], "ms_per_step": m["e2e_ms_step"],
                    "what": "2-bit packed reads (u64 words + u32 lengths) in pinned host memory -> cdg_upload_packed -> iteration -> cdg_alns_compact "
                            "(kept alignments, %d B each, %d of them) + cdg_patches (%d corrected bases) in pinned host memory; bytes summed over ranks" % (
                                4 + 5 * 2, last["n_alns"], last["n_patches"])},
            "headline": "e2e (SURVEY §8(d): host to host); value = value_resident explains it",
            "value_resident": m["value"],
            "gpu_launches": m["launches"], "clocks": clocks_summary(m["samples"])}
    if parity is not None:
        line["parity_check"] = parity
    # roofline of the dominant kernel: one LSD pass of sort 1 reads and writes every 16-byte record once; launch times from CUDA
    # events recorded around every onesweep launch on the library's stream (last timed step)
    launches1 = max(1, st["n_sort1_pass_launches"])
    per_launch_bytes = 2 * 16 * st["sort1_records"]
    per_launch_ms = st["ms_sort1_passes"] / launches1
    ach = per_launch_bytes / (per_launch_ms * 1e-3) / 1e9 if per_launch_ms > 0 else None
    ratio, ratio_src = traffic_from_profile()
    line["roofline"] = {"bound": "hbm", "kernel": "cdgsort::onesweep_kernel<256,16,true>%s (sort 1, %d launches per step, %.3f ms each, %d records x 32 B)" % (
                            "" if world == 1 else " on rank 0", launches1, per_launch_ms, st["sort1_records"]),
                        "achieved": ach, "peak": peak, "unit": "GB/s", "frac": (ach / peak) if ach else None,
                        "traffic": ratio * per_launch_bytes if (ratio and ach) else None,
                        "traffic_source": ratio_src, "peak_source": peak_src}
    if world == 1:
        P1, P2 = survey_passes(st["n_seqs"], 20)
        stf = dict(st); stf["sort1_passes"], stf["sort2_passes"] = P1, P2
        whole = alg_bytes(stf, 20) / (m["ms_step"] * 1e-3) / 1e9
        own = alg_bytes(st, 20) / (m["ms_step"] * 1e-3) / 1e9
        line["roofline"].update({"whole_iteration_alg_GBps": whole, "whole_iteration_frac": whole / peak,
                                 "whole_iteration_frac_e2e": alg_bytes(stf, 20) / (m["e2e_ms_step"] * 1e-3) / 1e9 / peak,
                                 "alg_passes": {"P1": P1, "P2": P2, "note": "SURVEY §8(d) formula; the implementation runs %d + %d passes" % (
                                     st["sort1_passes"], st["sort2_passes"])},
                                 "whole_iteration_frac_own_passes": own / peak})
        line["stage_ms"] = {k: st[k] for k in st if k.startswith("ms_")}
        line["counters"] = {k: st[k] for k in ("n_seqs", "n_bases", "n_records", "n_grouped", "n_hits", "n_alns", "sort1_passes", "sort2_passes")}
    else:
        line["phase_ms_rank0"] = {k: round(v, 2) for k, v in last.get("phase_ms", {}).items()}
        line["phase_ms_rank0_e2e"] = {k: round(v, 2) for k, v in last.get("phase_ms_e2e", {}).items()}
        line["exchange"] = {"records_sent_rank0": last.get("sent"), "grouped_sent_rank0": last.get("sent2"),
                            "nvlink_bytes_rank0_per_step": int(16 * (last.get("sent") or 0) + 8 * (last.get("sent2") or 0))}
        if not args.no_secondary:
            m2 = measure("config2", 10_000_000, max(3, args.steps // 2), min(args.warmup, 3), False)
            line["config2_weak"] = {"workload": "synthetic %d merged reads, 35-70 bp, %d bp genome (BASELINE configs[1] x %d, the round-1 scaling workload)" % (
                                        m2["n"], 50_000_000 * world, world),
                                    "value": m2["value"], "ms_per_step": m2["ms_step"], "e2e": m2["e2e"], "e2e_ms_per_step": m2["e2e_ms_step"],
                                    "steps": max(3, args.steps // 2), "phase_ms_rank0": {k: round(v, 2) for k, v in m2["last"].get("phase_ms", {}).items()}}
    if rank == 0 and not args.no_cpu_baseline:
        try:
            r = reference_arm(1, 0, world, args.cpu_sample, threads1_reads=100_000)
            line["cpu_baseline"] = cpu_baseline_obj(r)
        except Exception as ex:   # the baseline is reported, never required for the GPU numbers
            line["cpu_baseline"] = {"value": None, "unit": "Mreads/s", "cores": os.cpu_count(), "kind": "reference", "sample": "failed: %s" % str(ex)[:200]}
    if world > 1:
        dist.barrier()
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
