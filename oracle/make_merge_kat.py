#!/usr/bin/env python3
"""TEST INFRASTRUCTURE — regenerates the known answers of SURVEY §8(f) N3 (scoring front half of ancient_contig_merge):
tests/golden/merge_kat_{asm5,asm9,synth}.txt.gz and the synthetic contig set tests/golden/synth_contigs{,_aln}.gz.

Compiles oracle/ref_merge.cpp against the UNMODIFIED reference (objects of oracle/_ref/obj; ancientContigsResults.cpp is compiled
from where it lies because the comparator is file-local) and runs it on
  * the golden contig iterations of config 1 (correction_5 + aln_asm_5, correction_9 + aln_asm_9),
  * a seeded synthetic set of overlapping, lightly damaged contigs (both strands, a few N), whose alignment DB is produced by the
    reference binary's own kmermatcher + rescorediagonal with the contig-iteration argv of tests/golden/cfg1_stage_argv.log,
each in --unsafe 0 and --unsafe 1 (and the synthetic set also with --min-merge-seq-id 0.97, BASELINE configs[4]).
Needs /root/reference (headers) and oracle/_ref (python oracle/build_ref.py)."""
import gzip
import os
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from build_ref import DEFS, INC_SUBDIRS  # noqa: E402
from make_golden import canonical, load_db  # noqa: E402
from support import GOLD, load_golden, write_db  # noqa: E402

REF = "/root/reference"
_COMP = np.zeros(256, np.uint8)
for _a, _b in zip(b"ACGTN", b"TGCAN"):
    _COMP[_a] = _b


def synth_contigs(seed=20260118, genome_len=40000, n=120):
    rng = np.random.default_rng(seed)
    acgt = np.frombuffer(b"ACGT", np.uint8)
    genome = acgt[rng.integers(0, 4, genome_len)]
    seqs = []
    for _ in range(n):
        L = int(rng.integers(400, 2501))
        s = int(rng.integers(0, genome_len - L))
        r = genome[s:s + L].copy()
        if rng.random() < 0.5:
            r = _COMP[r[::-1]]
        err = (0.0, 0.001, 0.004, 0.012)[int(rng.integers(0, 4))]
        u = rng.random(L)
        r[(r == ord("C")) & (u < 0.004)] = ord("T")     # residual damage: C->T / G->A pairs against the other contigs
        r[(r == ord("G")) & (u > 0.996)] = ord("A")
        e = rng.random(L) < err
        if e.any():
            r[e] = acgt[rng.integers(0, 4, int(e.sum()))]
        if rng.random() < 0.15:
            r[rng.integers(0, L, 3)] = ord("N")
        seqs.append(r.tobytes())
    return seqs


def gz_write(path, blob):
    with open(path, "wb") as raw:
        with gzip.GzipFile(fileobj=raw, mode="wb", mtime=0, compresslevel=9) as f:
            f.write(blob)


def main():
    binary = os.path.join(HERE, "_ref", "carpedeam_ref")
    srcs = [l.strip() for l in open(os.path.join(HERE, "ref_sources.txt")) if l.strip()]
    objs = [os.path.join(HERE, "_ref", "obj", s.replace("/", "__") + ".o") for s in srcs
            if not s.endswith(("commons/Application.cpp", "assembler/ancientContigsResults.cpp"))]
    if not all(os.path.exists(o) for o in objs) or not os.path.exists(binary):
        sys.exit("build oracle/_ref first: python oracle/build_ref.py")
    exe = os.path.join(HERE, "_ref", "ref_merge")
    incs = ["-I" + os.path.join(REF, d) for d in INC_SUBDIRS] + ["-I" + os.path.join(HERE, "_ref", "generated")]
    cmd = ["g++", "-std=c++1y", "-O3", "-fsigned-char", "-march=x86-64-v3", "-w", "-fopenmp"] + DEFS + incs + \
          objs + [os.path.join(HERE, "ref_merge.cpp")] + ["-o", exe, "-latomic", "-lm", "-lz", "-lpthread"]
    subprocess.run(cmd, check=True)
    work = tempfile.mkdtemp(prefix="cdmerge")
    p5, p3 = os.path.join(GOLD, "dhigh5p.prof"), os.path.join(GOLD, "dhigh3p.prof")
    sets = []
    for tag, seqname, alnname in (("asm5", "correction_5", "aln_asm_5"), ("asm9", "correction_9", "aln_asm_9")):
        for name, dbtype in ((seqname, 1), (alnname, 5)):
            keys, payloads, exts = load_golden(name)
            write_db(os.path.join(work, name), keys, payloads, exts, dbtype)
        sets.append((tag, os.path.join(work, seqname), os.path.join(work, alnname), (("0", "0.99"), ("1", "0.99"))))
    # synthetic contigs: alignments from the reference's own stages (argv: tests/golden/cfg1_stage_argv.log, contig iterations)
    seqs = synth_contigs()
    sdb = os.path.join(work, "synth")
    write_db(sdb, list(range(len(seqs))), [s + b"\n" for s in seqs], [1] * len(seqs), 1)
    km = ["--sub-mat", "nucl:nucleotide.out,aa:blosum62.out", "--alph-size", "nucl:5,aa:13", "--min-seq-id", "0.9", "--kmer-per-seq", "200",
          "--spaced-kmer-mode", "0", "--kmer-per-seq-scale", "0.200", "--adjust-kmer-len", "0", "--mask", "0", "--mask-lower-case", "0", "--cov-mode", "1",
          "-k", "22", "-c", "0", "--max-seq-len", "200000", "--hash-shift", "67", "--split-memory-limit", "0", "--include-only-extendable", "1",
          "--ignore-multi-kmer", "1", "--threads", "4", "--compressed", "0", "-v", "1"]
    rs = ["--sub-mat", "nucl:nucleotide.out,aa:blosum62.out", "--rescore-mode", "3", "--wrapped-scoring", "0", "--filter-hits", "0", "-e", "0.001", "-c", "0",
          "-a", "0", "--cov-mode", "1", "--min-seq-id", "0.9", "--min-aln-len", "0", "--seq-id-mode", "0", "--add-self-matches", "0", "--sort-results", "0",
          "--db-load-mode", "0", "--threads", "4", "--compressed", "0", "-v", "1"]
    subprocess.run([binary, "kmermatcher", sdb, sdb + "_pref"] + km, check=True)
    subprocess.run([binary, "rescorediagonal", sdb, sdb, sdb + "_pref", sdb + "_aln"] + rs, check=True)
    gz_write(os.path.join(GOLD, "synth_contigs.gz"), canonical(load_db(sdb)))
    gz_write(os.path.join(GOLD, "synth_contigs_aln.gz"), canonical(load_db(sdb + "_aln")))
    sets.append(("synth", sdb, sdb + "_aln", (("0", "0.99"), ("1", "0.99"), ("1", "0.97"))))
    for tag, seqdb, alndb, runs in sets:
        out = []
        for unsafe, thr in runs:
            r = subprocess.run([exe, seqdb, alndb, p5, p3, unsafe, thr], check=True, capture_output=True, text=True)
            out.append("H\t%s\t%s\n" % (unsafe, thr) + r.stdout)
            lines = r.stdout.splitlines()
            print(tag, "unsafe", unsafe, "thr", thr, ":", sum(l.startswith("Q") for l in lines), "queries,",
                  sum(l.startswith("A") and l.split("\t")[2] == "1" for l in lines), "merge candidates,",
                  sum(l.startswith("A") and l.split("\t")[13] == "1" for l in lines), "queued;", r.stderr.strip().splitlines()[-1])
        gz_write(os.path.join(GOLD, "merge_kat_%s.txt.gz" % tag), "".join(out).encode())


if __name__ == "__main__":
    main()
