#!/usr/bin/env python3
"""TEST INFRASTRUCTURE — regenerates tests/golden/ from the UNMODIFIED reference binary.

Runs `oracle/_ref/carpedeam_ref ancient_assemble example/test_data.fq.gz ... --ancient-damage example/dhigh`
(BASELINE.json configs[0]) and stores, for selected iterations, the stage
databases the hot path reads and writes, in canonical per-key form (per-thread data-file order
is not stable in the reference, the per-key content is):

    for keys ascending:  b"%d\t%d\n" % (key, ext) + payload_without_final_NUL + b"\0"

gzip-compressed as tests/golden/cfg1_<name>.gz, plus the SHA-256 list MANIFEST.sha256 and the
damage profiles the run used.  Also dumps ALP known-answer vectors (see oracle/ref_probe.cpp).

Needs /root/reference (inputs) and oracle/_ref/carpedeam_ref (python oracle/build_ref.py).
"""
import glob
import gzip
import hashlib
import os
import shutil
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
GOLD = os.path.join(ROOT, "tests", "golden")
REF = "/root/reference"


def load_db(path):
    """MMseqs2 DB -> {key: (payload without trailing NUL, ext)} (DBReader.cpp:792-817 index format)."""
    files = sorted(glob.glob(path + ".[0-9]*"), key=lambda s: int(s.rsplit(".", 1)[1])) or [path]
    data = b"".join(open(f, "rb").read() for f in files)
    out = {}
    for line in open(path + ".index"):
        f = line.split()
        key, off, ln = int(f[0]), int(f[1]), int(f[2])
        ext = int(f[3]) if len(f) > 3 else 0
        out[key] = (data[off:off + ln - 1], ext)
    return out


def canonical(db):
    parts = []
    for key in sorted(db):
        payload, ext = db[key]
        parts.append(b"%d\t%d\n" % (key, ext) + payload + b"\0")
    return b"".join(parts)


def main():
    binary = os.path.join(HERE, "_ref", "carpedeam_ref")
    if not os.path.exists(binary):
        sys.exit("build oracle/_ref first: python oracle/build_ref.py")
    os.makedirs(GOLD, exist_ok=True)
    work = tempfile.mkdtemp(prefix="cdgold")
    subprocess.run([binary, "ancient_assemble", REF + "/example/test_data.fq.gz", work + "/out.fasta", work + "/tmp",
                    "--ancient-damage", REF + "/example/dhigh", "--threads", "8",
                    "--remove-tmp-files", "0", "--delete-tmp-inc", "0"], check=True,
                   stdout=open(work + "/run.log", "w"), stderr=subprocess.STDOUT)
    top = os.path.realpath(work + "/tmp/latest")
    it = os.path.realpath(top + "/nuclassembly_tmp/latest")
    want = {
        "nucl_reads": top + "/nucl_reads",
        "pref_0": it + "/pref_0", "aln_0": it + "/aln_0", "correction_0": it + "/correction_0",
        "assembly_reads_0": it + "/assembly_reads_0",
        "pref_1": it + "/pref_1", "aln_1": it + "/aln_1", "correction_1": it + "/correction_1",
        "assembly_reads_4": it + "/assembly_reads_4",
        "pref_asm_5": it + "/pref_asm_5", "aln_asm_5": it + "/aln_asm_5", "correction_5": it + "/correction_5",
        "assembly_contigs_8_noneCycle": it + "/assembly_contigs_8_noneCycle",
        "pref_asm_9": it + "/pref_asm_9", "aln_asm_9": it + "/aln_asm_9", "correction_9": it + "/correction_9",
    }
    manifest = []
    for name, path in want.items():
        if not os.path.exists(path + ".index"):
            print("missing", path)
            continue
        blob = canonical(load_db(path))
        manifest.append("%s  cfg1_%s" % (hashlib.sha256(blob).hexdigest(), name))
        with open(os.path.join(GOLD, "cfg1_%s.gz" % name), "wb") as raw:
            with gzip.GzipFile(fileobj=raw, mode="wb", mtime=0, compresslevel=9) as f:
                f.write(blob)
    manifest.append("%s  out.fasta" % hashlib.sha256(open(work + "/out.fasta", "rb").read()).hexdigest())
    open(os.path.join(GOLD, "MANIFEST.sha256"), "w").write("\n".join(manifest) + "\n")
    for p in ("dhigh5p.prof", "dhigh3p.prof", "test_data.fq.gz"):     # the run's inputs (example data), for the workflow test on the GPU box
        shutil.copy(REF + "/example/" + p, os.path.join(GOLD, p))
    # which kmermatcher input belongs to which pref: see the stage argv lines of run.log
    log = [l for l in open(work + "/run.log") if l.startswith(("kmermatcher ", "rescorediagonal ", "ancient_correction "))]
    open(os.path.join(GOLD, "cfg1_stage_argv.log"), "w").write("".join(l.replace(work, "$WORK") for l in log))
    print("\n".join(manifest))
    shutil.rmtree(work)


if __name__ == "__main__":
    main()
