#!/usr/bin/env python3
"""TEST INFRASTRUCTURE — regenerates tests/golden/extension_kat_{safe,unsafe}.txt.gz.

Compiles oracle/ref_extension.cpp against the UNMODIFIED reference objects (oracle/_ref/obj), writes the golden correction_1 /
aln_1 databases of config 1 back to disk and stores what the reference's consensusCaller, updateSeqIdConsensusReads and
r_s_pair return for every query with at least two alignment lines.  Needs /root/reference (headers) and oracle/_ref/obj."""
import gzip
import os
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from build_ref import DEFS, INC_SUBDIRS  # noqa: E402
from support import GOLD, load_golden, write_db  # noqa: E402

REF = "/root/reference"


def main():
    srcs = [l.strip() for l in open(os.path.join(HERE, "ref_sources.txt")) if l.strip()]
    objs = [os.path.join(HERE, "_ref", "obj", s.replace("/", "__") + ".o") for s in srcs if not s.endswith("commons/Application.cpp")]
    if not all(os.path.exists(o) for o in objs):
        sys.exit("build oracle/_ref first: python oracle/build_ref.py")
    exe = os.path.join(HERE, "_ref", "ref_extension")
    incs = ["-I" + os.path.join(REF, d) for d in INC_SUBDIRS] + ["-I" + os.path.join(HERE, "_ref", "generated")]
    cmd = ["g++", "-std=c++1y", "-O2", "-fsigned-char", "-march=x86-64-v3", "-w", "-fopenmp"] + DEFS + incs + \
          objs + [os.path.join(HERE, "ref_extension.cpp")] + ["-o", exe, "-latomic", "-lm", "-lz", "-lpthread"]
    subprocess.run(cmd, check=True)
    work = tempfile.mkdtemp(prefix="cdext")
    for name, dbtype in (("correction_1", 1), ("aln_1", 5)):
        keys, payloads, exts = load_golden(name)
        write_db(os.path.join(work, name), keys, payloads, exts, dbtype)
    for tag, unsafe in (("safe", "0"), ("unsafe", "1")):
        r = subprocess.run([exe, os.path.join(work, "correction_1"), os.path.join(work, "aln_1"), os.path.join(GOLD, "dhigh5p.prof"),
                            os.path.join(GOLD, "dhigh3p.prof"), unsafe], check=True, capture_output=True, text=True)
        with open(os.path.join(GOLD, "extension_kat_%s.txt.gz" % tag), "wb") as raw:
            with gzip.GzipFile(fileobj=raw, mode="wb", mtime=0, compresslevel=9) as f:
                f.write(r.stdout.encode())
        print(tag, len(r.stdout.splitlines()), "lines;", r.stderr.strip().splitlines()[-1])


if __name__ == "__main__":
    main()
