#!/usr/bin/env python3
"""TEST INFRASTRUCTURE — regenerates tests/golden/pairlik_kat.txt.

Compiles oracle/ref_pairlik.cpp against the UNMODIFIED reference objects that oracle/build_ref.py left in
oracle/_ref/obj (every object except the one holding main()), runs it on the shipped dhigh profiles and
stores the known answers of the reference's calcLikelihood / calcLikelihoodConsensus.
Needs /root/reference (headers) and oracle/_ref/obj.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from build_ref import DEFS, INC_SUBDIRS  # noqa: E402

REF = "/root/reference"
GOLD = os.path.join(os.path.dirname(HERE), "tests", "golden")


def main():
    # same link order as build_ref.py (the reference has duplicate inline definitions whose winner depends on it)
    srcs = [l.strip() for l in open(os.path.join(HERE, "ref_sources.txt")) if l.strip()]
    objs = [os.path.join(HERE, "_ref", "obj", s.replace("/", "__") + ".o") for s in srcs if not s.endswith("commons/Application.cpp")]
    if not all(os.path.exists(o) for o in objs):
        sys.exit("build oracle/_ref first: python oracle/build_ref.py")
    exe = os.path.join(HERE, "_ref", "ref_pairlik")
    incs = ["-I" + os.path.join(REF, d) for d in INC_SUBDIRS] + ["-I" + os.path.join(HERE, "_ref", "generated")]
    cmd = ["g++", "-std=c++1y", "-O2", "-fsigned-char", "-march=x86-64-v3", "-w", "-fopenmp"] + DEFS + incs + \
          objs + [os.path.join(HERE, "ref_pairlik.cpp")] + ["-o", exe, "-latomic", "-lm", "-lz", "-lpthread"]
    subprocess.run(cmd, check=True)
    out = subprocess.run([exe, os.path.join(GOLD, "dhigh5p.prof"), os.path.join(GOLD, "dhigh3p.prof")], check=True,
                         capture_output=True, text=True).stdout
    open(os.path.join(GOLD, "pairlik_kat.txt"), "w").write(out)
    print("wrote %d known answers" % len(out.splitlines()))


if __name__ == "__main__":
    main()
