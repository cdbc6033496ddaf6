#!/usr/bin/env python3
"""TEST INFRASTRUCTURE — builds the UNMODIFIED reference binary into oracle/_ref/.

Recipe (no cmake, no reference build system): g++/gcc are run directly on the reference
sources where they lie under /root/reference (list: oracle/ref_sources.txt, paths only),
objects + the binary go to oracle/_ref/ only.  The reference embeds a few data files
(substitution matrices, workflow shell scripts) as `xxd -i` byte arrays
(lib/mmseqs/cmake/MMseqsResourceCompiler.cmake); this script writes the equivalent
headers into oracle/_ref/generated/ (byte arrays of the reference's own data files, not
source code).  Nothing from /root/reference is copied into the tracked tree.

The binary is used (a) to validate oracle/ restatements and regenerate tests/golden/,
(b) as the CPU baseline leg of bench.py (`cpu_baseline.kind == "reference"`,
`bench.py --impl reference`).  It is never on the product path.

Usage: python oracle/build_ref.py [--ref /root/reference] [--jobs N] [--march x86-64-v3]
       python oracle/build_ref.py --gpu-stages      (afterwards: the same reference linked with carpedeam_b200/binding/stages_gpu.cpp)
"""
import argparse
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref")

INC_SUBDIRS = [
    "lib", "src/commons", "lib/mmseqs/lib/zstd/lib", "lib/mmseqs/lib/tinyexpr", "lib/mmseqs/lib/microtar",
    "lib/mmseqs/lib/simde", "lib/mmseqs/lib", "lib/mmseqs/lib/simd", "lib/mmseqs/lib/gzstream",
    "lib/mmseqs/lib/alp", "lib/mmseqs/lib/cacode", "lib/mmseqs/lib/ksw2", "lib/mmseqs/lib/xxhash",
    "lib/mmseqs/lib/ips4o", "lib/mmseqs/src/alignment", "lib/mmseqs/src/clustering", "lib/mmseqs/src/commons",
    "lib/mmseqs/src/multihit", "lib/mmseqs/src/prefiltering", "lib/mmseqs/src/linclust",
    "lib/mmseqs/src/taxonomy", "lib/mmseqs/src/util", "lib/mmseqs/src", "lib/mmseqs/lib/zstd/lib/common",
    "src", "src/assembler", "src/workflow", "src/util", "lib/flash", "lib/kerasify", "lib/libgab",
]
DEFS = ["-DENABLE_IPS4O=1", "-DHAVE_POSIX_FADVISE=1", "-DHAVE_POSIX_MADVISE=1", "-DHAVE_ZLIB=1", "-DOPENMP=1",
        "-DGIT_SHA1=oracle_ref", "-DNDEBUG"]


def xxd_header(path, out_dir):
    """Equivalent of `xxd -i NAME | sed 's/unsigned char/static const unsigned char/'`."""
    name = os.path.basename(path)
    var = name.replace("/", "_").replace(".", "_").replace("-", "_")
    data = open(path, "rb").read()
    lines = ["static const unsigned char %s[] = {" % var]
    for i in range(0, len(data), 12):
        lines.append("\t" + " ".join("0x%02x," % b for b in data[i:i + 12]))
    lines.append("};")
    lines.append("unsigned int %s_len = %d;" % (var, len(data)))
    with open(os.path.join(out_dir, name + ".h"), "w") as f:
        f.write("\n".join(lines) + "\n")


def generate_headers(ref, gen):
    os.makedirs(gen, exist_ok=True)
    roots = [os.path.join(ref, "data"), os.path.join(ref, "lib/mmseqs/data")]
    for root in roots:
        for dirpath, _dirs, files in os.walk(root):
            for fn in files:
                if fn.endswith((".sh", ".out", ".lib", ".model", ".html", ".zst")):
                    xxd_header(os.path.join(dirpath, fn), gen)


GPU_ENTRY_POINTS = {   # object -> (mangled entry point, name it gets so that carpedeam_b200/binding/stages_gpu.cpp can take its place)
    "lib/mmseqs/src/linclust/kmermatcher.cpp": ("_Z11kmermatcheriPPKcRK7Command", "_Z15kmermatcher_cpuiPPKcRK7Command"),
    "lib/mmseqs/src/alignment/rescorediagonal.cpp": ("_Z15rescorediagonaliPPKcRK7Command", "_Z19rescorediagonal_cpuiPPKcRK7Command"),
    "src/assembler/correction.cpp": ("_Z10correctioniPPKcRK7Command", "_Z14correction_cpuiPPKcRK7Command"),
    "src/assembler/cyclecheck.cpp": ("_Z10cyclecheckiPPKcRK7Command", "_Z14cyclecheck_cpuiPPKcRK7Command"),
}


def build_gpu_stages(ref, march):
    """oracle/_ref/carpedeam_ref_gpu: the reference with the in-tree binding (carpedeam_b200/binding/stages_gpu.cpp, compiled against the
    reference's headers) in place of the entry points of kmermatcher, rescorediagonal, ancient_correction and cyclecheck.  No reference source is changed:
    the three original entry points are renamed in COPIES of their objects (objcopy --redefine-sym) and stay reachable as *_cpu for calls
    outside the GPU path.  Needs oracle/_ref/obj (a previous run of this script) and the built libcarpedeam_gpu.so / libcarpedeam_host.so."""
    root = os.path.dirname(HERE)
    pkg = os.path.join(root, "carpedeam_b200")
    obj, obj_gpu = os.path.join(OUT, "obj"), os.path.join(OUT, "obj_gpu")
    os.makedirs(obj_gpu, exist_ok=True)
    srcs = [l.strip() for l in open(os.path.join(HERE, "ref_sources.txt")) if l.strip()]
    objs = []
    for s in srcs:
        o = os.path.join(obj, s.replace("/", "__") + ".o")
        if not os.path.exists(o):
            print("missing", o, "- run oracle/build_ref.py first")
            return 1
        if s in GPU_ENTRY_POINTS:
            old, new = GPU_ENTRY_POINTS[s]
            o2 = os.path.join(obj_gpu, os.path.basename(o))
            subprocess.run(["objcopy", "--redefine-sym", "%s=%s" % (old, new), o, o2], check=True)
            o = o2
        objs.append(o)
    incs = ["-I" + os.path.join(ref, d) for d in INC_SUBDIRS] + ["-I" + os.path.join(OUT, "generated"), "-I" + os.path.join(root, "include")]
    bind_o = os.path.join(obj_gpu, "stages_gpu.o")
    cmd = ["g++", "-std=c++1y", "-fno-exceptions", "-O2", "-fsigned-char", "-march=" + march, "-w", "-fopenmp"] + DEFS + incs + \
          ["-c", os.path.join(pkg, "binding", "stages_gpu.cpp"), "-o", bind_o]
    p = subprocess.run(cmd, capture_output=True, text=True)
    if p.returncode != 0:
        print(p.stderr[-6000:], file=sys.stderr)
        return 2
    binary = os.path.join(OUT, "carpedeam_ref_gpu")
    link = ["g++", "-fopenmp", "-o", binary] + objs + [bind_o, "-L" + pkg, "-lcarpedeam_gpu", "-lcarpedeam_host", "-Wl,-rpath," + pkg,
                                                        "-latomic", "-lm", "-lz", "-lpthread"]
    p = subprocess.run(link, capture_output=True, text=True)
    if p.returncode != 0:
        print(p.stderr[-6000:], file=sys.stderr)
        return 3
    print("built", binary)
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ref", default="/root/reference")
    ap.add_argument("--jobs", type=int, default=os.cpu_count() or 4)
    ap.add_argument("--march", default="x86-64-v3")
    ap.add_argument("--force", action="store_true")
    ap.add_argument("--gpu-stages", action="store_true", help="build oracle/_ref/carpedeam_ref_gpu (the reference + the in-tree GPU binding)")
    a = ap.parse_args()
    if a.gpu_stages:
        if not os.path.isdir(a.ref):
            print("reference tree %s absent" % a.ref)
            return 1
        return build_gpu_stages(a.ref, a.march)
    binary = os.path.join(OUT, "carpedeam_ref")
    if os.path.exists(binary) and not a.force:
        print("oracle/_ref/carpedeam_ref present")
        return 0
    if not os.path.isdir(a.ref):
        print("reference tree %s absent: cannot build oracle/_ref (prebuilt binary expected)" % a.ref)
        return 1
    gen = os.path.join(OUT, "generated")
    obj = os.path.join(OUT, "obj")
    os.makedirs(obj, exist_ok=True)
    generate_headers(a.ref, gen)
    incs = ["-I" + os.path.join(a.ref, d) for d in INC_SUBDIRS] + ["-I" + gen]
    common = ["-O3", "-fsigned-char", "-march=" + a.march, "-w", "-fopenmp"] + DEFS + incs
    srcs = [l.strip() for l in open(os.path.join(HERE, "ref_sources.txt")) if l.strip()]
    jobs = []
    for s in srcs:
        o = os.path.join(obj, s.replace("/", "__") + ".o")
        if "lib/zstd/" in s:
            # zstd carries its own xxhash.h: keep the other trees off its include path
            zi = ["-I" + os.path.join(a.ref, "lib/mmseqs/lib/zstd/lib"),
                  "-I" + os.path.join(a.ref, "lib/mmseqs/lib/zstd/lib/common")]
            cmd = ["gcc", "-O3", "-fsigned-char", "-march=" + a.march, "-w", "-DNDEBUG"] + zi + \
                  ["-c", os.path.join(a.ref, s), "-o", o]
        elif s.endswith(".c"):
            cmd = ["gcc"] + common + ["-c", os.path.join(a.ref, s), "-o", o]
        else:
            cmd = ["g++", "-std=c++1y", "-fno-exceptions"] + common + ["-c", os.path.join(a.ref, s), "-o", o]
        if "lib/alp" in s or "lib/kerasify" in s or "lib/flash" in s or "lib/cacode" in s:
            cmd = [c for c in cmd if c != "-fno-exceptions"]
        jobs.append((o, cmd))

    def run(job):
        o, cmd = job
        if os.path.exists(o) and not a.force:
            return 0, ""
        p = subprocess.run(cmd, capture_output=True, text=True)
        return p.returncode, p.stderr[-2000:]

    failed = 0
    with ThreadPoolExecutor(a.jobs) as ex:
        for (o, _), (rc, err) in zip(jobs, ex.map(run, jobs)):
            if rc != 0:
                failed += 1
                print("FAILED", o, err, file=sys.stderr)
    if failed:
        return 2
    link = ["g++", "-fopenmp", "-o", binary] + [o for o, _ in jobs] + ["-latomic", "-lm", "-lz", "-lpthread"]
    p = subprocess.run(link, capture_output=True, text=True)
    if p.returncode != 0:
        print(p.stderr[-4000:], file=sys.stderr)
        return 3
    print("built", binary)
    return 0


if __name__ == "__main__":
    sys.exit(main())
