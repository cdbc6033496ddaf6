#!/usr/bin/env python3
"""TEST INFRASTRUCTURE — regenerates the golden DBs of SURVEY §8(f) N4 (linclust's Hamming pre-clustering step):
tests/golden/cfg1_lin_pref_9.gz (input), cfg1_lin_rescore_9.gz and cfg1_lin_rescore_wrapped_9.gz (outputs).

Runs the UNMODIFIED reference binary (oracle/_ref/carpedeam_ref) on the golden sequence set correction_9 of config 1 with the argv
`carpedeam linclust` generates for a nucleotide DB (lib/mmseqs/data/workflow/linclust.sh:24-30, lib/mmseqs/src/workflow/Linclust.cpp:96-112;
captured from a linclust run of the same binary): kmermatcher (-k 0 -> 17, --kmer-per-seq 21, -c 0.8, --ignore-multi-kmer 0) and
rescorediagonal --rescore-mode 0 (HAMMING_PAR: -c 0.8 --cov-mode 0 --min-seq-id 0.9 --seq-id-mode 0), once with --wrapped-scoring 0 and
once with --wrapped-scoring 1.  Only the entries of the first N_KEEP queries are stored (the rescoring is per prefilter line)."""
import os
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from make_golden import canonical, load_db  # noqa: E402
from make_merge_kat import gz_write  # noqa: E402
from support import GOLD, load_golden, write_db  # noqa: E402

N_KEEP = 6000
KM = ("--sub-mat nucl:nucleotide.out,aa:blosum62.out --alph-size nucl:5,aa:13 --min-seq-id 0.9 --kmer-per-seq 21 --spaced-kmer-mode 0 "
      "--kmer-per-seq-scale nucl:0.200,aa:0.000 --adjust-kmer-len 0 --mask 0 --mask-lower-case 0 --cov-mode 0 -k 0 -c 0.8 --max-seq-len 65535 "
      "--hash-shift 67 --split-memory-limit 0 --include-only-extendable 0 --ignore-multi-kmer 0 --threads 4 --compressed 0 -v 1").split()
RS = ("--sub-mat nucl:nucleotide.out,aa:blosum62.out --rescore-mode 0 --filter-hits 0 -e 0.001 -c 0.8 -a 0 --cov-mode 0 --min-seq-id 0.9 "
      "--min-aln-len 0 --seq-id-mode 0 --add-self-matches 0 --sort-results 0 --db-load-mode 0 --threads 4 --compressed 0 -v 1").split()


def main():
    binary = os.path.join(HERE, "_ref", "carpedeam_ref")
    if not os.path.exists(binary):
        sys.exit("build oracle/_ref first: python oracle/build_ref.py")
    work = tempfile.mkdtemp(prefix="cdham")
    keys, payloads, exts = load_golden("correction_9")
    db = os.path.join(work, "seq")
    write_db(db, keys, payloads, exts, 1)
    subprocess.run([binary, "kmermatcher", db, db + "_pref"] + KM, check=True)
    subprocess.run([binary, "rescorediagonal", db, db, db + "_pref", db + "_r0"] + RS + ["--wrapped-scoring", "0"], check=True)
    subprocess.run([binary, "rescorediagonal", db, db, db + "_pref", db + "_r1"] + RS + ["--wrapped-scoring", "1"], check=True,
                   stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)   # one warning per skipped target
    # the k-mer matcher of the assembler's own final linclust call (-k 20 --kmer-per-seq 200 --ignore-multi-kmer 1, data/nuclassemble.sh -> linclust) with
    # its coverage filter (-c 0.99 --cov-mode 1) and with plain linclust's (-c 0.8 --cov-mode 0): assignGroup's canBeCovered (kmermatcher.cpp:526-529)
    km2 = [x for x in KM]
    for flag, val in (("--kmer-per-seq", "200"), ("-k", "20"), ("--ignore-multi-kmer", "1")):
        km2[km2.index(flag) + 1] = val
    for tag, c, mode in (("c099m1", "0.99", "1"), ("c08m0", "0.8", "0")):
        a = [x for x in km2]
        a[a.index("-c") + 1] = c; a[a.index("--cov-mode") + 1] = mode
        subprocess.run([binary, "kmermatcher", db, db + "_pref_" + tag] + a, check=True)
    keep = set(sorted(keys)[:N_KEEP])
    for tag in ("c099m1", "c08m0"):
        d = {k: v for k, v in load_db(db + "_pref_" + tag).items() if k in keep}
        gz_write(os.path.join(GOLD, "cfg1_lin_pref_%s_9.gz" % tag), canonical(d))
        print("lin_pref_" + tag, len(d), "entries,", sum(v[0].count(b"\n") for v in d.values()), "lines")
    for name, path in (("lin_pref_9", db + "_pref"), ("lin_rescore_9", db + "_r0"), ("lin_rescore_wrapped_9", db + "_r1")):
        d = {k: v for k, v in load_db(path).items() if k in keep}
        gz_write(os.path.join(GOLD, "cfg1_%s.gz" % name), canonical(d))
        print(name, len(d), "entries,", sum(v[0].count(b"\n") for v in d.values()), "lines, dbtype",
              int.from_bytes(open(path + ".dbtype", "rb").read(), "little"))


if __name__ == "__main__":
    main()
