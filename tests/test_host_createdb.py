"""SURVEY §8(f) N4, host side: createdb (FASTA / FASTQ [.gz] -> the resident sequence set) against the reference's own output, and the
host library's export list against include/carpedeam_host.h."""
import ctypes
import gzip
import os
import re
import subprocess

import numpy as np
import pytest

from support import GOLD, ROOT, load_golden, read_db

REF_BIN = os.path.join(ROOT, "oracle", "_ref", "carpedeam_ref")


def test_host_library_exports_every_declared_symbol():
    import carpedeam_b200 as cd
    cd.build()
    hdr = open(os.path.join(ROOT, "include", "carpedeam_host.h")).read()
    names = sorted(set(re.findall(r"\b(cdh_[a-z0-9_]+)\s*\(", hdr)))
    assert len(names) >= 15
    H = ctypes.CDLL(cd.HOST_LIB_PATH)
    for n in names:
        assert hasattr(H, n), n


def test_createdb_reproduces_the_reference_sequence_db():
    """the example reads (tests/golden/test_data.fq.gz = example/test_data.fq.gz) through cdh_createdb: keys, order (--shuffle 1: 32 round-robin
    splits, concatenated and renumbered) and entry bytes of the nucl_reads DB the reference's createdb wrote (tests/golden/cfg1_nucl_reads.gz)"""
    import carpedeam_b200 as cd
    data, off, lens, headers = cd.createdb([os.path.join(GOLD, "test_data.fq.gz")])
    keys, pay, ext = load_golden("nucl_reads")
    assert len(lens) == len(keys) == 14996 and keys == list(range(len(keys)))
    for i, p in enumerate(pay):
        o = int(off[i])
        assert data[o:o + int(lens[i]) + 1].tobytes() == p, i
    assert int(lens.sum()) == 766712   # config 1 size
    # the packed view cdg_upload_packed takes
    words, nmask, has_n = cd.pack_seqs(data, off, lens)
    assert len(words) == int(((lens.astype(np.int64) + 31) // 32).sum()) and not has_n


@pytest.mark.skipif(not os.path.exists(REF_BIN), reason="oracle/_ref/carpedeam_ref not built")
@pytest.mark.parametrize("shuffle", [1, 0])
def test_createdb_files_equal_the_reference_binarys(tmp_path, shuffle):
    """FASTA with multi-line records, comments, lower case and N, plus a FASTQ with multi-line sequence and '@' inside the qualities, in two
    input files: sequence DB and header DB entries per key equal what `carpedeam createdb` writes (both shuffle modes), .source equal"""
    import carpedeam_b200 as cd
    rng = np.random.default_rng(5)
    fa = tmp_path / "a.fasta"
    with open(fa, "w") as f:
        for i in range(150):
            s = "".join(rng.choice(list("ACGTNacgt"), int(rng.integers(1, 400))))
            name = ("seq%d", "sp|P%05d|NAME_X", "gi|%d|ref|NC_1.1|", "consensus_tr|Q%d|Y", "gnl|db|id%d", "uc%d", "lcl|%d")[i % 7] % i
            f.write(">%s some comment %d\n" % (name, i) if i % 3 else ">%s\n" % name)
            for j in range(0, len(s), 60):
                f.write(s[j:j + 60] + "\n")
    fq = tmp_path / "b.fq.gz"
    with gzip.open(fq, "wt") as f:
        for i in range(90):
            s = "".join(rng.choice(list("ACGT"), int(rng.integers(30, 120))))
            q = "".join(rng.choice(list("@+>I5#"), len(s)))
            if i % 4 == 0:
                f.write("@r%d/1 x\n%s\n%s\n+\n%s\n%s\n" % (i, s[:20], s[20:], q[:33], q[33:]))
            else:
                f.write("@r%d\n%s\n+r%d\n%s\n" % (i, s, i, q))
    mine = str(tmp_path / "mine")
    data, off, lens, headers = cd.createdb([str(fa), str(fq)], shuffle=bool(shuffle), out_db=mine)
    ref = str(tmp_path / "ref")
    subprocess.run([REF_BIN, "createdb", str(fa), str(fq), ref, "--shuffle", str(shuffle), "-v", "1"], check=True)
    for suffix in ("", "_h"):
        a, b = read_db(mine + suffix), read_db(ref + suffix)
        assert set(a) == set(b) == set(range(240))
        for k in a:
            assert a[k][0] == b[k][0], (suffix, k, a[k][0][:80], b[k][0][:80])
        assert open(mine + suffix + ".dbtype", "rb").read() == open(ref + suffix + ".dbtype", "rb").read()
    assert open(mine + ".source").read() == open(ref + ".source").read()
    assert open(mine + ".lookup").read() == open(ref + ".lookup").read()   # identifiers by Util::parseFastaHeader (sp|, gi|, gnl|, consensus_ ...)
    assert len(lens) == 240


def test_stage_binary_createdb(tmp_path):
    """carpedeam_gpu createdb (host only, no device needed) writes the reference's nucl_reads DB"""
    import carpedeam_b200 as cd
    cd.build()
    out = str(tmp_path / "nucl_reads")
    subprocess.run([cd.STAGE_BIN, "createdb", os.path.join(GOLD, "test_data.fq.gz"), out, "--dbtype", "0", "--shuffle", "1", "--createdb-mode", "1",
                    "--write-lookup", "0", "--id-offset", "0", "--compressed", "0", "-v", "3"], check=True)
    keys, pay, ext = load_golden("nucl_reads")
    mine = read_db(out)
    assert sorted(mine) == keys and all(mine[k] == (p, 0) for k, p in zip(keys, pay))
    assert open(out + ".dbtype", "rb").read() == (1).to_bytes(4, "little")
