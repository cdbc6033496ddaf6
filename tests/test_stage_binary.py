"""The drop-in boundary at process level: carpedeam_gpu <stage> reads and writes the same MMseqs2 databases as the
reference's modules.  Output DBs must equal, per key, the DBs the unmodified reference binary wrote (tests/golden)."""
import os
import subprocess

import pytest

from support import GOLD, load_golden, read_db, write_db

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("seqdb,pref,aln,corr,k,only_ext", [
    ("nucl_reads", "pref_0", "aln_0", "correction_0", 20, 0),
    ("assembly_reads_4", "pref_asm_5", "aln_asm_5", "correction_5", 22, 1),
])
def test_stage_binary_matches_reference_dbs(tmp_path, seqdb, pref, aln, corr, k, only_ext):
    import carpedeam_b200 as cd
    cd.build()
    exe = cd.STAGE_BIN
    keys, payloads, exts = load_golden(seqdb)
    d = str(tmp_path)
    write_db(d + "/seq", keys, payloads, exts, 1)
    # argv as the workflow passes it (tests/golden/cfg1_stage_argv.log)
    subprocess.run([exe, "kmermatcher", d + "/seq", d + "/pref", "--sub-mat", "nucl:nucleotide.out,aa:blosum62.out", "--alph-size", "nucl:5,aa:13",
                    "--min-seq-id", "0.9", "--kmer-per-seq", "200", "--spaced-kmer-mode", "0", "--kmer-per-seq-scale", "0.200", "--adjust-kmer-len", "0",
                    "--mask", "0", "--mask-lower-case", "0", "--cov-mode", "1", "-k", str(k), "-c", "0", "--max-seq-len", "200000", "--hash-shift", "67",
                    "--split-memory-limit", "0", "--include-only-extendable", str(only_ext), "--ignore-multi-kmer", "1", "--threads", "8",
                    "--compressed", "0", "-v", "3"], check=True)
    subprocess.run([exe, "rescorediagonal", d + "/seq", d + "/seq", d + "/pref", d + "/aln", "--sub-mat", "nucl:nucleotide.out,aa:blosum62.out",
                    "--rescore-mode", "3", "--wrapped-scoring", "0", "--filter-hits", "0", "-e", "0.001", "-c", "0", "-a", "0", "--cov-mode", "1",
                    "--min-seq-id", "0.9", "--min-aln-len", "0", "--seq-id-mode", "0", "--add-self-matches", "0", "--sort-results", "0",
                    "--db-load-mode", "0", "--threads", "8", "--compressed", "0", "-v", "3"], check=True)
    subprocess.run([exe, "ancient_correction", d + "/seq", d + "/aln", d + "/corr", "--min-seq-id", "0.9", "--max-seq-len", "200000", "--keep-target", "1",
                    "--threads", "8", "-v", "3", "--rescore-mode", "3", "--min-ryseq-id-corr-reads", "0.99", "--min-seqid-corr-reads", "0.9",
                    "--ancient-damage", os.path.join(GOLD, "dhigh"), "--unsafe", "0", "--min-cov-safe", "5"], check=True)
    for name, out in ((pref, "pref"), (aln, "aln"), (corr, "corr")):
        gk, gp, ge = load_golden(name)
        mine = read_db(d + "/" + out)
        assert sorted(mine) == gk
        bad = [kk for kk, p, e in zip(gk, gp, ge) if mine[kk] != (p, e)]
        assert not bad, (name, len(bad), mine[bad[0]][0][:200], gp[gk.index(bad[0])][:200])


def test_stage_binary_error_behaviour(tmp_path):
    import carpedeam_b200 as cd
    cd.build()
    r = subprocess.run([cd.STAGE_BIN, "kmermatcher", str(tmp_path / "nope"), str(tmp_path / "out"), "-k", "20"], capture_output=True)
    assert r.returncode == 1


HAMMING_ARGV = ["--sub-mat", "nucl:nucleotide.out,aa:blosum62.out", "--rescore-mode", "0", "--filter-hits", "0", "-e", "0.001", "-c", "0.8", "-a", "0",
                "--cov-mode", "0", "--min-seq-id", "0.9", "--min-aln-len", "0", "--seq-id-mode", "0", "--add-self-matches", "0", "--sort-results", "0",
                "--db-load-mode", "0", "--threads", "4", "--compressed", "0", "-v", "3"]   # HAMMING_PAR as `carpedeam linclust` generates it


@pytest.mark.parametrize("wrapped,golden", [(0, "lin_rescore_9"), (1, "lin_rescore_wrapped_9")])
def test_stage_binary_hamming_rescore_matches_reference_db(tmp_path, wrapped, golden):
    """carpedeam_gpu rescorediagonal --rescore-mode 0 on the pref DB of the reference's linclust kmermatcher == the DB the
    reference's rescorediagonal wrote (keys, entries, dbtype 14)"""
    import carpedeam_b200 as cd
    cd.build()
    d = str(tmp_path)
    write_db(d + "/seq", *load_golden("correction_9"), 1)
    write_db(d + "/pref", *load_golden("lin_pref_9"), 14)
    subprocess.run([cd.STAGE_BIN, "rescorediagonal", d + "/seq", d + "/seq", d + "/pref", d + "/out"] + HAMMING_ARGV + ["--wrapped-scoring", str(wrapped)], check=True)
    gk, gp, ge = load_golden(golden)
    mine = read_db(d + "/out")
    assert sorted(mine) == gk
    bad = [kk for kk, p in zip(gk, gp) if mine[kk][0] != p]
    assert not bad, (len(bad), mine[bad[0]], gp[gk.index(bad[0])])
    assert open(d + "/out.dbtype", "rb").read() == (14).to_bytes(4, "little")


@pytest.mark.parametrize("chop", [0, 1])
def test_stage_binary_cyclecheck_matches_reference(tmp_path, chop):
    """carpedeam_gpu cyclecheck writes the entries the reference reported (tests/golden/cycle_kat.txt), whole or cut at the diagonal"""
    import carpedeam_b200 as cd
    from support import load_cycle_kat
    cd.build()
    s, want = load_cycle_kat()
    d = str(tmp_path)
    write_db(d + "/seq", *load_golden("synth_cycle", ""), 1)
    subprocess.run([cd.STAGE_BIN, "cyclecheck", d + "/seq", d + "/cyc", "--max-seq-len", "200000", "--chop-cycle", str(chop), "--threads", "4", "-v", "3"], check=True)
    mine = read_db(d + "/cyc")
    assert sorted(mine) == [int(k) for k in range(s.n) if want[k]]
    for k, (payload, ext) in mine.items():
        assert payload == s.seq(k)[:int(want[k]) if chop else None] + b"\n" and ext == 0, k
