"""The in-tree binding (north_star: host code stays inside the reference's modules and calls CUDA through the C ABI):
carpedeam_b200/binding/stages_gpu.cpp is compiled against the reference's own headers and linked with the UNMODIFIED reference objects in place of
the entry points of kmermatcher / rescorediagonal / ancient_correction / cyclecheck (oracle/build_ref.py --gpu-stages -> oracle/_ref/carpedeam_ref_gpu).

CPU (needs /root/reference and oracle/_ref/obj, i.e. the build container): the binary builds; a call on the GPU path ends with the error of a missing
device (no CPU fallback); a call outside the GPU path is delegated to the reference's own implementation and writes what the reference writes.
GPU: the three stages of the binding binary on config 1's first iteration against the golden DBs, and the fused `iteration` stage of carpedeam_gpu (both
ran green on B200 with the last GPU seconds of the round, profiles/r02v_last_check.log); the binding's cyclecheck entry point and the contig case of the
fused stage have not run on a device: xfail(strict=False)."""
import os
import subprocess

import pytest

from support import GOLD, ROOT, load_golden, read_db, write_db

REF = os.path.join(ROOT, "oracle", "_ref", "carpedeam_ref")
BIND = os.path.join(ROOT, "oracle", "_ref", "carpedeam_ref_gpu")
HAVE_REF_TREE = os.path.isdir("/root/reference") and os.path.isdir(os.path.join(ROOT, "oracle", "_ref", "obj"))

KM = ("--sub-mat nucl:nucleotide.out,aa:blosum62.out --alph-size nucl:5,aa:13 --min-seq-id 0.9 --kmer-per-seq 200 --spaced-kmer-mode 0 --kmer-per-seq-scale 0.200 "
      "--adjust-kmer-len 0 --mask 0 --mask-lower-case 0 --cov-mode 1 -k 20 -c 0 --max-seq-len 200000 --hash-shift 67 --split-memory-limit 0 "
      "--include-only-extendable 0 --threads 4 --compressed 0 -v 3").split()
RS = ("--sub-mat nucl:nucleotide.out,aa:blosum62.out --wrapped-scoring 0 --filter-hits 0 -e 0.001 -c 0 -a 0 --cov-mode 1 --min-seq-id 0.9 --min-aln-len 0 "
      "--seq-id-mode 0 --add-self-matches 0 --sort-results 0 --db-load-mode 0 --threads 4 --compressed 0 -v 3").split()


# the binding binary finds libcarpedeam_gpu.so / libcarpedeam_host.so through its rpath (the path of the build container); wherever the repository
# lies at test time, the loader also looks in its carpedeam_b200/
ENV = dict(os.environ, LD_LIBRARY_PATH=os.path.join(ROOT, "carpedeam_b200") + os.pathsep + os.environ.get("LD_LIBRARY_PATH", ""))


def _have_gpu():
    import torch
    return torch.cuda.is_available()


@pytest.fixture(scope="module")
def binding():
    if not HAVE_REF_TREE:
        pytest.skip("needs /root/reference and oracle/_ref/obj (the build container)")
    import carpedeam_b200 as cd
    cd.build()
    r = subprocess.run(["python", os.path.join(ROOT, "oracle", "build_ref.py"), "--gpu-stages"], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    return BIND


def test_binding_compiles_against_the_reference_and_has_no_cpu_fallback(binding, tmp_path):
    d = str(tmp_path)
    write_db(d + "/seq", *load_golden("nucl_reads"), 1)
    if _have_gpu():
        pytest.skip("GPU present: see the GPU test")
    r = subprocess.run([binding, "kmermatcher", d + "/seq", d + "/pref"] + KM + ["--ignore-multi-kmer", "1"], capture_output=True, text=True, env=ENV)
    assert r.returncode != 0 and "No usable CUDA device" in r.stdout + r.stderr
    assert "GPU k-mer matcher" in r.stdout + r.stderr            # the call was on the GPU path (the reference's own code prints no such line)
    assert not os.path.exists(d + "/pref.index") or os.path.getsize(d + "/pref.index") == 0
    write_db(d + "/contigs", *load_golden("synth_cycle", ""), 1)
    r = subprocess.run([binding, "cyclecheck", d + "/contigs", d + "/cyc", "--max-seq-len", "200000", "--chop-cycle", "1", "--threads", "4", "-v", "3"],
                       capture_output=True, text=True, env=ENV)
    assert r.returncode != 0 and "No usable CUDA device" in r.stdout + r.stderr


def test_binding_delegates_calls_outside_the_gpu_path_to_the_reference(binding, tmp_path):
    """--ignore-multi-kmer 0 (k-mer matcher) and --rescore-mode 1 (rescorediagonal) are not on the GPU path: the binding hands them to the
    reference's own entry points (renamed *_cpu at link time) and the output equals the unmodified binary's"""
    d = str(tmp_path)
    write_db(d + "/seq", *load_golden("nucl_reads"), 1)
    for exe, tag in ((binding, "b"), (REF, "r")):
        subprocess.run([exe, "kmermatcher", d + "/seq", d + "/pref_" + tag] + KM + ["--ignore-multi-kmer", "0"], check=True, capture_output=True, env=ENV)
        subprocess.run([exe, "rescorediagonal", d + "/seq", d + "/seq", d + "/pref_" + tag, d + "/resc_" + tag] + RS + ["--rescore-mode", "1"], check=True,
                       capture_output=True, env=ENV)
    for name in ("pref_", "resc_"):
        a, b = read_db(d + "/" + name + "b"), read_db(d + "/" + name + "r")
        assert a == b and len(a) == 14996, name


@pytest.mark.gpu
def test_binding_stages_reproduce_the_golden_dbs_on_the_gpu(tmp_path):
    """the reference binary with the in-tree binding: kmermatcher, rescorediagonal and ancient_correction of config 1's first iteration write the
    golden DBs (ran green on B200 with the last GPU seconds of the round: profiles/r02v_last_check.log)"""
    if not os.path.exists(BIND) or not os.path.exists(REF):
        pytest.skip("oracle/_ref/carpedeam_ref_gpu not present")
    if "avx2" not in open("/proc/cpuinfo").read():
        pytest.skip("the reference objects were built for x86-64-v3")
    d = str(tmp_path)
    write_db(d + "/seq", *load_golden("nucl_reads"), 1)
    subprocess.run([BIND, "kmermatcher", d + "/seq", d + "/pref"] + KM + ["--ignore-multi-kmer", "1"], check=True, timeout=180, env=ENV)
    subprocess.run([BIND, "rescorediagonal", d + "/seq", d + "/seq", d + "/pref", d + "/aln"] + RS + ["--rescore-mode", "3"], check=True, timeout=180, env=ENV)
    subprocess.run([BIND, "ancient_correction", d + "/seq", d + "/aln", d + "/corr", "--min-seq-id", "0.9", "--max-seq-len", "200000", "--keep-target", "1",
                    "--threads", "4", "-v", "3", "--rescore-mode", "3", "--min-ryseq-id-corr-reads", "0.99", "--min-seqid-corr-reads", "0.9",
                    "--ancient-damage", os.path.join(GOLD, "dhigh"), "--unsafe", "0", "--min-cov-safe", "5"], check=True, timeout=180, env=ENV)
    for name, out in (("pref_0", "pref"), ("aln_0", "aln"), ("correction_0", "corr")):
        gk, gp, ge = load_golden(name)
        mine = read_db(d + "/" + out)
        assert sorted(mine) == gk, name
        assert not [k for k, p, e in zip(gk, gp, ge) if mine[k] != (p, e)], name


@pytest.mark.gpu
@pytest.mark.xfail(strict=False, reason="the cyclecheck entry point of the binding binary never ran on a device (GPU budget of the round spent)")
def test_binding_cyclecheck_on_the_gpu(tmp_path):
    if not os.path.exists(BIND) or "avx2" not in open("/proc/cpuinfo").read():
        pytest.skip("oracle/_ref/carpedeam_ref_gpu not present / not runnable")
    from support import load_cycle_kat
    d = str(tmp_path)
    s, want = load_cycle_kat()
    write_db(d + "/contigs", *load_golden("synth_cycle", ""), 1)
    subprocess.run([BIND, "cyclecheck", d + "/contigs", d + "/cyc", "--max-seq-len", "200000", "--chop-cycle", "1", "--threads", "4", "-v", "3"], check=True, timeout=180, env=ENV)
    mine = read_db(d + "/cyc")
    assert sorted(mine) == [int(k) for k in range(s.n) if want[k]]
    assert all(payload == s.seq(k)[:int(want[k])] + b"\n" for k, (payload, ext) in mine.items())


@pytest.mark.gpu
@pytest.mark.parametrize("seqdb,pref,aln,corr,k,only_ext", [
    ("nucl_reads", "pref_0", "aln_0", "correction_0", 20, 0),       # ran green on B200 (profiles/r02v_last_check.log)
    pytest.param("assembly_reads_4", "pref_asm_5", "aln_asm_5", "correction_5", 22, 1,
                 marks=pytest.mark.xfail(strict=False, reason="the contig case of the fused stage never ran on a device (GPU budget of the round spent)")),
])
def test_stage_binary_fused_iteration_writes_the_three_reference_dbs(tmp_path, seqdb, pref, aln, corr, k, only_ext):
    """`carpedeam_gpu iteration` = kmermatcher + rescorediagonal + ancient_correction in one process and one upload, emitting the
    same three databases"""
    import carpedeam_b200 as cd
    cd.build()
    d = str(tmp_path)
    write_db(d + "/seq", *load_golden(seqdb), 1)
    subprocess.run([cd.STAGE_BIN, "iteration", d + "/seq", d + "/pref", d + "/aln", d + "/corr", "--sub-mat", "nucl:nucleotide.out,aa:blosum62.out", "--min-seq-id", "0.9",
                    "--kmer-per-seq", "200", "--kmer-per-seq-scale", "0.200", "--cov-mode", "1", "-k", str(k), "-c", "0", "--hash-shift", "67",
                    "--include-only-extendable", str(only_ext), "--ignore-multi-kmer", "1", "--rescore-mode", "3", "-e", "0.001", "--min-ryseq-id-corr-reads", "0.99",
                    "--ancient-damage", os.path.join(GOLD, "dhigh"), "--threads", "8", "-v", "3"], check=True, timeout=180)
    for name, out in ((pref, "pref"), (aln, "aln"), (corr, "corr")):
        gk, gp, ge = load_golden(name)
        mine = read_db(d + "/" + out)
        assert sorted(mine) == gk, name
        assert not [kk for kk, p, e in zip(gk, gp, ge) if mine[kk] != (p, e)], name
