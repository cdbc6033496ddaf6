"""SURVEY §8(f) N3 on the CPU: the oracle's restatement of the contig-merge scores and the product's host-side queue order against
what the reference's own functions returned (tests/golden/merge_kat_*.txt.gz, written by oracle/make_merge_kat.py from
oracle/ref_merge.cpp linked against the unmodified reference)."""
import os

import numpy as np
import pytest

from support import GOLD, Oracle, check_merge_against_kat, load_merge_kat, load_merge_set


@pytest.fixture(scope="module")
def oracle():
    o = Oracle()
    o.load_damage(os.path.join(GOLD, "dhigh"))
    return o


@pytest.mark.parametrize("tag", ["synth", "asm9", "asm5"])
def test_oracle_merge_scores_match_reference_functions(oracle, tag):
    s, (off, t, sid, qs, qe, ds, de) = load_merge_set(tag)
    for (unsafe, thr), kat in load_merge_kat(tag).items():
        got = oracle.merge(s, off, t, qs, qe, ds, de, unsafe=unsafe, merge_thr=np.float32(thr))
        n_queued, n_exact = check_merge_against_kat(off, got, kat, 1e-6)
        assert n_queued > 100 and n_exact == n_queued
        for q, k in kat.items():
            l0 = int(off[q])
            idx = [i for i, a in enumerate(k["lines"]) if a["queued"]]
            if idx:
                order = oracle.merge_queue_order(got["aln_len_cons"][l0:][idx], got["deam"][l0:][idx])
                assert [idx[j] for j in order] == k["pops"], (tag, unsafe, thr, q)


@pytest.mark.parametrize("tag", ["synth", "asm5"])
def test_host_queue_order_matches_reference_priority_queue(tag):
    """cdh_merge_less / cdh_merge_queue_order (carpedeam_b200/host/merge.cpp) on the reference's own (alnLengthCons, deamMatch) values:
    the lines leave the queue in the order CompareNuclResultByScoreContigs + std::priority_queue produced in the reference"""
    import carpedeam_b200 as cd
    n_multi = 0
    for (unsafe, thr), kat in load_merge_kat(tag).items():
        for q, k in kat.items():
            idx = [i for i, a in enumerate(k["lines"]) if a["queued"]]
            if not idx:
                continue
            alc = [k["lines"][i]["aln_len_cons"] for i in idx]
            deam = [k["lines"][i]["deam"] for i in idx]
            order = cd.merge_queue_order(alc, deam)
            assert [idx[j] for j in order] == k["pops"], (tag, unsafe, thr, q)
            n_multi += len(idx) > 1
    assert n_multi > 20


def test_host_merge_model_matches_oracle_damage_tables(oracle):
    import carpedeam_b200 as cd
    m = cd.load_merge_model(os.path.join(GOLD, "dhigh"))
    D, Drev = oracle.get_damage()
    got = np.array(m.match_lik[:]).reshape(2, 4, 4)
    assert np.array_equal(got[0], np.asarray(D).reshape(11, 4, 4)[5]) and np.array_equal(got[1], np.asarray(Drev).reshape(11, 4, 4)[5])
    assert got[0][1][3] > 0 and got[0][2][0] > 0


def test_oracle_cyclecheck_matches_reference_binary(oracle):
    """the oracle's cyclecheck on the synthetic contig set == the split diagonals the reference binary reported (oracle/make_cycle_golden.py)"""
    from support import load_cycle_kat
    s, want = load_cycle_kat()
    assert np.array_equal(oracle.cyclecheck(s), want) and (want > 0).sum() > 50


def test_cyclecheck_matching_bodies_on_the_cpu_including_the_long_subrun_path(tmp_path):
    """the host/device bodies of the cyclecheck matching (cdg_cycle.cuh), run on the CPU by tests/cycle_host_harness.cu: with the shipped cap every
    record walks back itself (the path the GPU tests cover); with a cap of 3 almost every repeated k-mer goes through the forward walk of its
    sub-run head (the path for k-mers repeated thousands of times in one sequence) — both must give the reference's split diagonals"""
    import shutil
    import subprocess
    from support import ROOT, load_cycle_kat
    if shutil.which("nvcc") is None:
        pytest.skip("nvcc not on PATH")
    exe = str(tmp_path / "cycle_host_harness")
    subprocess.run(["nvcc", "-O2", "-std=c++17", "-x", "cu", "-o", exe, os.path.join(ROOT, "tests", "cycle_host_harness.cu")], check=True)
    s, want = load_cycle_kat()
    inp = str(tmp_path / "set.bin")
    with open(inp, "wb") as f:
        f.write(np.uint32(s.n).tobytes()); f.write(s.len.astype(np.uint32).tobytes())
        for i in range(s.n):
            f.write(s.seq(i))
    for cap in (4096, 3, 0):
        out = subprocess.run([exe, inp, "22", "200000", str(cap)], check=True, capture_output=True, text=True).stdout.split()
        assert np.array_equal(np.array(out, np.uint32), want), cap
