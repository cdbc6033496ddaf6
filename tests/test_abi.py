"""CPU-side checks: the C-ABI library loads and exports every symbol include/carpedeam_gpu.h declares,
the host-side damage builder agrees with the oracle, and there is no CPU fallback."""
import ctypes
import os
import re

import numpy as np
import pytest

from support import GOLD, ROOT


def _declared():
    hdr = open(os.path.join(ROOT, "include", "carpedeam_gpu.h")).read()
    return sorted(set(re.findall(r"\b(cdg_[a-z0-9_]+)\s*\(", hdr)))


def test_library_exports_every_declared_symbol():
    import carpedeam_b200 as cd
    cd.build()
    L = ctypes.CDLL(cd.LIB_PATH)
    names = _declared()
    assert len(names) >= 15
    for n in names:
        assert hasattr(L, n), n


def test_no_cpu_fallback():
    import torch
    import carpedeam_b200 as cd
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(cd.CdgError):
        cd.Context(0)


def test_product_never_imports_oracle():
    bad = []
    for d, _dirs, files in os.walk(os.path.join(ROOT, "carpedeam_b200")):
        for f in files:
            if f.endswith((".py", ".cpp", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(d, f), errors="ignore").read()
                if re.search(r"#include[^\n]*oracle|import[^\n]*oracle|CDLL\([^\n]*oracle|dlopen\([^\n]*oracle|\borc_[a-z0-9_]+\s*\(|libcarpedeam_oracle", txt):
                    bad.append(f)
    assert not bad, bad


def test_host_damage_tables_match_oracle(oracle):
    import carpedeam_b200 as cd
    cd.build()
    for pfx in (os.path.join(GOLD, "dhigh"), ""):
        d = cd.load_damage(pfx)
        oracle.load_damage(pfx)
        D, R = oracle.get_damage()
        logd = np.array(d.log_d).reshape(11, 4, 4); logr = np.array(d.log_drev).reshape(11, 4, 4)
        assert np.array_equal(logd, np.log(np.maximum(D, 1e-3)))
        assert np.array_equal(logr, np.log(np.maximum(R, 1e-3)))
        se = np.array(d.log_seq_err).reshape(4, 4)
        assert abs(se[0][0] - np.log(0.99)) < 1e-15 and abs(se[0][1] - np.log(0.01 / 3)) < 1e-15


def test_damage_profile_errors():
    import carpedeam_b200 as cd
    with pytest.raises(cd.CdgError):
        cd.load_damage("/nonexistent/prefix")


def test_host_pair_model_matches_oracle_columns(oracle):
    """cdh_build_pair_model tabulates the per-column term of calcLikelihood: a one-column pair through the oracle
    (restatement pinned by tests/golden/pairlik_kat.txt) must give the same log for every (strand, class, qBase, tBase)."""
    import carpedeam_b200 as cd
    cd.build()
    pfx = os.path.join(GOLD, "dhigh")
    m = cd.load_pair_model(pfx)
    oracle.load_damage(pfx)
    tab = np.array(m.log_lik).reshape(2, 11, 4, 4)
    assert m.log_excess == np.log(np.float32(0.0625)) and m.log_rand == np.log(np.float32(0.85))
    comp = {"A": "T", "C": "G", "G": "C", "T": "A"}
    dbLen = 30
    for rev in (0, 1):
        for cls in range(11):
            pos = cls if cls < 5 else (12 if cls == 5 else dbLen - 5 + (cls - 6))
            for qi, qb in enumerate("ACGT"):
                for ti, tb in enumerate("ACGT"):
                    t = ["A"] * dbLen
                    t[pos] = tb
                    stored = "".join(comp[c] for c in reversed(t)) if rev else "".join(t)
                    query = "A" * pos + qb
                    # one aligned column at oriented target position pos; maxAln = 1 -> no excess term
                    a, _ = oracle.pair_loglik(query.encode(), stored.encode(), rev, pos, pos, 1, 1)
                    assert a == tab[rev, cls, qi, ti], (rev, cls, qb, tb, a, tab[rev, cls, qi, ti])
