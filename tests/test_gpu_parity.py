"""GPU parity tests: the CUDA path through the C ABI (carpedeam_b200 -> libcarpedeam_gpu.so) against the
oracle on the same inputs, and against the reference's own golden DBs.  Bit-exact for hits, diagonals,
integer scores, coordinates and corrected bases; 1e-5 relative for log-likelihoods (north_star)."""
import os

import numpy as np
import pytest

from support import GOLD, SeqSet, format_aln, format_pref, load_golden, parse_aln, parse_pref, synth_reads
from support import seqid_text as seqid_text_bytes

pytestmark = pytest.mark.gpu

CASES = [
    ("nucl_reads", "pref_0", "aln_0", "correction_0", 20, 0),
    ("assembly_reads_0", "pref_1", "aln_1", "correction_1", 20, 0),
    ("assembly_reads_4", "pref_asm_5", "aln_asm_5", "correction_5", 22, 1),
    ("assembly_contigs_8_noneCycle", "pref_asm_9", "aln_asm_9", "correction_9", 22, 1),
]


@pytest.fixture(scope="module")
def ctx():
    import carpedeam_b200 as cd
    c = cd.Context(0)
    yield c
    c.close()


def test_hash_and_revcomp(ctx, oracle):
    rng = np.random.default_rng(1)
    for k in (20, 22, 6, 31):
        x = rng.integers(0, 1 << (2 * k), 5000, dtype=np.uint64)
        h, r = ctx.test_hash_revcomp(x, k, 67)
        for i in range(0, 5000, 97):
            assert int(h[i]) == oracle.L.orc_xxh64_u64(int(x[i]), 67)
            assert int(r[i]) == oracle.L.orc_revcomp(int(x[i]), k)


@pytest.mark.parametrize("n", [0, 1, 31, 4096, 4097, 100003, 3000001])
@pytest.mark.parametrize("pairs", [False, True])
def test_radix_sort(ctx, n, pairs):
    rng = np.random.default_rng(n + 7)
    keys = rng.integers(0, 1 << 62, n, dtype=np.uint64)
    if n > 10:
        keys[: n // 3] &= np.uint64(0xFFFF)      # skewed digits / many duplicates
    vals = np.arange(n, dtype=np.uint64) if pairs else None
    for lo, hi in ((0, 40), (0, 63), (3, 17)):
        k2, v2 = ctx.test_radix_sort(keys, vals, lo, hi)
        mask = np.uint64(((1 << hi) - 1) ^ ((1 << lo) - 1))
        order = np.argsort(keys & mask, kind="stable")
        assert np.array_equal(k2, keys[order])
        if pairs:
            assert np.array_equal(v2, vals[order])   # stability


@pytest.mark.parametrize("variant", [0, 1, 2, 3, 4, 5, 6, 7, 8])
def test_radix_sort_all_tile_shapes(ctx, variant):
    import carpedeam_b200 as cd
    L = cd.lib()
    rng = np.random.default_rng(variant)
    try:
        L.cdg_test_set_sort_variant(variant, variant)
        for n in (5, 2047, 2049, 300001, 2500003):
            keys = rng.integers(0, 1 << 63, n, dtype=np.uint64)
            keys[: n // 2] &= np.uint64(0xFFFFF)
            vals = np.arange(n, dtype=np.uint64)
            for pairs in (False, True):
                k2, v2 = ctx.test_radix_sort(keys, vals if pairs else None, 0, 44)
                order = np.argsort(keys & np.uint64((1 << 44) - 1), kind="stable")
                assert np.array_equal(k2, keys[order])
                if pairs:
                    assert np.array_equal(v2, vals[order])
    finally:
        L.cdg_test_set_sort_variant(8, 0)     # the library defaults


def _hits_equal(a, b):
    for f in ("rep", "tgt", "score", "diag"):
        assert np.array_equal(a[f], b[f]), f
    assert a["R1"] == b["R1"] and a["R2"] == b["R2"]


@pytest.mark.parametrize("seqdb,pref,aln,corr,k,only_ext", CASES)
def test_kmermatch_vs_oracle_and_reference(ctx, oracle, seqdb, pref, aln, corr, k, only_ext):
    import carpedeam_b200 as cd
    s = SeqSet.from_golden(seqdb)
    ctx.upload(s.blob, s.off, s.len, s.ext)
    got = ctx.kmermatch(cd.kmer_params(k=k, include_only_extendable=only_ext))
    want = oracle.kmermatch(s, k=k, only_ext=only_ext)
    _hits_equal(got, want)
    keys, payloads, _ = load_golden(pref)
    mine = format_pref(s, got)
    assert all(mine[kk] == p for kk, p in zip(keys, payloads))


@pytest.mark.parametrize("seqdb,pref,aln,corr,k,only_ext", CASES)
def test_rescore_standalone_vs_reference(ctx, oracle, seqdb, pref, aln, corr, k, only_ext):
    s = SeqSet.from_golden(seqdb)
    keys, payloads, _ = load_golden(pref)
    q, t, sc, dg, off = parse_pref(s, payloads)
    ctx.upload(s.blob, s.off, s.len, s.ext)
    ctx.set_pref_lines(off, t, sc, dg)
    got = ctx.rescore()
    want = oracle.rescore(s, q, t, sc, dg)
    for f in ("keep", "score", "bits", "qs", "qe", "ds", "de"):
        assert np.array_equal(got[f], want[f]), f
    assert np.array_equal(got["seqid"], want["seqid"])
    assert np.allclose(got["evalue"], want["evalue"], rtol=1e-12, atol=0)
    mine = format_aln(s, got["line_off"], got["tgt"], got)
    akeys, apay, _ = load_golden(aln)
    bad = [kk for kk, p in zip(akeys, apay) if mine[kk] != p]
    assert not bad, (len(bad), mine[bad[0]][:200], apay[akeys.index(bad[0])][:200])


@pytest.mark.parametrize("seqdb,pref,aln,corr,k,only_ext", CASES)
def test_correction_standalone_vs_reference(ctx, oracle, seqdb, pref, aln, corr, k, only_ext):
    import carpedeam_b200 as cd
    s = SeqSet.from_golden(seqdb)
    akeys, apay, _ = load_golden(aln)
    off, t, sid, qs, qe, ds, de = parse_aln(s, apay)
    ctx.upload(s.blob, s.off, s.len, s.ext)
    ctx.set_damage(cd.load_damage(os.path.join(GOLD, "dhigh")))
    ctx.set_alignments(off, t, sid, qs, qe, ds, de)
    out, changed = ctx.correct()
    oracle.load_damage(os.path.join(GOLD, "dhigh"))
    want, lik_want = oracle.correct(s, off, t, sid, qs, qe, ds, de, want_lik=True)
    assert np.array_equal(out, want)
    ckeys, cpay, _ = load_golden(corr)
    for i, p in enumerate(cpay):
        o = int(s.off[i]); L = int(s.len[i])
        assert out[o:o + L].tobytes() + b"\n" == p
    assert changed == int((out != s.blob).sum())
    # log-likelihoods: within 1e-5 relative of the reference's long double sums
    lik = ctx.correct_likelihoods(s.total)
    m = ~np.isnan(lik_want)
    assert np.array_equal(m, ~np.isnan(lik))
    assert np.allclose(lik[m], lik_want[m], rtol=1e-5, atol=0)
    assert m.sum() > 0


@pytest.mark.parametrize("seqdb,pref,aln,corr,k,only_ext", CASES)
def test_fused_iteration_vs_reference(ctx, seqdb, pref, aln, corr, k, only_ext):
    import carpedeam_b200 as cd
    s = SeqSet.from_golden(seqdb)
    ctx.upload(s.blob, s.off, s.len, s.ext)
    ctx.set_damage(cd.load_damage(os.path.join(GOLD, "dhigh")))
    hits, alns, out, changed = ctx.iteration(kmer=cd.kmer_params(k=k, include_only_extendable=only_ext), fetch_hits=True)
    keys, payloads, _ = load_golden(pref)
    mine = format_pref(s, hits)
    assert all(mine[kk] == p for kk, p in zip(keys, payloads))
    akeys, apay, _ = load_golden(aln)
    mine = format_aln(s, alns["line_off"], alns["tgt"], alns)
    assert all(mine[kk] == p for kk, p in zip(akeys, apay))
    ckeys, cpay, _ = load_golden(corr)
    for i, p in enumerate(cpay):
        o = int(s.off[i]); L = int(s.len[i])
        assert out[o:o + L].tobytes() + b"\n" == p
    st = ctx.stats()
    assert st["n_hits"] == len(hits["rep"]) and st["kernel_launches"] > 0


def _oracle_iteration(oracle, s, k=20, only_ext=0):
    hits = oracle.kmermatch(s, k=k, only_ext=only_ext)
    n = s.n
    H = len(hits["rep"])
    lo = np.searchsorted(hits["rep"], np.arange(n + 1))
    line_off = (np.arange(n + 1) + lo).astype(np.uint64)
    M = n + H
    q = np.zeros(M, np.uint32); t = np.zeros(M, np.uint32); sc = np.zeros(M, np.int32); dg = np.zeros(M, np.uint16)
    self_pos = line_off[:n].astype(np.int64)
    q[self_pos] = np.arange(n); t[self_pos] = np.arange(n)
    hp = np.arange(H) + hits["rep"].astype(np.int64) + 1
    q[hp] = hits["rep"]; t[hp] = hits["tgt"]; sc[hp] = hits["score"]; dg[hp] = hits["diag"].astype(np.uint16)
    r = oracle.rescore(s, q, t, sc, dg)
    kept = r["keep"].astype(bool)
    from support import seqid_text
    sid = np.array([np.float32(float(seqid_text(x))) for x in r["seqid"][kept]], np.float32)
    aoff = np.zeros(n + 1, np.uint64)
    np.add.at(aoff, q[kept].astype(np.int64) + 1, 1)
    aoff = np.cumsum(aoff).astype(np.uint64)
    out, _ = oracle.correct(s, aoff, t[kept], sid, r["qs"][kept], r["qe"][kept], r["ds"][kept], r["de"][kept])
    return hits, r, line_off, out


@pytest.mark.parametrize("n,glen,nfrac,seed", [(3000, 20000, 0.0, 11), (20000, 100000, 0.002, 12), (500, 3000, 0.05, 13)])
def test_fused_iteration_vs_oracle_synthetic(ctx, oracle, n, glen, nfrac, seed):
    import carpedeam_b200 as cd
    seqs = synth_reads(n, glen, seed, n_frac=nfrac)
    if seed == 13:   # ragged: lower case, IUPAC, tiny and long sequences, duplicates, reverse-complement duplicates
        seqs[3] = seqs[3].lower(); seqs[5] = seqs[5][:10] + b"R" + seqs[5][11:]; seqs[7] = b"ACGTACGTAC"; seqs[9] = seqs[8]
        seqs[11] = b"A" * 60; seqs[13] = (seqs[12] + seqs[14] + seqs[15] + seqs[16] + seqs[17] + seqs[18]) * 2
    s = SeqSet(seqs)
    ctx.upload(s.blob, s.off, s.len, s.ext)
    ctx.set_damage(cd.load_damage(os.path.join(GOLD, "dhigh")))
    oracle.load_damage(os.path.join(GOLD, "dhigh"))
    hits, alns, out, changed = ctx.iteration(fetch_hits=True)
    ohits, r, line_off, oout = _oracle_iteration(oracle, s)
    _hits_equal(hits, ohits)
    assert np.array_equal(alns["line_off"], line_off)
    for f in ("keep", "score", "bits", "qs", "qe", "ds", "de"):
        assert np.array_equal(alns[f], r[f]), f
    assert np.array_equal(alns["seqid"], r["seqid"])
    assert np.array_equal(out, oout)
    assert changed == int((out != s.blob).sum())


def test_full_size_properties(ctx):
    """size-independent properties at a size the oracle cannot do in seconds (2M reads)."""
    import carpedeam_b200 as cd
    from bench import synth_reads_fast
    s = synth_reads_fast(2_000_000, 10_000_000, 99)
    ctx.upload(s["blob"], s["off"], s["len"], None)
    ctx.set_damage(cd.load_damage(os.path.join(GOLD, "dhigh")))
    hits, alns, out, changed = ctx.iteration(fetch_hits=True)
    rep, tgt = hits["rep"].astype(np.int64), hits["tgt"].astype(np.int64)
    pair = rep * (1 << 32) + tgt
    assert np.all(np.diff(pair) > 0)                       # hit list strictly sorted by (rep,target): no duplicates
    assert np.all(rep != tgt)
    L = s["len"].astype(np.int64)
    assert np.all(L[rep] >= L[tgt])                        # representative = longest member of its k-mer group
    assert np.all(np.abs(hits["score"]) >= 1)
    # self alignments: always kept, full length, score 2L
    lo = alns["line_off"].astype(np.int64)
    selfl = lo[:-1]
    assert np.all(alns["keep"][selfl] == 1) and np.all(alns["score"][selfl] == 2 * L) and np.all(alns["qe"][selfl] == L - 1)
    k = alns["keep"].astype(bool)
    assert np.all(alns["seqid"][k] >= np.float32(0.9) - np.float32(1.2e-7))
    # correction only rewrites base bytes, lengths and separators untouched, and is idempotent w.r.t. unchanged reads
    diff = np.nonzero(out != s["blob"])[0]
    assert len(diff) == changed and changed > 0
    assert set(np.unique(out[diff]).tolist()) <= set(b"ACGT")
    hits2, alns2, out2, changed2 = ctx.iteration(fetch_hits=True)
    assert np.array_equal(out, out2) and np.array_equal(hits["score"], hits2["score"])   # deterministic


def test_correction_generic_kernel_path(oracle, monkeypatch):
    """queries with > 65535 counted targets go through the generic 32-bit-counter kernel; force that path with the
    test knob and require identical corrected reads."""
    import carpedeam_b200 as cd
    monkeypatch.setenv("CDG_WIDE_THRESHOLD", "3")
    c = cd.Context(0)
    try:
        s = SeqSet.from_golden("nucl_reads")
        akeys, apay, _ = load_golden("aln_0")
        off, t, sid, qs, qe, ds, de = parse_aln(s, apay)
        c.upload(s.blob, s.off, s.len, s.ext)
        c.set_damage(cd.load_damage(os.path.join(GOLD, "dhigh")))
        c.set_alignments(off, t, sid, qs, qe, ds, de)
        out, changed = c.correct()
        ckeys, cpay, _ = load_golden("correction_0")
        for i, p in enumerate(cpay):
            o = int(s.off[i]); L = int(s.len[i])
            assert out[o:o + L].tobytes() + b"\n" == p
        assert changed == 3348
    finally:
        c.close()


@pytest.mark.parametrize("k,only_ext", [(22, 1), (20, 0)])
def test_long_contigs_and_u16_diagonal_wrap(ctx, oracle, k, only_ext):
    """contigs of 40-150 kb next to reads cut from them: exercises the block-per-sequence extraction with an active hash
    threshold, diagonals beyond +-32767 (stored as u16 and re-expanded by trying +-65536 multiples,
    DistanceCalculator.h:93-113) and ext-flagged queries/targets in the correction."""
    import carpedeam_b200 as cd
    rng = np.random.default_rng(5 + k)
    acgt = np.frombuffer(b"ACGT", np.uint8)
    contigs = [acgt[rng.integers(0, 4, L)].tobytes() for L in (150000, 70001, 40000)]
    seqs, ext = [], []
    for c in contigs:
        seqs.append(c); ext.append(1)
    comp = bytes.maketrans(b"ACGT", b"TGCA")
    for c in contigs:
        for _ in range(400):
            L = int(rng.integers(60, 400)); st = int(rng.integers(0, len(c) - L))
            r = bytearray(c[st:st + L])
            for _m in range(int(rng.integers(0, 3))):
                r[int(rng.integers(0, L))] = int(acgt[rng.integers(0, 4)])
            r = bytes(r)
            if rng.random() < 0.5:
                r = r.translate(comp)[::-1]
            seqs.append(r); ext.append(0)
    order = rng.permutation(len(seqs))
    seqs = [seqs[i] for i in order]; ext = [ext[i] for i in order]
    s = SeqSet(seqs, ext=ext)
    ctx.upload(s.blob, s.off, s.len, s.ext)
    ctx.set_damage(cd.load_damage(os.path.join(GOLD, "dhigh")))
    oracle.load_damage(os.path.join(GOLD, "dhigh"))
    hits, alns, out, changed = ctx.iteration(kmer=cd.kmer_params(k=k, include_only_extendable=only_ext), fetch_hits=True)
    ohits, r, line_off, oout = _oracle_iteration(oracle, s, k=k, only_ext=only_ext)
    _hits_equal(hits, ohits)
    assert np.abs(ohits["diag"].astype(np.int64)).max() > 0 and len(ohits["rep"]) > 500
    for f in ("keep", "score", "bits", "qs", "qe", "ds", "de"):
        assert np.array_equal(alns[f], r[f]), f
    if not only_ext:   # (with include-only-extendable reads inside a contig are dropped, kmermatcher.cpp:526-536)
        assert (np.abs(alns["qs"][alns["keep"] == 1].astype(np.int64)).max()) > 40000   # diagonals beyond the u16 range were recovered
    assert np.array_equal(out, oout)


def test_pair_likelihood_known_answers(ctx, oracle):
    """a19: cdg_pair_loglik / cdg_pair_loglik_consensus against the outputs of the reference's own calcLikelihood /
    calcLikelihoodConsensus (tests/golden/pairlik_kat.txt) and against the oracle.  Tolerance: north_star asks for 1e-5
    relative; the fp64 warp reduction agrees with the reference's long double sums to ~1e-15, asserted at 1e-12."""
    import carpedeam_b200 as cd
    from support import load_pairlik_kat
    kat = load_pairlik_kat()
    seqs = []
    for k in kat:   # consensus-mode queries: the middle third of the consensus is the query read ('N' holes -> 'A', only its length matters)
        q = k["query"] if k["mode"] == "L" else k["query"][k["qlen"]:2 * k["qlen"]].replace(b"N", b"A")
        seqs += [q, k["target"]]
    s = SeqSet(seqs)
    ctx.upload(s.blob, s.off, s.len, s.ext)
    ctx.set_pair_model(cd.load_pair_model(os.path.join(GOLD, "dhigh")))
    oracle.load_damage(os.path.join(GOLD, "dhigh"))
    for mode in ("L", "C"):
        idx = [i for i, k in enumerate(kat) if k["mode"] == mode]
        sub = [kat[i] for i in idx]
        q = [2 * i for i in idx]; t = [2 * i + 1 for i in idx]
        args = (q, t, [k["rev"] for k in sub], [k["qs"] for k in sub], [k["qe"] for k in sub], [k["ds"] for k in sub],
                [k["de"] for k in sub], [k["max_aln"] for k in sub])
        if mode == "L":
            a, b = ctx.pair_loglik(*args)
        else:
            blob = b"".join(k["query"] for k in sub)
            off = np.cumsum([0] + [len(k["query"]) for k in sub[:-1]]).astype(np.uint64)
            a, b = ctx.pair_loglik(*args, consensus=np.frombuffer(blob, np.uint8), cons_off=off)
        want_a = np.array([k["slen"] for k in sub]); want_b = np.array([k["sratio"] for k in sub])
        assert np.all(np.abs(a - want_a) <= 1e-12 * np.abs(want_a)), np.abs(a / want_a - 1).max()
        assert np.all(np.abs(b - want_b) <= 1e-10 * np.abs(want_b) + 1e-300), (b, want_b)
        assert (want_b > 0.5).sum() >= 0 and (want_a > -1e9).sum() > 100      # mostly real (non-wrapped) cases
        for j in range(0, len(sub), 17):   # oracle cross-check on the same inputs
            k = sub[j]
            if mode == "L":
                oa, ob = oracle.pair_loglik(k["query"], k["target"], k["rev"], k["qs"], k["ds"], k["qe"] - k["qs"] + 1, k["max_aln"])
            else:
                oa, ob = oracle.pair_loglik_consensus(k["query"], k["qlen"], k["target"], k["rev"], k["qs"], k["qe"], k["ds"], k["de"], k["max_aln"])
            assert abs(a[j] - oa) <= 1e-12 * abs(oa)


def test_pair_likelihood_argument_errors(ctx):
    import carpedeam_b200 as cd
    s = SeqSet([b"ACGTACGTACGTACGTACGT", b"ACGTACGTACGTACGTACGTAA"])
    ctx.upload(s.blob, s.off, s.len, s.ext)
    ctx.set_pair_model(cd.load_pair_model(""))
    with pytest.raises(cd.CdgError):
        ctx.pair_loglik([0], [7], [0], [0], [9], [0], [9], [10])          # target index out of range
    with pytest.raises(cd.CdgError):
        ctx.pair_loglik([0], [1], [0], [0], [9], [0], [8], [10])          # not an ungapped alignment
    a, b = ctx.pair_loglik([0], [1], [0], [0], [19], [0], [19], [20])     # no damage, identical columns
    assert abs(a[0] - (20 * np.log(0.999 + 0.001 / 3 * 3 * 1e-3))) < 1e-9


@pytest.mark.parametrize("ct0,ga0", [(0.7, 0.65), (0.0, 0.0)])
def test_correction_with_unusual_damage_profiles(ctx, oracle, tmp_path, ct0, ga0):
    """The correction kernel skips columns where every counted target agrees with the query only when the damage tables
    make that provably safe; rates above 0.5 (first case) switch the shortcut off, an all-zero profile (second case,
    nuclassembleUtil.cpp:824-832 semantics through explicit zero rows) keeps it on.  Both must equal the oracle."""
    import carpedeam_b200 as cd
    hdr = "A>C\tA>G\tA>T\tC>A\tC>G\tC>T\tG>A\tG>C\tG>T\tT>A\tT>C\tT>G\n"
    rows5 = [[0.0] * 12 for _ in range(5)]; rows3 = [[0.0] * 12 for _ in range(5)]
    for i, f in enumerate((1.0, 0.6, 0.4, 0.3, 0.2)):
        rows5[i][5] = ct0 * f; rows3[4 - i][6] = ga0 * f
    pfx = str(tmp_path / "odd")
    for name, rows in (("5p.prof", rows5), ("3p.prof", rows3)):
        with open(pfx + name, "w") as f:
            f.write(hdr + "".join("\t".join("%g" % v for v in r) + "\n" for r in rows))
    s = SeqSet(synth_reads(6000, 30000, 21))
    ctx.upload(s.blob, s.off, s.len, s.ext)
    ctx.set_damage(cd.load_damage(pfx))
    oracle.load_damage(pfx)
    hits, alns, out, changed = ctx.iteration(fetch_hits=True)
    ohits, r, line_off, oout = _oracle_iteration(oracle, s)
    assert np.array_equal(out, oout)
    assert changed == int((out != s.blob).sum())
    if ct0 > 0:
        assert changed > 0
    lik = ctx.correct_likelihoods(s.total)     # the likelihood export evaluates every covered column
    assert np.isfinite(lik).any()


def _metagenome_reads(n, n_genomes, seed, rates5, rates3, interior, mu=44.0, sigma=0.2, lmin=30, lmax=150):
    """BASELINE configs[2]/[3] shape at test size: log-normal read lengths clipped to [30,150], several genomes at
    log-normal relative abundance (mixed coverage), terminal C->T / G->A damage with the given rates."""
    rng = np.random.default_rng(seed)
    acgt = np.frombuffer(b"ACGT", np.uint8)
    comp = np.zeros(256, np.uint8)
    for a, b in zip(b"ACGT", b"TGCA"):
        comp[a] = b
    sizes = np.exp(rng.uniform(np.log(4000), np.log(40000), n_genomes)).astype(int)
    genomes = [acgt[rng.integers(0, 4, s)] for s in sizes]
    ab = rng.lognormal(0, 1, n_genomes) * sizes
    which = rng.choice(n_genomes, n, p=ab / ab.sum())
    lens = np.clip(np.round(rng.lognormal(np.log(mu), sigma, n)), lmin, lmax).astype(int)
    seqs = []
    for i in range(n):
        g = genomes[which[i]]; L = int(lens[i])
        st = int(rng.integers(0, len(g) - L))
        r = g[st:st + L].copy()
        if rng.random() < 0.5:
            r = comp[r[::-1]]
        u = rng.random(L)
        p_ct = np.full(L, interior); p_ct[:5] = rates5
        p_ga = np.full(L, interior); p_ga[L - 5:] = rates3[::-1]
        ct = (r == ord("C")) & (u < p_ct); ga = (r == ord("G")) & (u < p_ga)
        r[ct] = ord("T"); r[ga] = ord("A")
        e = rng.random(L) < 0.001
        if e.any():
            r[e] = acgt[rng.integers(0, 4, int(e.sum()))]
        seqs.append(r.tobytes())
    return seqs


@pytest.mark.parametrize("name,rates,interior", [
    ("config 3 (dhigh damage)", (.329405, .221745, .187678, .161196, .144011), 0.01),
    ("config 4 (low damage)", (.05, .03, .02, .015, .01), 0.002)])
def test_metagenome_shaped_reads_vs_oracle(ctx, oracle, tmp_path, name, rates, interior):
    """BASELINE configs[2] and [3] at test size: reads of 30-150 bp (the 4- and 8-windows-per-lane extraction kernels, queries
    longer than the compact-column limit of the correction), mixed coverage from 1x to >100x, dhigh and low-damage profiles."""
    import carpedeam_b200 as cd
    hdr = "A>C\tA>G\tA>T\tC>A\tC>G\tC>T\tG>A\tG>C\tG>T\tT>A\tT>C\tT>G\n"
    rows5 = [[0.0] * 12 for _ in range(5)]; rows3 = [[0.0] * 12 for _ in range(5)]
    for i in range(5):
        rows5[i][5] = rates[i]; rows3[4 - i][6] = rates[i]
    pfx = str(tmp_path / "prof")
    for fn, rows in (("5p.prof", rows5), ("3p.prof", rows3)):
        with open(pfx + fn, "w") as f:
            f.write(hdr + "".join("\t".join("%g" % v for v in r) + "\n" for r in rows))
    # sigma 0.5 instead of the config's 0.2: at 12 k reads the 0.2 tail stops at ~90 bp, the wider one reaches the 150 bp clip
    s = SeqSet(_metagenome_reads(12000, 12, 31, np.array(rates), np.array(rates), interior, sigma=0.5))
    assert s.len.max() >= 140 and s.len.min() <= 32 and (s.len > 96).sum() > 100
    ctx.upload(s.blob, s.off, s.len, s.ext)
    ctx.set_damage(cd.load_damage(pfx))
    oracle.load_damage(pfx)
    hits, alns, out, changed = ctx.iteration(fetch_hits=True)
    ohits, r, line_off, oout = _oracle_iteration(oracle, s)
    _hits_equal(hits, ohits)
    for f in ("keep", "score", "bits", "qs", "qe", "ds", "de"):
        assert np.array_equal(alns[f], r[f]), f
    assert np.array_equal(out, oout)
    assert changed == int((out != s.blob).sum()) and changed > 0
    cover = np.bincount(hits["rep"], minlength=s.n)
    assert cover.max() > 50 and (cover == 0).sum() > 0      # deep and shallow genomes in one set


def _run_both(ctx, oracle, seqs, k=20):
    import carpedeam_b200 as cd
    s = SeqSet(seqs)
    ctx.upload(s.blob, s.off, s.len, s.ext)
    ctx.set_damage(cd.load_damage(os.path.join(GOLD, "dhigh")))
    oracle.load_damage(os.path.join(GOLD, "dhigh"))
    hits, alns, out, changed = ctx.iteration(kmer=cd.kmer_params(k=k), fetch_hits=True)
    ohits, r, line_off, oout = _oracle_iteration(oracle, s, k=k)
    _hits_equal(hits, ohits)
    assert np.array_equal(alns["line_off"], line_off)
    for f in ("keep", "score", "bits", "qs", "qe", "ds", "de"):
        assert np.array_equal(alns[f], r[f]), f
    assert np.array_equal(out, oout)
    return s, hits, alns, out, changed


def test_degenerate_sets(ctx, oracle):
    """edge cases: one sequence; every sequence shorter than k (no k-mer records at all, only identity records); reads that are
    all N; exact duplicates and reverse-complement duplicates only (identity-record groups, kmermatcher.cpp:244-262)."""
    rng = np.random.default_rng(3)
    acgt = np.frombuffer(b"ACGT", np.uint8)
    one = acgt[rng.integers(0, 4, 61)].tobytes()
    s, hits, alns, out, changed = _run_both(ctx, oracle, [one])
    assert len(hits["rep"]) == 0 and changed == 0 and alns["keep"].tolist() == [1]
    short = [acgt[rng.integers(0, 4, int(L))].tobytes() for L in rng.integers(1, 20, 300)]
    s, hits, alns, out, changed = _run_both(ctx, oracle, short)
    assert hits["R1"] == 300 and changed == 0                      # identity records only
    s, hits, alns, out, changed = _run_both(ctx, oracle, [b"N" * 40, b"N" * 40, one, b"NNNNNNNNNNNNNNNNNNNNNNNNA"])
    comp = bytes.maketrans(b"ACGT", b"TGCA")
    dup = [one] * 7 + [one.translate(comp)[::-1]] * 5 + [one[:40]] * 3
    s, hits, alns, out, changed = _run_both(ctx, oracle, dup)
    assert len(hits["rep"]) > 0


def test_deep_pileup_uses_wide_counters(ctx, oracle):
    """70 000 copies of one locus with damage: the representative's columns see more than 65 535 counted targets, so the
    16-bit counters of the fast correction kernel hand the query to the generic kernel (no test knob involved)."""
    rng = np.random.default_rng(8)
    acgt = np.frombuffer(b"ACGT", np.uint8)
    locus = acgt[rng.integers(0, 4, 64)]
    seqs = []
    for i in range(70000):
        r = locus.copy()
        u = rng.random(64)
        r[(r == ord("C")) & (u < 0.02)] = ord("T")
        if i % 9 == 0:
            r = r[: 50 + i % 14]
        seqs.append(r.tobytes())
    s, hits, alns, out, changed = _run_both(ctx, oracle, seqs)
    lo = alns["line_off"].astype(np.int64)
    assert (np.diff(lo)).max() > 65536
    # the compact path emits patches straight from the fast kernel and never builds bytes for a packed upload; a query for the
    # wide kernel makes it fall back to the byte mode — same result
    import carpedeam_b200 as cd
    words, nmask, _ = cd.pack_seqs(s.blob, s.off, s.len)
    ctx.upload_packed(words, nmask, s.len, s.ext)
    ip = cd.IterParams(cd.kmer_params(), cd.RescoreParams(0.9, 0.001, 0), cd.CorrParams(0.9, 0.99))
    ca, pa = cd.AlnsCompact(), cd.Patches()
    assert ctx.iteration_compact(ip, ca, pa) == changed
    assert np.array_equal(cd.apply_patches(s.blob, pa), out)


def test_empty_set_is_an_error_or_empty_result():
    """an empty sequence DB must not crash or hang: either a clean CdgError (with a message) or empty outputs"""
    import carpedeam_b200 as cd
    c = cd.Context(0)
    try:
        c.set_damage(cd.load_damage(os.path.join(GOLD, "dhigh")))
        try:
            c.upload(np.zeros(0, np.uint8), np.zeros(0, np.uint64), np.zeros(0, np.uint32), np.zeros(0, np.uint8))
            hits, alns, out, changed = c.iteration(fetch_hits=True)
            assert len(hits["rep"]) == 0 and changed == 0 and len(alns["tgt"]) == 0
        except cd.CdgError as e:
            assert len(str(e)) > 10
    finally:
        c.close()


def test_caller_provided_lists_are_validated(ctx):
    """stand-alone rescorediagonal / ancient_correction take lists parsed from text DBs: indices outside the resident set, broken
    offsets and alignments whose coordinates lie outside the sequences must be rejected or ignored, never dereferenced"""
    import carpedeam_b200 as cd
    s = SeqSet(synth_reads(50, 2000, 4))
    ctx.upload(s.blob, s.off, s.len, s.ext)
    ctx.set_damage(cd.load_damage(os.path.join(GOLD, "dhigh")))
    off = np.arange(s.n + 1, dtype=np.uint64)
    tgt = np.arange(s.n, dtype=np.uint32)
    bad = tgt.copy(); bad[7] = 1 << 30
    with pytest.raises(cd.CdgError):
        ctx.set_pref_lines(off, bad, np.zeros(s.n, np.int32), np.zeros(s.n, np.uint16))
    boff = off.copy(); boff[5] = 99
    with pytest.raises(cd.CdgError):
        ctx.set_pref_lines(boff, tgt, np.zeros(s.n, np.int32), np.zeros(s.n, np.uint16))
    with pytest.raises(cd.CdgError):
        ctx.set_alignments(off, bad, np.ones(s.n, np.float32), np.zeros(s.n, np.int32), s.len.astype(np.int32) - 1, np.zeros(s.n, np.int32),
                           s.len.astype(np.int32) - 1)
    # coordinates far outside the sequences: accepted as records, but never counted and never loaded
    two = np.repeat(np.arange(s.n + 1, dtype=np.uint64) * 2, 1)
    t2 = np.repeat(tgt, 2); t2[1::2] = (t2[1::2] + 1) % s.n
    qs = np.zeros(2 * s.n, np.int32); qe = np.repeat(s.len.astype(np.int32) - 1, 2); ds = qs.copy(); de = qe.copy()
    qs[1::2] = -5000; qe[1::2] = 100000; ds[1::2] = -7; de[1::2] = 2_000_000_000
    ctx.set_alignments(two, t2, np.ones(2 * s.n, np.float32), qs, qe, ds, de)
    out, changed = ctx.correct()
    assert changed == 0 and np.array_equal(out, s.blob)


def test_async_upload_pipeline_equals_synchronous_upload(oracle):
    """cdg_upload_seqs_async + cdg_iteration (chunked copy, per-chunk pack and extraction) must give exactly what the synchronous
    upload gives; also when the first consumer is not the extraction (stand-alone rescore on caller-provided lines)."""
    import torch
    import carpedeam_b200 as cd
    seqs = synth_reads(30000, 150000, 77, n_frac=0.001)
    seqs[5] = seqs[5].lower()
    s = SeqSet(seqs)
    c = cd.Context(0)
    try:
        c.set_damage(cd.load_damage(os.path.join(GOLD, "dhigh")))
        c.upload(s.blob, s.off, s.len, s.ext)
        hits, alns, out, changed = c.iteration(fetch_hits=True)
        pin = lambda a: torch.from_numpy(a).pin_memory()
        hb, ho, hl, he = pin(s.blob), pin(s.off.view(np.int64)), pin(s.len.view(np.int32)), pin(s.ext)
        for rep in range(2):
            c.upload_raw_async(hb.data_ptr(), hb.numel(), ho.data_ptr(), hl.data_ptr(), he.data_ptr(), s.n)
            hits2, alns2, out2, changed2 = c.iteration(fetch_hits=True)
            for f in ("rep", "tgt", "score", "diag"):
                assert np.array_equal(hits[f], hits2[f]), f
            for f in ("keep", "score", "bits", "qs", "qe", "ds", "de", "seqid"):
                assert np.array_equal(alns[f], alns2[f]), f
            assert np.array_equal(out, out2) and changed == changed2
        # entries NOT in index order in the data file (index sorted by key over a shuffled data file): same set, same answers
        perm = np.random.default_rng(1).permutation(s.n)
        blob2 = np.zeros_like(s.blob); off2 = np.zeros_like(s.off)
        pos = 0
        for i in perm.tolist():
            L = int(s.len[i]) + 2
            blob2[pos:pos + L] = s.blob[int(s.off[i]):int(s.off[i]) + L]; off2[i] = pos; pos += L
        hb2, ho2 = pin(blob2), pin(off2.view(np.int64))
        c.upload_raw_async(hb2.data_ptr(), hb2.numel(), ho2.data_ptr(), hl.data_ptr(), he.data_ptr(), s.n)
        hits4, alns4, out4, changed4 = c.iteration(fetch_hits=True)
        for f in ("rep", "tgt", "score", "diag"):
            assert np.array_equal(hits[f], hits4[f]), f
        assert changed4 == changed
        for i in range(0, s.n, 97):
            assert np.array_equal(out4[int(off2[i]):int(off2[i]) + int(s.len[i])], out[int(s.off[i]):int(s.off[i]) + int(s.len[i])])
        # first consumer = rescore on provided lines
        c.upload_raw_async(hb.data_ptr(), hb.numel(), ho.data_ptr(), hl.data_ptr(), he.data_ptr(), s.n)
        lo = alns["line_off"]
        q = np.repeat(np.arange(s.n), np.diff(lo.astype(np.int64)))
        pref = np.zeros(len(q), np.int32); dg = np.zeros(len(q), np.uint16)
        hp = np.arange(len(hits["rep"])) + hits["rep"].astype(np.int64) + 1
        pref[hp] = hits["score"]; dg[hp] = hits["diag"].astype(np.uint16)
        c.set_pref_lines(lo, alns["tgt"], pref, dg)
        alns3 = c.rescore()
        for f in ("keep", "score", "qs", "qe", "ds", "de"):
            assert np.array_equal(alns[f], alns3[f]), f
    finally:
        c.close()


@pytest.mark.parametrize("world", [2, 3])
@pytest.mark.parametrize("case", ["golden reads k=20", "golden contigs k=22 ext", "synthetic with N"])
def test_virtual_ranks_equal_unsplit_run(ctx, world, case):
    """hash-range passes on one GPU (virtual ranks = threads with one context each, multigpu.run_virtual_ranks): every rank sorts and
    groups one hash range, grouped records go to the rank owning the representative, the head run crosses rank boundaries —
    hits, alignment records and corrected bytes must equal the unsplit single-context run bit for bit.  This is the multi-GPU
    algorithm (tests/mg_worker.py runs it over NCCL / peer memory on 2-8 GPUs) inside the single-GPU suite."""
    import carpedeam_b200 as cd
    from carpedeam_b200.multigpu import run_virtual_ranks
    if case == "golden reads k=20":
        s, kp = SeqSet.from_golden("nucl_reads"), cd.kmer_params()
    elif case == "golden contigs k=22 ext":
        s, kp = SeqSet.from_golden("assembly_contigs_8_noneCycle"), cd.kmer_params(k=22, include_only_extendable=1)
    else:
        s, kp = SeqSet(synth_reads(20000, 100000, 21, n_frac=0.001)), cd.kmer_params()
    dmg = cd.load_damage(os.path.join(GOLD, "dhigh"))
    ctx.upload(s.blob, s.off, s.len, s.ext)
    ctx.set_damage(dmg)
    hits, alns, out, changed = ctx.iteration(kmer=kp, fetch_hits=True)
    parts = run_virtual_ranks(0, world, s.blob, s.off, s.len, s.ext, dmg, kmer=kp)
    for f in ("rep", "tgt", "score", "diag"):
        assert np.array_equal(np.concatenate([p["hits"][f] for p in parts]), hits[f]), f
    for f in ("tgt", "keep", "score", "bits", "qs", "qe", "ds", "de", "seqid"):
        assert np.array_equal(np.concatenate([p["alns"][f] for p in parts]), alns[f]), f
    merged = s.blob.copy()
    for p in parts:
        merged[p["range_offset"]:p["range_offset"] + p["range_bytes"]] = p["out"]
    assert np.array_equal(merged, out) and sum(p["changed"] for p in parts) == changed
    assert max(p["records_received"] for p in parts) < 0.75 * hits["R1"] or world == 1      # no rank sorted the whole record set


def _kept(alns):
    """the kept lines of a full alignment export, re-grouped by query (what cdh_expand_alns must reproduce)"""
    k = alns["keep"].astype(bool)
    lo = alns["line_off"].astype(np.int64)
    csum = np.concatenate([[0], np.cumsum(k)])
    out = {f: alns[f][k] for f in ("tgt", "score", "bits", "evalue", "seqid", "qs", "qe", "ds", "de")}
    out["line_off"] = csum[lo].astype(np.uint64)
    return out


@pytest.mark.parametrize("case", ["golden reads", "golden contigs k=22", "synthetic with N and lower case", "synthetic with N", "long contigs (32-bit fields)"])
def test_compact_outputs_equal_full_outputs(ctx, case):
    """cdg_iteration_compact (kept alignments as 14-byte records with implicit self lines, corrected reads as patches) expanded
    on the host == cdg_iteration's full-fidelity outputs, field for field; the packed upload (2-bit words + N mask in, DB bytes
    rebuilt on the device) == the byte upload of the same set."""
    import ctypes as C
    import carpedeam_b200 as cd
    kp = cd.kmer_params()
    if case == "golden reads":
        s = SeqSet.from_golden("nucl_reads")
    elif case == "golden contigs k=22":
        s, kp = SeqSet.from_golden("assembly_contigs_8_noneCycle"), cd.kmer_params(k=22, include_only_extendable=1)
    elif case == "synthetic with N and lower case":
        seqs = synth_reads(20000, 100000, 31, n_frac=0.002)
        seqs[7] = seqs[7].lower(); seqs[11] = b"N" * 40; seqs[12] = seqs[13]
        s = SeqSet(seqs)
    elif case == "synthetic with N":      # packed upload of a set with N: the device rebuilds the DB bytes (byte paths of rescore / correction)
        seqs = synth_reads(20000, 100000, 32, n_frac=0.003)
        seqs[11] = b"N" * 40; seqs[12] = seqs[13]
        s = SeqSet(seqs)
    else:
        rng = np.random.default_rng(5)
        g = np.frombuffer(b"ACGT", np.uint8)[rng.integers(0, 4, 120000)]
        seqs = [g[a:a + L].tobytes() for a, L in ((0, 40000), (30000, 50000), (70000, 45000), (100, 39000))] + synth_reads(2000, 100000, 6)
        s = SeqSet(seqs)
    ctx.set_damage(cd.load_damage(os.path.join(GOLD, "dhigh")))
    ctx.upload(s.blob, s.off, s.len, s.ext)
    hits, alns, out, changed = ctx.iteration(kmer=kp, fetch_hits=True)
    ip = cd.IterParams(kp, cd.RescoreParams(0.9, 0.001, 0), cd.CorrParams(0.9, 0.99))
    uploads = [("bytes", lambda: ctx.upload(s.blob, s.off, s.len, s.ext))]
    plain = all(set(s.seq(i)) <= set(b"ACGTN") for i in range(s.n))
    if plain:
        words, nmask, has_n = cd.pack_seqs(s.blob, s.off, s.len)
        uploads.append(("packed", lambda: ctx.upload_packed(words, nmask, s.len, s.ext)))
    else:
        with pytest.raises(cd.CdgError):
            cd.pack_seqs(s.blob, s.off, s.len)
    want = _kept(alns)
    for name, up in uploads:
        up()
        ca, pa = cd.AlnsCompact(), cd.Patches()
        ch = ctx.iteration_compact(ip, ca, pa)
        assert ca.width == (2 if s.len.max() < 32768 else 4)
        got = cd.expand_alns(ca, s.len)
        assert np.array_equal(got["line_off"], want["line_off"]), name
        for f in ("tgt", "score", "bits", "seqid", "qs", "qe", "ds", "de"):
            assert np.array_equal(got[f], want[f]), (name, f)
        assert np.allclose(got["evalue"], want["evalue"], rtol=1e-12, atol=0), name      # glibc vs CUDA exp/erfc
        assert ch == changed and pa.n == changed
        assert np.array_equal(cd.apply_patches(s.blob, pa), out), name
        # fewer bytes than the full export, and implicit self lines really are implied
        assert ca.n_alns <= int(alns["keep"].sum())
        if case == "golden reads":
            assert ca.n_alns == int(alns["keep"].sum()) - s.n
        if name == "packed":      # a byte-returning call after the packed upload: the library rebuilds the DB bytes on demand
            _, alns3, out3, ch3 = ctx.iteration(kmer=kp)
            assert np.array_equal(out3, out) and ch3 == changed and np.array_equal(alns3["keep"], alns["keep"])
    # stand-alone stage exports after cdg_rescore / cdg_correct
    ctx.upload(s.blob, s.off, s.len, s.ext)
    ctx.kmermatch(kp, fetch=False)
    ctx.rescore(fetch=False)
    got = cd.expand_alns(ctx.export_alns_compact(), s.len)
    # (the stand-alone stages do not round seqId through the alignment text, the kept lines are the same)
    for f in ("tgt", "score", "qs", "qe", "ds", "de", "seqid"):
        assert np.array_equal(got[f], want[f]), f
    out2, ch2 = ctx.correct()
    assert np.array_equal(cd.apply_patches(s.blob, ctx.export_patches()), out2)


def test_packed_upload_rejects_inconsistent_arguments(ctx):
    import carpedeam_b200 as cd
    s = SeqSet(synth_reads(100, 2000, 3))
    words, nmask, _ = cd.pack_seqs(s.blob, s.off, s.len)
    with pytest.raises(cd.CdgError):
        ctx.upload_packed(words[:-1], None, s.len)
    # dirty padding / N positions in the words are sanitised on the device: results equal the clean upload
    ctx.set_damage(cd.load_damage(os.path.join(GOLD, "dhigh")))
    ctx.upload_packed(words, nmask, s.len)
    h1 = ctx.kmermatch()
    dirty = words.copy()
    wo = np.concatenate([[0], np.cumsum((s.len.astype(np.int64) + 31) // 32)])
    for i in range(s.n):
        L = int(s.len[i])
        if L % 32:
            dirty[wo[i + 1] - 1] |= np.uint64((1 << (2 * (32 - L % 32))) - 1)
    ctx.upload_packed(dirty, nmask, s.len)
    h2 = ctx.kmermatch()
    for f in ("rep", "tgt", "score", "diag"):
        assert np.array_equal(h1[f], h2[f])


def test_multi_gpu_worker_under_torchrun():
    """tests/mg_worker.py (one process per GPU, NCCL, peer-memory and NCCL exchanges, byte and packed-block uploads, full and
    compact outputs == the single-GPU run) — runs whenever the box has at least two GPUs."""
    import subprocess
    import sys
    import torch
    n = min(torch.cuda.device_count(), 4)
    if n < 2:
        pytest.skip("needs >= 2 GPUs (the driver's 1-GPU box runs the same phases through ThreadComm: test_virtual_ranks_equal_unsplit_run)")
    here = os.path.dirname(os.path.abspath(__file__))
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(n), "--master-addr", "127.0.0.1",
                        "--master-port", "29613", os.path.join(here, "mg_worker.py")], capture_output=True, text=True, timeout=900)
    assert r.returncode == 0 and "MULTIGPU PARITY OK" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]


@pytest.mark.parametrize("seqdb,pref,aln,corr,k,only_ext", CASES)
def test_two_word_grouped_records_forced_on_golden_sets(ctx, monkeypatch, seqdb, pref, aln, corr, k, only_ext):
    """the two-word grouped records (rep in its own word: sets where 2*idBits + diagBits + 1 > 64) forced on the reference's own
    iterations: same pref DB as the reference wrote, incl. the rep-boundary bleed of config 1"""
    import carpedeam_b200 as cd
    monkeypatch.setenv("CDG_FORCE_WIDE_KEY2", "1")
    s = SeqSet.from_golden(seqdb)
    ctx.upload(s.blob, s.off, s.len, s.ext)
    got = ctx.kmermatch(cd.kmer_params(k=k, include_only_extendable=only_ext))
    keys, payloads, _ = load_golden(pref)
    mine = format_pref(s, got)
    assert all(mine[kk] == p for kk, p in zip(keys, payloads))
    id_bits = int(np.ceil(np.log2(s.n))); diag_bits = int(np.ceil(np.log2(2 * int(s.len.max()) + 1)))
    assert ctx.stats()["sort2_passes"] == -(-(id_bits + diag_bits) // 8) + -(-id_bits // 8)      # (id, diag) passes + rep passes of the two pair sorts


def test_config4_shaped_set_needs_more_than_64_grouped_record_bits(ctx, oracle):
    """BASELINE configs[3] shape (contig iterations of a 5*10^7-sequence set: k=22, include-only-extendable, ext-flagged contigs of
    8 kb and more next to reads): 2^25 + sequences give 26-bit ids, 8-16 kb contigs 16-bit diagonals -> 2*26 + 16 + 1 = 69 bits,
    the single-word grouped record does not exist.  Most of the 3.4*10^7 entries are 1-base fillers (no k-mers) so that the oracle
    finishes in seconds (13-base fillers, all different: no k-mer, no identity group); hits are compared in full."""
    import carpedeam_b200 as cd
    rng = np.random.default_rng(44)
    acgt = np.frombuffer(b"ACGT", np.uint8)
    comp = bytes.maketrans(b"ACGT", b"TGCA")
    contigs = [acgt[rng.integers(0, 4, L)].tobytes() for L in (16000, 9000, 8200, 12000)]
    real, ext = [], []
    for c in contigs:
        real.append(c); ext.append(1)
        for _ in range(600):
            L = int(rng.integers(35, 120)); st = int(rng.integers(-20, len(c) - L + 20))
            r = c[max(0, st):st + L]
            if st < 0:
                r = acgt[rng.integers(0, 4, -st)].tobytes() + r        # overhanging reads stay extendable
            if st + L > len(c):
                r = r + acgt[rng.integers(0, 4, st + L - len(c))].tobytes()
            if rng.random() < 0.5:
                r = r.translate(comp)[::-1]
            real.append(r); ext.append(0)
    # two overlapping contigs (long shared diagonal run) and the fillers in between: real sequences sit at both ends of the id range
    real.append(contigs[0][-3000:] + acgt[rng.integers(0, 4, 6000)].tobytes()); ext.append(1)
    half = len(real) // 2
    n_fill = (1 << 25) + 5
    head, tail = SeqSet(real[:half], ext=ext[:half]), SeqSet(real[half:], ext=ext[half:])
    s = SeqSet.__new__(SeqSet)
    s.n = head.n + n_fill + tail.n
    fill = np.zeros((n_fill, 15), np.uint8)
    idx = np.arange(n_fill, dtype=np.uint64)
    for j in range(13):
        fill[:, j] = acgt[((idx >> np.uint64(2 * j)) & np.uint64(3)).astype(np.int64)]
    fill[:, 13] = ord("\n")
    s.blob = np.concatenate([head.blob, fill.ravel(), tail.blob])
    s.len = np.concatenate([head.len, np.full(n_fill, 13, np.uint32), tail.len])
    s.off = np.zeros(s.n, np.uint64); s.off[1:] = np.cumsum(s.len[:-1].astype(np.uint64) + 2)
    s.ext = np.concatenate([head.ext, np.zeros(n_fill, np.uint8), tail.ext])
    s.keys = np.arange(s.n, dtype=np.uint32); s.total = int(s.len.sum())
    assert 2 * 26 + int(np.ceil(np.log2(2 * int(s.len.max()) + 1))) + 1 > 64
    ctx.upload(s.blob, s.off, s.len, s.ext)
    got = ctx.kmermatch(cd.kmer_params(k=22, include_only_extendable=1))
    want = oracle.kmermatch(s, k=22, only_ext=1)
    _hits_equal(got, want)
    assert len(got["rep"]) > 1000 and got["rep"].max() > n_fill and got["rep"].min() < half
    # the rest of the iteration on the same context (rescore + correction of 3.4*10^7 entries): the alignments of the real sequences
    ctx.set_damage(cd.load_damage(os.path.join(GOLD, "dhigh")))
    a = ctx.rescore()
    q = np.repeat(np.arange(s.n, dtype=np.uint32), np.diff(a["line_off"].astype(np.int64)))
    sel = q != a["tgt"]
    r = oracle.rescore(s, q[sel], a["tgt"][sel], got["score"], (got["diag"].astype(np.int64) & 0xFFFF).astype(np.uint16))
    for f in ("keep", "score", "qs", "qe", "ds", "de"):
        assert np.array_equal(a[f][sel], r[f]), f


@pytest.mark.parametrize("seqdb,pref,aln,corr,k,only_ext", CASES)
@pytest.mark.parametrize("passes", [3, 7])
def test_hash_range_passes_inside_the_call_equal_the_unsplit_result(ctx, seqdb, pref, aln, corr, k, only_ext, passes):
    """cdg_kmermatch with split_passes = M cuts the 16-bit hash space into M ranges from the 65536-bin histogram
    (setupKmerSplits, kmermatcher.cpp:742-784), extracts / sorts / groups range by range on ONE resident copy of the set and reduces
    once: the pref DB must be the one the reference wrote without splitting (its own forced 3-way split gives the same DBs on
    config 1)."""
    import carpedeam_b200 as cd
    s = SeqSet.from_golden(seqdb)
    ctx.upload(s.blob, s.off, s.len, s.ext)
    got = ctx.kmermatch(cd.kmer_params(k=k, include_only_extendable=only_ext, split_passes=passes))
    assert ctx.stats()["split_passes"] == passes
    keys, payloads, _ = load_golden(pref)
    mine = format_pref(s, got)
    assert all(mine[kk] == p for kk, p in zip(keys, payloads))
    whole = ctx.kmermatch(cd.kmer_params(k=k, include_only_extendable=only_ext, split_passes=1))
    _hits_equal(got, whole)


def test_hash_range_passes_on_repeats_long_contigs_and_wide_records(ctx, oracle, monkeypatch):
    """the split loop on the awkward inputs: reads with repeated k-mers (exact replay path), contigs with an active hash threshold,
    sequences shorter than k, N; once more with the two-word grouped records"""
    import carpedeam_b200 as cd
    rng = np.random.default_rng(77)
    acgt = np.frombuffer(b"ACGT", np.uint8)
    contig = acgt[rng.integers(0, 4, 30000)].tobytes()
    seqs = synth_reads(6000, 30000, 78, n_frac=0.002) + [contig, contig[1000:9000], b"ACGTACGTACGTACGTACGTACGTACGTACGTACGTACGTACGT", b"A" * 60, b"ACGT", b"N" * 50]
    seqs += [contig[i:i + 90] for i in range(0, 28000, 37)] + [(b"ACGGT" * 20)[:77], (b"ACGGT" * 20)[:77]]
    s = SeqSet(seqs)
    ctx.upload(s.blob, s.off, s.len, s.ext)
    want = oracle.kmermatch(s)
    for env in (None, "1"):
        if env:
            monkeypatch.setenv("CDG_FORCE_WIDE_KEY2", env)
        got = ctx.kmermatch(cd.kmer_params(split_passes=4))
        assert ctx.stats()["split_passes"] == 4
        _hits_equal(got, want)


@pytest.mark.parametrize("seqdb,k,only_ext", [("nucl_reads", 20, 0), ("assembly_contigs_8_noneCycle", 22, 1)])
def test_sub_range_call_is_one_split_of_the_reference(ctx, oracle, seqdb, k, only_ext):
    """hash_lo / hash_hi: the partial result of one split (kmermatcher.cpp:665-675; records filtered at :244 and :315)"""
    import carpedeam_b200 as cd
    s = SeqSet.from_golden(seqdb)
    ctx.upload(s.blob, s.off, s.len, s.ext)
    union = 0
    for lo, hi in ((0, 21845), (21846, 43690), (43691, 65535)):
        got = ctx.kmermatch(cd.kmer_params(k=k, include_only_extendable=only_ext, hash_lo=lo, hash_hi=hi))
        want = oracle.kmermatch(s, k=k, only_ext=only_ext, hash_lo=lo, hash_hi=hi)
        _hits_equal(got, want)
        union += len(got["rep"])
    assert union >= len(oracle.kmermatch(s, k=k, only_ext=only_ext)["rep"])      # a pair's k-mers are spread over the splits


def test_memory_driven_number_of_hash_range_passes(ctx):
    """split_passes = 0: the number of passes follows from the free device memory (kmermatcher.cpp:615-627).  With most of the HBM
    taken by someone else the call must split by itself and still return the unsplit result."""
    import torch
    import carpedeam_b200 as cd
    from bench import synth_reads_fast
    s = synth_reads_fast(2_000_000, 10_000_000, 5)
    ctx.upload(s["blob"], s["off"], s["len"], None)
    whole = ctx.kmermatch(cd.kmer_params(split_passes=1))
    assert ctx.stats()["split_passes"] == 1
    c2 = cd.Context(0)        # a fresh context: nothing reserved yet, so the free memory decides
    hog = None
    try:
        c2.upload(s["blob"], s["off"], s["len"], None)
        torch.cuda.synchronize()
        free, _ = torch.cuda.mem_get_info()
        hog = torch.empty(max(0, free - (2600 << 20)), dtype=torch.uint8, device="cuda")      # the unsplit arrays need ~3.3 GB
        got = c2.kmermatch(cd.kmer_params())
        assert c2.stats()["split_passes"] > 1
    finally:
        del hog
        c2.close()
        torch.cuda.empty_cache()
    _hits_equal(got, whole)


@pytest.mark.parametrize("tag,unsafe", [("safe", 0), ("unsafe", 1)])
def test_extension_scores_vs_reference_functions(ctx, tag, unsafe):
    """cdg_extension_scores (orientation, identity recompute, candidates, consensusCaller,
    updateSeqIdConsensusReads, r_s_pair) against what the reference's own functions returned on config 1's second iteration
    (tests/golden/extension_kat_*.txt.gz, oracle/ref_extension.cpp): flags, oriented coordinates, float identities and consensus
    strings bit-exact, likelihood scores within 1e-5 relative (north_star; measured ~1e-13)."""
    import carpedeam_b200 as cd
    from support import check_extension_against_kat, load_extension_kat
    s = SeqSet.from_golden("correction_1")
    akeys, apay, _ = load_golden("aln_1")
    off, t, sid, qs, qe, ds, de = parse_aln(s, apay)
    ctx.upload(s.blob, s.off, s.len, s.ext)
    ctx.set_alignments(off, t, sid, qs, qe, ds, de)
    ctx.set_pair_model(cd.load_pair_model(os.path.join(GOLD, "dhigh")))
    got = ctx.extension_scores(unsafe=unsafe, keys=s.keys)
    assert np.array_equal(got["consensus_off"], np.concatenate([[0], np.cumsum(3 * s.len.astype(np.uint64))]).astype(np.uint64))
    n_scored = check_extension_against_kat(s, off, got, load_extension_kat(tag), rtol=1e-5)
    assert n_scored > 4000


def test_extension_scores_vs_oracle_on_synthetic_sets(ctx, oracle):
    """same against the oracle on a synthetic iteration (reads with N, ext-flagged contigs as targets and queries, unsafe consensus
    with enough coverage to call flank bases): every line of every query"""
    import carpedeam_b200 as cd
    rng = np.random.default_rng(91)
    seqs = synth_reads(6000, 12000, 92, n_frac=0.001)
    acgt = np.frombuffer(b"ACGT", np.uint8)
    seqs += [acgt[rng.integers(0, 4, 900)].tobytes(), acgt[rng.integers(0, 4, 400)].tobytes()]
    ext = [0] * 6000 + [1, 1]
    s = SeqSet(seqs, ext=ext)
    ctx.set_damage(cd.load_damage(os.path.join(GOLD, "dhigh")))
    ctx.upload(s.blob, s.off, s.len, s.ext)
    hits, alns, out, changed = ctx.iteration(fetch_hits=True)
    k = alns["keep"].astype(bool)
    lo = alns["line_off"].astype(np.int64)
    off = np.concatenate([[0], np.cumsum(k)])[lo].astype(np.uint64)
    # the aln DB text rounds the identity to three decimals (Util.cpp:278-307): what ancient_read_assemble reads back
    sid = np.array([np.float32(float(seqid_text_bytes(x))) for x in alns["seqid"][k]], np.float32)
    args = (off, alns["tgt"][k], sid, alns["qs"][k], alns["qe"][k], alns["ds"][k], alns["de"][k])
    oracle.load_damage(os.path.join(GOLD, "dhigh"))
    ctx.set_alignments(*args)
    ctx.set_pair_model(cd.load_pair_model(os.path.join(GOLD, "dhigh")))
    for unsafe, min_cov in ((0, 5), (1, 2)):
        want = oracle.extension(s, *args, unsafe=unsafe, min_cov_safe=min_cov)
        got = ctx.extension_scores(unsafe=unsafe, min_cov_safe=min_cov)
        for f in ("flags", "qs", "qe", "ds", "de", "seqid", "ry", "max_left", "max_right", "consensus"):
            assert np.array_equal(got[f], want[f]), (unsafe, f)
        sc = (want["flags"] & 4) != 0
        assert sc.sum() > 1000
        assert np.allclose(got["slen"][sc], want["slen"][sc], rtol=1e-5, atol=0) and np.allclose(got["sratio"], want["sratio"], rtol=1e-5, atol=1e-300)
    assert (want["consensus"] != ord("N")).sum() > s.total      # the unsafe run called bases in the flanks


def _synth_contig_set(seed, n, genome_len, lmin=600, lmax=6000):
    """overlapping contigs of one random genome, both strands, light residual damage and errors, a few N (cf. oracle/make_merge_kat.py)"""
    rng = np.random.default_rng(seed)
    acgt = np.frombuffer(b"ACGT", np.uint8)
    comp = np.zeros(256, np.uint8)
    for a, b in zip(b"ACGTN", b"TGCAN"):
        comp[a] = b
    genome = acgt[rng.integers(0, 4, genome_len)]
    seqs = []
    for _ in range(n):
        L = int(rng.integers(lmin, lmax + 1))
        st = int(rng.integers(0, genome_len - L))
        r = genome[st:st + L].copy()
        if rng.random() < 0.5:
            r = comp[r[::-1]]
        err = (0.0, 0.001, 0.004, 0.012)[int(rng.integers(0, 4))]
        u = rng.random(L)
        r[(r == ord("C")) & (u < 0.004)] = ord("T")
        r[(r == ord("G")) & (u > 0.996)] = ord("A")
        e = rng.random(L) < err
        if e.any():
            r[e] = acgt[rng.integers(0, 4, int(e.sum()))]
        if rng.random() < 0.15:
            r[rng.integers(0, L, 3)] = ord("N")
        seqs.append(r.tobytes())
    return seqs


@pytest.mark.parametrize("tag", ["synth", "asm9", "asm5"])
def test_merge_scores_vs_reference_functions(ctx, tag):
    """cdg_merge_scores (orientation, identity recompute, merge candidates, consensusCaller, updateSeqIdConsensus,
    queue filter, ancientMatchCount / deamMatches) against what the reference's own functions returned on the contig iterations of
    config 1 and on a synthetic contig set (tests/golden/merge_kat_*.txt.gz, oracle/ref_merge.cpp), --unsafe 0 and 1 (and
    --min-merge-seq-id 0.97 on the synthetic set): flags, oriented coordinates, float identities, alnLengthCons bit-exact; deamMatch (a
    float built from double posteriors) bit-exact as well — the tolerance 1e-6 only guards the table-free parts of the arithmetic —
    and the host-side queue order (cdh_merge_queue_order) equal to the order the reference's priority queue popped."""
    import carpedeam_b200 as cd
    from support import check_merge_against_kat, load_merge_kat, load_merge_set
    s, (off, t, sid, qs, qe, ds, de) = load_merge_set(tag)
    ctx.upload(s.blob, s.off, s.len, s.ext)
    ctx.set_alignments(off, t, sid, qs, qe, ds, de)
    ctx.set_merge_model(cd.load_merge_model(os.path.join(GOLD, "dhigh")))
    for (unsafe, thr), kat in load_merge_kat(tag).items():
        got = ctx.merge_scores(min_merge_seq_id=thr, unsafe=unsafe, keys=s.keys)
        n_queued, n_exact = check_merge_against_kat(off, got, kat, 1e-6)
        assert n_queued > 100 and n_exact == n_queued, (tag, unsafe, thr, n_queued, n_exact)
        for q, k in kat.items():
            l0 = int(off[q])
            idx = [i for i, a in enumerate(k["lines"]) if a["queued"]]
            if idx:
                order = cd.merge_queue_order(got["aln_len_cons"][l0:][idx], got["deam"][l0:][idx])
                assert [idx[j] for j in order] == k["pops"], (tag, unsafe, thr, q)


def test_merge_scores_vs_oracle_on_synthetic_contigs(ctx, oracle):
    """same against the oracle on a larger synthetic contig iteration whose alignments come from the device's own k=22 kmermatch + rescore
    (600 contigs of 0.6-6 kb, both strands, N, mixed error rates; unsafe consensus with --min-cov-safe 2 so that flank bases are called)"""
    import carpedeam_b200 as cd
    seqs = _synth_contig_set(20260119, 600, 150000)
    s = SeqSet(seqs, ext=[1] * len(seqs))
    ctx.set_damage(cd.load_damage(os.path.join(GOLD, "dhigh")))
    ctx.upload(s.blob, s.off, s.len, s.ext)
    hits, alns, out, changed = ctx.iteration(kmer=cd.kmer_params(k=22, include_only_extendable=1), fetch_hits=True)
    k = alns["keep"].astype(bool)
    lo = alns["line_off"].astype(np.int64)
    off = np.concatenate([[0], np.cumsum(k)])[lo].astype(np.uint64)
    args = (off, alns["tgt"][k], alns["seqid"][k], alns["qs"][k], alns["qe"][k], alns["ds"][k], alns["de"][k])
    # the merge reads the corrected sequences (correction_N): make them the resident set, as the workflow does
    s2 = SeqSet([bytes(out[int(o):int(o) + int(l)]) for o, l in zip(s.off, s.len)], ext=[1] * len(seqs))
    ctx.upload(s2.blob, s2.off, s2.len, s2.ext)
    ctx.set_alignments(*args)
    oracle.load_damage(os.path.join(GOLD, "dhigh"))
    ctx.set_merge_model(cd.load_merge_model(os.path.join(GOLD, "dhigh")))
    oargs = (args[0], args[1]) + args[3:]
    for unsafe, min_cov, thr in ((0, 5, 0.99), (1, 2, 0.99), (1, 1, 0.97)):
        want = oracle.merge(s2, *oargs, unsafe=unsafe, merge_thr=np.float32(thr), min_cov_safe=min_cov)
        got = ctx.merge_scores(min_merge_seq_id=thr, unsafe=unsafe, min_cov_safe=min_cov)
        for f in ("flags", "qs", "qe", "ds", "de", "seqid0", "ry0", "seqid", "ry", "aln_len_cons"):
            assert np.array_equal(got[f], want[f]), (unsafe, f)
        qd = (want["flags"] & 4) != 0
        assert qd.sum() > 300
        assert np.allclose(got["deam"], want["deam"], rtol=1e-6, atol=0)
        assert (got["deam"] == want["deam"]).all()


@pytest.mark.parametrize("wrapped", [0, 1])
def test_hamming_rescore_vs_reference(ctx, oracle, wrapped):
    """cdg_rescore_hamming (rescorediagonal --rescore-mode 0, linclust's pre-clustering step, with and without
    --wrapped-scoring) on the prefilter lines the reference's linclust kmermatcher wrote for config 1's last sequence set: every entry
    equals what the reference binary wrote (tests/golden/cfg1_lin_rescore[_wrapped]_9.gz), every line equals the oracle."""
    from support import format_hamming, load_hamming_golden
    s, (q, t, sc, dg, off), want = load_hamming_golden()
    ctx.upload(s.blob, s.off, s.len, s.ext)
    ctx.set_pref_lines(off, t, sc, dg)
    got = ctx.rescore_hamming(wrapped=wrapped)
    ref = oracle.rescore_hamming(s, q, t, sc, dg, wrapped=wrapped)
    for f in ("keep", "score", "diag"):
        assert np.array_equal(got[f], ref[f]), f
    k = ref["keep"].astype(bool)
    assert np.array_equal(got["seqid"][k], ref["seqid"][k]) and got["n_kept"] == int(k.sum())
    mine = format_hamming(s, q, got["tgt"], got["keep"], got["score"], got["diag"])
    assert set(mine) == set(want[wrapped])
    bad = [kk for kk in mine if mine[kk] != want[wrapped][kk]]
    assert not bad, (len(bad), mine[bad[0]], want[wrapped][bad[0]])
    assert int(k.sum()) > 9000 and (ref["score"] < 0).sum() > 500


def test_hamming_rescore_on_ragged_sets_vs_oracle(ctx, oracle):
    """same against the oracle on the device's own prefilter lines: reads with N and lower-case / IUPAC letters (byte path), a 70 kb and a
    40 kb contig sharing a 30 kb overlap (diagonals beyond 16 bits), every coverage mode and identity mode, wrapped and not"""
    import carpedeam_b200 as cd
    rng = np.random.default_rng(77)
    seqs = synth_reads(3000, 9000, 78, n_frac=0.002)
    for i in range(0, 3000, 17):
        b = bytearray(seqs[i]); b[len(b) // 2] = ord("acgtRYn"[i % 7]); seqs[i] = bytes(b)
    acgt = np.frombuffer(b"ACGT", np.uint8)
    g = acgt[rng.integers(0, 4, 80000)]
    seqs += [g[:70000].tobytes(), g[40000:80000].tobytes(), g[100:900].tobytes()]
    s = SeqSet(seqs)
    ctx.upload(s.blob, s.off, s.len, s.ext)
    hits = ctx.kmermatch(cd.kmer_params(k=17, kmers_per_seq=60))
    assert len(hits["rep"]) > 3000
    first = None
    for wrapped in (0, 1):
        for cov_mode, cov_thr, sid_mode, min_len in ((0, 0.8, 0, 0), (1, 0.5, 1, 0), (2, 0.5, 2, 40), (3, 0.6, 0, 0), (4, 0.6, 0, 0), (5, 0.7, 0, 0)):
            got = ctx.rescore_hamming(cov_thr=cov_thr, cov_mode=cov_mode, min_seq_id=0.9, seq_id_mode=sid_mode, min_aln_len=min_len, wrapped=wrapped)
            lo = got["line_off"].astype(np.int64)
            q = np.repeat(np.arange(s.n, dtype=np.uint32), np.diff(lo))
            if first is None:
                first = (q, got["tgt"].copy())
                # prefilter lines as cdg_rescore sees them: self line first, then the hits of the representative
                sc = np.zeros(len(q), np.int32); dg = np.zeros(len(q), np.uint16)
                hit_lines = q != got["tgt"]
                assert hit_lines.sum() == len(hits["rep"])
                sc[hit_lines] = hits["score"]; dg[hit_lines] = hits["diag"].astype(np.uint16)
            ref = oracle.rescore_hamming(s, first[0], first[1], sc, dg, wrapped=wrapped, cov_thr=cov_thr, cov_mode=cov_mode, seqid_mode=sid_mode, aln_len_thr=min_len)
            for f in ("keep", "score", "diag"):
                assert np.array_equal(got[f], ref[f]), (wrapped, cov_mode, f)
            assert ref["keep"].sum() > s.n


@pytest.mark.parametrize("lo", [55, 40])
def test_identity_sort_fixup_orders_mixed_runs(ctx, oracle, monkeypatch, lo):
    """the identity records are radix-sorted by key bits [23, 63) and the runs that agree there but not below are ordered by a fix-up
    kernel: with the test knob CDG_ID_SORT_LO the radix part shrinks to a few bits, so that every run is mixed and goes through the
    fix-up (duplicated sequences included: equal full keys must keep their (length desc, id asc) order) — the hit list must not change"""
    import carpedeam_b200 as cd
    seqs = synth_reads(3000, 6000, 123)
    seqs += seqs[100:400] + seqs[100:200]      # identical sequences: identity groups of 2 and 3
    s = SeqSet(seqs)
    ctx.upload(s.blob, s.off, s.len, s.ext)
    monkeypatch.setenv("CDG_ID_SORT_LO", str(lo))
    got = ctx.kmermatch()
    monkeypatch.delenv("CDG_ID_SORT_LO")
    _hits_equal(got, ctx.kmermatch())
    _hits_equal(got, oracle.kmermatch(s))


def test_cyclecheck_vs_reference(ctx, oracle):
    """cdg_cyclecheck against the split diagonals the reference's cyclecheck reported (tests/golden/cycle_kat.txt, chopped entry
    lengths of `carpedeam cyclecheck --chop-cycle 1`) on linear contigs, circular genomes read past their start (exact / mutated overlaps), tandem
    repeats, homopolymers, N runs and sequences around the three-k-mer limit; and against the oracle with other k and a --max-seq-len that skips"""
    from support import load_cycle_kat
    s, want = load_cycle_kat()
    ctx.upload(s.blob, s.off, s.len, s.ext)
    got = ctx.cyclecheck()
    assert np.array_equal(got, want), np.nonzero(got != want)[0][:10]
    assert (want > 0).sum() > 50
    for k, mx in ((22, 4000), (15, 200000), (27, 200000)):
        assert np.array_equal(ctx.cyclecheck(k=k, max_seq_len=mx), oracle.cyclecheck(s, k=k, max_seq_len=mx)), (k, mx)
    # the k-mer matcher still works on the same context afterwards (cyclecheck reuses its record arrays)
    _hits_equal(ctx.kmermatch(), oracle.kmermatch(s))


@pytest.mark.parametrize("tag,cov_thr,cov_mode", [("c099m1", 0.99, 1), ("c08m0", 0.8, 0)])
def test_kmermatch_coverage_filter_vs_reference(ctx, oracle, tag, cov_thr, cov_mode):
    """cdg_kmermatch with linclust's coverage filter (cov_thr / cov_mode = -c / --cov-mode, kmermatcher.cpp:526-529) against
    the pref DB the reference's k-mer matcher wrote with the same setting, and against the oracle hit by hit"""
    import carpedeam_b200 as cd
    s = SeqSet.from_golden("correction_9")
    ctx.upload(s.blob, s.off, s.len, s.ext)
    got = ctx.kmermatch(cd.kmer_params(k=20, cov_thr=cov_thr, cov_mode=cov_mode))
    oracle.set_kmermatch_cov(np.float32(cov_thr), cov_mode)
    try:
        _hits_equal(got, oracle.kmermatch(s, k=20))
    finally:
        oracle.set_kmermatch_cov()
    mine = format_pref(s, got)
    gk, gp, _ = load_golden("lin_pref_%s_9" % tag)
    assert not [k for k, p in zip(gk, gp) if mine[k] != p]
