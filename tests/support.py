"""Test support: golden-DB loading, oracle (ctypes) wrapper, synthetic read generator, text formats.

TEST INFRASTRUCTURE — this module is the only place the oracle library is loaded from Python
(tests/, __graft_entry__.smoke(), bench.py's cpu_baseline leg).
"""
import ctypes as C
import gzip
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")
ORACLE_DIR = os.path.join(ROOT, "oracle")
ORACLE_SO = os.path.join(ORACLE_DIR, "_build", "libcarpedeam_oracle.so")


# ------------------------------------------------------------------------------------------------
# canonical golden DBs  (format: oracle/make_golden.py)
def load_golden(name, prefix="cfg1_"):
    """-> (keys[int], payloads[bytes], exts[int]) sorted by key."""
    blob = gzip.open(os.path.join(GOLD, "%s%s.gz" % (prefix, name)), "rb").read()
    keys, payloads, exts = [], [], []
    pos = 0
    while pos < len(blob):
        nl = blob.index(b"\n", pos)
        k, e = blob[pos:nl].split(b"\t")
        end = blob.index(b"\0", nl + 1)
        keys.append(int(k)); exts.append(int(e)); payloads.append(blob[nl + 1:end])
        pos = end + 1
    return keys, payloads, exts


class SeqSet:
    """Flat sequence set as the C-ABI takes it: MMseqs2 data-file layout `BASES\\n\\0` per entry
    (lib/mmseqs/src/commons/DBReader.h)."""

    def __init__(self, seqs, keys=None, ext=None):
        n = len(seqs)
        self.n = n
        self.keys = np.arange(n, dtype=np.uint32) if keys is None else np.asarray(keys, dtype=np.uint32)
        self.ext = np.zeros(n, dtype=np.uint8) if ext is None else np.asarray(ext, dtype=np.uint8)
        self.len = np.array([len(s) for s in seqs], dtype=np.uint32)
        self.off = np.zeros(n, dtype=np.uint64)
        if n:
            self.off[1:] = np.cumsum(self.len[:-1].astype(np.uint64) + 2)
        self.blob = np.frombuffer(b"".join(s + b"\n\0" for s in seqs), dtype=np.uint8).copy()
        self.total = int(self.len.sum())

    @classmethod
    def from_golden(cls, name, prefix="cfg1_"):
        keys, payloads, exts = load_golden(name, prefix)
        assert keys == sorted(keys)
        return cls([p.rstrip(b"\n") for p in payloads], keys, exts)

    def seq(self, i):
        o = int(self.off[i])
        return self.blob[o:o + int(self.len[i])].tobytes()

    def key_to_index(self):
        return {int(k): i for i, k in enumerate(self.keys)}


def load_extension_kat(tag):
    """tests/golden/extension_kat_<tag>.txt.gz (oracle/ref_extension.cpp): per query with >= 2 alignment lines what the reference's
    consensusCaller / updateSeqIdConsensusReads / r_s_pair return -> {query index: dict}"""
    out = {}
    cur = None
    for line in gzip.open(os.path.join(GOLD, "extension_kat_%s.txt.gz" % tag), "rt"):
        f = line.rstrip("\n").split("\t")
        if f[0] == "Q":
            cur = dict(max_left=int(f[3]), max_right=int(f[4]), consensus=f[5].encode(), lines=[])
            out[int(f[1])] = cur
        else:
            cur["lines"].append(dict(cand=int(f[2]), rev=int(f[3]), qs=int(f[4]), qe=int(f[5]), ds=int(f[6]), de=int(f[7]), seqid=np.float32(f[8]), ry=np.float32(f[9]),
                                     scored=int(f[10]), slen=float(f[11]), sratio=float(f[12])))
    return out


def check_extension_against_kat(s, aln_off, got, kat, rtol):
    """got: dict of per-line arrays (flags, qs, qe, ds, de, seqid, ry, slen, sratio) + per-query max_left / max_right / consensus"""
    cons_off = np.concatenate([[0], np.cumsum(3 * s.len.astype(np.int64))])
    n_scored = 0
    for q, k in kat.items():
        l0 = int(aln_off[q])
        assert int(aln_off[q + 1]) - l0 == len(k["lines"]), q
        assert got["max_left"][q] == k["max_left"] and got["max_right"][q] == k["max_right"], q
        assert got["consensus"][cons_off[q]:cons_off[q + 1]].tobytes() == k["consensus"], q
        for i, a in enumerate(k["lines"]):
            l = l0 + i
            fl = int(got["flags"][l])
            assert (fl & 1) == a["cand"] and ((fl >> 1) & 1) == a["rev"] and ((fl >> 2) & 1) == a["scored"], (q, i, fl, a)
            assert (got["qs"][l], got["qe"][l], got["ds"][l], got["de"][l]) == (a["qs"], a["qe"], a["ds"], a["de"]), (q, i)
            assert got["seqid"][l] == a["seqid"] and got["ry"][l] == a["ry"], (q, i, got["seqid"][l], a["seqid"], got["ry"][l], a["ry"])
            if a["scored"]:
                n_scored += 1
                assert np.isclose(got["slen"][l], a["slen"], rtol=rtol, atol=0) and np.isclose(got["sratio"][l], a["sratio"], rtol=rtol, atol=1e-300), (q, i)
    return n_scored


MERGE_SETS = {"asm5": ("correction_5", "aln_asm_5", "cfg1_"), "asm9": ("correction_9", "aln_asm_9", "cfg1_"), "synth": ("synth_contigs", "synth_contigs_aln", "")}


def load_merge_kat(tag):
    """tests/golden/merge_kat_<tag>.txt.gz (oracle/ref_merge.cpp): per run (unsafe, --min-merge-seq-id) and per query with >= 2 alignment lines what the
    reference's consensusCaller / updateSeqIdConsensus / ancientMatchCount return and the order its priority queue pops the queued lines in
    -> {(unsafe, thr): {query index: dict}}"""
    out = {}
    run = cur = None
    for line in gzip.open(os.path.join(GOLD, "merge_kat_%s.txt.gz" % tag), "rt"):
        f = line.rstrip("\n").split("\t")
        if f[0] == "H":
            run = out.setdefault((int(f[1]), float(f[2])), {})
        elif f[0] == "Q":
            cur = dict(cons_hash=int(f[3], 16), lines=[], pops=[])
            run[int(f[1])] = cur
        elif f[0] == "A":
            cur["lines"].append(dict(cand=int(f[2]), rev=int(f[3]), qs=int(f[4]), qe=int(f[5]), ds=int(f[6]), de=int(f[7]), seqid0=np.float32(f[8]), ry0=np.float32(f[9]),
                                     seqid=np.float32(f[10]), ry=np.float32(f[11]), aln_len_cons=int(f[12]), queued=int(f[13]), deam=np.float32(f[14])))
        else:
            cur["pops"] = [int(x) for x in f[1:]]
    return out


def load_merge_set(tag):
    """sequence set + alignment CSR arrays (parse_aln) of one merge_kat set"""
    seqname, alnname, prefix = MERGE_SETS[tag]
    s = SeqSet.from_golden(seqname, prefix)
    akeys, apay, _ = load_golden(alnname, prefix)
    amap = dict(zip(akeys, apay))
    return s, parse_aln(s, [amap.get(int(k), b"") for k in s.keys])


def check_merge_against_kat(aln_off, got, kat, deam_rtol):
    """got: per-line arrays flags, qs, qe, ds, de, seqid0, ry0, seqid, ry, aln_len_cons, deam (+ optional per-query cons_hash).
    -> (queued lines, lines whose deamMatch is bit-equal)"""
    n_queued = n_exact = 0
    for q, k in kat.items():
        l0 = int(aln_off[q])
        assert int(aln_off[q + 1]) - l0 == len(k["lines"]), q
        if "cons_hash" in got:
            assert int(got["cons_hash"][q]) == k["cons_hash"], q
        for i, a in enumerate(k["lines"]):
            l = l0 + i
            fl = int(got["flags"][l])
            assert (fl & 1) == a["cand"] and ((fl >> 1) & 1) == a["rev"] and ((fl >> 2) & 1) == a["queued"], (q, i, fl, a)
            assert (got["qs"][l], got["qe"][l], got["ds"][l], got["de"][l]) == (a["qs"], a["qe"], a["ds"], a["de"]), (q, i)
            assert got["seqid0"][l] == a["seqid0"] and got["ry0"][l] == a["ry0"], (q, i, got["seqid0"][l], a["seqid0"])
            assert got["seqid"][l] == a["seqid"] and got["ry"][l] == a["ry"], (q, i, got["seqid"][l], a["seqid"], got["ry"][l], a["ry"])
            assert int(got["aln_len_cons"][l]) == a["aln_len_cons"], (q, i)
            if a["queued"]:
                n_queued += 1
                n_exact += int(got["deam"][l] == a["deam"])
                assert np.isclose(got["deam"][l], a["deam"], rtol=deam_rtol, atol=0), (q, i, got["deam"][l], a["deam"])
    return n_queued, n_exact


def load_cycle_kat():
    """tests/golden/synth_cycle.gz + cycle_kat.txt (oracle/make_cycle_golden.py): sequence set and the split diagonal the reference's cyclecheck
    reported per key (0 = not circular)"""
    s = SeqSet.from_golden("synth_cycle", "")
    want = np.zeros(s.n, np.uint32)
    for line in open(os.path.join(GOLD, "cycle_kat.txt")):
        k, d = line.split()
        want[int(k)] = int(d)
    return s, want


def load_pairlik_kat():
    """tests/golden/pairlik_kat.txt (oracle/ref_pairlik.cpp): outputs of the reference's calcLikelihood[Consensus]."""
    out = []
    for line in open(os.path.join(GOLD, "pairlik_kat.txt")):
        f = line.rstrip("\n").split("\t")
        if f[0] == "L":
            out.append(dict(mode="L", rev=int(f[1]), qs=int(f[2]), qe=int(f[3]), ds=int(f[4]), de=int(f[5]), max_aln=int(f[6]),
                            query=f[7].encode(), target=f[8].encode(), slen=float(f[9]), sratio=float(f[10])))
        else:
            out.append(dict(mode="C", rev=int(f[1]), qlen=int(f[2]), qs=int(f[3]), qe=int(f[4]), ds=int(f[5]), de=int(f[6]),
                            max_aln=int(f[7]), query=f[8].encode(), target=f[9].encode(), slen=float(f[10]), sratio=float(f[11])))
    return out


# ------------------------------------------------------------------------------------------------
def build_oracle():
    subprocess.run(["make", "-s", "-C", ORACLE_DIR], check=True)
    return ORACLE_SO


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t))


class Oracle:
    def __init__(self):
        if not os.path.exists(ORACLE_SO):
            build_oracle()
        L = C.CDLL(ORACLE_SO)
        L.orc_kmermatch.restype = C.c_long
        L.orc_kmermatch.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_int, C.c_float,
                                    C.c_uint64, C.c_int, C.c_int]
        L.orc_xxh64_u64.restype = C.c_uint64
        L.orc_xxh64_u64.argtypes = [C.c_uint64, C.c_uint64]
        L.orc_revcomp.restype = C.c_uint64
        L.orc_revcomp.argtypes = [C.c_uint64, C.c_int]
        L.orc_evalue.argtypes = [C.c_double, C.c_double, C.c_double, C.POINTER(C.c_double), C.POINTER(C.c_double)]
        L.orc_load_damage.argtypes = [C.c_char_p, C.c_char_p]
        self.L = L

    def set_kmermatch_cov(self, cov_thr=0.0, cov_mode=1):
        """-c / --cov-mode for the following kmermatch calls (the assembler's setting is the default)"""
        self.L.orc_set_kmermatch_cov.restype = None
        self.L.orc_set_kmermatch_cov(C.c_float(cov_thr), C.c_int(cov_mode))

    def kmermatch(self, s, k=20, kmers_per_seq=200, scale=0.2, seed=67, only_ext=0, keep_records=False, hash_lo=0, hash_hi=65535):
        """hash_lo / hash_hi: one split of the reference's split loop (kmermatcher.cpp:665-675)"""
        self.L.orc_kmermatch_range.restype = C.c_long
        self.L.orc_kmermatch_range.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_int, C.c_float, C.c_uint64, C.c_int, C.c_int,
                                               C.c_uint, C.c_uint]
        H = self.L.orc_kmermatch_range(s.blob.ctypes.data, s.off.ctypes.data, s.len.ctypes.data, s.n, k, kmers_per_seq, scale,
                                       seed, only_ext, 1 if keep_records else 0, hash_lo, hash_hi)
        rep = np.zeros(H, np.uint32); tgt = np.zeros(H, np.uint32); sc = np.zeros(H, np.int32); dg = np.zeros(H, np.int16)
        self.L.orc_km_hits(rep.ctypes.data, tgt.ctypes.data, sc.ctypes.data, dg.ctypes.data)
        r1 = C.c_uint64(); r2 = C.c_uint64()
        self.L.orc_km_counts(C.byref(r1), C.byref(r2))
        return dict(rep=rep, tgt=tgt, score=sc, diag=dg, R1=r1.value, R2=r2.value)

    def km_records(self, stage, count):
        kmer = np.zeros(count, np.uint64); ids = np.zeros(count, np.uint32)
        sl = np.zeros(count, np.int32); pos = np.zeros(count, np.int32)
        self.L.orc_km_records(stage, kmer.ctypes.data, ids.ctypes.data, sl.ctypes.data, pos.ctypes.data)
        return kmer, ids, sl, pos

    def evalue(self, score, qlen, dbres):
        e = C.c_double(); b = C.c_double()
        self.L.orc_evalue(score, qlen, dbres, C.byref(e), C.byref(b))
        return e.value, b.value

    def rescore(self, s, q_idx, t_idx, pref_score, diag16, db_res=None, eval_thr=0.001, seqid_thr=0.9):
        M = len(q_idx)
        q_idx = np.ascontiguousarray(q_idx, np.uint32); t_idx = np.ascontiguousarray(t_idx, np.uint32)
        pref_score = np.ascontiguousarray(pref_score, np.int32); diag16 = np.ascontiguousarray(diag16, np.uint16)
        out = dict(keep=np.zeros(M, np.uint8), score=np.zeros(M, np.int32), bits=np.zeros(M, np.int32),
                   evalue=np.zeros(M, np.float64), seqid=np.zeros(M, np.float32), qs=np.zeros(M, np.int32),
                   qe=np.zeros(M, np.int32), ds=np.zeros(M, np.int32), de=np.zeros(M, np.int32),
                   idcnt=np.zeros(M, np.int32), alnlen=np.zeros(M, np.int32))
        if db_res is None:
            db_res = s.total
        self.L.orc_rescore(C.c_void_p(s.blob.ctypes.data), C.c_void_p(s.off.ctypes.data), C.c_void_p(s.len.ctypes.data),
                           C.c_size_t(s.n), C.c_void_p(q_idx.ctypes.data), C.c_void_p(t_idx.ctypes.data),
                           C.c_void_p(pref_score.ctypes.data), C.c_void_p(diag16.ctypes.data), C.c_size_t(M),
                           C.c_double(db_res), C.c_double(eval_thr), C.c_float(seqid_thr),
                           *[C.c_void_p(out[k].ctypes.data) for k in
                             ("keep", "score", "bits", "evalue", "seqid", "qs", "qe", "ds", "de", "idcnt", "alnlen")])
        return out

    def rescore_hamming(self, s, q_idx, t_idx, pref_score, diag16, wrapped=0, cov_thr=0.8, cov_mode=0, seqid_thr=0.9, seqid_mode=0, aln_len_thr=0):
        """rescorediagonal --rescore-mode 0 per prefilter line (orc_rescore_hamming)"""
        M = len(q_idx)
        q_idx = np.ascontiguousarray(q_idx, np.uint32); t_idx = np.ascontiguousarray(t_idx, np.uint32)
        pref_score = np.ascontiguousarray(pref_score, np.int32); diag16 = np.ascontiguousarray(diag16, np.uint16)
        out = dict(keep=np.zeros(M, np.uint8), score=np.zeros(M, np.int32), diag=np.zeros(M, np.int32), seqid=np.zeros(M, np.float32))
        self.L.orc_rescore_hamming.restype = None
        self.L.orc_rescore_hamming(C.c_void_p(s.blob.ctypes.data), C.c_void_p(s.off.ctypes.data), C.c_void_p(s.len.ctypes.data), C.c_void_p(q_idx.ctypes.data),
                                   C.c_void_p(t_idx.ctypes.data), C.c_void_p(pref_score.ctypes.data), C.c_void_p(diag16.ctypes.data), C.c_size_t(M),
                                   C.c_int(wrapped), C.c_float(cov_thr), C.c_int(cov_mode), C.c_float(seqid_thr), C.c_int(seqid_mode), C.c_int(aln_len_thr),
                                   *[C.c_void_p(out[k].ctypes.data) for k in ("keep", "score", "diag", "seqid")])
        return out

    def cyclecheck(self, s, k=22, max_seq_len=200000):
        out = np.zeros(s.n, np.uint32)
        self.L.orc_cyclecheck.restype = None
        self.L.orc_cyclecheck(C.c_void_p(s.blob.ctypes.data), C.c_void_p(s.off.ctypes.data), C.c_void_p(s.len.ctypes.data), C.c_size_t(s.n), C.c_int(k),
                              C.c_uint(max_seq_len), C.c_void_p(out.ctypes.data))
        return out

    def load_damage(self, prefix):
        if prefix:
            self.L.orc_load_damage((prefix + "5p.prof").encode(), (prefix + "3p.prof").encode())
        else:
            self.L.orc_load_damage(b"", b"")

    def get_damage(self):
        D = np.zeros((11, 4, 4)); R = np.zeros((11, 4, 4))
        self.L.orc_get_damage(C.c_void_p(D.ctypes.data), C.c_void_p(R.ctypes.data))
        return D, R

    def pair_loglik(self, query, target, rev, q_start, db_start, aln_len, max_aln, seq_err=0.001, rand=0.85, excess=0.0625):
        """calcLikelihood (nuclassembleUtil.cpp:95-200); target = stored bytes, coordinates in the oriented frame."""
        a = C.c_double(); b = C.c_double()
        self.L.orc_pair_loglik(query, target, C.c_uint(len(target)), C.c_int(rev), C.c_int(q_start), C.c_int(db_start),
                               C.c_uint(aln_len), C.c_uint(max_aln), C.c_double(seq_err), C.c_float(rand), C.c_float(excess),
                               C.byref(a), C.byref(b))
        return a.value, b.value

    def pair_loglik_consensus(self, consensus, q_len, target, rev, q_start, q_end, db_start, db_end, max_aln,
                              seq_err=0.001, rand=0.85, excess=0.0625):
        """calcLikelihoodConsensus (nuclassembleUtil.cpp:203-374)."""
        a = C.c_double(); b = C.c_double()
        self.L.orc_pair_loglik_consensus(consensus, C.c_uint(q_len), target, C.c_uint(len(target)), C.c_int(rev), C.c_int(q_start),
                                         C.c_int(q_end), C.c_int(db_start), C.c_int(db_end), C.c_uint(q_end - q_start + 1),
                                         C.c_uint(max_aln), C.c_double(seq_err), C.c_float(rand), C.c_float(excess),
                                         C.byref(a), C.byref(b))
        return a.value, b.value

    def extension(self, s, aln_off, t_idx, seqid, qs, qe, ds, de, unsafe=0, seqid_thr=0.9, ry_thr=0.99, min_cov_safe=5, seq_err=0.001, rand=0.85,
                  excess=0.0625):
        """scoring front half of ancient_read_assemble (orc_extension) over all queries; arrays per alignment line"""
        aln_off = np.ascontiguousarray(aln_off, np.uint64); t_idx = np.ascontiguousarray(t_idx, np.uint32); seqid = np.ascontiguousarray(seqid, np.float32)
        qs, qe, ds, de = [np.ascontiguousarray(x, np.int32) for x in (qs, qe, ds, de)]
        M = len(t_idx)
        out = dict(flags=np.zeros(M, np.uint8), qs=np.zeros(M, np.int32), qe=np.zeros(M, np.int32), ds=np.zeros(M, np.int32), de=np.zeros(M, np.int32),
                   seqid=np.zeros(M, np.float32), ry=np.zeros(M, np.float32), slen=np.zeros(M, np.float64), sratio=np.zeros(M, np.float64),
                   max_left=np.zeros(s.n, np.uint32), max_right=np.zeros(s.n, np.uint32), consensus=np.zeros(3 * s.total, np.uint8))
        keys = np.ascontiguousarray(s.keys, np.uint32)
        self.L.orc_extension.restype = None
        self.L.orc_extension(C.c_void_p(s.blob.ctypes.data), C.c_void_p(s.off.ctypes.data), C.c_void_p(s.len.ctypes.data), C.c_void_p(s.ext.ctypes.data),
                             C.c_void_p(keys.ctypes.data), C.c_size_t(s.n), C.c_void_p(aln_off.ctypes.data), C.c_void_p(t_idx.ctypes.data),
                             C.c_void_p(qs.ctypes.data), C.c_void_p(qe.ctypes.data), C.c_void_p(ds.ctypes.data), C.c_void_p(de.ctypes.data),
                             C.c_void_p(seqid.ctypes.data), C.c_float(seqid_thr), C.c_float(ry_thr), C.c_int(unsafe), C.c_uint(min_cov_safe),
                             C.c_double(seq_err), C.c_float(rand), C.c_float(excess),
                             *[C.c_void_p(out[k].ctypes.data) for k in ("flags", "qs", "qe", "ds", "de", "seqid", "ry", "slen", "sratio", "max_left", "max_right",
                                                                        "consensus")])
        return out

    def merge(self, s, aln_off, t_idx, qs, qe, ds, de, unsafe=0, merge_thr=0.99, ry_thr=0.99, min_cov_safe=5):
        """scoring front half of ancient_contig_merge (orc_merge) over all queries; arrays per alignment line"""
        aln_off = np.ascontiguousarray(aln_off, np.uint64); t_idx = np.ascontiguousarray(t_idx, np.uint32)
        qs, qe, ds, de = [np.ascontiguousarray(x, np.int32) for x in (qs, qe, ds, de)]
        M = len(t_idx)
        out = dict(flags=np.zeros(M, np.uint8), qs=np.zeros(M, np.int32), qe=np.zeros(M, np.int32), ds=np.zeros(M, np.int32), de=np.zeros(M, np.int32),
                   seqid0=np.zeros(M, np.float32), ry0=np.zeros(M, np.float32), seqid=np.zeros(M, np.float32), ry=np.zeros(M, np.float32),
                   aln_len_cons=np.zeros(M, np.uint32), deam=np.zeros(M, np.float32), cons_hash=np.zeros(s.n, np.uint64))
        keys = np.ascontiguousarray(s.keys, np.uint32)
        self.L.orc_merge.restype = None
        self.L.orc_merge(C.c_void_p(s.blob.ctypes.data), C.c_void_p(s.off.ctypes.data), C.c_void_p(s.len.ctypes.data), C.c_void_p(keys.ctypes.data), C.c_size_t(s.n),
                         C.c_void_p(aln_off.ctypes.data), C.c_void_p(t_idx.ctypes.data), C.c_void_p(qs.ctypes.data), C.c_void_p(qe.ctypes.data),
                         C.c_void_p(ds.ctypes.data), C.c_void_p(de.ctypes.data), C.c_float(merge_thr), C.c_float(ry_thr), C.c_int(unsafe), C.c_uint(min_cov_safe),
                         *[C.c_void_p(out[k].ctypes.data) for k in ("flags", "qs", "qe", "ds", "de", "seqid0", "ry0", "seqid", "ry", "aln_len_cons", "deam",
                                                                    "cons_hash")])
        return out

    def merge_queue_order(self, aln_len_cons, deam):
        alc = np.ascontiguousarray(aln_len_cons, np.uint32); d = np.ascontiguousarray(deam, np.float32)
        order = np.zeros(len(alc), np.uint32)
        self.L.orc_merge_queue_order.restype = None
        self.L.orc_merge_queue_order(C.c_void_p(alc.ctypes.data), C.c_void_p(d.ctypes.data), C.c_size_t(len(alc)), C.c_void_p(order.ctypes.data))
        return order

    def correct(self, s, aln_off, t_idx, seqid, qs, qe, ds, de, seqid_thr=0.9, ry_thr=0.99, seq_err=0.01, want_lik=False):
        aln_off = np.ascontiguousarray(aln_off, np.uint64); t_idx = np.ascontiguousarray(t_idx, np.uint32)
        seqid = np.ascontiguousarray(seqid, np.float32)
        qs, qe, ds, de = [np.ascontiguousarray(x, np.int32) for x in (qs, qe, ds, de)]
        out = s.blob.copy()
        lik_off = np.zeros(s.n + 1, np.uint64); lik_off[1:] = np.cumsum(s.len.astype(np.uint64))
        lik = np.full((s.total, 4), np.nan) if want_lik else None
        self.L.orc_correct(C.c_void_p(s.blob.ctypes.data), C.c_void_p(s.off.ctypes.data), C.c_void_p(s.len.ctypes.data),
                           C.c_void_p(s.ext.ctypes.data), C.c_size_t(s.n), C.c_void_p(aln_off.ctypes.data),
                           C.c_void_p(t_idx.ctypes.data), C.c_void_p(seqid.ctypes.data), C.c_void_p(qs.ctypes.data),
                           C.c_void_p(qe.ctypes.data), C.c_void_p(ds.ctypes.data), C.c_void_p(de.ctypes.data),
                           C.c_float(seqid_thr), C.c_float(ry_thr), C.c_double(seq_err), C.c_void_p(out.ctypes.data),
                           C.c_void_p(lik_off.ctypes.data), C.c_void_p(lik.ctypes.data) if want_lik else None)
        return out, lik


# ------------------------------------------------------------------------------------------------
# text formats between the stages (QueryMatcher.h:81-126, Matcher.cpp:274-401, Util.cpp:278-307)
def parse_pref(s, payloads):
    """pref DB entries -> flat lines (q_idx, t_idx, prefScore, diag16) in entry order, + per-query offsets."""
    k2i = s.key_to_index()
    q, t, sc, dg, offs = [], [], [], [], [0]
    for qi, p in enumerate(payloads):
        for line in p.split(b"\n"):
            if not line:
                continue
            a, b, c = line.split(b"\t")
            q.append(qi); t.append(k2i[int(a)]); sc.append(int(b)); dg.append(int(c) & 0xFFFF)
        offs.append(len(q))
    return (np.array(q, np.uint32), np.array(t, np.uint32), np.array(sc, np.int32), np.array(dg, np.uint16),
            np.array(offs, np.uint64))


def format_pref(s, hits):
    """Oracle / device hit list (rep index ascending, target ascending) -> {key: payload} like the pref DB."""
    out = {}
    rep, tgt, sc, dg = hits["rep"], hits["tgt"], hits["score"], hits["diag"]
    keys = s.keys
    lines = {}
    for r, t, c, d in zip(rep.tolist(), tgt.tolist(), sc.tolist(), dg.tolist()):
        lines.setdefault(r, []).append(b"%d\t%d\t%d\n" % (int(keys[t]), c, d))
    for i in range(s.n):
        k = int(keys[i])
        out[k] = b"%d\t0\t0\n" % k + b"".join(lines.get(i, []))
    return out


def load_hamming_golden():
    """sequence set, prefilter lines of the first 6000 queries (linclust's kmermatcher) and what the reference's rescorediagonal --rescore-mode 0
    wrote for them without / with --wrapped-scoring (oracle/make_hamming_golden.py) -> (s, (q, t, score, diag16, off), {wrapped: {query key: payload}})"""
    s = SeqSet.from_golden("correction_9")
    pk, pp, _ = load_golden("lin_pref_9")
    pmap = dict(zip(pk, pp))
    lines = parse_pref(s, [pmap.get(int(k), b"") for k in s.keys])
    want = {}
    for w, name in ((0, "lin_rescore_9"), (1, "lin_rescore_wrapped_9")):
        k, p, _ = load_golden(name)
        want[w] = dict(zip(k, p))
    return s, lines, want


def format_hamming(s, q, t, keep, score, diag):
    """kept lines -> {query key: pref-format payload} (QueryMatcher::prefilterHitToBuffer) for the queries that have prefilter lines"""
    out = {}
    for qi, ti, k, sc, dg in zip(q, t, keep, score, diag):
        e = out.setdefault(int(s.keys[qi]), [])
        if k:
            e.append(b"%d\t%d\t%d\n" % (int(s.keys[ti]), int(sc), int(dg)))
    return {k: b"".join(v) for k, v in out.items()}


def seqid_text(x):
    x = np.float32(x)
    if x == np.float32(1.0):
        return b"1.00"
    s = b"0."
    if x < 0.10:
        s += b"0"
    if x < 0.01:
        s += b"0"
    return s + b"%d" % int(np.float32(x * np.float32(1000)))


def format_aln(s, q_off, t_idx, r):
    """rescore outputs (per pref line) -> {key: payload} like the aln DB (kept lines only, pref order)."""
    out = {}
    keys = s.keys
    for qi in range(s.n):
        lines = []
        for h in range(int(q_off[qi]), int(q_off[qi + 1])):
            if not r["keep"][h]:
                continue
            t = int(t_idx[h])
            lines.append(b"%d\t%d\t%s\t%s\t%d\t%d\t%d\t%d\t%d\t%d\n" % (
                int(keys[t]), int(r["bits"][h]), seqid_text(r["seqid"][h]), ("%.3E" % r["evalue"][h]).encode(),
                int(r["qs"][h]), int(r["qe"][h]), int(s.len[qi]), int(r["ds"][h]), int(r["de"][h]), int(s.len[t])))
        out[int(keys[qi])] = b"".join(lines)
    return out


def parse_aln(s, payloads):
    """aln DB entries -> CSR arrays as Matcher::parseAlignmentRecord yields them."""
    k2i = s.key_to_index()
    off, t, sid, qs, qe, ds, de = [0], [], [], [], [], [], []
    for p in payloads:
        for line in p.split(b"\n"):
            if not line:
                continue
            f = line.split(b"\t")
            t.append(k2i[int(f[0])]); sid.append(np.float32(float(f[2])))
            qs.append(int(f[4])); qe.append(int(f[5])); ds.append(int(f[7])); de.append(int(f[8]))
        off.append(len(t))
    return (np.array(off, np.uint64), np.array(t, np.uint32), np.array(sid, np.float32), np.array(qs, np.int32),
            np.array(qe, np.int32), np.array(ds, np.int32), np.array(de, np.int32))


# ------------------------------------------------------------------------------------------------
# synthetic damaged reads (SURVEY §8(d) config 2 shape, any size)
DHIGH5 = (.329405, .221745, .187678, .161196, .144011)
DHIGH3 = (.32891, .223405, .188599, .164419, .146352)   # positions L-1 .. L-5 are rows 4..0 of dhigh3p.prof
_COMP = np.zeros(256, np.uint8)
for _a, _b in zip(b"ACGT", b"TGCA"):
    _COMP[_a] = _b


def synth_reads(n, genome_len, seed, lmin=35, lmax=70, interior=0.01, sub_err=0.001, n_frac=0.0):
    rng = np.random.default_rng(seed)
    genome = rng.integers(0, 4, genome_len, dtype=np.uint8)
    acgt = np.frombuffer(b"ACGT", np.uint8)
    genome = acgt[genome]
    lens = rng.integers(lmin, lmax + 1, n)
    starts = rng.integers(0, genome_len - lmax, n)
    strand = rng.random(n) < 0.5
    seqs = []
    for i in range(n):
        L = int(lens[i])
        r = genome[starts[i]:starts[i] + L].copy()
        if strand[i]:
            r = _COMP[r[::-1]]
        u = rng.random(L)
        p_ct = np.full(L, interior); p_ct[:5] = DHIGH5
        p_ga = np.full(L, interior); p_ga[L - 5:] = DHIGH3[::-1]
        ct = (r == ord("C")) & (u < p_ct)
        ga = (r == ord("G")) & (u < p_ga)
        r[ct] = ord("T"); r[ga] = ord("A")
        e = rng.random(L) < sub_err
        if e.any():
            r[e] = acgt[rng.integers(0, 4, int(e.sum()))]
        if n_frac > 0:
            m = rng.random(L) < n_frac
            r[m] = ord("N")
        seqs.append(r.tobytes())
    return seqs


# ------------------------------------------------------------------------------------------------
# MMseqs2 DB files on disk (for the drop-in stage binary tests)
def write_db(path, keys, payloads, exts, dbtype):
    """payloads without the trailing NUL; writes <path>, <path>.index (4 columns), <path>.dbtype"""
    import struct
    with open(path, "wb") as f, open(path + ".index", "w") as ix:
        off = 0
        for k, p, e in zip(keys, payloads, exts):
            f.write(p + b"\0")
            ix.write("%d\t%d\t%d\t%d\n" % (k, off, len(p) + 1, e))
            off += len(p) + 1
    open(path + ".dbtype", "wb").write(struct.pack("<i", dbtype))


def read_db(path):
    import glob
    files = sorted(glob.glob(path + ".[0-9]*"), key=lambda s: int(s.rsplit(".", 1)[1])) or [path]
    data = b"".join(open(f, "rb").read() for f in files)
    out = {}
    for line in open(path + ".index"):
        f = line.split()
        out[int(f[0])] = (data[int(f[1]):int(f[1]) + int(f[2]) - 1], int(f[3]) if len(f) > 3 else 0)
    return out
