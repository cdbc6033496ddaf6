"""carpedeam_b200 — B200-native (sm_100a) drop-in for CarpeDeam's read-level hot path.

Python is plumbing only: this module loads libcarpedeam_gpu.so (hand-written CUDA behind the C ABI
of include/carpedeam_gpu.h) through ctypes and mirrors the three reference stage modules

    kmermatcher         lib/mmseqs/src/linclust/kmermatcher.cpp
    rescorediagonal     lib/mmseqs/src/alignment/rescorediagonal.cpp
    ancient_correction  src/assembler/correction.cpp

There is no CPU fallback: if the CUDA library or a device is missing, everything here raises.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libcarpedeam_gpu.so")
HOST_LIB_PATH = os.path.join(HERE, "libcarpedeam_host.so")
STAGE_BIN = os.path.join(HERE, "bin", "carpedeam_gpu")
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-shared", "-Xcompiler", "-fPIC"]


def build(force=False, verbose=False):
    """Compile the CUDA library (and the host-side C++ helpers) in-tree for sm_100a."""
    src = os.path.join(HERE, "csrc", "cdg_api.cu")
    deps = [os.path.join(HERE, "csrc", f) for f in os.listdir(os.path.join(HERE, "csrc"))]
    deps.append(os.path.join(os.path.dirname(HERE), "include", "carpedeam_gpu.h"))
    if force or not os.path.exists(LIB_PATH) or any(os.path.getmtime(d) > os.path.getmtime(LIB_PATH) for d in deps):
        cmd = ["nvcc"] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", LIB_PATH, src]
        subprocess.run(cmd, check=True)
    hsrc_dir = os.path.join(HERE, "host")
    if os.path.isdir(hsrc_dir):
        hsrc = sorted(os.path.join(hsrc_dir, f) for f in os.listdir(hsrc_dir) if f.endswith(".cpp") and not f.startswith("main_"))
        hdeps = [os.path.join(hsrc_dir, f) for f in os.listdir(hsrc_dir)]
        if hsrc and (force or not os.path.exists(HOST_LIB_PATH) or any(os.path.getmtime(d) > os.path.getmtime(HOST_LIB_PATH) for d in hdeps)):
            subprocess.run(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-fopenmp", "-ffp-contract=off", "-o", HOST_LIB_PATH] + hsrc + ["-lz"], check=True)
        # the drop-in stage binary (same sub-commands and DB arguments as the reference's modules)
        msrc = os.path.join(hsrc_dir, "main_stages.cpp")
        if os.path.exists(msrc) and (force or not os.path.exists(STAGE_BIN) or any(os.path.getmtime(d) > os.path.getmtime(STAGE_BIN) for d in hdeps + [LIB_PATH])):
            os.makedirs(os.path.dirname(STAGE_BIN), exist_ok=True)
            subprocess.run(["g++", "-O2", "-std=c++17", "-fopenmp", "-ffp-contract=off", "-o", STAGE_BIN, msrc] + hsrc +
                           ["-L" + HERE, "-lcarpedeam_gpu", "-lz", "-Wl,-rpath," + HERE], check=True)
    return LIB_PATH


class CdgError(RuntimeError):
    pass


class KmerParams(C.Structure):
    _fields_ = [("k", C.c_int32), ("kmers_per_seq", C.c_int32), ("kmers_per_seq_scale", C.c_float), ("hash_shift", C.c_uint64),
                ("include_only_extendable", C.c_int32), ("hash_lo", C.c_uint32), ("hash_hi", C.c_uint32), ("split_passes", C.c_int32),
                ("cov_thr", C.c_float), ("cov_mode", C.c_int32)]


class Hits(C.Structure):
    _fields_ = [("n_hits", C.c_uint64), ("rep", C.POINTER(C.c_uint32)), ("target", C.POINTER(C.c_uint32)),
                ("score", C.POINTER(C.c_int32)), ("diag", C.POINTER(C.c_int16)), ("n_records", C.c_uint64), ("n_grouped", C.c_uint64)]


class RescoreParams(C.Structure):
    _fields_ = [("min_seq_id", C.c_float), ("evalue_thr", C.c_double), ("db_residues", C.c_uint64)]


class Alns(C.Structure):
    _fields_ = [("n_lines", C.c_uint64), ("n_queries", C.c_uint64), ("line_off", C.POINTER(C.c_uint64)), ("target", C.POINTER(C.c_uint32)),
                ("keep", C.POINTER(C.c_uint8)), ("score", C.POINTER(C.c_int32)), ("bits", C.POINTER(C.c_int32)),
                ("evalue", C.POINTER(C.c_double)), ("seq_id", C.POINTER(C.c_float)), ("q_start", C.POINTER(C.c_int32)),
                ("q_end", C.POINTER(C.c_int32)), ("db_start", C.POINTER(C.c_int32)), ("db_end", C.POINTER(C.c_int32))]


class Damage(C.Structure):
    _fields_ = [("log_d", C.c_double * 176), ("log_drev", C.c_double * 176), ("log_seq_err", C.c_double * 16)]


class CorrParams(C.Structure):
    _fields_ = [("min_seqid_corr_reads", C.c_float), ("min_ryseq_id_corr_reads", C.c_float)]


class IterParams(C.Structure):
    _fields_ = [("kmer", KmerParams), ("rescore", RescoreParams), ("corr", CorrParams)]


class PairModel(C.Structure):
    _fields_ = [("log_lik", C.c_double * 352), ("log_excess", C.c_float), ("log_rand", C.c_float)]


class Pairs(C.Structure):
    _fields_ = [("n", C.c_uint64), ("query", C.c_void_p), ("target", C.c_void_p), ("rev", C.c_void_p), ("q_start", C.c_void_p),
                ("q_end", C.c_void_p), ("db_start", C.c_void_p), ("db_end", C.c_void_p), ("max_aln", C.c_void_p)]


class AlnsCompact(C.Structure):
    _fields_ = [("n_queries", C.c_uint64), ("n_alns", C.c_uint64), ("first_query", C.c_uint64), ("count", C.POINTER(C.c_uint32)),
                ("target", C.POINTER(C.c_uint32)), ("width", C.c_int32), ("score", C.c_void_p), ("id_cnt", C.c_void_p), ("q_start", C.c_void_p),
                ("db_start", C.c_void_p), ("aln_len_rev", C.c_void_p), ("db_residues", C.c_uint64)]


class Patches(C.Structure):
    _fields_ = [("n", C.c_uint64), ("entry", C.POINTER(C.c_uint64))]


class ExtParams(C.Structure):
    _fields_ = [("min_seq_id", C.c_float), ("min_ryseq_id", C.c_float), ("unsafe", C.c_int32), ("min_cov_safe", C.c_uint32), ("keys", C.c_void_p)]


class ExtScores(C.Structure):
    _fields_ = [("n_lines", C.c_uint64), ("n_queries", C.c_uint64), ("flags", C.POINTER(C.c_uint8)), ("q_start", C.POINTER(C.c_int32)),
                ("q_end", C.POINTER(C.c_int32)), ("db_start", C.POINTER(C.c_int32)), ("db_end", C.POINTER(C.c_int32)), ("seq_id", C.POINTER(C.c_float)),
                ("ry_seq_id", C.POINTER(C.c_float)), ("s_len_norm", C.POINTER(C.c_double)), ("s_ratio", C.POINTER(C.c_double)),
                ("max_aln_left", C.POINTER(C.c_uint32)), ("max_aln_right", C.POINTER(C.c_uint32)), ("consensus", C.POINTER(C.c_uint8)),
                ("consensus_off", C.POINTER(C.c_uint64))]


class MergeModel(C.Structure):
    _fields_ = [("match_lik", C.c_double * 32)]


class MergeParams(C.Structure):
    _fields_ = [("min_merge_seq_id", C.c_float), ("min_ryseq_id", C.c_float), ("unsafe", C.c_int32), ("min_cov_safe", C.c_uint32), ("keys", C.c_void_p)]


class MergeOut(C.Structure):
    _fields_ = [("n_lines", C.c_uint64), ("flags", C.POINTER(C.c_uint8)), ("q_start", C.POINTER(C.c_int32)), ("q_end", C.POINTER(C.c_int32)),
                ("db_start", C.POINTER(C.c_int32)), ("db_end", C.POINTER(C.c_int32)), ("seq_id_query", C.POINTER(C.c_float)),
                ("ry_seq_id_query", C.POINTER(C.c_float)), ("seq_id", C.POINTER(C.c_float)), ("ry_seq_id", C.POINTER(C.c_float)),
                ("deam_match", C.POINTER(C.c_float)), ("aln_len_cons", C.POINTER(C.c_uint32))]


class HammingParams(C.Structure):
    _fields_ = [("cov_thr", C.c_float), ("cov_mode", C.c_int32), ("min_seq_id", C.c_float), ("seq_id_mode", C.c_int32), ("min_aln_len", C.c_int32),
                ("wrapped_scoring", C.c_int32)]


class HammingOut(C.Structure):
    _fields_ = [("n_lines", C.c_uint64), ("n_queries", C.c_uint64), ("n_kept", C.c_uint64), ("line_off", C.POINTER(C.c_uint64)), ("target", C.POINTER(C.c_uint32)),
                ("keep", C.POINTER(C.c_uint8)), ("pref_score", C.POINTER(C.c_int32)), ("diag", C.POINTER(C.c_int32)), ("seq_id", C.POINTER(C.c_float))]


class SeqSetC(C.Structure):
    _fields_ = [("n", C.c_uint64), ("data_bytes", C.c_uint64), ("header_bytes", C.c_uint64), ("total_residues", C.c_uint64), ("data", C.POINTER(C.c_uint8)),
                ("offsets", C.POINTER(C.c_uint64)), ("lengths", C.POINTER(C.c_uint32)), ("headers", C.POINTER(C.c_uint8)), ("header_offsets", C.POINTER(C.c_uint64))]


class CycleParams(C.Structure):
    _fields_ = [("k", C.c_int32), ("max_seq_len", C.c_uint32)]


class CycleOut(C.Structure):
    _fields_ = [("n", C.c_uint64), ("n_cyclic", C.c_uint64), ("n_kmers", C.c_uint64), ("split_diagonal", C.POINTER(C.c_uint32))]


class Stats(C.Structure):
    _fields_ = [("n_seqs", C.c_uint64), ("n_bases", C.c_uint64), ("n_records", C.c_uint64), ("n_grouped", C.c_uint64),
                ("n_hits", C.c_uint64), ("n_alns", C.c_uint64), ("sort1_passes", C.c_uint32), ("sort2_passes", C.c_uint32),
                ("kernel_launches", C.c_uint64), ("ms_extract", C.c_float), ("ms_sort1", C.c_float), ("ms_group", C.c_float),
                ("ms_sort2", C.c_float), ("ms_reduce", C.c_float), ("ms_rescore", C.c_float), ("ms_correct", C.c_float),
                ("ms_sort1_passes", C.c_float), ("n_sort1_pass_launches", C.c_uint32), ("sort1_records", C.c_uint64), ("split_passes", C.c_uint32)]


_lib = None
_hostlib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise CdgError("libcarpedeam_gpu.so is not built (python -c 'import __graft_entry__ as g; g.build()'); no CPU fallback exists")
        L = C.CDLL(LIB_PATH)
        L.cdg_create.argtypes = [C.c_int, C.POINTER(C.c_void_p)]
        L.cdg_destroy.argtypes = [C.c_void_p]
        L.cdg_last_error.argtypes = [C.c_void_p]; L.cdg_last_error.restype = C.c_char_p
        L.cdg_stream.argtypes = [C.c_void_p]; L.cdg_stream.restype = C.c_void_p
        L.cdg_upload_seqs.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64]
        L.cdg_upload_seqs_async.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64]
        L.cdg_kmermatch.argtypes = [C.c_void_p, C.POINTER(KmerParams), C.POINTER(Hits)]
        L.cdg_set_pref_lines.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64]
        L.cdg_rescore.argtypes = [C.c_void_p, C.POINTER(RescoreParams), C.POINTER(Alns)]
        L.cdg_set_damage.argtypes = [C.c_void_p, C.POINTER(Damage)]
        L.cdg_set_alignments.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64] + [C.c_void_p] * 6 + [C.c_uint64]
        L.cdg_correct.argtypes = [C.c_void_p, C.POINTER(CorrParams), C.c_void_p, C.POINTER(C.c_uint64)]
        L.cdg_correct_likelihoods.argtypes = [C.c_void_p, C.POINTER(CorrParams), C.c_void_p]
        L.cdg_iteration.argtypes = [C.c_void_p, C.POINTER(IterParams), C.POINTER(Hits), C.POINTER(Alns), C.c_void_p, C.POINTER(C.c_uint64)]
        L.cdg_set_pair_model.argtypes = [C.c_void_p, C.POINTER(PairModel)]
        L.cdg_pair_loglik.argtypes = [C.c_void_p, C.POINTER(Pairs), C.c_void_p, C.c_void_p]
        L.cdg_pair_loglik_consensus.argtypes = [C.c_void_p, C.POINTER(Pairs), C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p, C.c_void_p]
        L.cdg_upload_packed.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint64]
        L.cdg_iteration_compact.argtypes = [C.c_void_p, C.POINTER(IterParams), C.POINTER(AlnsCompact), C.POINTER(Patches), C.POINTER(C.c_uint64)]
        L.cdg_export_alns_compact.argtypes = [C.c_void_p, C.POINTER(AlnsCompact)]
        L.cdg_export_patches.argtypes = [C.c_void_p, C.POINTER(Patches)]
        L.cdg_extension_scores.argtypes = [C.c_void_p, C.POINTER(ExtParams), C.POINTER(ExtScores)]
        L.cdg_get_stats.argtypes = [C.c_void_p, C.POINTER(Stats)]
        L.cdg_test_radix_sort.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_int, C.c_int]
        L.cdg_test_hash_revcomp.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_int, C.c_uint64, C.c_void_p, C.c_void_p]
        L.cdg_test_set_sort_variant.argtypes = [C.c_int, C.c_int]
        _lib = L
    return _lib


def hostlib():
    global _hostlib
    if _hostlib is None:
        if not os.path.exists(HOST_LIB_PATH):
            raise CdgError("libcarpedeam_host.so is not built")
        H = C.CDLL(HOST_LIB_PATH)
        H.cdh_build_damage.argtypes = [C.c_char_p, C.c_double, C.POINTER(Damage), C.c_char_p, C.c_size_t]
        H.cdh_build_damage.restype = C.c_int
        H.cdh_build_pair_model.argtypes = [C.c_char_p, C.c_double, C.c_float, C.c_float, C.POINTER(PairModel), C.c_char_p, C.c_size_t]
        H.cdh_build_pair_model.restype = C.c_int
        H.cdh_build_merge_model.argtypes = [C.c_char_p, C.POINTER(MergeModel), C.c_char_p, C.c_size_t]
        H.cdh_build_merge_model.restype = C.c_int
        H.cdh_merge_less.argtypes = [C.c_uint32, C.c_float, C.c_uint32, C.c_float]; H.cdh_merge_less.restype = C.c_int
        H.cdh_merge_queue_order.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]; H.cdh_merge_queue_order.restype = None
        H.cdh_createdb.argtypes = [C.POINTER(C.c_char_p), C.c_int, C.c_int, C.c_int, C.c_char_p, C.POINTER(SeqSetC), C.c_char_p, C.c_size_t]; H.cdh_createdb.restype = C.c_int
        H.cdh_free_seqset.argtypes = [C.POINTER(SeqSetC)]; H.cdh_free_seqset.restype = None
        H.cdh_packed_words.argtypes = [C.c_void_p, C.c_uint64]; H.cdh_packed_words.restype = C.c_uint64
        H.cdh_pack_seqs.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p, C.POINTER(C.c_int), C.c_char_p, C.c_size_t]
        H.cdh_pack_seqs.restype = C.c_int
        H.cdh_expanded_lines.argtypes = [C.POINTER(AlnsCompact)]; H.cdh_expanded_lines.restype = C.c_uint64
        H.cdh_expand_alns.argtypes = [C.POINTER(AlnsCompact)] + [C.c_void_p] * 12; H.cdh_expand_alns.restype = C.c_int
        H.cdh_apply_patches.argtypes = [C.c_void_p, C.c_uint64, C.POINTER(Patches)]; H.cdh_apply_patches.restype = C.c_int
        _hostlib = H
    return _hostlib


def pack_seqs(blob, off, lens, with_mask=True):
    """DB bytes -> (words u64, nmask u32 or None, has_n): the 2-bit packed set cdg_upload_packed takes (host/pack.cpp).
    Raises CdgError when a letter other than upper-case A,C,G,T,N occurs."""
    blob = np.ascontiguousarray(blob, np.uint8); off = np.ascontiguousarray(off, np.uint64); lens = np.ascontiguousarray(lens, np.uint32)
    H = hostlib()
    nw = int(H.cdh_packed_words(_ptr(lens), len(lens)))
    words = np.zeros(nw, np.uint64)
    nmask = np.zeros(nw, np.uint32) if with_mask else None
    has_n = C.c_int(0)
    err = C.create_string_buffer(512)
    if H.cdh_pack_seqs(_ptr(blob), _ptr(off), _ptr(lens), len(lens), _ptr(words), _ptr(nmask), C.byref(has_n), err, 512) != 0:
        raise CdgError("pack_seqs: " + err.value.decode())
    return words, nmask, bool(has_n.value)


def expand_alns(compact, lens):
    """cdg_alns_compact -> the dict Context._alns gives for the same alignments restricted to keep == 1 (host/pack.cpp)."""
    lens = np.ascontiguousarray(lens, np.uint32)
    H = hostlib()
    M = int(H.cdh_expanded_lines(C.byref(compact)))
    nq = int(compact.n_queries)
    out = dict(line_off=np.zeros(nq + 1, np.uint64), tgt=np.zeros(M, np.uint32), keep=np.zeros(M, np.uint8), score=np.zeros(M, np.int32),
               bits=np.zeros(M, np.int32), evalue=np.zeros(M, np.float64), seqid=np.zeros(M, np.float32), qs=np.zeros(M, np.int32),
               qe=np.zeros(M, np.int32), ds=np.zeros(M, np.int32), de=np.zeros(M, np.int32))
    rc = H.cdh_expand_alns(C.byref(compact), _ptr(lens), *[_ptr(out[k]) for k in ("line_off", "tgt", "keep", "score", "bits", "evalue", "seqid", "qs",
                                                                                   "qe", "ds", "de")])
    if rc != 0:
        raise CdgError("cdh_expand_alns failed (%d)" % rc)
    return out


def apply_patches(blob, patches):
    """corrected reads = the input bytes with the patches applied (host/pack.cpp)"""
    out = np.ascontiguousarray(blob, np.uint8).copy()
    if hostlib().cdh_apply_patches(_ptr(out), out.nbytes, C.byref(patches)) != 0:
        raise CdgError("cdh_apply_patches: offset outside the data")
    return out


def _ptr(a):
    return None if a is None else C.c_void_p(a.ctypes.data)


def _view(ptr, n, dtype):
    if n == 0:
        return np.zeros(0, dtype)
    return np.ctypeslib.as_array(ptr, shape=(int(n),)).view(dtype).copy()


def kmer_params(k=20, kmers_per_seq=200, scale=0.2, hash_shift=67, include_only_extendable=0, hash_lo=0, hash_hi=65535, split_passes=0, cov_thr=0.0,
                cov_mode=1):
    return KmerParams(k, kmers_per_seq, scale, hash_shift, include_only_extendable, hash_lo, hash_hi, split_passes, cov_thr, cov_mode)


def load_damage(prefix, seq_err=0.01):
    """--ancient-damage PFX -> cdg_damage, built by the C++ host code (carpedeam_b200/host/damage.cpp)."""
    d = Damage()
    err = C.create_string_buffer(512)
    rc = hostlib().cdh_build_damage((prefix or "").encode(), seq_err, C.byref(d), err, 512)
    if rc != 0:
        raise CdgError("damage profile: " + err.value.decode())
    return d


def load_pair_model(prefix, seq_err=0.001, rand_align_penal=0.85, excess_penal=0.0625):
    """--ancient-damage PFX (+ --ext-random-align, --excess-penalty) -> cdg_pair_model (carpedeam_b200/host/damage.cpp)."""
    m = PairModel()
    err = C.create_string_buffer(512)
    rc = hostlib().cdh_build_pair_model((prefix or "").encode(), seq_err, rand_align_penal, excess_penal, C.byref(m), err, 512)
    if rc != 0:
        raise CdgError("damage profile: " + err.value.decode())
    return m


def createdb(files, shuffle=True, out_db=None, write_lookup=True):
    """FASTA / FASTQ [.gz] files -> (data u8, offsets u64[n], lengths u32[n], headers [bytes]) in the reference's createdb key order
    (carpedeam_b200/host/createdb.cpp); with out_db also written as MMseqs2 database files"""
    arr = (C.c_char_p * len(files))(*[f.encode() for f in files])
    s = SeqSetC()
    err = C.create_string_buffer(512)
    if hostlib().cdh_createdb(arr, len(files), 1 if shuffle else 0, 1 if write_lookup else 0, out_db.encode() if out_db else None, C.byref(s), err, 512) != 0:
        raise CdgError("createdb: " + err.value.decode())
    try:
        n = int(s.n)
        data = _view(s.data, s.data_bytes, np.uint8); off = _view(s.offsets, n + 1, np.uint64); lens = _view(s.lengths, n, np.uint32)
        hb = _view(s.headers, s.header_bytes, np.uint8).tobytes(); ho = _view(s.header_offsets, n + 1, np.uint64)
        headers = [hb[int(ho[i]):int(ho[i + 1]) - 2] for i in range(n)]
    finally:
        hostlib().cdh_free_seqset(C.byref(s))
    return data, off[:n].copy(), lens, headers


def load_merge_model(prefix):
    """--ancient-damage PFX -> cdg_merge_model (carpedeam_b200/host/merge.cpp)."""
    m = MergeModel()
    err = C.create_string_buffer(512)
    if hostlib().cdh_build_merge_model((prefix or "").encode(), C.byref(m), err, 512) != 0:
        raise CdgError("damage profile: " + err.value.decode())
    return m


def merge_queue_order(aln_len_cons, deam_match):
    """order in which the reference's priority queue (CompareNuclResultByScoreContigs) pops lines pushed in the given order"""
    alc = np.ascontiguousarray(aln_len_cons, np.uint32); d = np.ascontiguousarray(deam_match, np.float32)
    order = np.zeros(len(alc), np.uint32)
    hostlib().cdh_merge_queue_order(_ptr(alc), _ptr(d), len(alc), _ptr(order))
    return order


class Context:
    """One GPU, one stream.  Mirrors the reference's stage modules on a resident sequence set."""

    def __init__(self, device=0):
        self._h = C.c_void_p()
        rc = lib().cdg_create(device, C.byref(self._h))
        if rc != 0:
            raise CdgError("cdg_create(device=%d) failed (%d): CUDA device required, no CPU fallback" % (device, rc))
        self._keep = []
        self._stream_mapped = None      # multigpu.ShardedIteration: sizes of the peer-mapped receive areas

    def close(self):
        if self._h:
            lib().cdg_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc, what):
        if rc != 0:
            raise CdgError("%s failed (%d): %s" % (what, rc, lib().cdg_last_error(self._h).decode()))

    @property
    def stream(self):
        return lib().cdg_stream(self._h)

    def upload(self, blob, off, lens, ext=None):
        blob = np.ascontiguousarray(blob, np.uint8); off = np.ascontiguousarray(off, np.uint64); lens = np.ascontiguousarray(lens, np.uint32)
        ext = None if ext is None else np.ascontiguousarray(ext, np.uint8)
        self.n = len(lens); self.data_bytes = blob.nbytes
        self._ck(lib().cdg_upload_seqs(self._h, _ptr(blob), blob.nbytes, _ptr(off), _ptr(lens), _ptr(ext), len(lens)), "cdg_upload_seqs")

    def upload_raw(self, blob_ptr, nbytes, off_ptr, len_ptr, ext_ptr, n):
        """pointer-level upload (pinned host buffers owned by the caller, e.g. torch pinned tensors)"""
        self.n = n; self.data_bytes = nbytes
        self._ck(lib().cdg_upload_seqs(self._h, C.c_void_p(blob_ptr), nbytes, C.c_void_p(off_ptr), C.c_void_p(len_ptr),
                                       C.c_void_p(ext_ptr) if ext_ptr else None, n), "cdg_upload_seqs")

    def upload_raw_async(self, blob_ptr, nbytes, off_ptr, len_ptr, ext_ptr, n):
        """cdg_upload_seqs_async: returns once the copies are queued; the (pinned) host buffers must stay valid until the next call
        on this context has returned (kmermatch / iteration consume the pieces as they arrive)"""
        self.n = n; self.data_bytes = nbytes
        self._ck(lib().cdg_upload_seqs_async(self._h, C.c_void_p(blob_ptr), nbytes, C.c_void_p(off_ptr), C.c_void_p(len_ptr),
                                             C.c_void_p(ext_ptr) if ext_ptr else None, n), "cdg_upload_seqs_async")

    def upload_packed(self, words, nmask, lens, ext=None):
        """cdg_upload_packed from numpy arrays (kept alive on the context until the next upload)"""
        words = np.ascontiguousarray(words, np.uint64); lens = np.ascontiguousarray(lens, np.uint32)
        nmask = None if nmask is None else np.ascontiguousarray(nmask, np.uint32)
        ext = None if ext is None else np.ascontiguousarray(ext, np.uint8)
        self._keep = [words, nmask, lens, ext]
        self.n = len(lens); self.data_bytes = int(lens.astype(np.uint64).sum()) + 2 * len(lens)
        self._ck(lib().cdg_upload_packed(self._h, _ptr(words), _ptr(nmask), _ptr(lens), _ptr(ext), len(lens), len(words)), "cdg_upload_packed")

    def upload_packed_raw(self, words_ptr, nmask_ptr, len_ptr, ext_ptr, n, n_words):
        self.n = n
        self._ck(lib().cdg_upload_packed(self._h, C.c_void_p(words_ptr), C.c_void_p(nmask_ptr) if nmask_ptr else None, C.c_void_p(len_ptr),
                                         C.c_void_p(ext_ptr) if ext_ptr else None, n, n_words), "cdg_upload_packed")

    def iteration_compact(self, ip, alns=None, patches=None):
        """cdg_iteration_compact; alns / patches are AlnsCompact / Patches structures filled with views of the library's pinned buffers"""
        ch = C.c_uint64()
        self._ck(lib().cdg_iteration_compact(self._h, C.byref(ip), C.byref(alns) if alns is not None else None,
                                             C.byref(patches) if patches is not None else None, C.byref(ch)), "cdg_iteration_compact")
        return ch.value

    def export_alns_compact(self):
        a = AlnsCompact()
        self._ck(lib().cdg_export_alns_compact(self._h, C.byref(a)), "cdg_export_alns_compact")
        return a

    def export_patches(self):
        p = Patches()
        self._ck(lib().cdg_export_patches(self._h, C.byref(p)), "cdg_export_patches")
        return p

    @staticmethod
    def _hits(h):
        H = h.n_hits
        return dict(rep=_view(h.rep, H, np.uint32), tgt=_view(h.target, H, np.uint32), score=_view(h.score, H, np.int32),
                    diag=_view(h.diag, H, np.int16), R1=h.n_records, R2=h.n_grouped)

    @staticmethod
    def _alns(a):
        M = a.n_lines
        return dict(line_off=_view(a.line_off, a.n_queries + 1, np.uint64), tgt=_view(a.target, M, np.uint32), keep=_view(a.keep, M, np.uint8),
                    score=_view(a.score, M, np.int32), bits=_view(a.bits, M, np.int32), evalue=_view(a.evalue, M, np.float64),
                    seqid=_view(a.seq_id, M, np.float32), qs=_view(a.q_start, M, np.int32), qe=_view(a.q_end, M, np.int32),
                    ds=_view(a.db_start, M, np.int32), de=_view(a.db_end, M, np.int32))

    def kmermatch(self, params=None, fetch=True):
        p = params or kmer_params()
        h = Hits()
        self._ck(lib().cdg_kmermatch(self._h, C.byref(p), C.byref(h) if fetch else None), "cdg_kmermatch")
        return self._hits(h) if fetch else None

    def set_pref_lines(self, line_off, tgt, pref_score, diag16):
        line_off = np.ascontiguousarray(line_off, np.uint64); tgt = np.ascontiguousarray(tgt, np.uint32)
        pref_score = np.ascontiguousarray(pref_score, np.int32); diag16 = np.ascontiguousarray(diag16, np.uint16)
        self._ck(lib().cdg_set_pref_lines(self._h, _ptr(line_off), len(line_off) - 1, _ptr(tgt), _ptr(pref_score), _ptr(diag16), len(tgt)),
                 "cdg_set_pref_lines")

    def rescore(self, min_seq_id=0.9, evalue_thr=0.001, db_residues=0, fetch=True):
        p = RescoreParams(min_seq_id, evalue_thr, db_residues)
        a = Alns()
        self._ck(lib().cdg_rescore(self._h, C.byref(p), C.byref(a) if fetch else None), "cdg_rescore")
        return self._alns(a) if fetch else None

    def set_damage(self, damage):
        self._ck(lib().cdg_set_damage(self._h, C.byref(damage)), "cdg_set_damage")

    def set_alignments(self, aln_off, tgt, seqid, qs, qe, ds, de):
        aln_off = np.ascontiguousarray(aln_off, np.uint64); tgt = np.ascontiguousarray(tgt, np.uint32); seqid = np.ascontiguousarray(seqid, np.float32)
        qs, qe, ds, de = [np.ascontiguousarray(x, np.int32) for x in (qs, qe, ds, de)]
        self._ck(lib().cdg_set_alignments(self._h, _ptr(aln_off), len(aln_off) - 1, _ptr(tgt), _ptr(seqid), _ptr(qs), _ptr(qe), _ptr(ds), _ptr(de),
                                          len(tgt)), "cdg_set_alignments")

    def correct(self, min_seqid=0.9, min_ry=0.99, out=None):
        p = CorrParams(min_seqid, min_ry)
        if out is None:
            out = np.zeros(self.data_bytes, np.uint8)
        ch = C.c_uint64()
        self._ck(lib().cdg_correct(self._h, C.byref(p), _ptr(out), C.byref(ch)), "cdg_correct")
        return out, ch.value

    def correct_likelihoods(self, total_bases, min_seqid=0.9, min_ry=0.99):
        p = CorrParams(min_seqid, min_ry)
        lik = np.zeros((total_bases, 4), np.float64)
        self._ck(lib().cdg_correct_likelihoods(self._h, C.byref(p), _ptr(lik)), "cdg_correct_likelihoods")
        return lik

    def iteration(self, kmer=None, min_seq_id=0.9, evalue_thr=0.001, db_residues=0, min_seqid_corr=0.9, min_ry=0.99,
                  fetch_hits=False, fetch_alns=True, out=None, out_ptr=None):
        ip = IterParams(kmer or kmer_params(), RescoreParams(min_seq_id, evalue_thr, db_residues), CorrParams(min_seqid_corr, min_ry))
        h = Hits(); a = Alns(); ch = C.c_uint64()
        if out is None and out_ptr is None:
            out = np.zeros(self.data_bytes, np.uint8)
        optr = C.c_void_p(out_ptr) if out_ptr else _ptr(out)
        self._ck(lib().cdg_iteration(self._h, C.byref(ip), C.byref(h) if fetch_hits else None, C.byref(a) if fetch_alns else None, optr,
                                     C.byref(ch)), "cdg_iteration")
        return (self._hits(h) if fetch_hits else None, self._alns(a) if fetch_alns else None, out, ch.value)

    def iteration_raw(self, ip, hits, alns, out_ptr):
        """no-copy variant for the benchmark: results stay in the library's pinned buffers"""
        ch = C.c_uint64()
        self._ck(lib().cdg_iteration(self._h, C.byref(ip), C.byref(hits) if hits is not None else None,
                                     C.byref(alns) if alns is not None else None, C.c_void_p(out_ptr) if out_ptr else None, C.byref(ch)),
                 "cdg_iteration")
        return ch.value

    def set_pair_model(self, model):
        self._ck(lib().cdg_set_pair_model(self._h, C.byref(model)), "cdg_set_pair_model")

    def pair_loglik(self, query, target, rev, qs, qe, ds, de, max_aln, consensus=None, cons_off=None):
        """calcLikelihood (consensus=None) / calcLikelihoodConsensus over n pairs of the resident set -> (sLenNorm, sRatio)"""
        query = np.ascontiguousarray(query, np.uint32); target = np.ascontiguousarray(target, np.uint32); rev = np.ascontiguousarray(rev, np.uint8)
        qs, qe, ds, de = [np.ascontiguousarray(x, np.int32) for x in (qs, qe, ds, de)]
        max_aln = np.ascontiguousarray(max_aln, np.uint32)
        n = len(query)
        P = Pairs(n, *[x.ctypes.data for x in (query, target, rev, qs, qe, ds, de, max_aln)])
        a = np.zeros(n, np.float64); b = np.zeros(n, np.float64)
        if consensus is None:
            self._ck(lib().cdg_pair_loglik(self._h, C.byref(P), _ptr(a), _ptr(b)), "cdg_pair_loglik")
        else:
            consensus = np.ascontiguousarray(consensus, np.uint8); cons_off = np.ascontiguousarray(cons_off, np.uint64)
            self._ck(lib().cdg_pair_loglik_consensus(self._h, C.byref(P), _ptr(consensus), consensus.nbytes, _ptr(cons_off), _ptr(a), _ptr(b)),
                     "cdg_pair_loglik_consensus")
        return a, b

    def extension_scores(self, min_seq_id=0.9, min_ryseq_id=0.99, unsafe=0, min_cov_safe=5, keys=None):
        """cdg_extension_scores on the context's alignments -> dict of numpy arrays (copies), same names as Oracle.extension"""
        keys = None if keys is None else np.ascontiguousarray(keys, np.uint32)
        p = ExtParams(min_seq_id, min_ryseq_id, unsafe, min_cov_safe, keys.ctypes.data if keys is not None else None)
        o = ExtScores()
        self._ck(lib().cdg_extension_scores(self._h, C.byref(p), C.byref(o)), "cdg_extension_scores")
        M, n = o.n_lines, o.n_queries
        cons_off = _view(o.consensus_off, n + 1, np.uint64)
        return dict(flags=_view(o.flags, M, np.uint8), qs=_view(o.q_start, M, np.int32), qe=_view(o.q_end, M, np.int32), ds=_view(o.db_start, M, np.int32),
                    de=_view(o.db_end, M, np.int32), seqid=_view(o.seq_id, M, np.float32), ry=_view(o.ry_seq_id, M, np.float32),
                    slen=_view(o.s_len_norm, M, np.float64), sratio=_view(o.s_ratio, M, np.float64), max_left=_view(o.max_aln_left, n, np.uint32),
                    max_right=_view(o.max_aln_right, n, np.uint32), consensus=_view(o.consensus, int(cons_off[-1]), np.uint8), consensus_off=cons_off)

    def rescore_hamming(self, cov_thr=0.8, cov_mode=0, min_seq_id=0.9, seq_id_mode=0, min_aln_len=0, wrapped=0):
        """cdg_rescore_hamming (rescorediagonal --rescore-mode 0) on the context's prefilter lines -> dict of numpy arrays (copies)"""
        p = HammingParams(cov_thr, cov_mode, min_seq_id, seq_id_mode, min_aln_len, wrapped)
        o = HammingOut()
        self._ck(lib().cdg_rescore_hamming(self._h, C.byref(p), C.byref(o)), "cdg_rescore_hamming")
        M, n = o.n_lines, o.n_queries
        return dict(line_off=_view(o.line_off, n + 1, np.uint64), tgt=_view(o.target, M, np.uint32), keep=_view(o.keep, M, np.uint8),
                    score=_view(o.pref_score, M, np.int32), diag=_view(o.diag, M, np.int32), seqid=_view(o.seq_id, M, np.float32), n_kept=int(o.n_kept))

    def cyclecheck(self, k=22, max_seq_len=200000):
        """cdg_cyclecheck on the resident set -> split diagonal per sequence (0 = not circular)"""
        p = CycleParams(k, max_seq_len)
        o = CycleOut()
        self._ck(lib().cdg_cyclecheck(self._h, C.byref(p), C.byref(o)), "cdg_cyclecheck")
        return _view(o.split_diagonal, o.n, np.uint32)

    def set_merge_model(self, model):
        self._ck(lib().cdg_set_merge_model(self._h, C.byref(model)), "cdg_set_merge_model")

    def merge_scores(self, min_merge_seq_id=0.99, min_ryseq_id=0.99, unsafe=0, min_cov_safe=5, keys=None):
        """cdg_merge_scores on the context's alignments -> dict of numpy arrays (copies), same names as Oracle.merge"""
        keys = None if keys is None else np.ascontiguousarray(keys, np.uint32)
        p = MergeParams(min_merge_seq_id, min_ryseq_id, unsafe, min_cov_safe, keys.ctypes.data if keys is not None else None)
        o = MergeOut()
        self._ck(lib().cdg_merge_scores(self._h, C.byref(p), C.byref(o)), "cdg_merge_scores")
        M = o.n_lines
        return dict(flags=_view(o.flags, M, np.uint8), qs=_view(o.q_start, M, np.int32), qe=_view(o.q_end, M, np.int32), ds=_view(o.db_start, M, np.int32),
                    de=_view(o.db_end, M, np.int32), seqid0=_view(o.seq_id_query, M, np.float32), ry0=_view(o.ry_seq_id_query, M, np.float32),
                    seqid=_view(o.seq_id, M, np.float32), ry=_view(o.ry_seq_id, M, np.float32), deam=_view(o.deam_match, M, np.float32),
                    aln_len_cons=_view(o.aln_len_cons, M, np.uint32))

    def stats(self):
        s = Stats()
        self._ck(lib().cdg_get_stats(self._h, C.byref(s)), "cdg_get_stats")
        return {f[0]: getattr(s, f[0]) for f in Stats._fields_}

    def test_radix_sort(self, keys, vals, bit_lo, bit_hi):
        keys = np.ascontiguousarray(keys, np.uint64).copy()
        vals = None if vals is None else np.ascontiguousarray(vals, np.uint64).copy()
        self._ck(lib().cdg_test_radix_sort(self._h, _ptr(keys), _ptr(vals), len(keys), bit_lo, bit_hi), "cdg_test_radix_sort")
        return keys, vals

    def test_hash_revcomp(self, x, k, seed):
        x = np.ascontiguousarray(x, np.uint64)
        h = np.zeros_like(x); r = np.zeros_like(x)
        self._ck(lib().cdg_test_hash_revcomp(self._h, _ptr(x), len(x), k, seed, _ptr(h), _ptr(r)), "cdg_test_hash_revcomp")
        return h, r
