/*
 * carpedeam_gpu.h — C ABI of libcarpedeam_gpu.so (CUDA, sm_100a only; no CPU fallback).
 *
 * Drop-in boundary for CarpeDeam's read-level hot path.  The reference has no FFI: each stage is
 * a C++ module `int stage(int argc, const char **argv, const Command&)` that reads and writes
 * MMseqs2 databases.  These entry points are what the bodies of those modules call once the
 * DBReader has mapped the databases (binding sketch: INTEGRATION.md):
 *
 *   kmermatcher        lib/mmseqs/src/linclust/kmermatcher.cpp:596 (kmermatcherInner)   -> cdg_kmermatch
 *   rescorediagonal    lib/mmseqs/src/alignment/rescorediagonal.cpp:45 (doRescorediagonal) -> cdg_rescore
 *   ancient_correction src/assembler/correction.cpp:128 (doCorrection)                   -> cdg_correct
 *   one fused reads-only iteration (data/nuclassemble.sh:105-126)                         -> cdg_iteration
 * and, for the modules either side of the loop:
 *   ancient_read_assemble, before its queue   src/assembler/ancientReadsResults.cpp:200-370     -> cdg_extension_scores, cdg_pair_loglik*
 *   ancient_contig_merge, before its queue    src/assembler/ancientContigsResults.cpp:187-262   -> cdg_merge_scores
 *   cyclecheck                                src/assembler/cyclecheck.cpp:72-262               -> cdg_cyclecheck
 *   rescorediagonal --rescore-mode 0          rescorediagonal.cpp:194-334 (HAMMING branch)      -> cdg_rescore_hamming
 * The in-tree form of the bindings (compiled against the reference's own classes): carpedeam_b200/binding/stages_gpu.cpp.
 *
 * Conventions: every function returns 0 on success or a negative cdg_status; cdg_last_error()
 * gives the message.  Input buffers are caller-owned host memory (pinned memory makes the copies
 * asynchronous).  Output views (cdg_hits, cdg_alns) point at library-owned pinned host memory
 * that stays valid until the next call that produces the same kind of output, or cdg_destroy.
 * Sequence ids are INDICES into the uploaded set (0..n-1); the caller uploads entries in ascending
 * dbKey order (the order of an MMseqs2 .index file), so index order == dbKey order, which is what
 * the reference's comparators use (kmermatcher.h:76-114, id = dbKey).
 * One host thread per context.  A context owns one CUDA device and one stream.
 */
#ifndef CARPEDEAM_GPU_H
#define CARPEDEAM_GPU_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct cdg_ctx cdg_ctx;

typedef enum {
    CDG_OK = 0,
    CDG_ERR_CUDA = -1,        /* CUDA runtime error (message has the call and the error string) */
    CDG_ERR_ARG = -2,         /* invalid argument / call order */
    CDG_ERR_UNSUPPORTED = -3, /* input outside what this build handles (message says which limit) */
    CDG_ERR_NOMEM = -4
} cdg_status;

/* ---- context ------------------------------------------------------------------------------- */
int cdg_create(int device, cdg_ctx **out);
void cdg_destroy(cdg_ctx *ctx);
const char *cdg_last_error(const cdg_ctx *ctx);
/* the CUDA stream (cudaStream_t) all work of this context is enqueued on — for event timing */
void *cdg_stream(cdg_ctx *ctx);

/* ---- sequence set (DBReader view: lib/mmseqs/src/commons/DBReader.h:169-237) -----------------
 * data   : the DB data file bytes; entry i is data[offsets[i] .. offsets[i]+lens[i]) (then "\n\0")
 * lens   : sequence lengths without the trailing "\n\0"  (DBReader::getSeqLen)
 * ext    : CarpeDeam's 4th index column, "was extended" (DBReader::getExtData, DBReader.cpp:807-817)
 * The set stays resident on the device until the next upload.  Packs the bases to 2 bit + N mask.
 * The four arrays may live in host memory (pinned memory makes the copy asynchronous) or in device memory of this GPU
 * (multi-GPU: every rank reads 1/world of the DB from the host and the ranks all-gather over NVLink, multigpu.py). */
int cdg_upload_seqs(cdg_ctx *ctx, const uint8_t *data, uint64_t data_bytes, const uint64_t *offsets,
                    const uint32_t *lens, const uint8_t *ext, uint64_t n);
/* Same, but returns as soon as the copies are queued: the DB bytes travel in pieces on a copy stream and the next call on the
 * context consumes them as they arrive (cdg_kmermatch / cdg_iteration pack and extract piece c while piece c+1 is still on the
 * PCIe bus).  The four arrays must be HOST memory (pinned for real overlap) and stay valid until that next call has returned. */
int cdg_upload_seqs_async(cdg_ctx *ctx, const uint8_t *data, uint64_t data_bytes, const uint64_t *offsets,
                          const uint32_t *lens, const uint8_t *ext, uint64_t n);

/* 2-bit packed upload (SURVEY §8(b) proposal `cdg_upload_seqs(packed2bit, nmask_words, ...)`; §8(d): the metric is quoted from
 * PACKED sequences in pinned host memory).  For sets whose letters are upper-case A,C,G,T,N only (what createdb writes for
 * FASTQ reads; carpedeam_b200/host/pack.cpp `cdh_pack_seqs` builds the arrays and refuses anything else):
 *   words : u64, 32 bases per word, sequence i owns ceil(lens[i]/32) consecutive words in index order, first base in the most
 *           significant bits (the order of Indexer::computeKmerIdx, Indexer.h:136-143), codes A0 C1 T2 G3, N and padding 0
 *   nmask : u32 per word, bit 31-j set = base j of the word is N; NULL = the set holds no N
 * The library rebuilds the DB byte layout (`BASES\n\0` per entry, offsets = prefix sum of lens+2) on the device, so every later
 * call behaves exactly as after cdg_upload_seqs of that DB.  Asynchronous like cdg_upload_seqs_async (arrays: host memory that
 * stays valid until the next call on the context returned, or device memory of this GPU). */
int cdg_upload_packed(cdg_ctx *ctx, const uint64_t *words, const uint32_t *nmask, const uint32_t *lens, const uint8_t *ext,
                      uint64_t n, uint64_t n_words);

/* ---- kmermatcher ----------------------------------------------------------------------------- */
typedef struct {
    int32_t k;                       /* -k (20 reads / 22 contigs) */
    int32_t kmers_per_seq;           /* --kmer-per-seq 200 */
    float kmers_per_seq_scale;       /* --kmer-per-seq-scale 0.2 */
    uint64_t hash_shift;             /* --hash-shift 67 (XXH64 seed) */
    int32_t include_only_extendable; /* --include-only-extendable */
    uint32_t hash_lo, hash_hi;       /* 16-bit hash range owned by this call (0..65535 = all; a sub-range yields that range's
                                        partial result like one split of the reference, kmermatcher.cpp:665-675) */
    int32_t split_passes;            /* hash-range passes INSIDE the call with unsplit results (--split-memory-limit): 0 = as many as
                                        the free device memory asks for (kmermatcher.cpp:615-627), 1 = never split, M > 1 = M ranges
                                        cut from the 65536-bin hash histogram (setupKmerSplits, :742-784) */
    float cov_thr;                   /* -c and --cov-mode: with --include-only-extendable 0 a member stays in its group only if */
    int32_t cov_mode;                /* Util::canBeCovered(cov_thr, cov_mode, repLen, memberLen) (kmermatcher.cpp:526-529); the assembler
                                        passes -c 0 (always true), its final linclust -c 0.99 --cov-mode 1, plain linclust -c 0.8 --cov-mode 0 */
} cdg_kmer_params;

/* prefilter hits in pref-DB order: for rep index ascending, target index ascending
 * (writeKmerMatcherResult, kmermatcher.cpp:815-930).  score = +-#shared k-mers (sign = strand),
 * diag = best diagonal truncated to int16 as prefilterHitToBuffer writes it (QueryMatcher.h:114-126).
 * The `key 0 0` self lines are implicit: every sequence has one. */
typedef struct {
    uint64_t n_hits;
    const uint32_t *rep;
    const uint32_t *target;
    const int32_t *score;
    const int16_t *diag;
    /* work counters: emitted k-mer records, records after group assignment */
    uint64_t n_records, n_grouped;
} cdg_hits;

/* Runs extraction -> sort -> group -> sort -> reduce on the resident set.  The hit list stays on the
 * device for cdg_rescore; `out` (may be NULL) receives a host copy. */
int cdg_kmermatch(cdg_ctx *ctx, const cdg_kmer_params *p, cdg_hits *out);

/* ---- rescorediagonal (mode 3, nucleotides, same query/target DB) ------------------------------ */
typedef struct {
    float min_seq_id;      /* --min-seq-id 0.9 */
    double evalue_thr;     /* -e 0.001 */
    uint64_t db_residues;  /* DBReader::getAminoAcidDBSize() of the target DB; 0 = resident set's total */
} cdg_rescore_params;

/* one record per prefilter line (self lines included), in pref order, grouped by query:
 * lines of query q are [line_off[q], line_off[q+1]) */
typedef struct {
    uint64_t n_lines;
    uint64_t n_queries;
    const uint64_t *line_off; /* n_queries+1 */
    const uint32_t *target;
    const uint8_t *keep;      /* 1 = the reference writes this alignment (rescorediagonal.cpp:314) */
    const int32_t *score;     /* raw ungapped score (DistanceCalculator.h:204-220) */
    const int32_t *bits;      /* int(bitScore+0.5) */
    const double *evalue;
    const float *seq_id;      /* idCnt/alnLen (0 when the E-value gate failed, as in the reference) */
    const int32_t *q_start, *q_end, *db_start, *db_end; /* as written: q_start > q_end = reverse strand */
} cdg_alns;

/* Replace the device-resident prefilter lines by caller-provided ones (stand-alone rescorediagonal:
 * lines parsed from a pref DB with QueryMatcher::parsePrefilterHits, QueryMatcher.h:81-104).
 * Lines must be grouped by query in ascending query index; every query has >= 0 lines. */
int cdg_set_pref_lines(cdg_ctx *ctx, const uint64_t *line_off, uint64_t n_queries, const uint32_t *target,
                       const int32_t *pref_score, const uint16_t *diag, uint64_t n_lines);

int cdg_rescore(cdg_ctx *ctx, const cdg_rescore_params *p, cdg_alns *out);

/* ---- ancient_correction ----------------------------------------------------------------------- */
/* Damage model, host-built exactly as the reference builds it (nuclassembleUtil.cpp:821-1007,
 * correction.cpp:51-75,93-107): natural logs, base order A0 C1 G2 T3.
 *   log_d[l][q][t]    = log(max((double)D[l].p[q][t], 1e-3)),  log_drev likewise for Drev
 *   log_seq_err[q][b] = (double)logl(seqErr.p[q][b])
 * carpedeam_b200/host/damage.cpp builds these from the --ancient-damage .prof files. */
typedef struct {
    double log_d[11][4][4];
    double log_drev[11][4][4];
    double log_seq_err[4][4];
} cdg_damage;
int cdg_set_damage(cdg_ctx *ctx, const cdg_damage *d);

typedef struct {
    float min_seqid_corr_reads;   /* par.seqIdThr in correction.cpp:359 (--min-seq-id 0.9) */
    float min_ryseq_id_corr_reads;/* --min-ryseq-id-corr-reads 0.99 */
} cdg_corr_params;

/* Replace the device-resident alignments by caller-provided ones (stand-alone ancient_correction:
 * records parsed from the aln DB with Matcher::parseAlignmentRecord, Matcher.cpp:274-353). */
int cdg_set_alignments(cdg_ctx *ctx, const uint64_t *aln_off, uint64_t n_queries, const uint32_t *target,
                       const float *seq_id, const int32_t *q_start, const int32_t *q_end,
                       const int32_t *db_start, const int32_t *db_end, uint64_t n_alns);

/* Writes the corrected sequences into out_data, same layout as the uploaded data (only base bytes
 * are written; copy the input first or pass a buffer of data_bytes — separators are copied too).
 * n_changed (may be NULL) receives the number of bases that differ from the input. */
int cdg_correct(cdg_ctx *ctx, const cdg_corr_params *p, uint8_t *out_data, uint64_t *n_changed);

/* Per-column log-likelihoods of the last cdg_correct (north_star: within 1e-5 relative of the reference's
 * long double sums): lik[(base_offset(q) + column) * 4 + qBase], NaN where mostLikeliBaseRead was not
 * evaluated; base_offset = prefix sum of lens.  Debug/validation entry point. */
int cdg_correct_likelihoods(cdg_ctx *ctx, const cdg_corr_params *p, double *lik);

/* ---- damage-aware pair likelihood (extension scoring) ---------------------------------
 * calcLikelihood (src/assembler/nuclassembleUtil.cpp:95-200) and calcLikelihoodConsensus (:203-374) as r_s_pair calls
 * them (src/assembler/ancientReadsResults.cpp:48-70).  The per-column term depends only on (strand, position class,
 * query base, target base); the host tabulates it (carpedeam_b200/host/damage.cpp: cdh_build_pair_model):
 *   log_lik[rev][l][q][t] = log( sum_x seqErr.p[x][t] * max(1e-3, (rev ? Drev : D)[l].p[q][x]) ),  A0 C1 G2 T3
 *   log_excess = logf(--excess-penalty 0.0625f), log_rand = logf(--ext-random-align 0.85f)  (float, as the reference) */
typedef struct {
    double log_lik[2][11][4][4];
    float log_excess;
    float log_rand;
} cdg_pair_model;
int cdg_set_pair_model(cdg_ctx *ctx, const cdg_pair_model *m);

/* n pairs over the resident set; coordinates are the reference's working frame after ancientReadsResults.cpp:225-236:
 * q_start <= q_end, and for rev[i] != 0 the target is reverse-complemented (getNuclRevFragment) and db_start/db_end index
 * that oriented target.  max_aln = the maxAln argument (longest left/right overlap of the query). */
typedef struct {
    uint64_t n;
    const uint32_t *query, *target;
    const uint8_t *rev;
    const int32_t *q_start, *q_end, *db_start, *db_end;
    const uint32_t *max_aln;
} cdg_pairs;
/* calcLikelihood: columns q_start..q_end of the query read against db_start.. of the oriented target.
 * Outputs (host, n doubles each): scorePerRes::sLenNorm and ::sRatio. */
int cdg_pair_loglik(cdg_ctx *ctx, const cdg_pairs *pairs, double *s_len_norm, double *s_ratio);
/* calcLikelihoodConsensus: the query side is the caller's consensus string of pair i, consensus + cons_off[i], 3*qLen bytes
 * ('N' = no call; layout of consensusCaller, nuclassembleUtil.cpp:570-702), qLen = length of resident sequence query[i]. */
int cdg_pair_loglik_consensus(cdg_ctx *ctx, const cdg_pairs *pairs, const uint8_t *consensus, uint64_t consensus_bytes,
                              const uint64_t *cons_off, double *s_len_norm, double *s_ratio);

/* ---- fused iteration: kmermatch -> rescore -> correct with no host round trip ------------------ */
typedef struct {
    cdg_kmer_params kmer;
    cdg_rescore_params rescore;
    cdg_corr_params corr;
} cdg_iter_params;
int cdg_iteration(cdg_ctx *ctx, const cdg_iter_params *p, cdg_hits *hits, cdg_alns *alns, uint8_t *out_data,
                  uint64_t *n_changed);

/* ---- compact results (what has to cross PCIe; SURVEY §8(d): "corrected packed sequences + alignment SoA") ----------------
 * Alignments: only what rescorediagonal WRITES (keep == 1), in pref order, grouped by query.  An ungapped alignment is fully
 * described by (target, raw score, identical columns, first query / target column, length, strand):
 *   q_end = reverse ? q_start - (aln_len-1) : q_start + (aln_len-1),  db_end = db_start + aln_len - 1,
 *   seq_id = score ? (float) id_cnt / (float) aln_len : 0,  bits / evalue from (score, query length, db residues)
 * (carpedeam_b200/host/expand.cpp `cdh_expand_alns` rebuilds the cdg_alns arrays / the alignment DB text from these with the
 * reference's own formulas).  The self alignment of query q (always kept, rescorediagonal.cpp:304-314) is IMPLICIT when q is
 * plain upper-case ACGT: target q, score 2*len, id_cnt len, columns 0..len-1; for any other query it is the first exported
 * record of q (target == q).  width = 2: the five per-alignment arrays are int16/uint16 (longest sequence < 32768), else 4. */
typedef struct {
    uint64_t n_queries, n_alns;
    uint64_t first_query;        /* index of the first query of this context's range (multi-GPU), else 0 */
    const uint32_t *count;       /* n_queries: exported alignments of each query */
    const uint32_t *target;      /* n_alns */
    int32_t width;
    const void *score, *id_cnt;  /* uint16 / uint32 */
    const void *q_start, *db_start; /* int16 / int32 (a score-0 self alignment has start = -1, DistanceCalculator.h:42-55) */
    const void *aln_len_rev;     /* uint16 / uint32: aln_len << 1 | reverse strand */
    uint64_t db_residues;        /* the E-value's database size used on the device */
} cdg_alns_compact;
/* Corrected reads as a patch list: entry = byte offset in the DB data layout << 8 | new base letter; unordered, offsets distinct.
 * (correction.cpp:465-470 rewrites the whole sequence; a caller that holds the input applies the patches to its copy.) */
typedef struct { uint64_t n; const uint64_t *entry; } cdg_patches;

/* cdg_iteration with compact outputs (either may be NULL).  Results are identical to cdg_iteration's (tests/test_gpu_parity.py
 * test_compact_outputs_equal_full_outputs). */
int cdg_iteration_compact(cdg_ctx *ctx, const cdg_iter_params *p, cdg_alns_compact *alns, cdg_patches *patches, uint64_t *n_changed);
/* compact views of the last cdg_rescore / cdg_correct of the stand-alone stages */
int cdg_export_alns_compact(cdg_ctx *ctx, cdg_alns_compact *out);
int cdg_export_patches(cdg_ctx *ctx, cdg_patches *out);

/* ---- scoring front half of ancient_read_assemble ---------------------------------------------------------
 * For the alignment lines of the context (cdg_set_alignments with the aln DB of the iteration, or the fused cdg_rescore) and
 * the resident — corrected — sequences: everything doNuclAssembly1 computes per query before the greedy extension
 * (src/assembler/ancientReadsResults.cpp:200-370): orientation of reverse alignments (:200-236), identity / RY identity against
 * the query (:247-293), the extension candidates `notContig` (:294-315), consensusCaller (nuclassembleUtil.cpp:570-702; --unsafe 0
 * copies the query between two N flanks, --unsafe 1 calls the flanks from the candidates with --min-cov-safe),
 * updateSeqIdConsensusReads (:377-470: identities against the consensus, longest left / right overlap per query) and the scores
 * of r_s_pair / calcLikelihoodConsensus (:48-70, :203-374) for the lines that pass the final filter (:351-352).  Needs
 * cdg_set_pair_model.  The host keeps the priority queue: push line l when flags[l] & 4 and s_ratio[l] > --likelihood-ratio-threshold. */
typedef struct {
    float min_seq_id;          /* --min-seq-id (par.seqIdThr, 0.9 in the assembler workflow) */
    float min_ryseq_id;        /* --min-ryseq-id (par.rySeqIdThr, 0.99) */
    int32_t unsafe;            /* --unsafe */
    uint32_t min_cov_safe;     /* --min-cov-safe (5) */
    const uint32_t *keys;      /* DB keys of the sequences (n entries, host) or NULL when key == index: the reference compares a
                                  target INDEX with the query KEY at ancientReadsResults.cpp:263 */
} cdg_ext_params;
typedef struct {               /* pinned host memory owned by the context, valid until the next call */
    uint64_t n_lines, n_queries;
    const uint8_t *flags;      /* bit 0 extension candidate, bit 1 isRevToAlignment, bit 2 scored */
    const int32_t *q_start, *q_end, *db_start, *db_end;   /* oriented frame */
    const float *seq_id, *ry_seq_id;
    const double *s_len_norm, *s_ratio;                   /* 0 unless scored */
    const uint32_t *max_aln_left, *max_aln_right;         /* per query */
    const uint8_t *consensus; const uint64_t *consensus_off;   /* query q: 3*len letters at consensus + consensus_off[q] */
} cdg_ext_scores;
int cdg_extension_scores(cdg_ctx *ctx, const cdg_ext_params *p, cdg_ext_scores *out);

/* ---- scoring front half of ancient_contig_merge -----------------------------------------------------------
 * For the alignment lines of the context (the aln DB of a contig iteration through cdg_set_alignments, or the fused cdg_rescore) and the
 * resident — corrected — sequences: everything doNuclAssembly2 computes per query before the greedy merge
 * (src/assembler/ancientContigsResults.cpp:187-262): orientation of reverse alignments and identity / RY identity against the query for
 * every line (:195-222), the merge candidates `contigs` (:224-229: both identities above their thresholds, not the query itself),
 * consensusCaller (nuclassembleUtil.cpp:570-702, as in cdg_extension_scores), updateSeqIdConsensus (:704-790: identities over the non-N
 * columns of the padded target against the consensus, alnLengthCons), the queue filter (:246-250: overlaps shorter than 500 need
 * min(500, 0.2 * dbLen)) and ancientMatchCount (:1050-1180: expected matches from the consensus identity plus the posterior-weighted
 * C->T / G->A columns of deamMatches :1011-1047, interior damage class, D or Drev by strand).
 * The host keeps the priority queue: push line l when flags[l] & 4, ordered by cdh_merge_less (carpedeam_b200/host/merge.cpp =
 * CompareNuclResultByScoreContigs, ancientContigsResults.cpp:24-70) over (aln_len_cons, deam_match). */
typedef struct {
    double match_lik[2][4][4];   /* [isRevToAlignment][consensus base][target base], A0 C1 G2 T3: D[5] / Drev[5] of initDeamProbabilities as
                                    doubles (carpedeam_b200/host/merge.cpp: cdh_build_merge_model); only C->T and G->A are read */
} cdg_merge_model;
int cdg_set_merge_model(cdg_ctx *ctx, const cdg_merge_model *m);
typedef struct {
    float min_merge_seq_id;    /* --min-merge-seq-id (par.mergeSeqIdThr, 0.99) */
    float min_ryseq_id;        /* --min-ryseq-id (par.rySeqIdThr, 0.99) */
    int32_t unsafe;            /* --unsafe */
    uint32_t min_cov_safe;     /* --min-cov-safe (5) */
    const uint32_t *keys;      /* DB keys of the sequences (n entries, host) or NULL when key == index */
} cdg_merge_params;
typedef struct {               /* pinned host memory owned by the context, valid until the next call */
    uint64_t n_lines;
    const uint8_t *flags;      /* bit 0 merge candidate, bit 1 isRevToAlignment, bit 2 queued */
    const int32_t *q_start, *q_end, *db_start, *db_end;   /* oriented frame */
    const float *seq_id_query, *ry_seq_id_query;          /* identities against the query (:213-222), every line */
    const float *seq_id, *ry_seq_id;                      /* after updateSeqIdConsensus (= the former unless a candidate touches an end) */
    const float *deam_match;                              /* ancientMatchCount, 0 unless queued */
    const uint32_t *aln_len_cons;                         /* alnLengthCons, 0 unless candidate */
} cdg_merge_out;
int cdg_merge_scores(cdg_ctx *ctx, const cdg_merge_params *p, cdg_merge_out *out);

/* ---- cyclecheck (src/assembler/cyclecheck.cpp:31-269, called after every contig iteration: data/nuclassemble.sh:19-60) ----
 * Per resident sequence shorter than max_seq_len: the diagonal at which it overlaps itself (a circular genome read past its start), or 0.
 * k-mers (forward strand, windows with X included, Indexer.h:21-84) of the front / middle / back third are matched (first occurrence in the
 * earlier part against all in the later part), hits counted per diagonal >= L/3, and the first diagonal whose 1 % band has more than 0.24
 * hits per possible k-mer is reported (:212-236).  The host writes the reported entries (whole, or cut at the diagonal with --chop-cycle). */
typedef struct {
    int32_t k;                 /* 22 (setCycleCheckDefaults, cyclecheck.cpp:26-29) */
    uint32_t max_seq_len;      /* --max-seq-len (200000 in the assembler workflow): longer sequences are skipped */
} cdg_cycle_params;
typedef struct {               /* pinned host memory owned by the context, valid until the next call */
    uint64_t n, n_cyclic, n_kmers;
    const uint32_t *split_diagonal;   /* per sequence, 0 = not reported */
} cdg_cycle_out;
int cdg_cyclecheck(cdg_ctx *ctx, const cdg_cycle_params *p, cdg_cycle_out *out);

/* ---- rescorediagonal --rescore-mode 0 (linclust's Hamming pre-clustering) -------------------------------------
 * The HAMMING branch of doRescorediagonal (lib/mmseqs/src/alignment/rescorediagonal.cpp:194-334, :251-254, :323-329) over the prefilter lines
 * of the context (cdg_kmermatch, or cdg_set_pref_lines with a pref DB: lib/mmseqs/data/workflow/linclust.sh:28-32): number of equal
 * bytes on the best real diagonal (DistanceCalculator.h:93-175, :275-295) or, with --wrapped-scoring, on the best rotation of the
 * circular query (:57-91; targets longer than the query are skipped as in the reference), identity by --seq-id-mode, then the
 * coverage / identity / length filters (Util.cpp:533-567).  Output per prefilter line in pref order: what the reference writes for
 * kept lines — `target \t int(100*seqId) (negative on the reverse strand) \t diagonal (as short)` — and keep == 0 for dropped ones. */
typedef struct {
    float cov_thr;             /* -c (linclust: max(0.5, -c) = 0.8) */
    int32_t cov_mode;          /* --cov-mode 0..5 */
    float min_seq_id;          /* --min-seq-id (linclust: max(0.5, --min-seq-id) = 0.9) */
    int32_t seq_id_mode;       /* --seq-id-mode 0 alignment length, 1 shorter, 2 longer sequence */
    int32_t min_aln_len;       /* --min-aln-len */
    int32_t wrapped_scoring;   /* --wrapped-scoring */
} cdg_hamming_params;
typedef struct {               /* pinned host memory owned by the context, valid until the next call */
    uint64_t n_lines, n_queries, n_kept;
    const uint64_t *line_off;  /* n_queries + 1 */
    const uint32_t *target;
    const uint8_t *keep;
    const int32_t *pref_score; /* int(100 * seqId), negated for reverse-strand lines */
    const int32_t *diag;       /* the value the pref line prints */
    const float *seq_id;
} cdg_hamming_out;
int cdg_rescore_hamming(cdg_ctx *ctx, const cdg_hamming_params *p, cdg_hamming_out *out);

/* ---- multi-GPU phases ------------------------------------------------------------
 * One context per rank, the whole sequence set uploaded on every rank.  K-mer records are owned by hash range like
 * kmermatcher's split passes (kmermatcher.cpp:742-784): rank r owns hash16 in [r*65536/world, (r+1)*65536/world).
 * The caller exchanges the device buffers between the phases (carpedeam_b200/multigpu.py, NCCL through torch.distributed):
 *   cdg_mg_extract  -> all-to-all of the 16-byte records (send_counts[o] records for rank o, contiguous, in rank order); the identity
 *                      records of the rank's slice of sequences (one per sequence, kmermatcher.cpp:244-262) travel in the same exchange
 *                      (id_send_counts / id_send_keys / id_send_vals), so no rank hashes sequences outside its slice
 *   cdg_mg_recv_records / cdg_mg_sort -> min key of the rank (the ranks compare 8 bytes each: first-run quirk, :466)
 *   cdg_mg_group    -> all-to-all of the 8-byte grouped records to the rank owning the representative's index range
 *   cdg_mg_recv_grouped / cdg_mg_sort2 / cdg_mg_head_area -> the leading records with one target id (any number) go to rank-1
 *                      (rep-boundary bleed, :886-888)
 *   cdg_mg_finish   -> hits, alignments and corrected bytes of the rank's query range [rank*ceil(n/world), ...) */
/* partition != 0: the records are partitioned by owner into contiguous send buffers for an NCCL all-to-all (baseline exchange).
 * partition == 0: only the per-owner counts are returned; cdg_mg_scatter_records then writes the partition straight into the
 * peers' receive buffers over NVLink (CUDA IPC mappings from cdg_mg_ipc_export / cdg_mg_ipc_open) — partition pass and transfer
 * are one kernel. */
int cdg_mg_extract(cdg_ctx *ctx, const cdg_kmer_params *p, int world, int rank, int partition, uint64_t *send_counts, void **send_keys,
                   void **send_vals, uint64_t *id_send_counts, void **id_send_keys, void **id_send_vals);
/* Streamed record exchange (peer-memory mode): extraction in `chunks` pieces, the routing pass of piece c writes into the peers'
 * receive areas while piece c+1 is extracted.  Every owner's area = world regions of cap_per_source records (one per source, rank
 * order); cdg_mg_plan returns the region size every rank derives from the same resident set.  Order of calls per iteration:
 *   cdg_mg_plan -> cdg_mg_recv_records(world * cap, ...) + handle exchange when the areas had to grow -> cdg_mg_extract_stream
 *   -> the ranks all-gather send_counts / id_send_counts / overflow -> cdg_mg_recv_fill(counts of every source for this rank)
 *   -> cdg_mg_scatter_records(NULL, id offsets) -> barrier -> cdg_mg_sort ...  (overflow on any rank: redo the phase with the
 *   counted exchange, cdg_mg_extract + cdg_mg_scatter_records) */
int cdg_mg_plan(cdg_ctx *ctx, const cdg_kmer_params *p, int world, int rank, uint64_t *cap_per_source);
int cdg_mg_extract_stream(cdg_ctx *ctx, const cdg_kmer_params *p, int chunks, uint64_t cap_per_source, uint64_t *send_counts,
                          uint64_t *id_send_counts, int *overflow);
int cdg_mg_recv_fill(cdg_ctx *ctx, const uint64_t *counts_from_source, uint64_t cap_per_source);
int cdg_mg_recv_records(cdg_ctx *ctx, uint64_t total, uint64_t id_total, void **recv_keys, void **recv_vals, void **id_recv_keys,
                        void **id_recv_vals);
/* receive buffer `which`: 0 k-mer keys, 1 k-mer values, 2 identity keys, 3 identity values, 4 grouped records.  handle = 72 bytes. */
int cdg_mg_ipc_export(cdg_ctx *ctx, int which, void *handle_out);
int cdg_mg_ipc_open(cdg_ctx *ctx, int which, int peer, const void *handle);
/* a receive buffer must not be reallocated while a peer maps it: capacity (bytes) so that the ranks can agree to close first */
int cdg_mg_ipc_capacity(cdg_ctx *ctx, int which, uint64_t *bytes);
int cdg_mg_ipc_close(cdg_ctx *ctx, int which, int peer);
int cdg_mg_scatter_records(cdg_ctx *ctx, const uint64_t *dst_off, const uint64_t *id_dst_off);
int cdg_mg_scatter_grouped(cdg_ctx *ctx, const uint64_t *dst_off);
int cdg_mg_sort(cdg_ctx *ctx, const cdg_kmer_params *p, uint64_t *min_key);
int cdg_mg_group(cdg_ctx *ctx, const cdg_kmer_params *p, int has_global_first_run, int partition, uint64_t *send_counts, void **send_key2);
int cdg_mg_recv_grouped(cdg_ctx *ctx, uint64_t total, void **recv_key2);
int cdg_mg_sort2(cdg_ctx *ctx, void **sorted, uint64_t *head_count, uint64_t *head_capacity);
/* device address the next rank's head run of n_incoming records must be written to (directly behind this rank's sorted records;
 * the array is moved to a larger buffer when the run does not fit the spare area — head runs are never cut) */
int cdg_mg_head_area(cdg_ctx *ctx, uint64_t n_incoming, void **dst);
int cdg_mg_finish(cdg_ctx *ctx, const cdg_rescore_params *rp, const cdg_corr_params *cp, uint64_t n_appended, cdg_hits *hits, cdg_alns *alns,
                  uint8_t *out_range, uint64_t *range_offset, uint64_t *range_bytes, uint64_t *n_changed);

/* cdg_mg_finish with the compact outputs of cdg_iteration_compact for the rank's query range */
int cdg_mg_finish_compact(cdg_ctx *ctx, const cdg_rescore_params *rp, const cdg_corr_params *cp, uint64_t n_appended, cdg_alns_compact *alns,
                          cdg_patches *patches, uint64_t *range_offset, uint64_t *range_bytes, uint64_t *n_changed);

/* ---- measurement ------------------------------------------------------------------------------ */
typedef struct {
    uint64_t n_seqs, n_bases;           /* N, S */
    uint64_t n_records, n_grouped;      /* R1, R2 */
    uint64_t n_hits, n_alns;            /* H, A (kept alignments incl. self) */
    uint32_t sort1_passes, sort2_passes;/* P1, P2 */
    uint64_t kernel_launches;           /* kernels launched by this context since creation */
    float ms_extract, ms_sort1, ms_group, ms_sort2, ms_reduce, ms_rescore, ms_correct; /* last run, CUDA events */
    /* dominant kernel: the LSD passes of sort 1 alone (events around every onesweep launch), records per launch */
    float ms_sort1_passes; uint32_t n_sort1_pass_launches; uint64_t sort1_records;
    uint32_t split_passes;              /* hash-range passes of the last cdg_kmermatch (1 = unsplit) */
} cdg_stats;
int cdg_get_stats(cdg_ctx *ctx, cdg_stats *out);

/* ---- unit-test entry points (exercise single kernels through the ABI) -------------------------- */
/* in-house LSD radix sort of (key,value) pairs on bits [bit_lo, bit_hi) of the key; values may be NULL */
int cdg_test_radix_sort(cdg_ctx *ctx, uint64_t *keys, uint64_t *vals, uint64_t n, int bit_lo, int bit_hi);
/* XXH64 of one u64 (kmermatcher.cpp:33-38) and 2-bit reverse complement (Util.cpp:601-638) on the device */
int cdg_test_hash_revcomp(cdg_ctx *ctx, const uint64_t *in, uint64_t n, int k, uint64_t seed, uint64_t *hash_out,
                          uint64_t *rc_out);

/* sort micro-benchmark on device-generated random keys (ms per full sort incl. histogram); tuning knob for the tile shape */
int cdg_test_sort_bench(cdg_ctx *ctx, uint64_t n, int bit_lo, int bit_hi, int has_val, int variant, int reps, float *ms_per_sort, int *passes);
int cdg_test_set_sort_variant(int keys_only_variant, int pairs_variant);

#ifdef __cplusplus
}
#endif
#endif
