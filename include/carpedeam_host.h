/*
 * carpedeam_host.h — C entry points of libcarpedeam_host.so: the host side of the drop-in boundary (plain C++, no CUDA).
 *
 * What the reference does on the CPU around the hot path and the device library (include/carpedeam_gpu.h) needs as input or hands
 * back as output: damage tables from `--ancient-damage PFX`, the text formats between the stages, the MMseqs2 database files,
 * 2-bit packing, createdb, and the ordering of the contig-merge priority queue.  Every function cites what it replaces.
 * The drop-in stage binary (carpedeam_b200/host/main_stages.cpp) is built from the same sources.
 */
#ifndef CARPEDEAM_HOST_H
#define CARPEDEAM_HOST_H

#include <stddef.h>
#include <stdint.h>

#include "carpedeam_gpu.h"

#ifdef __cplusplus
extern "C" {
#endif

/* ---- damage model: readNucSubstitionRatesFreq (src/assembler/nuclassembleUtil.h:53-102), initDeamProbabilities
 * (nuclassembleUtil.cpp:821-1007), getSeqErrorProf (:49-65).  prefix = the --ancient-damage argument ("" = no damage); errbuf receives
 * the message of a failure (return -1). */
int cdh_build_damage(const char *prefix, double seq_err, cdg_damage *out, char *errbuf, size_t errlen);            /* log tables of mostLikeliBaseRead, correction.cpp:51-107 */
int cdh_build_pair_model(const char *prefix, double seq_err, float rand_align_penal, float excess_penal, cdg_pair_model *out, char *errbuf,
                         size_t errlen);                                                                          /* calcLikelihood[Consensus], nuclassembleUtil.cpp:95-374 */
int cdh_build_merge_model(const char *prefix, cdg_merge_model *out, char *errbuf, size_t errlen);                  /* ancientMatchCount, nuclassembleUtil.cpp:1117,1151 */
int cdh_damage_matrices(const char *prefix, double *D /*[11][4][4]*/, double *Drev /*[11][4][4]*/, char *errbuf, size_t errlen);

/* ---- contig-merge priority queue: CompareNuclResultByScoreContigs (src/assembler/ancientContigsResults.cpp:24-70) over the
 * (aln_len_cons, deam_match) pairs cdg_merge_scores returns; queue_order = the order std::priority_queue pops lines pushed in index order */
int cdh_merge_less(uint32_t aln_len_cons1, float deam_match1, uint32_t aln_len_cons2, float deam_match2);
void cdh_merge_queue_order(const uint32_t *aln_len_cons, const float *deam_match, size_t m, uint32_t *order);

/* ---- text formats between the stages: QueryMatcher::prefilterHitToBuffer / parsePrefilterHits (lib/mmseqs/src/prefiltering/QueryMatcher.h:81-126),
 * Matcher::resultToBuffer / parseAlignmentRecord (lib/mmseqs/src/alignment/Matcher.cpp:274-401) */
size_t cdh_pref_roundtrip(const char *entry, char *out, size_t cap);
size_t cdh_format_aln_line(char *buf, uint32_t target_key, int32_t bits, float seq_id, double evalue, int32_t q_start, int32_t q_end, int32_t q_len,
                           int32_t db_start, int32_t db_end, int32_t db_len);
size_t cdh_parse_aln_entry(const char *entry, size_t index, uint32_t *target_key, float *seq_id, int32_t *coords4);
uint64_t cdh_format_selftest(uint64_t n, uint64_t seed);   /* the hand-written formatters / parsers against snprintf / strtod on n seeded values: mismatches */

/* ---- MMseqs2 database files: DBReader / DBWriter (lib/mmseqs/src/commons/DBReader.cpp:792-817, DBWriter.cpp:360-426); reads a DB
 * (split data files, unordered index, ext column) and writes it back; returns the number of entries */
long cdh_db_roundtrip(const char *in_path, const char *out_path, int *dbtype_out);
long cdh_db_roundtrip_parts(const char *in_path, const char *out_path, int parts);   /* through `parts` per-thread writers closed into one DB */

/* ---- 2-bit packing for cdg_upload_packed, and the host views of the compact results (carpedeam_gpu.h: cdg_alns_compact, cdg_patches) */
uint64_t cdh_packed_words(const uint32_t *lens, uint64_t n);
int cdh_pack_seqs(const uint8_t *data, const uint64_t *off, const uint32_t *lens, uint64_t n, uint64_t *words, uint32_t *nmask, int *has_n, char *errbuf,
                  size_t errlen);
uint64_t cdh_expanded_lines(const cdg_alns_compact *c);
int cdh_expand_alns(const cdg_alns_compact *c, const uint32_t *lens, uint64_t *line_off, uint32_t *target, uint8_t *keep, int32_t *score, int32_t *bits,
                    double *evalue, float *seq_id, int32_t *q_start, int32_t *q_end, int32_t *db_start, int32_t *db_end);
int cdh_apply_patches(uint8_t *data, uint64_t data_bytes, const cdg_patches *p);

/* ---- createdb: FASTA / FASTQ [.gz] -> sequence set in DB layout (lib/mmseqs/src/util/createdb.cpp, hard split mode; kseq reader
 * lib/mmseqs/lib/kseq/kseq.h).  shuffle != 0 reproduces --shuffle 1: key order = (input index % 32, input index / 32).  With out_db
 * the set is also written as <out_db>{,.index,.dbtype}, <out_db>_h{,.index,.dbtype}, <out_db>.source and, with write_lookup (--write-lookup 1, the
 * default), <out_db>.lookup (key, identifier parsed from the header by Util::parseFastaHeader, input file number).  Arrays are malloc'ed: cdh_free_seqset. */
typedef struct {
    uint64_t n, data_bytes, header_bytes, total_residues;
    uint8_t *data;             /* entries `BASES\n\0` in key order (keys 0..n-1) */
    uint64_t *offsets;         /* n + 1 */
    uint32_t *lengths;         /* bases per entry */
    uint8_t *headers;          /* entries `name[ comment]\n\0` in key order */
    uint64_t *header_offsets;  /* n + 1 */
} cdh_seqset;
int cdh_createdb(const char *const *files, int n_files, int shuffle, int write_lookup, const char *out_db, cdh_seqset *out, char *errbuf, size_t errlen);
void cdh_free_seqset(cdh_seqset *s);

#ifdef __cplusplus
}
#endif
#endif
