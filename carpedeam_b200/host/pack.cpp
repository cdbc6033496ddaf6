// Host side of the packed hand-off (include/carpedeam_gpu.h: cdg_upload_packed, cdg_alns_compact, cdg_patches):
//   cdh_pack_seqs       DB bytes -> 2-bit words + N mask  (what a createdb that feeds the GPU stages keeps resident;
//                       letter codes of Sequence::mapSequence, Sequence.cpp:476-489 / nucleotide.out:4: A0 C1 T2 G3)
//   cdh_expand_alns     compact alignment records -> the full cdg_alns arrays (Matcher::result_t fields, Matcher.h:33-56)
//   cdh_apply_patches   corrected-base patches -> corrected DB bytes (correction.cpp:465-470)
// OpenMP over sequences / queries like the reference's per-entry loops (rescorediagonal.cpp:125-146).
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <vector>

#include "../../include/carpedeam_gpu.h"
#include "../csrc/cdg_alp.h"

extern "C" uint64_t cdh_packed_words(const uint32_t *lens, uint64_t n) {
    uint64_t w = 0;
#pragma omp parallel for reduction(+ : w) schedule(static)
    for (int64_t i = 0; i < (int64_t) n; i++) w += (lens[i] + 31u) >> 5;
    return w;
}

// returns 0, or -1 with a message when a letter other than upper-case A,C,G,T,N occurs (such sets take cdg_upload_seqs).
// nmask may be NULL when the caller knows there is no N: then an N is an error too.  *has_n (may be NULL) = any N seen.
extern "C" int cdh_pack_seqs(const uint8_t *data, const uint64_t *off, const uint32_t *lens, uint64_t n, uint64_t *words, uint32_t *nmask,
                             int *has_n, char *errbuf, size_t errlen) {
    std::vector<uint64_t> wordOff(n + 1, 0);
    for (uint64_t i = 0; i < n; i++) wordOff[i + 1] = wordOff[i] + ((lens[i] + 31u) >> 5);
    int bad = 0, anyN = 0;
    int64_t badSeq = -1;
#pragma omp parallel for schedule(static) reduction(| : bad, anyN)
    for (int64_t i = 0; i < (int64_t) n; i++) {
        const uint8_t *s = data + off[i];
        const uint32_t L = lens[i];
        uint64_t *w = words + wordOff[i];
        uint32_t *m = nmask ? nmask + wordOff[i] : nullptr;
        for (uint32_t b = 0; b < L; b += 32) {
            const uint32_t c = L - b < 32 ? L - b : 32;
            uint64_t word = 0; uint32_t xm = 0;
            for (uint32_t j = 0; j < c; j++) {
                uint64_t code;
                switch (s[b + j]) {
                    case 'A': code = 0; break;
                    case 'C': code = 1; break;
                    case 'T': code = 2; break;
                    case 'G': code = 3; break;
                    case 'N': code = 0; xm |= 1u << (31 - j); break;
                    default: code = 0; bad = 1;
#pragma omp critical
                        if (badSeq < 0) badSeq = i;
                }
                word |= code << (62 - 2 * j);
            }
            w[b >> 5] = word;
            if (m) m[b >> 5] = xm;
            if (xm) anyN = 1;
        }
    }
    if (has_n) *has_n = anyN;
    if (bad || (anyN && !nmask)) {
        if (errbuf && errlen) snprintf(errbuf, errlen, bad ? "sequence %lld holds a letter other than upper-case A,C,G,T,N: use cdg_upload_seqs for this set"
                                                           : "the set holds N but no mask array was given%.0lld", (long long) badSeq);
        return -1;
    }
    return 0;
}

template <class T, class TS>
static void expand_range(const cdg_alns_compact *c, const uint32_t *lens, const uint64_t *lineOff, const uint64_t *cmpOff, uint32_t *target,
                         uint8_t *keep, int32_t *score, int32_t *bits, double *evalue, float *seqId, int32_t *qS, int32_t *qE, int32_t *dS, int32_t *dE) {
    const T *sc = (const T *) c->score, *idc = (const T *) c->id_cnt, *alr = (const T *) c->aln_len_rev;
    const TS *qs = (const TS *) c->q_start, *ds = (const TS *) c->db_start;
    const double dbRes = (double) c->db_residues;
#pragma omp parallel for schedule(dynamic, 4096)
    for (int64_t ql = 0; ql < (int64_t) c->n_queries; ql++) {
        const uint32_t q = (uint32_t) (c->first_query + (uint64_t) ql);
        const double qLen = (double) lens[q];
        uint64_t l = lineOff[ql];
        auto put = [&](uint32_t t, int32_t s, uint32_t id, int32_t a, int32_t b, uint32_t alnLen, bool rev) {
            target[l] = t; keep[l] = 1; score[l] = s;
            evalue[l] = cdg_evalue((double) s, qLen, dbRes);
            bits[l] = (int32_t) (cdg_bitscore((double) s) + 0.5);
            seqId[l] = s ? (float) id / (float) alnLen : 0.f;          // Util::computeSeqId; score 0 = the default LocalAlignment (start = end = -1)
            qS[l] = a; qE[l] = rev ? a - (int32_t) (alnLen - 1) : a + (int32_t) (alnLen - 1);
            dS[l] = b; dE[l] = b + (int32_t) (alnLen - 1);
            l++;
        };
        if (c->count[ql] & 0x80000000u) put(q, 2 * (int32_t) lens[q], lens[q], 0, 0, lens[q], false);   // implicit self alignment
        const uint32_t cnt = c->count[ql] & 0x7FFFFFFFu;
        for (uint64_t i = cmpOff[ql]; i < cmpOff[ql] + cnt; i++)
            put(c->target[i], (int32_t) sc[i], (uint32_t) idc[i], (int32_t) qs[i], (int32_t) ds[i], (uint32_t) alr[i] >> 1, ((uint32_t) alr[i] & 1u) != 0);
    }
}

// number of alignment lines cdh_expand_alns produces (exported + implicit self alignments)
extern "C" uint64_t cdh_expanded_lines(const cdg_alns_compact *c) {
    uint64_t m = c->n_alns;
#pragma omp parallel for reduction(+ : m) schedule(static)
    for (int64_t q = 0; q < (int64_t) c->n_queries; q++) m += c->count[q] >> 31;
    return m;
}

// lens: lengths of ALL sequences of the set (indexed by sequence index).  line_off: n_queries + 1; the other arrays hold
// cdh_expanded_lines(c) elements.  Output = the kept lines of cdg_alns (same values as cdg_rescore's export with keep == 1).
extern "C" int cdh_expand_alns(const cdg_alns_compact *c, const uint32_t *lens, uint64_t *line_off, uint32_t *target, uint8_t *keep, int32_t *score,
                               int32_t *bits, double *evalue, float *seq_id, int32_t *q_start, int32_t *q_end, int32_t *db_start, int32_t *db_end) {
    if (!c || !lens || !line_off) return -1;
    std::vector<uint64_t> cmpOff(c->n_queries + 1, 0);
    line_off[0] = 0;
    for (uint64_t q = 0; q < c->n_queries; q++) {
        const uint32_t cnt = c->count[q] & 0x7FFFFFFFu;
        cmpOff[q + 1] = cmpOff[q] + cnt;
        line_off[q + 1] = line_off[q] + cnt + (c->count[q] >> 31);
    }
    if (cmpOff[c->n_queries] != c->n_alns) return -2;
    if (c->width == 2) expand_range<uint16_t, int16_t>(c, lens, line_off, cmpOff.data(), target, keep, score, bits, evalue, seq_id, q_start, q_end, db_start, db_end);
    else expand_range<uint32_t, int32_t>(c, lens, line_off, cmpOff.data(), target, keep, score, bits, evalue, seq_id, q_start, q_end, db_start, db_end);
    return 0;
}

// data: the caller's copy of the DB bytes the set was uploaded from (or the canonical layout `BASES\n\0`, offsets = prefix sum
// of lens + 2, after cdg_upload_packed)
extern "C" int cdh_apply_patches(uint8_t *data, uint64_t data_bytes, const cdg_patches *p) {
    int bad = 0;
#pragma omp parallel for schedule(static) reduction(| : bad)
    for (int64_t i = 0; i < (int64_t) p->n; i++) {
        const uint64_t e = p->entry[i], pos = e >> 8;
        if (pos >= data_bytes) { bad = 1; continue; }
        data[pos] = (uint8_t) (e & 0xFF);
    }
    return bad ? -1 : 0;
}
