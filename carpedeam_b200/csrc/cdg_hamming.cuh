// rescorediagonal --rescore-mode 0 (Hamming) on the device, with and without --wrapped-scoring (linclust's
// pre-clustering step, lib/mmseqs/data/workflow/linclust.sh:28-32).
//   per-line body                     lib/mmseqs/src/alignment/rescorediagonal.cpp:194-334 (HAMMING branch :251-254, output :323-329)
//   computeUngappedAlignment          lib/mmseqs/src/alignment/DistanceCalculator.h:93-113, ungappedAlignmentByDiagonal :115-175
//   computeUngappedWrappedAlignment   :57-91 — the query is doubled in the reference; here the circular read wraps at most once
//                                     because wrapped scoring skips targets longer than the query (rescorediagonal.cpp:217-222)
//   computeInverseHammingDistance     :275-295 (number of equal BYTES: on 2-bit packed words for plain ACGT sequences, bytes otherwise)
//   Util::canBeCovered / hasCoverage / computeSeqId   lib/mmseqs/src/commons/Util.cpp:533-567, :588-598
// One thread per prefilter line, like rescore_kernel.
#pragma once
#include "cdg_align.cuh"

namespace cdgham {
using namespace cdgalign;

struct HammingParams { float covThr, seqIdThr; int covMode, seqIdMode, alnLenThr, wrapped; };

CDG_D bool can_be_covered(float covThr, int covMode, float q, float t) {
    switch (covMode) {
        case 0: return (q / t >= covThr) && (t / q >= covThr);
        case 1: return (q / t) >= covThr;                          // COV_MODE_TARGET
        case 2: return (t / q) >= covThr;                          // COV_MODE_QUERY
        case 3: return ((t / q) >= covThr) && (t / q) <= 1.0f;
        case 4: return ((q / t) >= covThr) && (q / t) <= 1.0f;
        case 5: return (fminf(t, q) / fmaxf(t, q)) >= covThr;
        default: return true;
    }
}
CDG_D bool has_coverage(float covThr, int covMode, float qCov, float tCov) {
    switch (covMode) {
        case 0: return qCov >= covThr && tCov >= covThr;
        case 2: return qCov >= covThr;
        case 1: return tCov >= covThr;
        default: return true;
    }
}

// equal bytes of q[qOff .. qOff+n) (forward or reverse-complement view of the query) and t[tOff .. tOff+n)
CDG_D u32 equal_count(bool fast, const u64 *qw, const u8 *qa, u32 qLen, bool qRev, u32 qOff, const u64 *tw, const u8 *ta, u32 tOff, u32 n) {
    if (n == 0) return 0;
    if (fast) { i64 s; u32 m; score_packed(qw, qLen, qRev, qOff, tw, tOff, n, s, m); return m; }
    u32 m = 0;
    for (u32 p = 0; p < n; p++) m += q_char(qa, qLen, qRev, qOff + p) == ta[tOff + p] ? 1u : 0u;
    return m;
}

__global__ void __launch_bounds__(256) hamming_kernel(
    const u8 *__restrict__ data, const SeqMeta *__restrict__ meta, const u64 *__restrict__ words, const u32 *__restrict__ lineQ,
    const u32 *__restrict__ lineT, const i32 *__restrict__ linePref, const u16 *__restrict__ lineDiag, u64 M, HammingParams P,
    u8 *__restrict__ keep, i32 *__restrict__ scoreOut, i32 *__restrict__ diagOut, float *__restrict__ seqIdOut, unsigned long long *__restrict__ keptCount) {
    const u64 l = (u64) blockIdx.x * blockDim.x + threadIdx.x;
    u32 kept = 0;
    if (l < M) {
        const u32 q = lineQ[l], t = lineT[l];
        const SeqMeta qm = meta[q], tm = meta[t];
        const u32 qLen = qm.lenFlags & META_LEN, tLen = tm.lenFlags & META_LEN;
        const bool isRev = linePref[l] < 0, isIdentity = q == t;
        const bool fast = !((qm.lenFlags | tm.lenFlags) & META_EXC);
        const u64 *qw = words + qm.wordOff, *tw = words + tm.wordOff;
        const u8 *qa = data + qm.off, *ta = data + tm.off;
        const u32 diag16 = lineDiag[l];
        u8 k = 0; i32 ps = 0, dg = 0; float sidf = 0.f;
        bool skip = !can_be_covered(P.covThr, P.covMode, (float) qLen, (float) tLen) || qLen == 0 || tLen == 0;
        if (P.wrapped && tLen > qLen) skip = true;   // "no valid wrapped scoring possible"
        if (!skip) {
            u32 best = 0, bestLen = 0; i32 bestDiag = 0;
            if (P.wrapped) {
                // starts realDiagonal = qLen - x for x = d*65536 - diag16 in [1, tLen), then diag16 + d*65536 below qLen: strict > keeps the first
                for (u32 d = 1; d * 65536u - diag16 < tLen; d++) {
                    const u32 start = qLen - (d * 65536u - diag16);
                    const u32 n1 = min(tLen, qLen - start);
                    const u32 m = equal_count(fast, qw, qa, qLen, isRev, start, tw, ta, 0, n1) + equal_count(fast, qw, qa, qLen, isRev, 0, tw, ta, n1, tLen - n1);
                    if (m > best) { best = m; bestDiag = (i32) start; }
                }
                for (u32 d = 0; d * 65536u + diag16 < qLen; d++) {
                    const u32 start = d * 65536u + diag16;
                    const u32 n1 = min(tLen, qLen - start);
                    const u32 m = equal_count(fast, qw, qa, qLen, isRev, start, tw, ta, 0, n1) + equal_count(fast, qw, qa, qLen, isRev, 0, tw, ta, n1, tLen - n1);
                    if (m > best) { best = m; bestDiag = (i32) start; }
                }
                bestLen = min(tLen, qLen);
            } else {
                for (u32 d = 1; d <= 1 + tLen / 32768; d++) {
                    const i32 diagonal = (i32) diag16 - (i32) (d * 65536);
                    const u32 ad = (u32) (-diagonal);
                    if (ad < tLen) {
                        const u32 n = min(tLen - ad, qLen);
                        const u32 m = equal_count(fast, qw, qa, qLen, isRev, 0, tw, ta, ad, n);
                        if (m > best) { best = m; bestDiag = diagonal; bestLen = n; }
                    }
                }
                for (u32 d = 0; d <= qLen / 65536; d++) {
                    const u32 ad = d * 65536u + diag16;
                    if (ad < qLen) {
                        const u32 n = min(tLen, qLen - ad);
                        const u32 m = equal_count(fast, qw, qa, qLen, isRev, ad, tw, ta, 0, n);
                        if (m > best) { best = m; bestDiag = (i32) ad; bestLen = n; }
                    }
                }
            }
            // a line without a single equal byte keeps the default LocalAlignment (diagonalLen 0, wrapped: min length): 0/0 identity, never kept
            const float targetCov = (float) bestLen / (float) tLen, queryCov = (float) bestLen / (float) qLen;
            if (P.seqIdMode == 1) sidf = (float) best / (float) min(qLen, tLen);
            else if (P.seqIdMode == 2) sidf = (float) best / (float) max(qLen, tLen);
            else sidf = (float) best / (float) bestLen;
            const bool hasCov = has_coverage(P.covThr, P.covMode, queryCov, targetCov);
            const bool hasSeqId = (double) sidf >= (double) (P.seqIdThr - FLT_EPSILON);
            const bool hasAlnLen = (i32) bestLen >= P.alnLenThr;
            if (isIdentity || (hasAlnLen && hasCov && hasSeqId)) {
                k = 1;
                const i32 s100 = sidf == sidf ? (i32) (100.0 * (double) sidf) : 0;   // `hit.prefScore = 100 * seqId` (double -> int)
                ps = isRev ? -s100 : s100;
                dg = (i32) (short) (u16) bestDiag;   // hit_t::diagonal is an unsigned short, written as short (QueryMatcher.h:114-126)
            }
        }
        keep[l] = k; scoreOut[l] = ps; diagOut[l] = dg; seqIdOut[l] = sidf;
        kept = k;
    }
    const u32 b = __ballot_sync(0xffffffffu, kept);
    if ((threadIdx.x & 31) == 0 && b) atomicAdd(keptCount, (unsigned long long) __popc(b));
}

}   // namespace cdgham
