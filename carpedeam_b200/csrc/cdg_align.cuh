// rescorediagonal (mode 3, nucleotides) and ancient_correction on the device.
//   lines_kernel     prefilter hits -> per-query line list (self line first, hits in target order)
//   rescore_kernel   one thread per line    (rescorediagonal.cpp:194-334, DistanceCalculator.h:93-220)
//   correct_kernel   one warp per query     (correction.cpp:7-123,128-490)
#pragma once
#include <float.h>
#include <math.h>
#include "cdg_common.cuh"

namespace cdgalign {

// ---- line list -----------------------------------------------------------------------------------
// pref entry of query q = `q 0 0` followed by its hits (kmermatcher.cpp:836,:710-730).  With hits sorted by
// (rep,target): self line of q at q + lower_bound(hitRep, q); hit h of rep r at h + r + 1.
// Queries [qLo, qLo+n) (multi-GPU: the rank's query range; single GPU: all); lineOff is indexed by q - qLo.
__global__ void __launch_bounds__(256) lines_self_kernel(const u32 *__restrict__ hitRep, u64 H, u64 qLo, u64 n, u64 *__restrict__ lineOff,
                                                         u32 *__restrict__ lineQ, u32 *__restrict__ lineT, i32 *__restrict__ linePref,
                                                         u16 *__restrict__ lineDiag, u32 *__restrict__ multiList = nullptr, u32 *__restrict__ multiCount = nullptr) {
    const u64 ql = (u64) blockIdx.x * blockDim.x + threadIdx.x;
    bool multi = false;
    if (ql <= n) {
        const u64 q = qLo + ql;
        u64 lo = 0, hi = H;
        while (lo < hi) { u64 mid = (lo + hi) >> 1; if (hitRep[mid] < q) lo = mid + 1; else hi = mid; }
        u64 l = ql + lo;
        lineOff[ql] = l;
        if (ql < n) { lineQ[l] = (u32) q; lineT[l] = (u32) q; linePref[l] = 0; lineDiag[l] = 0; multi = lo < H && hitRep[lo] == q; }
    }
    // the queries with more than one line (= at least one hit), for the correction: what multi_line_queries_kernel would list
    if (multiList) {
        const u32 b = __ballot_sync(0xffffffffu, multi);
        if (b) {
            const int lane = threadIdx.x & 31;
            u32 base = 0;
            if (lane == 0) base = atomicAdd(multiCount, (u32) __popc(b));
            base = __shfl_sync(0xffffffffu, base, 0);
            if (multi) multiList[base + __popc(b & ((1u << lane) - 1))] = (u32) ql;
        }
    }
}
__global__ void __launch_bounds__(256) lines_hits_kernel(const u32 *__restrict__ hitRep, const u32 *__restrict__ hitTgt,
                                                         const i32 *__restrict__ hitScore, const i32 *__restrict__ hitDiag, u64 H, u64 qLo,
                                                         u32 *__restrict__ lineQ, u32 *__restrict__ lineT, i32 *__restrict__ linePref,
                                                         u16 *__restrict__ lineDiag) {
    u64 h = (u64) blockIdx.x * blockDim.x + threadIdx.x;
    if (h >= H) return;
    u32 r = hitRep[h];
    u64 l = h + (r - qLo) + 1;
    lineQ[l] = r; lineT[l] = hitTgt[h]; linePref[l] = hitScore[h];
    lineDiag[l] = (u16) (short) hitDiag[h];   // (short) in writeKmerMatcherResult, parsed back as unsigned short
}
// for caller-provided lines: expand line_off into lineQ
__global__ void __launch_bounds__(256) lines_fillq_kernel(const u64 *__restrict__ lineOff, u64 n, u32 *__restrict__ lineQ) {
    u64 q = (u64) blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= n) return;
    for (u64 l = lineOff[q]; l < lineOff[q + 1]; l++) lineQ[l] = (u32) q;
}

// ---- packed helpers ----------------------------------------------------------------------------------
// c (1..32) bases starting at base p, left-aligned, rest zero.  Caller guarantees p + c <= L.
CDG_D u64 get_bases(const u64 *__restrict__ w, u32 p, u32 c) {
    u32 wi = p >> 5, sh = (p & 31) * 2;
    u64 v = w[wi] << sh;
    if (sh && (p & 31) + c > 32) v |= w[wi + 1] >> (64 - sh);
    return c == 32 ? v : (v & ~(~0ull >> (2 * c)));
}
// c bases of revcomp(seq) starting at revcomp position a
CDG_D u64 rev2bit64_dev(u64 x) {   // reverse the 32 2-bit groups: two BREV + one swap inside every group (cdg_rev2bit64 is the portable form)
    u32 lo = __brev((u32) (x >> 32)), hi = __brev((u32) x);
    lo = ((lo >> 1) & 0x55555555u) | ((lo & 0x55555555u) << 1);
    hi = ((hi >> 1) & 0x55555555u) | ((hi & 0x55555555u) << 1);
    return ((u64) hi << 32) | lo;
}
CDG_D u64 get_bases_rc(const u64 *__restrict__ w, u32 L, u32 a, u32 c) {
    u64 x = get_bases(w, L - a - c, c);            // forward bases, left aligned
    u64 r = rev2bit64_dev(x) << (2 * (32 - c));    // reversed, left aligned again
    r ^= 0xAAAAAAAAAAAAAAAAull;
    return c == 32 ? r : (r & ~(~0ull >> (2 * c)));
}

struct RescoreParams { float seqIdThr; double evalThr; double dbRes; };
struct LocalAln { i32 startPos, endPos; u32 score; u32 diagonalLen; u32 dist; i32 diagonal; u32 idCnt; };

// fast path: both sequences are plain upper-case ACGT. q is read forward or as its reverse complement.
CDG_D void score_packed(const u64 *__restrict__ qw, u32 qLen, bool qRev, u32 qOff, const u64 *__restrict__ tw, u32 tOff, u32 n,
                        i64 &score, u32 &matches) {
    u32 mism = 0;
    for (u32 o = 0; o < n; o += 32) {
        u32 c = n - o < 32 ? n - o : 32;
        u64 a = qRev ? get_bases_rc(qw, qLen, qOff + o, c) : get_bases(qw, qOff + o, c);
        u64 b = get_bases(tw, tOff + o, c);
        u64 x = a ^ b;
        mism += __popcll((x | (x >> 1)) & 0x5555555555555555ull);
    }
    matches = n - mism;
    score = 2ll * (i64) matches - 3ll * (i64) mism;
}
// exact byte path (IUPAC letters, lower case, '*'): createAsciiSubMat scores + case-folded identity
CDG_D unsigned char q_char(const u8 *q, u32 qLen, bool qRev, u32 i) {   // rescorediagonal.cpp:173-179
    return qRev ? (unsigned char) cdg_num2aa(cdg_rev_code(cdg_map_letter(q[qLen - 1 - i]))) : q[i];
}
CDG_D int sub_score(unsigned char a, unsigned char b) {
    int x = cdg_map_letter(a), y = cdg_map_letter(b);
    return (x == y && x != 4) ? 2 : -3;
}
CDG_D LocalAln by_diagonal(bool fast, const u64 *qw, const u8 *qa, u32 qLen, bool qRev, const u64 *tw, const u8 *ta, u32 tLen, i32 diagonal) {
    LocalAln r; r.startPos = -1; r.endPos = -1; r.score = 0; r.diagonalLen = 0; r.idCnt = 0;
    u32 ad = (u32) (diagonal < 0 ? -diagonal : diagonal);
    r.dist = ad; r.diagonal = diagonal;
    u32 n, qOff, tOff;
    if (diagonal >= 0 && ad < qLen) { n = min(tLen, qLen - ad); qOff = ad; tOff = 0; }
    else if (diagonal < 0 && ad < tLen) { n = min(tLen - ad, qLen); qOff = 0; tOff = ad; }
    else return r;
    r.diagonalLen = n;
    if (fast) {
        i64 s; u32 m;
        score_packed(qw, qLen, qRev, qOff, tw, tOff, n, s, m);
        r.startPos = 0; r.endPos = (i32) n - 1; r.score = (u32) (s > 0 ? s : 0); r.idCnt = m;
    } else {   // computeGlobalSubstitutionStartEndDistance (DistanceCalculator.h:204-220)
        unsigned char q0 = q_char(qa, qLen, qRev, qOff), t0 = ta[tOff];
        unsigned char q1 = q_char(qa, qLen, qRev, qOff + n - 1), t1 = ta[tOff + n - 1];
        u32 first = (q0 == '*' || t0 == '*') ? 1 : 0;
        u32 last = n - 1;
        if (last > 0 && (q1 == '*' || t1 == '*')) last--;
        i64 s = 0; u32 idc = 0;
        for (u32 p = first; p <= last; p++) {
            unsigned char a = q_char(qa, qLen, qRev, qOff + p), b = ta[tOff + p];
            s += sub_score(a, b);
            idc += ((a & (unsigned char) ~0x20) == (b & (unsigned char) ~0x20)) ? 1u : 0u;
        }
        r.startPos = (i32) first; r.endPos = (i32) last; r.score = (u32) (s > 0 ? s : 0); r.idCnt = idc;
    }
    return r;
}

__global__ void __launch_bounds__(256) rescore_kernel(
    const u8 *__restrict__ data, const SeqMeta *__restrict__ meta, const u64 *__restrict__ words,
    const u32 *__restrict__ lineQ, const u32 *__restrict__ lineT, const i32 *__restrict__ linePref, const u16 *__restrict__ lineDiag,
    u64 M, RescoreParams P,
    u8 *__restrict__ keep, i32 *__restrict__ score, i32 *__restrict__ bits, double *__restrict__ evalue, float *__restrict__ seqId,
    i32 *__restrict__ qS, i32 *__restrict__ qE, i32 *__restrict__ dS, i32 *__restrict__ dE, u32 *__restrict__ idCnt,
    unsigned long long *__restrict__ keptCount) {
    u64 l = (u64) blockIdx.x * blockDim.x + threadIdx.x;
    u32 kept = 0;
    if (l < M) {
        const u32 q = lineQ[l], t = lineT[l];
        const SeqMeta qm = meta[q], tm = meta[t];   // one 16-byte load per sequence (offset, first word, length, flags)
        const u32 qLen = qm.lenFlags & META_LEN, tLen = tm.lenFlags & META_LEN;
        const bool isRev = linePref[l] < 0;
        const bool isIdentity = q == t;
        const bool fast = !((qm.lenFlags | tm.lenFlags) & META_EXC);
        const u64 *qw = words + qm.wordOff, *tw = words + tm.wordOff;
        const u8 *qa = data + qm.off, *ta = data + tm.off;
        const u32 diag16 = lineDiag[l];
        LocalAln mx; mx.startPos = -1; mx.endPos = -1; mx.score = 0; mx.diagonalLen = 0; mx.dist = 0; mx.diagonal = 0; mx.idCnt = 0;
        if (qLen > 0 && tLen > 0) {
            for (u32 d = 1; d <= 1 + tLen / 32768; d++) {   // computeUngappedAlignment (DistanceCalculator.h:93-113)
                LocalAln tmp = by_diagonal(fast, qw, qa, qLen, isRev, tw, ta, tLen, (i32) diag16 - (i32) (d * 65536));
                if (tmp.score > mx.score) mx = tmp;
            }
            for (u32 d = 0; d <= qLen / 65536; d++) {
                LocalAln tmp = by_diagonal(fast, qw, qa, qLen, isRev, tw, ta, tLen, (i32) (d * 65536 + diag16));
                if (tmp.score > mx.score) mx = tmp;
            }
        }
        const i32 distance = (i32) mx.score;
        const double ev = cdg_evalue((double) distance, (double) qLen, P.dbRes);
        const i32 bitScore = (i32) (cdg_bitscore((double) distance) + 0.5);
        const i32 alnLen = (mx.endPos - mx.startPos) + 1;
        i32 qs, qe, ds, de;
        if (mx.diagonal >= 0) { qs = mx.startPos + (i32) mx.dist; qe = mx.endPos + (i32) mx.dist; ds = mx.startPos; de = mx.endPos; }
        else { qs = mx.startPos; qe = mx.endPos; ds = mx.startPos + (i32) mx.dist; de = mx.endPos + (i32) mx.dist; }
        float sid = 0.f;
        if ((ev <= P.evalThr || isIdentity) && mx.startPos >= 0) sid = (float) mx.idCnt / (float) alnLen;   // Util::computeSeqId
        if (isRev) { qs = (i32) qLen - qs - 1; qe = (i32) qLen - qe - 1; }
        const bool hasSeqId = (double) sid >= (double) (P.seqIdThr - FLT_EPSILON);
        const bool hasEvalue = ev <= P.evalThr;
        const bool k = isIdentity || (hasSeqId && hasEvalue);   // rescorediagonal.cpp:304-314 with -c 0, --min-aln-len 0
        keep[l] = k ? 1 : 0; score[l] = distance; bits[l] = bitScore; evalue[l] = ev; seqId[l] = sid;
        qS[l] = qs; qE[l] = qe; dS[l] = ds; dE[l] = de; idCnt[l] = mx.idCnt;
        kept = k ? 1u : 0u;
    }
    u32 b = __ballot_sync(0xffffffffu, kept);
    if ((threadIdx.x & 31) == 0 && b) atomicAdd(keptCount, (unsigned long long) __popc(b));
}

// ---- ancient_correction ----------------------------------------------------------------------------------
struct CorrParams { float seqIdThr; float ryThr; int roundSeqId; u32 wideThreshold; };   // roundSeqId: emulate the aln DB's 3-decimal text
struct DamageTables { double logD[11][4][4]; double logDrev[11][4][4]; double logSE[4][4]; };

CDG_D int nuc_map(unsigned char c) { return c == 'C' ? 1 : c == 'G' ? 2 : c == 'T' ? 3 : 0; }   // correction.cpp:170-174, operator[] -> 0
CDG_D int ry_map(unsigned char c) { return (c == 'C' || c == 'T') ? 1 : 0; }                      // :175-180
// target byte in alignment orientation: reverse targets go through getNuclRevFragment (nuclassembleUtil.cpp:67-76)
CDG_D unsigned char t_char(const u8 *t, u32 tLen, bool rev, u32 p) {
    if (!rev) return t[p];
    char r = cdg_num2aa(cdg_rev_code(cdg_map_letter(t[tLen - 1 - p])));
    return (unsigned char) (r == 'X' ? 'N' : r);
}
// Util::fastSeqIdToBuffer + strtod round trip (Util.cpp:278-307, Matcher.cpp:289)
CDG_D float seqid_text_roundtrip(float x) {
    if (x == 1.0f) return 1.0f;
    int i = (int) (x * 1000);
    return (float) ((double) i / 1000.0);
}

// mostLikeliBaseRead (correction.cpp:7-123).  cnt/rv: [t*11+l].  Sums are accumulated in the reference's
// loop order with separately rounded multiplies and adds (no FMA contraction).
CDG_D int most_likely_base(int b, u32 qIter, const u32 *cnt, const u32 *rv, const DamageTables &T, bool wasCorr, u32 qLen,
                           double *liks, bool *evaluated) {
    u32 cov[4];
#pragma unroll
    for (int t = 0; t < 4; t++) { u32 s = 0; for (int l = 0; l < 11; l++) s += cnt[t * 11 + l]; cov[t] = s; }
    *evaluated = false;
    if (!wasCorr) {
        u32 tot = cov[1] + cov[3] + cov[0] + cov[2];
        double ctRatio = (double) cov[3] / (double) tot, gaRatio = (double) cov[0] / (double) tot;
        if (ctRatio >= 0.4 || gaRatio >= 0.4) return b;   // :56-60
    }
    int cq = 5;
    if (qIter < 5) cq = (int) qIter;
    else if (qIter >= qLen - 5) cq = 11 - (int) (qLen - qIter);
    double lik[4];
#pragma unroll 1
    for (int q = 0; q < 4; q++) {
        const double logQ = wasCorr ? T.logSE[q][b] : T.logD[cq][q][b];
        double acc = 0.0;
        for (int t = 0; t < 4; t++) {
            if (cov[t] == 0) continue;
            const double base = __dadd_rn(T.logSE[t][b], logQ);
            for (int l = 0; l < 11; l++) {
                u32 c = cnt[t * 11 + l];
                if (c != 0) {
                    u32 nr = rv[t * 11 + l];
                    acc = __dadd_rn(acc, __dmul_rn((double) (i32) (c - nr), __dadd_rn(base, T.logD[l][q][t])));
                    acc = __dadd_rn(acc, __dmul_rn((double) (i32) nr, __dadd_rn(base, T.logDrev[l][q][t])));
                }
            }
        }
        lik[q] = acc;
    }
    *evaluated = true;
    if (liks) { liks[0] = lik[0]; liks[1] = lik[1]; liks[2] = lik[2]; liks[3] = lik[3]; }
    int best = 0;
    for (int q = 1; q < 4; q++) if (lik[q] > lik[best]) best = q;   // std::max_element: first maximum
    return best;
}

// Generic (slow) kernel: 32-bit counters in local memory, any coverage.  Runs only on the queries the fast kernel
// hands over through `list` (more than 65535 counted targets).
static constexpr int CORR_WARPS = 4;
__global__ void __launch_bounds__(CORR_WARPS * 32) correct_wide_kernel(
    const u32 *__restrict__ list, const u32 *__restrict__ listCount,
    const u8 *__restrict__ data, const u64 *__restrict__ off, const u32 *__restrict__ len, const u8 *__restrict__ ext, u64 qLo, u64 n,
    const u64 *__restrict__ lineOff, const u32 *__restrict__ lineT, const u8 *__restrict__ keep, const float *__restrict__ seqId,
    const i32 *__restrict__ qS_, const i32 *__restrict__ qE_, const i32 *__restrict__ dS_, const i32 *__restrict__ dE_,
    CorrParams P, const DamageTables *__restrict__ tablesG, u8 *__restrict__ lineSel, u8 *__restrict__ out,
    unsigned long long *__restrict__ changed, const u64 *__restrict__ likOff, double *__restrict__ likOut) {
    __shared__ DamageTables T;
    for (int i = threadIdx.x; i < (int) (sizeof(DamageTables) / sizeof(double)); i += blockDim.x) ((double *) &T)[i] = ((const double *) tablesG)[i];
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const u64 gwarp = ((u64) blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = ((u64) gridDim.x * blockDim.x) >> 5;
    u32 nChanged = 0;
    const u64 nList = *listCount;
    for (u64 li = gwarp; li < nList; li += nwarps) {
        const u64 qi = list[li];
        if (qi < qLo || qi >= qLo + n) continue;
        const u64 l0 = lineOff[qi - qLo], l1 = lineOff[qi - qLo + 1];
        const u32 qLen = len[qi];
        const u8 *q = data + off[qi];
        u8 *o = out + off[qi];
        // lines with nothing but (at most) the self alignment leave the sequence untouched: cov <= 1 everywhere
        if (l1 - l0 <= 1) continue;
        const bool qExt = ext[qi] != 0;
        // pass 1 (:220-245): avCov over the alignments the aln DB holds (= kept lines), float accumulate in order
        float avCov = 0.f;
        for (u64 l = l0; l < l1; l++) {
            if (!keep[l]) continue;
            i32 a = max(abs(qE_[l] - qS_[l]), abs(dE_[l] - dS_[l])) + 1;   // Matcher::computeAlnLength
            avCov += (float) a;
        }
        avCov = avCov / (float) qLen;
        // pass 2 (:273-323, :353-359): which targets are counted
        u32 nSel = 0;
        for (u64 l = l0; l < l1; l++) {
            u8 sel = 0;
            if (keep[l]) {
                const u32 t = lineT[l];
                if (!ext[t]) {
                    const u32 dL = len[t];
                    i32 qs = qS_[l], qe = qE_[l], ds = dS_[l], de = dE_[l];
                    const bool sane = (((u32) qs | (u32) qe | (u32) ds | (u32) de) >> 30) == 0;   // no negative / overflowing coordinate
                    const i32 aln = sane ? max(abs(qe - qs), abs(de - ds)) + 1 : 0;
                    bool rev = false;
                    if (qs > qe) { i32 x = qs; qs = qe; qe = x; i32 d0 = ds; ds = (i32) dL - de - 1; de = (i32) dL - d0 - 1; rev = true; }
                    const u8 *ts = data + off[t];
                    u32 dist = 0;
                    // A target counts only with aln >= 30 (:359), so shorter alignments need no RY distance; that also covers the
                    // records of score-0 alignments (all-N self hits: start = end = -1, DistanceCalculator.h:204-220), where the
                    // reference reads the byte in front of the sequence.  Coordinates outside the sequences (possible only in
                    // caller-provided alignments; undefined behaviour in the reference) never reach a load here.
                    const bool usable = aln >= 30 && qs >= 0 && ds >= 0 && (u32) (qs + aln) <= qLen && (u32) (ds + aln) <= dL;
                    if (usable)
                    for (i32 p = lane; p < aln; p += 32) dist += (ry_map(q[qs + p]) != ry_map(t_char(ts, dL, rev, (u32) (ds + p)))) ? 1u : 0u;
#pragma unroll
                    for (int s = 16; s > 0; s >>= 1) dist += __shfl_xor_sync(0xffffffffu, dist, s);
                    const float ry = (float) ((u32) aln - dist) / (float) aln;   // getRYSeqId
                    float thr = P.ryThr;
                    if (aln <= 100) { thr = ((float) aln - 1) / (float) aln; thr = floorf(thr * 1000) / 1000; }
                    const bool pick = (ry >= thr && ds == 0 && (u32) qe == qLen - 1) || (ry >= thr && qs == 0 && de == (i32) dL - 1) ||
                                      (ry >= thr && avCov < 50);
                    const float sid = P.roundSeqId ? seqid_text_roundtrip(seqId[l]) : seqId[l];
                    if (usable && pick && ry >= thr && sid >= P.seqIdThr) sel = rev ? 2 : 1;
                }
            }
            if (lane == 0) lineSel[l] = sel;
            nSel += sel ? 1u : 0u;
        }
        __syncwarp();
        if (nSel <= 1) continue;   // every column has coverage <= 1
        // pass 3+4, 32 query columns at a time (lane = column)
        for (u32 c0 = 0; c0 < qLen; c0 += 32) {
            const u32 col = c0 + lane;
            u32 cnt[44], rv[44];
#pragma unroll
            for (int i = 0; i < 44; i++) { cnt[i] = 0; rv[i] = 0; }
            u32 tot = 0;
            for (u64 l = l0; l < l1; l++) {
                const u8 sel = lineSel[l];
                if (!sel) continue;
                const u32 t = lineT[l];
                const u32 dL = len[t];
                i32 qs = qS_[l], qe = qE_[l], ds = dS_[l], de = dE_[l];
                const bool rev = sel == 2;
                if (rev) { i32 x = qs; qs = qe; qe = x; i32 d0 = ds; ds = (i32) dL - de - 1; de = (i32) dL - d0 - 1; }
                const i32 aln = max(abs(qe - qs), abs(de - ds)) + 1;
                if (col < qLen && (i32) col >= qs && (i32) col < qs + aln) {
                    const u32 tp = (u32) ds + (col - (u32) qs);
                    const int cl = ((i32) tp >= (i32) dL - 5) ? 6 + (int) (tp - (dL - 5)) : (tp < 5 ? (int) tp : 5);   // :361-371
                    const int bse = nuc_map(t_char(data + off[t], dL, rev, tp));
                    cnt[bse * 11 + cl] += 1; rv[bse * 11 + cl] += rev ? 1u : 0u; tot += 1;
                }
            }
            if (col < qLen) {
                const unsigned char orig = q[col];
                unsigned char res = orig;
                if (tot > 1) {
                    double liks[4]; bool evaluated;
                    int nb = most_likely_base(nuc_map(orig), col, cnt, rv, T, qExt, qLen, likOut ? liks : nullptr, &evaluated);
                    res = (unsigned char) "ACGT"[nb];
                    if (likOut && evaluated) { double *lo = likOut + (likOff[qi] + col) * 4; lo[0] = liks[0]; lo[1] = liks[1]; lo[2] = liks[2]; lo[3] = liks[3]; }
                }
                if (res != orig) { o[col] = res; nChanged++; }
            }
        }
    }
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) nChanged += __shfl_xor_sync(0xffffffffu, nChanged, s);
    if (lane == 0 && nChanged) atomicAdd(changed, (unsigned long long) nChanged);
}


// ---- fast kernel ---------------------------------------------------------------------------------------------
// One warp per query.  pass 1: exact integer sum of alignment lengths (lane per line).  pass 2: lane per line,
// RY identity from the 2-bit words (xor + popcount) unless a sequence holds exception bytes; the counted targets
// are compacted into a per-warp shared-memory line cache (up to 32; more fall back to re-reading the line arrays).
// pass 3: lane per query column; per-column counters count[t][l] | reverse[t][l] << 16 live in shared memory, a
// 44-bit mask remembers which are non-zero so that the likelihood visits only those, in the reference's (t,l) order.
static constexpr int CF_WARPS = 8;
static constexpr int CF_CACHE = 32;
using ::SeqMeta; using ::META_EXC; using ::META_EXT; using ::META_LEN;
struct LineCache { u64 tOff[CF_CACHE]; u32 wOff[CF_CACHE]; u32 dL[CF_CACHE]; i32 qs[CF_CACHE]; i32 aln[CF_CACHE]; i32 ds[CF_CACHE]; u32 rev[CF_CACHE]; };   // rev: bit 0 reverse strand, bit 1 target is plain ACGT (packed words usable)
static constexpr u32 CF_COMPACT_LEN = 96;   // queries up to this length evaluate only the columns where some counted target disagrees
struct CorrSmem { DamageTables T; u32 cnt[CF_WARPS][44][32]; LineCache lc[CF_WARPS]; u8 lut[256]; u8 colList[CF_WARPS][CF_COMPACT_LEN]; int domOK; };   // lut: bits 0-1 forward base index, 2-3 reverse-complement base index

__global__ void __launch_bounds__(256) meta_kernel(const u64 *__restrict__ off, const u64 *__restrict__ wordOff, const u32 *__restrict__ len,
                                                   const u8 *__restrict__ flags, const u8 *__restrict__ ext, u64 seqLo, u64 seqHi, SeqMeta *__restrict__ meta) {
    u64 i = seqLo + (u64) blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= seqHi) return;
    SeqMeta m; m.off = off[i]; m.wordOff = (u32) wordOff[i];
    m.lenFlags = (len[i] & META_LEN) | (flags[i] ? META_EXC : 0u) | (ext[i] ? META_EXT : 0u);
    meta[i] = m;
}

CDG_D u32 ry_dist_packed(const u64 *__restrict__ qw, u32 qs, const u64 *__restrict__ tw, u32 dL, bool rev, u32 ds, u32 aln) {
    u32 dist = 0;
    for (u32 o = 0; o < aln; o += 32) {
        u32 c = aln - o < 32 ? aln - o : 32;
        u64 a = get_bases(qw, qs + o, c);
        u64 b = rev ? get_bases_rc(tw, dL, ds + o, c) : get_bases(tw, ds + o, c);
        u64 x = a ^ b;
        dist += __popcll((x ^ (x >> 1)) & 0x5555555555555555ull);
    }
    return dist;
}

// RY distance (getRYSeqId) and, per 32-column query word, the columns where the oriented target differs from the query:
// bit 62-2j of m<w> = column 32w+j mismatches (words beyond the third are not recorded; the caller only uses the masks
// for queries of at most CF_COMPACT_LEN bases).  Same popcounts as ry_dist_packed, chunked by query word.
CDG_D u32 ry_dist_mismatch(const u64 *__restrict__ qw, u32 qs, const u64 *__restrict__ tw, u32 dL, bool rev, u32 ds, u32 aln, u64 &m0, u64 &m1, u64 &m2) {
    u32 dist = 0;
    const u32 qe = qs + aln;
    auto word = [&](u32 w, u64 &m, bool record) {
        const u32 lo = max(qs, w << 5), hi = min(qe, (w + 1) << 5), c = hi - lo;
        const u64 a = get_bases(qw, lo, c);
        const u32 tp = ds + (lo - qs);
        const u64 b = rev ? get_bases_rc(tw, dL, tp, c) : get_bases(tw, tp, c);
        const u64 x = a ^ b;
        dist += __popcll((x ^ (x >> 1)) & 0x5555555555555555ull);
        if (record) m = ((x | (x >> 1)) & 0x5555555555555555ull) >> (2 * (lo & 31));
    };
    u64 unused = 0;
    if (qs < 32 && qe > 0) word(0, m0, true);
    if (qs < 64 && qe > 32) word(1, m1, true);
    if (qs < 96 && qe > 64) word(2, m2, true);
    for (u32 w = max(3u, qs >> 5); (w << 5) < qe; w++) word(w, unused, false);
    return dist;
}

// one target column into the per-column counters; lut maps the raw target byte to its A0 C1 G2 T3 index
// (forward: correction.cpp:170-174; reverse: getNuclRevFragment, nuclassembleUtil.cpp:67-76, then the same map)
CDG_D void count_column(const u8 *__restrict__ data, u64 tOff, u32 dL, bool rev, u32 tp, u32 (*cnt)[32], int lane, const u8 *lut,
                        unsigned long long &mask) {
    const int cl = ((i32) tp >= (i32) dL - 5) ? 6 + (int) (tp - (dL - 5)) : (tp < 5 ? (int) tp : 5);   // correction.cpp:361-371
    const u32 e = lut[data[tOff + (rev ? dL - 1 - tp : tp)]];
    const int bse = rev ? (int) (e >> 2) : (int) (e & 3u);
    const int idx = bse * 11 + cl;
    cnt[idx][lane] += rev ? 0x10001u : 1u;
    mask |= 1ull << idx;
}

// same from the 2-bit words (plain ACGT targets): the words were just read for the RY distance, so this load hits L1 while the
// byte of the raw DB would come from DRAM.  Packed codes A0 C1 T2 G3 -> counter index A0 C1 G2 T3 = code ^ (code >> 1);
// complement = code ^ 2.
CDG_D void count_column_packed(const u64 *__restrict__ tw, u32 dL, bool rev, u32 tp, u32 (*cnt)[32], int lane, unsigned long long &mask) {
    const int cl = ((i32) tp >= (i32) dL - 5) ? 6 + (int) (tp - (dL - 5)) : (tp < 5 ? (int) tp : 5);   // correction.cpp:361-371
    const u32 sp = rev ? dL - 1 - tp : tp;
    u32 code = (u32) (tw[sp >> 5] >> (62 - 2 * (sp & 31))) & 3u;
    if (rev) code ^= 2u;
    const int idx = (int) (code ^ (code >> 1)) * 11 + cl;
    cnt[idx][lane] += rev ? 0x10001u : 1u;
    mask |= 1ull << idx;
}

// ---- per-line preparation of the correction (correction.cpp:273-323, :353-359), one thread per alignment line ------------
// Everything of pass 2 that does not depend on the query's average coverage: the RY identity of the pair (getRYSeqId), the
// thresholds, and — for plain ACGT pairs — the columns where the oriented target differs from the query.  The warp-per-query
// kernel then only combines lines (it used to do this work with one lane per line, mostly idle lanes).
//   prep[l]: bit 0 the line can be counted (usable, RY identity and sequence identity above their thresholds, target not
//            ext-flagged), bit 1 it satisfies one of the two end conditions (:303-311), bit 2 reverse strand, bit 3 a sequence of
//            the pair needs the byte path
//   alnLen[l]: Matcher::computeAlnLength of a kept line (0 otherwise) for the average coverage of pass 1 (:220-245)
//   mism[3l..3l+2]: mismatch columns of query words 0..2 (bit 62-2j = column 32w+j), plain pairs only
static constexpr u8 PREP_CAND = 1, PREP_END = 2, PREP_REV = 4, PREP_BYTES = 8;
__global__ void __launch_bounds__(256) corr_prep_kernel(
    const u8 *__restrict__ data, const SeqMeta *__restrict__ meta, const u64 *__restrict__ words, u64 qLo, const u64 *__restrict__ lineOff,
    const u32 *__restrict__ lineQ, const u32 *__restrict__ lineT, const u8 *__restrict__ keep, const float *__restrict__ seqId,
    const i32 *__restrict__ qS_, const i32 *__restrict__ qE_, const i32 *__restrict__ dS_, const i32 *__restrict__ dE_, u64 M,
    CorrParams P, u8 *__restrict__ prep, u32 *__restrict__ alnLen, u64 *__restrict__ mism) {
    const u64 l = (u64) blockIdx.x * blockDim.x + threadIdx.x;
    if (l >= M) return;
    u8 pr = 0; u32 al = 0;
    if (keep[l]) {
        i32 qs = qS_[l], qe = qE_[l], ds = dS_[l], de = dE_[l];
        al = (u32) (max(abs(qe - qs), abs(de - ds)) + 1);
        const u32 q = lineQ[l];
        const u64 ql = q - qLo;
        if (lineOff[ql + 1] - lineOff[ql] > 1) {   // a query with a single line is never touched (coverage <= 1)
            const SeqMeta qm = meta[q], tm = meta[lineT[l]];
            if (!(tm.lenFlags & META_EXT)) {
                const u32 qLen = qm.lenFlags & META_LEN, dL = tm.lenFlags & META_LEN;
                const bool sane = (((u32) qs | (u32) qe | (u32) ds | (u32) de) >> 30) == 0;   // no negative / overflowing coordinate
                const i32 aln = sane ? (i32) al : 0;
                bool rev = false;
                if (qs > qe) { i32 x = qs; qs = qe; qe = x; i32 d0 = ds; ds = (i32) dL - de - 1; de = (i32) dL - d0 - 1; rev = true; }
                const bool packed = !((qm.lenFlags | tm.lenFlags) & META_EXC);
                // (see correct_wide_kernel: alignments shorter than 30 never count, out-of-sequence coordinates never load)
                const bool usable = aln >= 30 && qs >= 0 && ds >= 0 && (u32) (qs + aln) <= qLen && (u32) (ds + aln) <= dL;
                u32 dist = 0;
                u64 m0 = 0, m1 = 0, m2 = 0;
                if (!usable) {}
                else if (packed) dist = ry_dist_mismatch(words + qm.wordOff, (u32) qs, words + tm.wordOff, dL, rev, (u32) ds, (u32) aln, m0, m1, m2);
                else {
                    const u8 *qb = data + qm.off, *ts = data + tm.off;
                    for (i32 p = 0; p < aln; p++) dist += (ry_map(qb[qs + p]) != ry_map(t_char(ts, dL, rev, (u32) (ds + p)))) ? 1u : 0u;
                }
                const float ry = (float) ((u32) aln - dist) / (float) aln;   // getRYSeqId (nuclassembleUtil.cpp:78-92)
                float thr = P.ryThr;
                if (aln <= 100) { thr = ((float) aln - 1) / (float) aln; thr = floorf(thr * 1000) / 1000; }
                const float sid = P.roundSeqId ? seqid_text_roundtrip(seqId[l]) : seqId[l];
                if (usable && ry >= thr && sid >= P.seqIdThr) {
                    pr = PREP_CAND | (rev ? PREP_REV : 0) | (packed ? 0 : PREP_BYTES);
                    if ((ds == 0 && (u32) qe == qLen - 1) || (qs == 0 && de == (i32) dL - 1)) pr |= PREP_END;
                    if (packed) { mism[3 * l] = m0; mism[3 * l + 1] = m1; mism[3 * l + 2] = m2; }
                }
            }
        }
    }
    prep[l] = pr; alnLen[l] = al;
}

// queries with more than one line (only those can change): compact list, order irrelevant
__global__ void __launch_bounds__(256) multi_line_queries_kernel(const u64 *__restrict__ lineOff, u64 n, u32 *__restrict__ list, u32 *__restrict__ count) {
    const u64 q = (u64) blockIdx.x * blockDim.x + threadIdx.x;
    const bool multi = q < n && lineOff[q + 1] - lineOff[q] > 1;
    const u32 b = __ballot_sync(0xffffffffu, multi);
    if (!b) return;
    const int lane = threadIdx.x & 31;
    u32 base = 0;
    if (lane == 0) base = atomicAdd(count, (u32) __popc(b));
    base = __shfl_sync(0xffffffffu, base, 0);
    if (multi) list[base + __popc(b & ((1u << lane) - 1))] = (u32) q;
}

__global__ void __launch_bounds__(CF_WARPS * 32, 4) correct_fast_kernel(
    const u32 *__restrict__ qList, const u32 *__restrict__ qCount,
    const u8 *__restrict__ data, const SeqMeta *__restrict__ meta, const u64 *__restrict__ words, u64 qLo, u64 n,
    const u64 *__restrict__ lineOff, const u32 *__restrict__ lineT, const u8 *__restrict__ keep, const float *__restrict__ seqId,
    const i32 *__restrict__ qS_, const i32 *__restrict__ qE_, const i32 *__restrict__ dS_, const i32 *__restrict__ dE_,
    CorrParams P, const DamageTables *__restrict__ tablesG, u8 *__restrict__ lineSel, u8 *__restrict__ out,
    unsigned long long *__restrict__ changed, const u64 *__restrict__ likOff, double *__restrict__ likOut,
    u32 *__restrict__ wideList, u32 *__restrict__ wideCount,
    u64 *__restrict__ patches /*out == nullptr: corrected bases go here as byte offset << 8 | letter*/, u64 patchCap, unsigned long long *__restrict__ patchCount,
    const u8 *__restrict__ prep, const u32 *__restrict__ alnLen, const u64 *__restrict__ mism /*corr_prep_kernel*/) {
    extern __shared__ __align__(16) unsigned char smemRaw[];
    CorrSmem &S = *reinterpret_cast<CorrSmem *>(smemRaw);
    for (int i = threadIdx.x; i < (int) (sizeof(DamageTables) / sizeof(double)); i += blockDim.x) ((double *) &S.T)[i] = ((const double *) tablesG)[i];
    for (int i = threadIdx.x; i < CF_WARPS * 44 * 32; i += blockDim.x) (&S.cnt[0][0][0])[i] = 0;
    for (int i = threadIdx.x; i < 256; i += blockDim.x) {
        const int fwd = nuc_map((unsigned char) i);
        const int rvi = nuc_map(t_char((const u8 *) &i, 1, true, 0));   // little-endian: the low byte of i is the character
        S.lut[i] = (u8) (fwd | (rvi << 2));
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        // A column whose counted target bases all equal the query base b keeps b whenever, for every other candidate x, each
        // term of lik[b] is >= the matching term of lik[x] and the query term is larger by a margin far above the rounding
        // error of the sums (mostLikeliBaseRead, correction.cpp:61-117; the :56-60 early return keeps b as well).  The tables
        // decide whether that holds (it fails only for substitution rates above 0.5); if not, every column is evaluated.
        bool ok = true;
        for (int b = 0; b < 4; b++)
            for (int x = 0; x < 4; x++) {
                if (x == b) continue;
                if (!(S.T.logSE[b][b] - S.T.logSE[x][b] >= 1e-6)) ok = false;
                for (int l = 0; l < 11; l++) {
                    if (!(S.T.logD[l][b][b] - S.T.logD[l][x][b] >= 1e-6)) ok = false;
                    if (!(S.T.logDrev[l][b][b] >= S.T.logDrev[l][x][b])) ok = false;
                }
            }
        S.domOK = ok ? 1 : 0;
    }
    __syncthreads();
    const DamageTables &T = S.T;
    const u8 *lut = S.lut;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    u32 (*cnt)[32] = S.cnt[warp];
    LineCache &LC = S.lc[warp];
    const u64 gwarp = ((u64) blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = ((u64) gridDim.x * blockDim.x) >> 5;
    u32 nChanged = 0;
    const u64 nList = *qCount;
    for (u64 li = gwarp; li < nList; li += nwarps) {
        const u64 ql = qList[li];   // local query index; queries with a single line are not in the list (coverage <= 1)
        if (ql >= n) continue;
        const u64 l0 = lineOff[ql], l1 = lineOff[ql + 1];
        const u64 qi = qLo + ql;
        const SeqMeta qm = meta[qi];
        const u32 qLen = qm.lenFlags & META_LEN;
        const u8 *q = data + qm.off;
        const u64 *qw = words + qm.wordOff;
        const bool qExc = (qm.lenFlags & META_EXC) != 0;
        const bool qExt = (qm.lenFlags & META_EXT) != 0;
        // pass 1 (correction.cpp:220-245): exact integer sum of the alignment lengths (corr_prep_kernel: 0 for lines not kept) -> avCov
        unsigned long long sum = 0;
        for (u64 l = l0 + lane; l < l1; l += 32) sum += (unsigned long long) alnLen[l];
#pragma unroll
        for (int s = 16; s > 0; s >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, s);
        float avCov;
        if (sum < (1ull << 24)) avCov = (float) sum;   // every partial sum of the reference's float accumulation is exact
        else {
            avCov = 0.f;
            for (u64 l = l0; l < l1; l++) if (keep[l]) avCov += (float) alnLen[l];
        }
        avCov = avCov / (float) qLen;
        // pass 2 (:273-323, :353-359), one line per lane: the per-line part comes from corr_prep_kernel, the coverage rule is applied here
        u32 nSel = 0;
        u64 mm0 = 0, mm1 = 0, mm2 = 0;   // columns where one of this lane's counted targets differs from the query
        bool compact = S.domOK && likOut == nullptr && !qExc && qLen <= CF_COMPACT_LEN;
        for (u64 lb = l0; lb < l1; lb += 32) {
            const u64 l = lb + lane;
            u8 sel = 0;
            const u8 pr = l < l1 ? prep[l] : (u8) 0;
            if ((pr & PREP_CAND) && ((pr & PREP_END) || avCov < 50)) sel = (pr & PREP_REV) ? 2 : 1;
            if (sel && (pr & PREP_BYTES)) compact = false;   // a counted target needs the byte path (made warp-uniform below)
            if (l < l1) lineSel[l] = sel;
            const u32 b = __ballot_sync(0xffffffffu, sel != 0);
            if (sel) {
                if (!(pr & PREP_BYTES)) { mm0 |= mism[3 * l]; mm1 |= mism[3 * l + 1]; mm2 |= mism[3 * l + 2]; }
                const u32 slot = nSel + __popc(b & ((1u << lane) - 1));
                if (slot < CF_CACHE) {
                    const SeqMeta tm = meta[lineT[l]];
                    const u32 dL = tm.lenFlags & META_LEN;
                    i32 qs = qS_[l], qe = qE_[l], ds = dS_[l], de = dE_[l];
                    const i32 aln = max(abs(qe - qs), abs(de - ds)) + 1;
                    if (sel == 2) { i32 x = qs; qs = qe; qe = x; ds = (i32) dL - de - 1; }
                    LC.tOff[slot] = tm.off; LC.dL[slot] = dL; LC.qs[slot] = qs; LC.aln[slot] = aln; LC.ds[slot] = ds;
                    LC.rev[slot] = (sel == 2 ? 1u : 0u) | ((tm.lenFlags & META_EXC) ? 0u : 2u); LC.wOff[slot] = tm.wordOff;
                }
            }
            nSel += __popc(b);
        }
        compact = __all_sync(0xffffffffu, compact);
        if (nSel <= 1) continue;
        if (nSel > P.wideThreshold) {   // 16-bit counters could overflow (> 65535): hand over to the generic kernel
            if (lane == 0) { u32 s = atomicAdd(wideCount, 1u); wideList[s] = (u32) qi; }
            continue;
        }
        __syncwarp();
        u8 *o = out + qm.off;
        // columns to evaluate: all of them, or (short plain queries, dominance holds) only those where a counted target disagrees
        u32 nCols = qLen;
        u8 *colList = S.colList[warp];
        if (compact) {
            nCols = 0;
#pragma unroll
            for (int w = 0; w < 3; w++) {
                const u64 mine = w == 0 ? mm0 : (w == 1 ? mm1 : mm2);
                const u32 lo32 = __reduce_or_sync(0xffffffffu, (u32) mine), hi32 = __reduce_or_sync(0xffffffffu, (u32) (mine >> 32));
                const u64 need = ((u64) hi32 << 32) | lo32;
                const u32 c = 32 * w + lane;
                const bool bit = ((need >> (62 - 2 * lane)) & 1ull) && c < qLen;
                const u32 b = __ballot_sync(0xffffffffu, bit);
                if (bit) colList[nCols + __popc(b & ((1u << lane) - 1))] = (u8) c;
                nCols += __popc(b);
            }
            __syncwarp();
        }
        for (u32 c0 = 0; c0 < nCols; c0 += 32) {
            const bool valid = c0 + lane < nCols;
            const u32 col = compact ? (valid ? colList[c0 + lane] : 0u) : c0 + lane;
            unsigned long long mask = 0;
            if (nSel <= CF_CACHE) {
#pragma unroll 4
                for (u32 i = 0; i < nSel; i++) {
                    const i32 qs = LC.qs[i], aln = LC.aln[i];
                    if (valid && (i32) col >= qs && (i32) col < qs + aln) {
                        const u32 rv = LC.rev[i];
                        if (rv & 2u) count_column_packed(words + LC.wOff[i], LC.dL[i], (rv & 1u) != 0, (u32) LC.ds[i] + (col - (u32) qs), cnt, lane, mask);
                        else count_column(data, LC.tOff[i], LC.dL[i], (rv & 1u) != 0, (u32) LC.ds[i] + (col - (u32) qs), cnt, lane, lut, mask);
                    }
                }
            } else {
                for (u64 l = l0; l < l1; l++) {
                    const u8 sel = lineSel[l];
                    if (!sel) continue;
                    const SeqMeta tm = meta[lineT[l]];
                    const u32 dL = tm.lenFlags & META_LEN;
                    i32 qs = qS_[l], qe = qE_[l], ds = dS_[l], de = dE_[l];
                    const bool rev = sel == 2;
                    if (rev) { i32 x = qs; qs = qe; qe = x; i32 d0 = ds; ds = (i32) dL - de - 1; de = (i32) dL - d0 - 1; }
                    const i32 aln = max(abs(qe - qs), abs(de - ds)) + 1;
                    if (valid && (i32) col >= qs && (i32) col < qs + aln) {
                        if (tm.lenFlags & META_EXC) count_column(data, tm.off, dL, rev, (u32) ds + (col - (u32) qs), cnt, lane, lut, mask);
                        else count_column_packed(words + tm.wordOff, dL, rev, (u32) ds + (col - (u32) qs), cnt, lane, mask);
                    }
                }
            }
            // coverage per target base from the non-zero counters
            u32 cov0 = 0, cov1 = 0, cov2 = 0, cov3 = 0;
            {
                unsigned long long m = mask;
                while (m) {
                    const int idx = __ffsll((long long) m) - 1;
                    m &= m - 1;
                    const u32 c = cnt[idx][lane] & 0xFFFFu;
                    if (idx < 11) cov0 += c; else if (idx < 22) cov1 += c; else if (idx < 33) cov2 += c; else cov3 += c;
                }
            }
            const u32 tot = cov0 + cov1 + cov2 + cov3;
            if (valid && tot > 1) {
                // plain queries: the letter comes from the 2-bit words already in cache (the DB bytes may not even exist: packed uploads)
                const unsigned char orig = qExc ? q[col] : (unsigned char) cdg_num2aa((int) ((qw[col >> 5] >> (62 - 2 * (col & 31))) & 3ull));
                const int b = nuc_map(orig);
                int nb = b;
                bool evaluate = true;
                if (!qExt) {   // mostLikeliBaseRead :56-60
                    const u32 sumc = cov1 + cov3 + cov0 + cov2;
                    const double ctRatio = (double) cov3 / (double) sumc, gaRatio = (double) cov0 / (double) sumc;
                    if (ctRatio >= 0.4 || gaRatio >= 0.4) evaluate = false;
                }
                if (evaluate) {
                    int cq = 5;
                    if (col < 5) cq = (int) col;
                    else if (col >= qLen - 5) cq = 11 - (int) (qLen - col);
                    double logQ[4], acc[4];
#pragma unroll
                    for (int x = 0; x < 4; x++) { logQ[x] = qExt ? T.logSE[x][b] : T.logD[cq][x][b]; acc[x] = 0.0; }
                    unsigned long long m = mask;
                    while (m) {   // ascending idx = t*11+l: the reference's loop order (:80-117)
                        const int idx = __ffsll((long long) m) - 1;
                        m &= m - 1;
                        const int t = idx / 11, l = idx - 11 * t;
                        const u32 v = cnt[idx][lane];
                        const u32 nr = v >> 16, fw = (v & 0xFFFFu) - nr;
                        const double lt = T.logSE[t][b];
#pragma unroll
                        for (int x = 0; x < 4; x++) {
                            const double base = __dadd_rn(lt, logQ[x]);
                            if (fw) acc[x] = __dadd_rn(acc[x], __dmul_rn((double) fw, __dadd_rn(base, T.logD[l][x][t])));
                            if (nr) acc[x] = __dadd_rn(acc[x], __dmul_rn((double) nr, __dadd_rn(base, T.logDrev[l][x][t])));
                        }
                    }
                    nb = 0;
#pragma unroll
                    for (int x = 1; x < 4; x++) if (acc[x] > acc[nb]) nb = x;   // first maximum (std::max_element)
                    if (likOut) { double *lo = likOut + (likOff[qi] + col) * 4; lo[0] = acc[0]; lo[1] = acc[1]; lo[2] = acc[2]; lo[3] = acc[3]; }
                }
                const unsigned char res = (unsigned char) "ACGT"[nb];
                if (res != orig) {
                    nChanged++;
                    if (out) o[col] = res;
                    else { const u64 slot = atomicAdd(patchCount, 1ull); if (slot < patchCap) patches[slot] = ((qm.off + col) << 8) | (u64) res; }
                }
            }
            while (mask) { const int idx = __ffsll((long long) mask) - 1; mask &= mask - 1; cnt[idx][lane] = 0; }
        }
        __syncwarp();   // the line cache is rewritten by the next query
    }
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) nChanged += __shfl_xor_sync(0xffffffffu, nChanged, s);
    if (lane == 0 && nChanged) atomicAdd(changed, (unsigned long long) nChanged);
}

}   // namespace cdgalign
