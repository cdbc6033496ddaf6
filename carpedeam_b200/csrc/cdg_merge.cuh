// Scoring front half of ancient_contig_merge on the device: everything doNuclAssembly2 computes per query between
// reading its alignment lines and pushing scored merge candidates into the priority queue (the greedy merge itself stays on the host;
// the queue's comparator is cdh_merge_less in carpedeam_b200/host/merge.cpp).
//   merge_lines_kernel    orientation, identity / RY identity against the query, merge-candidate filter
//                         (src/assembler/ancientContigsResults.cpp:187-236)
//   ext_consensus_kernel  consensusCaller (src/assembler/nuclassembleUtil.cpp:570-702) — the function N1 uses, same kernel; only in
//                         --unsafe 1: in safe mode the consensus is N^qLen | query | N^qLen and is never materialised
//   merge_score_kernel    updateSeqIdConsensus (:704-790), the queue filter (ancientContigsResults.cpp:246-250) and ancientMatchCount with
//                         deamMatches (nuclassembleUtil.cpp:1050-1180, :1011-1047)
// One warp per alignment line: contig overlaps are hundreds to tens of thousands of columns.
// flags per line: bit 0 merge candidate, bit 1 isRevToAlignment, bit 2 queued (same bit layout as cdg_ext.cuh).
#pragma once
#include "cdg_ext.cuh"

namespace cdgmerge {
using namespace cdgext;

static constexpr u8 MRG_CAND = EXT_CAND, MRG_REV = EXT_REV, MRG_QUEUED = 4;
static constexpr u32 MRG_MAX_LOG_LEN = 100000;   // `unsigned int maxLength = 1e5` (nuclassembleUtil.cpp:1018)

// host-computed pieces of deamMatches (glibc log, like the reference): logPower[len] = log(1.4e-9) - 3*log(len) for len <= 1e5
struct MergeModel {
    double matchLik[2][2];   // [isRevToAlignment][0: C->T, 1: G->A] = D / Drev [5].p[q][t] narrowed to double (:1117, :1151)
    double absLogMin, absLogMax;   // |logPower(10)|, |logPower(1e5)|
};
struct MergeParams { float mergeThr, ryThr; int unsafe; };

CDG_D u32 warp_sum(u32 v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

__global__ void __launch_bounds__(256) merge_lines_kernel(
    const u8 *__restrict__ data, const SeqMeta *__restrict__ meta, const u32 *__restrict__ keys /*null: key == index*/, const u32 *__restrict__ lineQ,
    const u32 *__restrict__ lineT, const u8 *__restrict__ keep, const i32 *__restrict__ qS_, const i32 *__restrict__ qE_, const i32 *__restrict__ dS_,
    const i32 *__restrict__ dE_, u64 M, MergeParams P, u8 *__restrict__ flags, i32 *__restrict__ oqS, i32 *__restrict__ oqE, i32 *__restrict__ odS,
    i32 *__restrict__ odE, float *__restrict__ seqId0, float *__restrict__ ry0) {
    const int lane = threadIdx.x & 31;
    const u64 gwarp = ((u64) blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = ((u64) gridDim.x * blockDim.x) >> 5;
    for (u64 l = gwarp; l < M; l += nwarps) {
        const u32 q = lineQ[l], t = lineT[l];
        const SeqMeta qm = meta[q], tm = meta[t];
        const u32 qLen = qm.lenFlags & META_LEN, dbLen = tm.lenFlags & META_LEN;
        i32 qs = qS_[l], qe = qE_[l], ds = dS_[l], de = dE_[l];
        const u32 alnLen = (u32) (max(abs(qe - qs), abs(de - ds)) + 1);   // Matcher::computeAlnLength
        u8 fl = 0;
        float sid = 0.f, ry = 0.f;
        if (keep[l]) {
            const bool rev = qs > qe;   // :195-204
            if (rev) { const i32 x = qs; qs = qe; qe = x; const i32 d0 = ds; ds = (i32) dbLen - de - 1; de = (i32) dbLen - d0 - 1; fl |= MRG_REV; }
            // :213-222; coordinates outside the sequences (undefined in the reference) never load
            if (qs >= 0 && qe >= qs && (u32) qe < qLen && ds >= 0 && (u32) (ds + (qe - qs)) < dbLen) {
                const u8 *qb = data + qm.off, *tb = data + tm.off;
                u32 idCnt = 0, idRy = 0;
                for (i32 i = qs + lane; i <= qe; i += 32) {
                    const unsigned char qc = qb[i], tc = t_char(tb, dbLen, rev, (u32) (ds + (i - qs)));
                    idCnt += qc == tc ? 1u : 0u; idRy += ry_map(qc) == ry_map(tc) ? 1u : 0u;
                }
                idCnt = warp_sum(idCnt); idRy = warp_sum(idRy);
                sid = (float) (i32) idCnt / (float) alnLen; ry = (float) (i32) idRy / (float) alnLen;
                // :224-229; an end overlap whose overhang is longer than the query would leave the 3*qLen consensus (undefined in the
                // reference; kmermatcher's representative is the longest member of its group, so it does not occur): no candidate
                bool right, left;
                ext_ends(qs, qe, ds, de, qLen, dbLen, right, left);
                if (sid >= P.mergeThr && ry >= P.ryThr && (keys ? keys[t] != keys[q] : t != q) && (!(right || left) || dbLen - alnLen <= qLen))
                    fl |= MRG_CAND;
            }
        }
        if (lane == 0) { flags[l] = fl; oqS[l] = qs; oqE[l] = qe; odS[l] = ds; odE[l] = de; seqId0[l] = sid; ry0[l] = ry; }
    }
}

// deamMatches (nuclassembleUtil.cpp:1011-1047) in the reference's operation order, every operation rounded on its own
CDG_D double deam_posterior(u32 alnLen, u32 scoreAln, double matchLik, const double *__restrict__ logPower, const MergeModel &Mo) {
    const double logLength = logPower[min(alnLen, MRG_MAX_LOG_LEN)];
    const double fractionLength = __ddiv_rn(__dsub_rn(fabs(logLength), Mo.absLogMax), __dsub_rn(Mo.absLogMin, Mo.absLogMax));
    const double priorAln = __dsub_rn(1.0, fractionLength);
    // 0.5f * ((((double) scoreAln + 3.0f * alnLength) / 5.0f + 0.9f) / (alnLength + 1)) + 0.5f * priorAln
    const double num = __dadd_rn(__ddiv_rn(__dadd_rn((double) scoreAln, (double) __fmul_rn(3.0f, (float) alnLen)), 5.0), (double) 0.9f);
    const double pMatch = __dadd_rn(__dmul_rn(0.5, __ddiv_rn(num, (double) (alnLen + 1u))), __dmul_rn(0.5, priorAln));
    const double likNoMatch = __dsub_rn(1.0, pMatch);
    const double oddsRatio = __ddiv_rn(likNoMatch, matchLik);
    const double odds = __ddiv_rn(__dsub_rn(1.0, pMatch), pMatch);
    return __ddiv_rn(1.0, __dadd_rn(1.0, __dmul_rn(oddsRatio, odds)));
}

// cons: materialised consensus (unsafe) or null (safe: N | query | N)
__global__ void __launch_bounds__(256) merge_score_kernel(
    const u8 *__restrict__ data, const SeqMeta *__restrict__ meta, const u64 *__restrict__ likOff, const u32 *__restrict__ lineQ, const u32 *__restrict__ lineT,
    u8 *__restrict__ flags, const i32 *__restrict__ oqS, const i32 *__restrict__ oqE, const i32 *__restrict__ odS, const i32 *__restrict__ odE,
    const float *__restrict__ seqId0, const float *__restrict__ ry0, const u8 *__restrict__ cons, const double *__restrict__ logPower, MergeModel Mo, u64 M,
    MergeParams P, float *__restrict__ seqId, float *__restrict__ ry, u32 *__restrict__ alnLenCons, float *__restrict__ deam) {
    const int lane = threadIdx.x & 31;
    const u64 gwarp = ((u64) blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = ((u64) gridDim.x * blockDim.x) >> 5;
    for (u64 l = gwarp; l < M; l += nwarps) {
        const u8 fl = flags[l];
        float sid = seqId0[l], r = ry0[l];
        u32 tot = 0; float dm = 0.f; bool queued = false;
        if (fl & MRG_CAND) {
            const u32 q = lineQ[l];
            const SeqMeta qm = meta[q], tm = meta[lineT[l]];
            const u32 qLen = qm.lenFlags & META_LEN, dbLen = tm.lenFlags & META_LEN;
            const i32 qs = oqS[l], qe = oqE[l], ds = odS[l], de = odE[l];
            const u32 alnLen = (u32) (max(qe - qs, de - ds) + 1);
            bool rightStart, leftStart;
            ext_ends(qs, qe, ds, de, qLen, dbLen, rightStart, leftStart);
            const bool rev = (fl & MRG_REV) != 0;
            u32 idCnt = 0, idRy = 0, nCT = 0, nGA = 0;
            if (leftStart || rightStart) {   // updateSeqIdConsensus :744-778 and the dimer counts of ancientMatchCount :1100-1166 share the walk
                const u32 pad = qLen - (dbLen - alnLen);
                const u32 base = leftStart ? pad : 3u * qLen - (dbLen + pad);   // the padded oriented target lies at consensus[base ..)
                const u8 *tb = data + tm.off, *qb = data + qm.off;
                const u8 *c = cons ? cons + 3ull * likOff[q] : nullptr;
                for (u32 i = lane; i < dbLen; i += 32) {
                    const u32 p = base + i;
                    const unsigned char cc = c ? c[p] : (p >= qLen && p < 2u * qLen ? qb[p - qLen] : (unsigned char) 'N');
                    const unsigned char tc = t_char(tb, dbLen, rev, i);
                    if (!(cc == 'N' || tc == 'N')) {
                        idCnt += cc == tc ? 1u : 0u; idRy += ry_map(cc) == ry_map(tc) ? 1u : 0u; tot++;
                        const int dimer = 4 * nuc_map(cc) + nuc_map(tc);
                        nCT += dimer == 7 ? 1u : 0u; nGA += dimer == 8 ? 1u : 0u;
                    }
                }
                idCnt = warp_sum(idCnt); idRy = warp_sum(idRy); tot = warp_sum(tot); nCT = warp_sum(nCT); nGA = warp_sum(nGA);
            }
            if (tot != 0) { sid = (float) (i32) idCnt / (float) tot; r = (float) (i32) idRy / (float) tot; }
            u32 minAlnLen = 500;   // ancientContigsResults.cpp:246-250
            if (alnLen < minAlnLen) minAlnLen = min(minAlnLen, (u32) __dmul_rn(0.2, (double) dbLen));
            if (sid >= P.mergeThr && r >= P.ryThr && alnLen >= minAlnLen) {
                queued = true;
                // ancientMatchCount :1066-1068 (unsigned wrap-around included), :1176
                const u32 mmCons = (u32) __dadd_rn((double) __fmul_rn(__fsub_rn(1.0f, sid), (float) tot), 0.5);
                const u32 mCons = tot - mmCons;
                const u32 scoreAln = mCons * 2u + mmCons * (u32) (-3);
                float mCT = 0.f, mGA = 0.f;
                const double likCT = Mo.matchLik[rev ? 1 : 0][0], likGA = Mo.matchLik[rev ? 1 : 0][1];
                if (likCT > 0 && nCT) {   // `mCT += deamMatches(..)`: a float accumulator fed with doubles, one rounding per dimer
                    const double post = deam_posterior(alnLen, scoreAln, likCT, logPower, Mo);
                    for (u32 i = 0; i < nCT; i++) mCT = __double2float_rn(__dadd_rn((double) mCT, post));
                }
                if (likGA > 0 && nGA) {
                    const double post = deam_posterior(alnLen, scoreAln, likGA, logPower, Mo);
                    for (u32 i = 0; i < nGA; i++) mGA = __double2float_rn(__dadd_rn((double) mGA, post));
                }
                dm = __fadd_rn(__fadd_rn(__fdiv_rn(__fadd_rn((float) scoreAln, __fmul_rn(3.0f, (float) tot)), 5.0f), mCT), mGA);
            }
        }
        if (lane == 0) { seqId[l] = sid; ry[l] = r; alnLenCons[l] = tot; deam[l] = dm; if (queued) flags[l] = fl | MRG_QUEUED; }
    }
}

}   // namespace cdgmerge
