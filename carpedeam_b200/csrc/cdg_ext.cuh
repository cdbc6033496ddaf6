// Scoring front half of ancient_read_assemble on the device: everything between reading a query's alignment
// lines and pushing scored extension candidates into the priority queue (the greedy extension itself stays on the host).
//   ext_lines_kernel      orientation, identity / RY identity against the query, candidate filter
//                         (src/assembler/ancientReadsResults.cpp:200-236, :247-293, :294-315)
//   ext_consensus_kernel  consensusCaller + calculateConsensus (src/assembler/nuclassembleUtil.cpp:570-702, :533-568)
//   ext_update_kernel     updateSeqIdConsensusReads (:377-470): identities against the consensus, longest left / right overlap
//   ext_select_kernel     the final filter and the overlap length r_s_pair passes on (ancientReadsResults.cpp:330-365, :48-70)
// followed by pair_loglik_kernel<1> (calcLikelihoodConsensus) over the selected lines.
// flags per line: bit 0 candidate (notContig), bit 1 isRevToAlignment, bit 2 scored.
#pragma once
#include "cdg_align.cuh"

namespace cdgext {
using namespace cdgalign;

struct ExtParams { float seqIdThr, ryThr; int unsafe; u32 minCovSafe; int roundSeqId; };
static constexpr u8 EXT_CAND = 1, EXT_REV = 2, EXT_SCORED = 4;

CDG_D void ext_ends(i32 qs, i32 qe, i32 ds, i32 de, u32 qLen, u32 dbLen, bool &right, bool &left) {
    right = (u32) ds == 0u && (u32) qe == qLen - 1u;
    left = (u32) qs == 0u && (u32) de == dbLen - 1u;
}

__global__ void __launch_bounds__(256) ext_lines_kernel(
    const u8 *__restrict__ data, const SeqMeta *__restrict__ meta, const u32 *__restrict__ keys /*null: key == index*/, const u64 *__restrict__ likOff,
    const u32 *__restrict__ lineQ, const u32 *__restrict__ lineT, const u8 *__restrict__ keep, const float *__restrict__ seqIdIn,
    const i32 *__restrict__ qS_, const i32 *__restrict__ qE_, const i32 *__restrict__ dS_, const i32 *__restrict__ dE_, u64 M, ExtParams P,
    u8 *__restrict__ flags, i32 *__restrict__ oqS, i32 *__restrict__ oqE, i32 *__restrict__ odS, i32 *__restrict__ odE, float *__restrict__ seqIdOut,
    float *__restrict__ ryOut, u64 *__restrict__ consOffLine) {
    const u64 l = (u64) blockIdx.x * blockDim.x + threadIdx.x;
    if (l >= M) return;
    const u32 q = lineQ[l], t = lineT[l];
    const SeqMeta qm = meta[q], tm = meta[t];
    const u32 qLen = qm.lenFlags & META_LEN, dbLen = tm.lenFlags & META_LEN;
    i32 qs = qS_[l], qe = qE_[l], ds = dS_[l], de = dE_[l];
    const u32 alnLen = (u32) (max(abs(qe - qs), abs(de - ds)) + 1);   // Matcher::computeAlnLength
    float seqId = P.roundSeqId ? seqid_text_roundtrip(seqIdIn[l]) : seqIdIn[l], ry = 0.f;
    u8 fl = 0;
    if (keep[l]) {
        bool right, left;
        ext_ends(qs, qe, ds, de, qLen, dbLen, right, left);
        bool rev = false;
        if (right || left) {   // :200-236
            if (qs > qe) { const i32 x = qs; qs = qe; qe = x; const i32 d0 = ds; ds = (i32) dbLen - de - 1; de = (i32) dbLen - d0 - 1; rev = true; }
        }
        ext_ends(qs, qe, ds, de, qLen, dbLen, right, left);
        const u32 qKey = keys ? keys[q] : q;
        // :247-293 (the target INDEX is compared with the query KEY there); coordinates outside the sequences never load
        if ((right || left) && t != qKey && qs >= 0 && qe >= qs && (u32) qe < qLen && ds >= 0 && (u32) (ds + (qe - qs)) < dbLen) {
            const u8 *qb = data + qm.off, *tb = data + tm.off;
            u32 idCnt = 0, idRy = 0;
            for (i32 i = qs; i <= qe; i++) {
                const unsigned char qc = qb[i], tc = t_char(tb, dbLen, rev, (u32) (ds + (i - qs)));
                idCnt += qc == tc ? 1u : 0u; idRy += ry_map(qc) == ry_map(tc) ? 1u : 0u;
            }
            seqId = (float) (i32) idCnt / (float) alnLen; ry = (float) (i32) idRy / (float) alnLen;
        }
        // :294-315; an overhang longer than the query would leave the 3*qLen consensus (undefined in the reference): no candidate
        const bool noOffset = dbLen - alnLen == 0u;
        if ((right || left) && !(tm.lenFlags & META_EXT) && alnLen >= 30u && seqId >= P.seqIdThr && !noOffset && dbLen >= alnLen && dbLen - alnLen <= qLen && alnLen <= qLen)
            fl |= EXT_CAND;
        if (rev) fl |= EXT_REV;
    }
    flags[l] = fl; oqS[l] = qs; oqE[l] = qe; odS[l] = ds; odE[l] = de; seqIdOut[l] = seqId; ryOut[l] = ry; consOffLine[l] = 3ull * likOff[q];
}

// one warp per query; consensus of query q = 3*qLen letters at cons + 3*likOff[q]
__global__ void __launch_bounds__(256) ext_consensus_kernel(
    const u8 *__restrict__ data, const SeqMeta *__restrict__ meta, const u32 *__restrict__ keys, const u64 *__restrict__ likOff, const u64 *__restrict__ lineOff, u64 n,
    const u32 *__restrict__ lineT, const u8 *__restrict__ flags, const i32 *__restrict__ oqS, const i32 *__restrict__ oqE, const i32 *__restrict__ odS,
    const i32 *__restrict__ odE, ExtParams P, u8 *__restrict__ cons) {
    const int lane = threadIdx.x & 31;
    const u64 gwarp = ((u64) blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = ((u64) gridDim.x * blockDim.x) >> 5;
    for (u64 q = gwarp; q < n; q += nwarps) {
        const SeqMeta qm = meta[q];
        const u32 qLen = qm.lenFlags & META_LEN;
        const u8 *qb = data + qm.off;
        u8 *c = cons + 3ull * likOff[q];
        const u64 l0 = lineOff[q], l1 = lineOff[q + 1];
        const u32 qKey = keys ? keys[q] : (u32) q;
        for (u32 p = lane; p < 3u * qLen; p += 32) {
            unsigned char out = 'N';
            if (P.unsafe && (p < qLen || p >= 2u * qLen)) {
                u32 cov[4] = {0, 0, 0, 0};
                for (u64 l = l0; l < l1; l++) {
                    if (!(flags[l] & EXT_CAND)) continue;
                    const u32 t = lineT[l];
                    const SeqMeta tm = meta[t];
                    const u32 dbLen = tm.lenFlags & META_LEN;
                    const i32 qs = oqS[l], qe = oqE[l], ds = odS[l], de = odE[l];
                    const bool rightStart = ds == 0 && de != (i32) dbLen - 1, leftStart = qs == 0 && qe != (i32) qLen - 1;   // :616-617
                    if (!((rightStart || leftStart) && (keys ? keys[t] : t) != qKey)) continue;
                    const u32 alnLen = (u32) (qe - qs + 1);
                    u32 start;
                    if ((u32) ds == 0u && (u32) qe == qLen - 1u) start = qLen + (u32) qs;                 // right extension (:651-657)
                    else if ((u32) qs == 0u && (u32) de == dbLen - 1u) start = qLen - (dbLen - alnLen);   // left extension (:658-664)
                    else continue;
                    if (p >= start && p - start < dbLen) cov[nuc_map(t_char(data + tm.off, dbLen, (flags[l] & EXT_REV) != 0, p - start))]++;
                }
                const u32 total = cov[0] + cov[1] + cov[2] + cov[3];
                if (total >= P.minCovSafe) {   // calculateConsensus (:533-568)
                    u32 mx = 0; int nMax = 0; unsigned char nuc = 'N';
                    for (int j = 0; j < 4; j++) {
                        if (cov[j] > mx) { mx = cov[j]; nuc = (unsigned char) "ACGT"[j]; nMax = 1; }
                        else if (cov[j] == mx && mx > 0) nMax++;
                    }
                    out = nMax > 1 ? (unsigned char) 'N' : nuc;
                }
            }
            if (p >= qLen && p < 2u * qLen) out = qb[p - qLen];   // the (corrected) query is trusted over the consensus (:669-673, :591-597)
            c[p] = out;
        }
    }
}

// one thread per candidate line: identities of the padded oriented target against the consensus
__global__ void __launch_bounds__(256) ext_update_kernel(
    const u8 *__restrict__ data, const SeqMeta *__restrict__ meta, const u32 *__restrict__ lineQ, const u32 *__restrict__ lineT, const u8 *__restrict__ flags,
    const i32 *__restrict__ oqS, const i32 *__restrict__ oqE, const i32 *__restrict__ odS, const i32 *__restrict__ odE, const u64 *__restrict__ consOffLine,
    const u8 *__restrict__ cons, u64 M, float *__restrict__ seqId, float *__restrict__ ry, u32 *__restrict__ maxLeft, u32 *__restrict__ maxRight) {
    const u64 l = (u64) blockIdx.x * blockDim.x + threadIdx.x;
    if (l >= M || !(flags[l] & EXT_CAND)) return;
    const u32 q = lineQ[l];
    const SeqMeta tm = meta[lineT[l]];
    const u32 qLen = meta[q].lenFlags & META_LEN, dbLen = tm.lenFlags & META_LEN;
    const i32 qs = oqS[l], qe = oqE[l], ds = odS[l], de = odE[l];
    const u32 alnLen = (u32) (qe - qs + 1);
    bool rightStart, leftStart;
    ext_ends(qs, qe, ds, de, qLen, dbLen, rightStart, leftStart);
    const u32 pad = qLen - (dbLen - alnLen);
    const u8 *c = cons + consOffLine[l];
    const u8 *tb = data + tm.off;
    const bool rev = (flags[l] & EXT_REV) != 0;
    u32 idCnt = 0, idRy = 0, tot = 0;
    if (leftStart || rightStart) {
        const u32 base = leftStart ? pad : 3u * qLen - (dbLen + pad);
        for (u32 i = 0; i < dbLen; i++) {
            const unsigned char cc = c[base + i], tc = t_char(tb, dbLen, rev, i);
            if (!(cc == 'N' || tc == 'N')) { idCnt += cc == tc ? 1u : 0u; idRy += ry_map(cc) == ry_map(tc) ? 1u : 0u; tot++; }
        }
    }
    if (tot != 0) { seqId[l] = (float) (i32) idCnt / (float) tot; ry[l] = (float) (i32) idRy / (float) tot; }
    if (leftStart) { if (tot) atomicMax(&maxLeft[q], tot); }
    else if (rightStart) { if (tot) atomicMax(&maxRight[q], tot); }
}

__global__ void __launch_bounds__(256) ext_select_kernel(
    const SeqMeta *__restrict__ meta, const u32 *__restrict__ keys, const u32 *__restrict__ lineQ, const u32 *__restrict__ lineT, u8 *__restrict__ flags,
    const i32 *__restrict__ oqS, const i32 *__restrict__ odS, const i32 *__restrict__ odE, const float *__restrict__ seqId, const float *__restrict__ ry,
    const u32 *__restrict__ maxLeft, const u32 *__restrict__ maxRight, u64 M, ExtParams P, u32 *__restrict__ maxAlnLine, u8 *__restrict__ revLine) {
    const u64 l = (u64) blockIdx.x * blockDim.x + threadIdx.x;
    if (l >= M) return;
    const u8 fl = flags[l];
    revLine[l] = (fl & EXT_REV) ? 1 : 0;
    maxAlnLine[l] = 0;
    if (!(fl & EXT_CAND)) return;
    const u32 q = lineQ[l], t = lineT[l];
    const u32 dbLen = meta[t].lenFlags & META_LEN;
    const i32 qs = oqS[l], ds = odS[l], de = odE[l];
    // notInside holds for every candidate (noOffset excluded them); (rightStart || leftStart) of :351-352 are the plain zero tests
    if ((ds == 0 || qs == 0) && (keys ? keys[t] != keys[q] : t != q) && ry[l] >= P.ryThr && seqId[l] >= P.seqIdThr) {
        flags[l] = fl | EXT_SCORED;
        maxAlnLine[l] = ((u32) qs == 0u && (u32) de == dbLen - 1u) ? maxLeft[q] : maxRight[q];   // r_s_pair (:56-60)
    }
}

}   // namespace cdgext
