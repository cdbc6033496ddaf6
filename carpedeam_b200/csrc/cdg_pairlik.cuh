// Damage-aware pair likelihood: calcLikelihood (src/assembler/nuclassembleUtil.cpp:95-200) and
// calcLikelihoodConsensus (:203-374) as called by r_s_pair (src/assembler/ancientReadsResults.cpp:48-70).
//
// Per aligned column the reference evaluates  lik = sum_x seqErr.p[x][tBase] * max(1e-3, D_or_Drev[class].p[qBase][x])
// — a function of (strand, class, qBase, tBase) only, so the host tabulates log(lik) once (352 doubles, shared memory)
// and a pair is a sum of table entries over its columns.  One warp per pair: lanes stride over the columns reading the
// raw DB bytes (coalesced), class from the column's position in the ORIENTED target, fixed-order fp64 warp reduction
// (the reference adds the same doubles sequentially in x87 long double: agreement ~1e-15 relative).
#pragma once
#include "cdg_align.cuh"

namespace cdgpair {
using namespace cdgalign;

struct PairModel { double logLik[2][11][4][4]; float logExcess; float logRand; };   // [rev][class][qBase][tBase]

CDG_D int pos_class(u32 pos, u32 dbLen) {   // subdeam_lookup (:126-136): the 3' loop overwrites the earlier ones
    if ((i32) pos >= (i32) dbLen - 5) return 6 + (int) (pos - (dbLen - 5));
    return pos < 5 ? (int) pos : 5;
}
CDG_D double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
// likMod += excess * log(excessPenal); ratio = 1/(1+exp(maxAln*log(randAlnPenal) - likMod)): both products are FLOAT
// products in the reference (unsigned * logf(float)), the unsigned `excess` wraps when counted > maxAln (:178-181, :331-337)
CDG_D void finish_pair(double sum, u32 maxAln, u32 counted, const PairModel &M, double *sLen, double *sRatio) {
    const u32 excess = maxAln - counted;
    const double likMod = sum + (double) __fmul_rn((float) excess, M.logExcess);
    const double randAln = (double) __fmul_rn((float) maxAln, M.logRand);
    *sLen = likMod;
    *sRatio = 1.0 / (1.0 + exp(randAln - likMod));
}

static constexpr int PL_WARPS = 8;

// mode 0: query = resident sequence qIdx[i]; columns [0, alnLen) at query qStart+c / oriented target dbStart+c.
// mode 1: query = consensus string cons + consOff[i] (3*qLen bytes); every oriented target position j is compared with
//         consensus[base + j]; columns where either is 'N' are skipped and the class is the rank of j among the
//         target's non-'N' bases (tIdx-1, :255-262).
template <int MODE>
__global__ void __launch_bounds__(PL_WARPS * 32) pair_loglik_kernel(
    const u8 *__restrict__ data, const SeqMeta *__restrict__ meta, u64 nPairs,
    const u32 *__restrict__ qIdx, const u32 *__restrict__ tIdx, const u8 *__restrict__ revFlag,
    const i32 *__restrict__ qS_, const i32 *__restrict__ qE_, const i32 *__restrict__ dS_, const i32 *__restrict__ dE_,
    const u32 *__restrict__ maxAln_, const u8 *__restrict__ cons, const u64 *__restrict__ consOff,
    const PairModel *__restrict__ modelG, double *__restrict__ sLen, double *__restrict__ sRatio,
    const u8 *__restrict__ selFlags = nullptr /*not null: only pairs with (selFlags[i] & selMask) are evaluated, the others get 0*/, u8 selMask = 0) {
    __shared__ PairModel M;
    {
        const u32 *src = (const u32 *) modelG; u32 *dst = (u32 *) &M;
        for (int i = threadIdx.x; i < (int) (sizeof(PairModel) / 4); i += blockDim.x) dst[i] = src[i];
    }
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (u64 i = (u64) blockIdx.x * PL_WARPS + warp; i < nPairs; i += (u64) gridDim.x * PL_WARPS) {
        if (selFlags && !(selFlags[i] & selMask)) { if (lane == 0) { sLen[i] = 0.0; sRatio[i] = 0.0; } continue; }
        const SeqMeta tm = meta[tIdx[i]];
        const u32 dbLen = tm.lenFlags & META_LEN;
        const u8 *t = data + tm.off;
        const bool rev = revFlag[i] != 0;
        const double (*tab)[4][4] = M.logLik[rev ? 1 : 0];
        const i32 qS = qS_[i], qE = qE_[i], dS = dS_[i], dE = dE_[i];
        const u32 alnLen = (u32) (qE - qS + 1);
        double acc = 0.0;
        u32 counted = 0;
        if (MODE == 0) {
            const u8 *q = data + meta[qIdx[i]].off;
            for (u32 c = lane; c < alnLen; c += 32) {
                const u32 tp = (u32) dS + c;
                acc += tab[pos_class(tp, dbLen)][nuc_map(q[(u32) qS + c])][nuc_map(t_char(t, dbLen, rev, tp))];
            }
            counted = alnLen;
        } else {
            const u32 qLen = meta[qIdx[i]].lenFlags & META_LEN;
            const bool rightStart = dS == 0 && (u32) qE == qLen - 1;
            const bool leftStart = qS == 0 && (u32) dE == dbLen - 1;
            if (leftStart || rightStart) {
                const u32 pad = qLen - (dbLen - alnLen);                       // consensusStart (:246, :289)
                const u32 base = leftStart ? pad : 3 * qLen - (dbLen + pad);
                const u8 *cq = cons + consOff[i];
                u32 seen = 0;                                                  // non-'N' target bases before this chunk
                for (u32 j0 = 0; j0 < dbLen; j0 += 32) {
                    const u32 j = j0 + lane;
                    unsigned char tc = 'N', cc = 'N';
                    if (j < dbLen) { tc = t_char(t, dbLen, rev, j); cc = cq[base + j]; }
                    const u32 nonN = __ballot_sync(0xffffffffu, tc != 'N');
                    const bool use = tc != 'N' && cc != 'N';
                    if (use) {
                        const u32 rank = seen + __popc(nonN & ((2u << lane) - 1u)) - 1u;   // tIdx - 1
                        acc += tab[pos_class(rank, dbLen)][nuc_map(cc)][nuc_map(tc)];
                    }
                    counted += __popc(__ballot_sync(0xffffffffu, use));
                    seen += __popc(nonN);
                }
            }
        }
        acc = warp_sum(acc);
        if (lane == 0) finish_pair(acc, maxAln_[i], counted, M, &sLen[i], &sRatio[i]);
    }
}

}   // namespace cdgpair
