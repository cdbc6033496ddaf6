// kmermatcher on the device (reference: lib/mmseqs/src/linclust/kmermatcher.cpp).
//   pack_kernel          ASCII -> 2-bit words + X mask            (Sequence::mapSequence, Sequence.cpp:476-489)
//   extract_short_kernel per-sequence k-mer selection, warp/seq   (fillKmerPositionArray :77-388)
//   extract_long_kernel  same for sequences where the hash threshold is active, block/seq
//   group_kernel         rep selection + diagonal per member      (sort :412 + assignGroup :453-562)
//   reduce kernels       best diagonal per (rep,target) + bleed   (sort :431 + writeKmerMatcherResult :815-930)
#pragma once
#include "cdg_common.cuh"
#include "cdg_sort.cuh"

namespace cdgkmer {

// ---------------------------------------------------------------------------------------------------
// Packed layout: sequence i owns words [wordOff[i], wordOff[i+1]) of `words` (u64, 32 bases per word,
// base j of a word in bits [62-2j, 63-2j], i.e. first base most significant, like the k-mer index of
// Indexer::computeKmerIdx, Indexer.h:136-143) and the same range of `xmask` (u32, bit 31-j set when base
// j maps to X).  flags[i] bit0: some byte is not one of upper-case A,C,G,T ("exception" sequence).
// One warp per sequence; per 32-base word the lanes read 32 consecutive bytes (one coalesced load) and the word is
// assembled from warp ballots of the two code bits.
CDG_D u64 interleave32(u32 hi, u32 lo) {   // bit j of hi -> bit 2j+1, bit j of lo -> bit 2j
    u64 a = hi, b = lo;
    a = (a | (a << 16)) & 0x0000FFFF0000FFFFull; b = (b | (b << 16)) & 0x0000FFFF0000FFFFull;
    a = (a | (a << 8)) & 0x00FF00FF00FF00FFull;  b = (b | (b << 8)) & 0x00FF00FF00FF00FFull;
    a = (a | (a << 4)) & 0x0F0F0F0F0F0F0F0Full;  b = (b | (b << 4)) & 0x0F0F0F0F0F0F0F0Full;
    a = (a | (a << 2)) & 0x3333333333333333ull;  b = (b | (b << 2)) & 0x3333333333333333ull;
    a = (a | (a << 1)) & 0x5555555555555555ull;  b = (b | (b << 1)) & 0x5555555555555555ull;
    return (a << 1) | b;
}
__global__ void __launch_bounds__(256) pack_kernel(const u8 *__restrict__ data, const u64 *__restrict__ off, const u32 *__restrict__ len,
                                                   const u64 *__restrict__ wordOff, u64 seqLo, u64 seqHi, u64 *__restrict__ words,
                                                   u32 *__restrict__ xmask, u8 *__restrict__ flags) {
    const int lane = threadIdx.x & 31;
    const u64 gwarp = ((u64) blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = ((u64) gridDim.x * blockDim.x) >> 5;
    for (u64 seq = seqLo + gwarp; seq < seqHi; seq += nwarps) {
        const u8 *s = data + off[seq];
        const u32 L = len[seq];
        const u64 w0 = wordOff[seq];
        const u32 nw = (L + 31) >> 5;
        u32 exc = 0;
        for (u32 w = 0; w < nw; w++) {
            const u32 p = (w << 5) + lane;
            int c = 0; u32 e = 0;
            if (p < L) { const unsigned char ch = s[p]; c = cdg_map_letter(ch); e = cdg_plain_code(ch) == 4; }
            // lane j holds base j of the word; base 0 must land in the most significant bits: reverse the ballots
            const u32 b1 = __brev(__ballot_sync(0xffffffffu, (c & 2) != 0));
            const u32 b0 = __brev(__ballot_sync(0xffffffffu, (c & 1) != 0));
            const u32 xm = __brev(__ballot_sync(0xffffffffu, c == 4));
            exc |= __ballot_sync(0xffffffffu, e);
            if (lane == 0) { words[w0 + w] = interleave32(b1, b0); xmask[w0 + w] = xm; }   // X is stored as code 0 (4 & 3)
        }
        if (lane == 0) flags[seq] = (u8) (exc ? 1 : 0);
    }
}

// 2k-bit window starting at base p of a packed sequence (caller guarantees p + k <= L)
CDG_D u64 window_bits(const u64 *__restrict__ w, u32 p, int k) {
    u32 wi = p >> 5, sh = (p & 31) * 2;
    u64 hi = w[wi];
    u64 v = hi << sh;
    if (sh && ((p & 31) + k > 32)) v |= w[wi + 1] >> (64 - sh);
    return v >> (64 - 2 * k);
}
CDG_D u32 window_xmask(const u32 *__restrict__ xm, u32 p, int k) {   // nonzero iff the window contains X
    u32 wi = p >> 5, sh = p & 31;
    u64 two = ((u64) xm[wi] << 32);
    if (sh + k > 32) two |= xm[wi + 1];
    u64 v = two << sh;
    return (u32) (v >> (64 - k)) ;
}
CDG_D int base_code(const u64 *__restrict__ w, const u32 *__restrict__ xm, u32 p) {   // A0 C1 T2 G3 X4
    if ((xm[p >> 5] >> (31 - (p & 31))) & 1u) return 4;
    return (int) ((w[p >> 5] >> (62 - 2 * (p & 31))) & 3ull);
}

struct ExtractParams {
    int k; int kmersPerSeq; float scale; u64 seed; u32 hashLo, hashHi;
    // multi-GPU slice: this call extracts the sequences whose (length desc, id asc) rank lies in [posLo, posHi), tags
    // every record with the owner of its hash (hash16 * world >> 16) and writes slots relative to slotBase.
    // identity: 1 = single GPU (world = 1, posLo = 0, posHi = n, slotBase = 0): records of the owned hash range at slot
    // rankOf[seq]; 2 = slice mode: one identity record per sequence of the slice at slot rank - posLo, tagged with its owner
    // (routed with the k-mer records); 0 = none.
    u32 world, posLo, posHi; u64 slotBase; int identity;
    u64 seqLo, seqHi;   // index-order iteration (single GPU): sequences [seqLo, seqHi) of this launch (chunked upload pipeline)
    u32 sentinelOwner = 0;   // owner tag of unused slots (streamed multi-GPU exchange: a discard destination instead of rank 0)
};
CDG_D u32 owner_of(u32 score, u32 world) { return (score * world) >> 16; }
// slots reserved for sequence of length L: upper bound of selected k-mers (kmermatcher.cpp:226)
CDG_HD u32 kmer_capacity(u32 L, int k, int kmersPerSeq, float scale) {
    u32 W = L >= (u32) k ? L - k + 1 : 0;
    float consF = (float) (kmersPerSeq - 1) + scale * (float) L;
    u64 cons = consF > 0.f ? (u64) consF : 0;
    return (u32) (cons < W ? cons : W);
}
CDG_HD bool is_short_seq(u32 L, int k, int kmersPerSeq, float scale) {   // no hash threshold, <= 256 windows
    u32 W = L >= (u32) k ? L - k + 1 : 0;
    return W <= 256 && kmer_capacity(L, k, kmersPerSeq, scale) >= W;
}

// powers of 31 (mod 2^64) for the identity hash Util::hash (Util.h:337-346): h = sum code_i * 31^(L-1-i)
__global__ void pow31_table_kernel(u64 *__restrict__ tab, u32 n) {   // tab[e] = 31^e, e < n (n is small: sequential)
    if (blockIdx.x == 0 && threadIdx.x == 0) { u64 r = 1; for (u32 e = 0; e < n; e++) { tab[e] = r; r *= 31; } }
}
static constexpr u32 POW31_TABLE = 1024;   // exponents below this come from the table (short sequences always do)
CDG_D u64 pow31(u32 e) {
    u64 r = 1, b = 31;
    while (e) { if (e & 1) r *= b; b *= b; e >>= 1; }
    return r;
}

struct WinRec { u64 kmer; u32 pos; u32 score; };   // kmer carries the strand bit 63

// reverse complement on the device: two BREV reverse the 64 bits, one swap of the bits inside every 2-bit group restores the codes
CDG_D u64 revcomp_kmer_dev(u64 idx, int k) {
    const u64 x = idx ^ 0xAAAAAAAAAAAAAAAAull;
    u32 lo = __brev((u32) (x >> 32)), hi = __brev((u32) x);
    lo = ((lo >> 1) & 0x55555555u) | ((lo & 0x55555555u) << 1);
    hi = ((hi >> 1) & 0x55555555u) | ((hi & 0x55555555u) << 1);
    return (((u64) hi << 32) | lo) >> (64 - 2 * k);
}
// canonical k-mer of window p; returns false for windows the reference skips (X, palindromes).
// SCORE = false: the 16-bit XXH64 score is not needed by the caller's fast path (no hash threshold, the whole hash range owned,
// one rank): r.score then holds a cheap 13-bit mixing of the k-mer (repeat-hint bitmap index only)
template <bool SCORE = true>
CDG_D bool make_window(const u64 *__restrict__ w, const u32 *__restrict__ xm, u32 p, u32 L, const ExtractParams &P, WinRec &r) {
    if (window_xmask(xm, p, P.k)) return false;                     // kmermatcher.cpp:146
    u64 idx = window_bits(w, p, P.k);
    u64 rc = revcomp_kmer_dev(idx, P.k);
    if (rc == idx) return false;                                    // :159
    bool useRc = rc < idx;
    u64 canon = useRc ? rc : idx;
    if (SCORE) r.score = (u32) (cdg_xxh64_u64(canon, P.seed) & 0xFFFFu);       // :164
    else r.score = (((u32) canon ^ (u32) (canon >> 21)) * 0x9E3779B1u) >> 19;
    r.kmer = useRc ? canon : (canon | CDG_B63);                     // :183
    r.pos = useRc ? (L - p - P.k) : p;                              // :185
    return true;
}

// order of SequencePosition::compareByScoreReverse (kmermatcher.h:29-46); full ties: forward first
CDG_D bool win_less(const WinRec &a, const WinRec &b) {
    if (a.score != b.score) return a.score < b.score;
    u64 ka = a.kmer | CDG_B63, kb = b.kmer | CDG_B63;
    if (ka != kb) return ka < kb;
    if (a.pos != b.pos) return a.pos < b.pos;
    return (a.kmer & CDG_B63) > (b.kmer & CDG_B63);
}

// The reference's sequential selection loop (kmermatcher.cpp:275-347) over records sorted by win_less.
// keep[i] = 1 for selected records.  Run by ONE thread; only for sequences with repeated k-mers or an
// active hash threshold.
CDG_D void select_sequential(const WinRec *sorted, u32 cnt, u32 considered, u32 threshold, int tooMuch, u8 *keep) {
    u32 selected = 0;
    for (u32 i = 0; i < cnt && selected < considered; i++) {
        u64 kmer = sorted[i].kmer | CDG_B63;
        if (i + 1 < cnt) {
            u64 next = sorted[i + 1].kmer | CDG_B63;
            if (kmer == next) {
                while (kmer == next && i < cnt) {
                    i++;
                    if (i >= cnt) break;
                    next = sorted[i].kmer | CDG_B63;
                }
            }
        }
        if (i >= cnt) break;
        if (sorted[i].score < threshold) {
            if (sorted[i].score == threshold - 1 && tooMuch) {
                tooMuch--;
                threshold -= (tooMuch == 0) ? 1 : 0;
            }
            selected++;
            keep[i] = 1;
        }
    }
}

// identity record (kmermatcher.cpp:133-137,244-262): key = XXH64(Util::hash(codes)), pos 0.
// Goes to idKeys/idVals[seq] unless its value (bit 63 masked) lies inside the k-mer key range, in
// which case it must be sorted among the k-mers: those go to the `extra` slots of the main array.
CDG_D void emit_identity(u64 seqHash, u64 seq, const ExtractParams &P, u64 *idKeys, u64 *idVals, u64 *mainKeys, u64 *mainVals,
                         u64 extraBase, u32 extraCap, u32 *extraCount, u32 *overflow, u32 myRank /*(length desc, id asc) rank of seq*/) {
    u32 h16 = (u32) (seqHash & 0xFFFFu);
    const bool slice = P.identity == 2;
    bool owned = slice || (h16 >= P.hashLo && h16 <= P.hashHi);
    u64 key = CDG_SENTINEL, val = 0;
    if (owned) { key = seqHash; val = cdg_pack_val((u32) seq, 0, slice ? owner_of(h16, P.world) : 0u); }
    if (owned && (seqHash & ~CDG_B63) < (1ull << (2 * P.k))) {
        u32 slot = atomicAdd(extraCount, 1u);
        if (slot < extraCap) { mainKeys[extraBase + slot] = key; mainVals[extraBase + slot] = val; }
        else atomicExch(overflow, 1u);
        key = CDG_SENTINEL; val = 0;
    }
    const u32 slot = myRank - (slice ? P.posLo : 0u);   // (length desc, id asc) order: after the stable sort the representative comes first
    idKeys[slot] = key; idVals[slot] = val;
}

static constexpr int EX_WARPS = 8;
static constexpr int EX_MAXW = 256;   // windows per short sequence
static constexpr int EX_BM_WORDS = 256;   // per-warp repeat-hint bitmap (8192 bits, indexed by the low 13 bits of the 16-bit hash)

// ---- short sequences: one warp per sequence ------------------------------------------------------------
// Fast path (no repeated k-mer inside the sequence): every valid window is emitted.  A per-warp bitmap over
// the 16-bit hash detects possible repeats (equal k-mers have equal hashes); on a hit the warp ranks its
// windows in shared... global scratch and lane 0 replays the reference's sequential loop.
template <int ITERS, bool SCORE = true>
__global__ void __launch_bounds__(EX_WARPS * 32, ITERS <= 2 ? 5 : 1) extract_short_kernel(
    const u64 *__restrict__ words, const u32 *__restrict__ xmask, const SeqMeta *__restrict__ meta,
    u64 n, ExtractParams P, const u64 *__restrict__ slotOff /*[n] first main slot of seq*/,
    u64 *__restrict__ mainKeys, u64 *__restrict__ mainVals, u64 *__restrict__ idKeys, u64 *__restrict__ idVals,
    u64 extraBase, u32 extraCap, u32 *extraCount, u32 *overflow, const u32 *__restrict__ rankOf, const u64 *__restrict__ pow31tab,
    WinRec *__restrict__ scratch /*[gridWarps][2*EX_MAXW]*/,
    u8 *__restrict__ scratchKeep /*[gridWarps][EX_MAXW]*/, unsigned long long *__restrict__ emitted,
    const u64 *__restrict__ order /*slice mode: sequence index by (length desc, id asc) rank, else null*/,
    const u64 *__restrict__ prefix /*slice mode: first slot by rank*/,
    const u32 *__restrict__ list = nullptr /*list mode: the iteration indices to process (extract_group_kernel's flagged sequences)*/,
    const u32 *__restrict__ listCount = nullptr) {
    __shared__ u32 bitmapAll[EX_WARPS * EX_BM_WORDS];   // 8192 bits per warp: small, so that registers bound the occupancy
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    u32 *bitmap = bitmapAll + warp * EX_BM_WORDS;
    for (int i = lane; i < EX_BM_WORDS; i += 32) bitmap[i] = 0;
    __syncwarp();
    const u64 gwarp = (u64) blockIdx.x * EX_WARPS + warp, nwarps = (u64) gridDim.x * EX_WARPS;
    WinRec *myScratch = scratch + gwarp * (2 * EX_MAXW);
    u8 *myKeep = scratchKeep + gwarp * EX_MAXW;
    unsigned long long localEmitted = 0;

    // Iteration space: all sequences in index order (single GPU), or the ranks [posLo, posHi) of this GPU's slice through
    // `order` (multi-GPU: the per-rank work does not grow with the number of ranks).  The loads that start a sequence
    // (index, meta, rank, slot) are issued one sequence ahead.
    const bool byRank = order != nullptr;
    const u64 itLo = list ? 0 : (byRank ? (u64) P.posLo : P.seqLo), itHi = list ? (u64) *listCount : (byRank ? (u64) P.posHi : (P.seqHi < n ? P.seqHi : n));
    SeqMeta smN; smN.off = 0; smN.wordOff = 0; smN.lenFlags = 0; u32 rankN = 0; u64 slotN = 0, seqN = 0;
    auto fetch = [&](u64 it) {
        if (list) it = list[it];
        if (byRank) { seqN = (u32) order[it]; rankN = (u32) it; slotN = prefix[it]; }
        else { seqN = it; rankN = rankOf[it]; slotN = slotOff[it]; }
        smN = meta[seqN];
    };
    if (itLo + gwarp < itHi) fetch(itLo + gwarp);
    for (u64 cur = itLo + gwarp; cur < itHi; cur += nwarps) {
        const SeqMeta sm = smN;   // one 16-byte load: length and first word
        const u32 myRank = rankN;
        const u64 slotAbs = slotN;
        const u64 seq = seqN;
        if (cur + nwarps < itHi) fetch(cur + nwarps);
        const u32 L = sm.lenFlags & META_LEN;
        if (!is_short_seq(L, P.k, P.kmersPerSeq, P.scale)) continue;   // handled by extract_long_kernel
        if (myRank < P.posLo || myRank >= P.posHi) continue;            // not in this rank's slice
        const u64 *w = words + sm.wordOff;
        const u32 *xm = xmask + sm.wordOff;
        const u32 W = L >= (u32) P.k ? L - P.k + 1 : 0;
        // identity hash
        u64 part = 0;
        {   // lane handles a contiguous chunk of bases (Horner), then weights it by 31^(bases after the chunk)
            const u32 c = (L + 31) >> 5;
            const u32 b0 = lane * c, b1 = min(b0 + c, L);
            if (b0 < L) {
                u64 h = 0;
                u32 curW = 0xFFFFFFFFu, xx = 0; u64 ww = 0;   // the chunk's words are loaded once, not per base
                for (u32 p = b0; p < b1; p++) {
                    const u32 wi = p >> 5, o = p & 31;
                    if (wi != curW) { ww = w[wi]; xx = xm[wi]; curW = wi; }
                    const u64 code = ((xx >> (31 - o)) & 1u) ? 4ull : ((ww >> (62 - 2 * o)) & 3ull);
                    h = h * 31 + code;
                }
                part = h * pow31tab[L - b1];   // L - b1 < 288 for short sequences
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
        if (lane == 0 && P.identity) emit_identity(cdg_xxh64_u64(part, P.seed), seq, P, idKeys, idVals, mainKeys, mainVals, extraBase, extraCap, extraCount, overflow, myRank);

        // windows: lane handles p = lane + 32*it
        WinRec rec[ITERS];
        u32 validMask = 0;   // bit it
        u32 flagged = 0;     // bit it: the window's bitmap bit was already set (possible repeat)
#pragma unroll
        for (int it = 0; it < ITERS; it++) {
            u32 p = lane + 32 * it;
            bool ok = false;
            if (p < W) ok = make_window<SCORE>(w, xm, p, L, P, rec[it]);
            if (ok) {
                validMask |= 1u << it;
                const u32 h = rec[it].score & (EX_BM_WORDS * 32 - 1);
                u32 bit = 1u << (h & 31);
                u32 old = atomicOr(&bitmap[h >> 5], bit);
                if (old & bit) flagged |= 1u << it;
            }
        }
        __syncwarp();
        // clear the bits again (cheaper than wiping the bitmap per sequence)
#pragma unroll
        for (int it = 0; it < ITERS; it++) if (validMask & (1u << it)) bitmap[(rec[it].score & (EX_BM_WORDS * 32 - 1)) >> 5] = 0;
        __syncwarp();
        // a set bit is only a hint (the bitmap index is 13 bits of the hash): confirm by comparing the flagged window's
        // k-mer with every other window of the sequence; only a true repeat takes the exact path
        u32 collide = 0;
        if (__any_sync(0xffffffffu, flagged)) {
#pragma unroll
            for (int it = 0; it < ITERS; it++) {
                u32 fl = __ballot_sync(0xffffffffu, (flagged >> it) & 1u);
                while (fl) {
                    const int src = __ffs(fl) - 1; fl &= fl - 1;
                    const u64 km = __shfl_sync(0xffffffffu, rec[it].kmer | CDG_B63, src);
                    bool eq = false;
#pragma unroll
                    for (int i2 = 0; i2 < ITERS; i2++)
                        eq |= ((validMask >> i2) & 1u) && (rec[i2].kmer | CDG_B63) == km && !(i2 == it && lane == src);
                    collide |= eq ? 1u : 0u;
                }
            }
            collide = __any_sync(0xffffffffu, collide);
        }

        const u64 slot0 = slotAbs - P.slotBase;
        const u32 cap = kmer_capacity(L, P.k, P.kmersPerSeq, P.scale);
        u32 written = 0;   // warp-uniform
        if (!collide) {
#pragma unroll
            for (int it = 0; it < ITERS; it++) {
                bool ok = (validMask >> it) & 1u;
                bool own = ok && (!SCORE || (rec[it].score >= P.hashLo && rec[it].score <= P.hashHi));
                u32 b = __ballot_sync(0xffffffffu, own);
                if (own) {
                    u64 dst = slot0 + written + __popc(b & ((1u << lane) - 1));
                    mainKeys[dst] = rec[it].kmer;
                    mainVals[dst] = cdg_pack_val((u32) seq, rec[it].pos, SCORE ? owner_of(rec[it].score, P.world) : 0u);
                }
                written += __popc(b);
            }
        } else {
            // exact path: materialise, rank (O(cnt^2/32)), sequential replay by lane 0
            u32 cnt = 0;
#pragma unroll
            for (int it = 0; it < ITERS; it++) {
                bool ok = (validMask >> it) & 1u;
                if (!SCORE && ok) rec[it].score = (u32) (cdg_xxh64_u64(rec[it].kmer & ~CDG_B63, P.seed) & 0xFFFFu);   // the replay orders by score
                u32 b = __ballot_sync(0xffffffffu, ok);
                if (ok) myScratch[cnt + __popc(b & ((1u << lane) - 1))] = rec[it];
                cnt += __popc(b);
            }
            __syncwarp();
            WinRec *sorted = myScratch + EX_MAXW;
            for (u32 i = lane; i < cnt; i += 32) {
                WinRec me = myScratch[i];
                u32 r = 0;
                for (u32 j = 0; j < cnt; j++) {
                    WinRec o = myScratch[j];
                    r += (win_less(o, me) || (!win_less(me, o) && j < i)) ? 1u : 0u;
                }
                sorted[r] = me;
                myKeep[i] = 0;
            }
            __syncwarp();
            if (lane == 0) select_sequential(sorted, cnt, cnt, 65536u, 0, myKeep);
            __syncwarp();
            for (u32 base = 0; base < cnt; base += 32) {
                u32 i = base + lane;
                bool own = i < cnt && myKeep[i] && sorted[i].score >= P.hashLo && sorted[i].score <= P.hashHi;
                u32 b = __ballot_sync(0xffffffffu, own);
                if (own) {
                    u64 dst = slot0 + written + __popc(b & ((1u << lane) - 1));
                    mainKeys[dst] = sorted[i].kmer;
                    mainVals[dst] = cdg_pack_val((u32) seq, sorted[i].pos, owner_of(sorted[i].score, P.world));
                }
                written += __popc(b);
            }
            __syncwarp();
        }
        for (u32 s = written + lane; s < cap; s += 32) { mainKeys[slot0 + s] = CDG_SENTINEL; mainVals[slot0 + s] = cdg_pack_val(0, 0, P.sentinelOwner); }
        localEmitted += written;
        // next sequence's packed words towards L1 while this one's records drain
        if (lane < 2 && cur + nwarps < itHi) {
            const void *pw = lane == 0 ? (const void *) (words + smN.wordOff) : (const void *) (xmask + smN.wordOff);
            asm volatile("prefetch.global.L1 [%0];" :: "l"(pw));
        }
    }
    if (lane == 0 && localEmitted) atomicAdd(emitted, localEmitted);
}

// ---- short sequences, GROUP lanes per sequence (32/GROUP sequences per warp) ------------------------------------------
// The fast path of extract_short_kernel for the common read lengths: a 35-70 bp read has 16-51 windows, so a full warp per read
// leaves half of the lanes idle and pays the per-sequence work (meta fetch, identity hash, slot bookkeeping) once per read; with
// 8 lanes per read four reads share every instruction.  Every valid window is emitted optimistically in window order; a blocked
// Bloom filter per group (128 words of 64 bits, a window sets two bits of one word with ONE atomic, so of two equal k-mers —
// even in the same instruction — the later one sees both bits set) flags a POSSIBLE repeated k-mer, and a flagged sequence is
// appended to `slowList` for extract_short_kernel's exact path, which rewrites all of its slots (false positives: < 1 % of reads).  Iteration index `cur` as in extract_short_kernel (sequence index, or the
// (length desc, id asc) rank in slice mode); slowList holds `cur`.
template <int GROUP, int ITERS, bool SCORE>
__global__ void __launch_bounds__(EX_WARPS * 32, GROUP == 8 ? 4 : 5) extract_group_kernel(
    const u64 *__restrict__ words, const u32 *__restrict__ xmask, const SeqMeta *__restrict__ meta,
    u64 n, ExtractParams P, const u64 *__restrict__ slotOff, u64 *__restrict__ mainKeys, u64 *__restrict__ mainVals,
    u64 *__restrict__ idKeys, u64 *__restrict__ idVals, u64 extraBase, u32 extraCap, u32 *extraCount, u32 *overflow,
    const u32 *__restrict__ rankOf, const u64 *__restrict__ pow31tab, unsigned long long *__restrict__ emitted,
    const u64 *__restrict__ order, const u64 *__restrict__ prefix, u32 *__restrict__ slowList, u32 *__restrict__ slowCount) {
    constexpr int GPW = 32 / GROUP;                   // groups (sequences) per warp
    constexpr u32 BM_WORDS = 128;                     // 64-bit words of the group's filter
    __shared__ unsigned long long bitmapAll[EX_WARPS * GPW * BM_WORDS];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane / GROUP, gl = lane % GROUP;
    unsigned long long *bm = bitmapAll + (warp * GPW + g) * BM_WORDS;
    for (int i = gl; i < (int) BM_WORDS; i += GROUP) bm[i] = 0;
    __syncwarp();
    const u32 gmask = (GROUP == 32 ? 0xffffffffu : ((1u << GROUP) - 1u)) << (g * GROUP);
    const u32 ltmask = gmask & ((1u << lane) - 1u);   // lanes of my group below me
    const u64 ggroup = ((u64) blockIdx.x * EX_WARPS + warp) * GPW + g, ngroups = (u64) gridDim.x * EX_WARPS * GPW;
    const bool byRank = order != nullptr;
    const u64 itLo = byRank ? (u64) P.posLo : P.seqLo, itHi = byRank ? (u64) P.posHi : (P.seqHi < n ? P.seqHi : n);
    unsigned long long localEmitted = 0;
    // all groups of a warp run the same number of rounds (the last ones with an inactive group)
    for (u64 base = itLo + (ggroup - g); base < itHi; base += ngroups) {
        const u64 cur = base + g;
        bool active = cur < itHi;
        u64 seq = 0, slotAbs = 0; u32 myRank = 0;
        SeqMeta sm; sm.off = 0; sm.wordOff = 0; sm.lenFlags = 0;
        if (active) {
            if (byRank) { seq = (u32) order[cur]; myRank = (u32) cur; slotAbs = prefix[cur]; }
            else { seq = cur; myRank = rankOf[cur]; slotAbs = slotOff[cur]; }
            sm = meta[seq];
        }
        const u32 L = sm.lenFlags & META_LEN;
        active = active && is_short_seq(L, P.k, P.kmersPerSeq, P.scale) && myRank >= P.posLo && myRank < P.posHi;
        const u64 *w = words + sm.wordOff;
        const u32 *xm = xmask + sm.wordOff;
        u32 W = (active && L >= (u32) P.k) ? L - P.k + 1 : 0;
        // a short sequence with more windows than GROUP * ITERS (the group size follows the typical read, not the longest one): its
        // identity record is written here, its windows by the exact kernel through the slow list
        const bool tooLong = W > (u32) (GROUP * ITERS);
        if (tooLong) W = 0;
        // identity hash (Util::hash, Util.h:337-346): sum of code_p * 31^(L-1-p), lane gl takes p = gl, gl+GROUP, ...
        u64 part = 0;
        if (active) {
            u32 curW = 0xFFFFFFFFu, xx = 0; u64 ww = 0;
            for (u32 p = gl; p < L; p += GROUP) {
                const u32 wi = p >> 5, o = p & 31;
                if (wi != curW) { ww = w[wi]; xx = xm[wi]; curW = wi; }
                const u64 code = ((xx >> (31 - o)) & 1u) ? 4ull : ((ww >> (62 - 2 * o)) & 3ull);
                part += code * pow31tab[L - 1 - p];   // L - 1 - p < 288 for short sequences
            }
        }
#pragma unroll
        for (int o = GROUP / 2; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
        if (active && gl == 0 && P.identity)
            emit_identity(cdg_xxh64_u64(part, P.seed), seq, P, idKeys, idVals, mainKeys, mainVals, extraBase, extraCap, extraCount, overflow, myRank);
        // windows
        const u64 slot0 = slotAbs - P.slotBase;
        u32 written = 0, flagged = 0;
        u32 hh[ITERS];   // the window's filter word, 0xFFFFFFFF = no valid window in this iteration
        const u32 maxW = __reduce_max_sync(0xffffffffu, W);   // over the groups of the warp: every lane skips the same iterations
#pragma unroll
        for (int it = 0; it < ITERS; it++) {   // unrolled: the loads and hashes of several iterations overlap
            hh[it] = 0xFFFFFFFFu;
            if ((u32) (it * GROUP) >= maxW) continue;     // warp-uniform
            const u32 p = gl + GROUP * it;
            WinRec r; bool ok = false;
            if (p < W) ok = make_window<SCORE>(w, xm, p, L, P, r);
            bool own = ok;
            if (SCORE) own = ok && r.score >= P.hashLo && r.score <= P.hashHi;
            if (ok) {
                const u64 canon = r.kmer & ~CDG_B63;
                const u32 mix = ((u32) canon ^ (u32) (canon >> 19)) * 0x9E3779B1u;
                const unsigned long long bits = (1ull << ((mix >> 19) & 63u)) | (1ull << ((mix >> 13) & 63u));
                hh[it] = mix >> 25;
                const unsigned long long old = atomicOr(&bm[hh[it]], bits);
                if ((old & bits) == bits) flagged = 1;
            }
            const u32 b = __ballot_sync(0xffffffffu, own) & gmask;
            if (own) {
                const u64 dst = slot0 + written + __popc(b & ltmask);
                mainKeys[dst] = r.kmer;
                mainVals[dst] = cdg_pack_val((u32) seq, r.pos, SCORE ? owner_of(r.score, P.world) : 0u);
            }
            written += __popc(b);
        }
        __syncwarp();
        // clear the touched filter words (cheaper than wiping 1 KB per sequence)
#pragma unroll
        for (int it = 0; it < ITERS; it++) if (hh[it] != 0xFFFFFFFFu) bm[hh[it]] = 0;
        const u32 flaggedLanes = __ballot_sync(0xffffffffu, flagged) & gmask;   // (taken by every lane: no short-circuit in front of a warp collective)
        const bool slow = tooLong || flaggedLanes != 0;
        if (active) {
            const u32 cap = kmer_capacity(L, P.k, P.kmersPerSeq, P.scale);
            if (!tooLong) for (u32 sI = written + gl; sI < cap; sI += GROUP) { mainKeys[slot0 + sI] = CDG_SENTINEL; mainVals[slot0 + sI] = cdg_pack_val(0, 0, P.sentinelOwner); }
            if (gl == 0) {
                if (slow) slowList[atomicAdd(slowCount, 1u)] = (u32) cur;   // the exact path rewrites the slots and counts the records
                else localEmitted += written;
            }
        }
        __syncwarp();
    }
    if (localEmitted) atomicAdd(emitted, localEmitted);
}

// ---- hash-range passes (kmermatcher.cpp:742-784 setupKmerSplits) ------------------------------------------------------------
// 65536-bin histogram of the 16-bit hashes of every valid window of the set: an upper bound of the records a hash range will hold
// (repeated k-mers and the per-sequence hash threshold only remove records), from which the host cuts contiguous ranges of about
// equal size.  One warp per sequence.
__global__ void __launch_bounds__(256) hash_hist_kernel(const u64 *__restrict__ words, const u32 *__restrict__ xmask, const SeqMeta *__restrict__ meta, u64 n,
                                                        ExtractParams P, unsigned long long *__restrict__ hist /*[65536]*/) {
    const int lane = threadIdx.x & 31;
    const u64 gwarp = ((u64) blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = ((u64) gridDim.x * blockDim.x) >> 5;
    for (u64 seq = gwarp; seq < n; seq += nwarps) {
        const SeqMeta sm = meta[seq];
        const u32 L = sm.lenFlags & META_LEN;
        if (L < (u32) P.k) continue;
        const u64 *w = words + sm.wordOff;
        const u32 *xm = xmask + sm.wordOff;
        for (u32 p = lane; p + P.k <= L; p += 32) {
            WinRec r;
            if (make_window<true>(w, xm, p, L, P, r)) atomicAdd(&hist[r.score], 1ull);
        }
    }
}
// records of a chunk buffer that are not sentinels, appended in order behind *base records of (outK, outV)
struct NotSentinelF { const u64 *keys; CDG_D u32 operator()(u64 i) const { return keys[i] != CDG_SENTINEL ? 1u : 0u; } };
__global__ void __launch_bounds__(cdgscan::BLOCK) compact_records_kernel(NotSentinelF f, const u64 *__restrict__ vals, u64 n, const unsigned long long *__restrict__ tileSums,
                                                                         const unsigned long long *__restrict__ base, u64 *__restrict__ outK, u64 *__restrict__ outV) {
    __shared__ u32 shWarp[cdgscan::BLOCK / 32 + 1];
    u64 prefix[cdgscan::ITEMS]; u32 val[cdgscan::ITEMS];
    cdgscan::tile_prefix(f, n, tileSums, shWarp, prefix, val);
    const u64 b = *base, tb = (u64) blockIdx.x * cdgscan::TILE;
#pragma unroll
    for (int i = 0; i < cdgscan::ITEMS; i++) {
        const u64 idx = tb + (u64) i * cdgscan::BLOCK + threadIdx.x;
        if (idx < n && val[i]) { outK[b + prefix[i]] = f.keys[idx]; outV[b + prefix[i]] = vals[idx]; }
    }
}
__global__ void add_u64_kernel(unsigned long long *dst, const unsigned long long *src) { *dst += *src; }
// length of the first run of equal k-mers of a sorted key array (strand bit ignored); sentinels never start a run
__global__ void __launch_bounds__(1024) first_run_kernel(const u64 *__restrict__ keys, u64 n, unsigned long long *__restrict__ out) {
    __shared__ unsigned long long firstOther;
    if (threadIdx.x == 0) firstOther = ~0ull;
    __syncthreads();
    if (n == 0 || keys[0] == CDG_SENTINEL) { if (threadIdx.x == 0) *out = 0; return; }
    const u64 k0 = keys[0] | CDG_B63;
    for (u64 b = 0; b < n; b += 1024) {
        const u64 i = b + threadIdx.x;
        if (i < n && (keys[i] == CDG_SENTINEL || (keys[i] | CDG_B63) != k0)) atomicMin(&firstOther, (unsigned long long) i);
        __syncthreads();
        if (firstOther != ~0ull) break;
    }
    if (threadIdx.x == 0) *out = firstOther == ~0ull ? n : firstOther;
}

// ---- long sequences (hash threshold active or > 256 windows): one block per sequence -------------------
// scoreDist/hier histograms (kmermatcher.cpp:128-129,224-241) in shared memory; candidates below the
// threshold are compacted to global scratch, bitonic-sorted by the block, replayed by thread 0.
static constexpr int EXL_BLOCK = 256;
__global__ void __launch_bounds__(EXL_BLOCK) extract_long_kernel(
    const u64 *__restrict__ words, const u32 *__restrict__ xmask, const u64 *__restrict__ wordOff, const u32 *__restrict__ len,
    const u32 *__restrict__ longList, u32 nLong, ExtractParams P, const u64 *__restrict__ slotOff,
    u64 *__restrict__ mainKeys, u64 *__restrict__ mainVals, u64 *__restrict__ idKeys, u64 *__restrict__ idVals,
    u64 extraBase, u32 extraCap, u32 *extraCount, u32 *overflow, const u32 *__restrict__ rankOf,
    WinRec *__restrict__ scratch, u8 *__restrict__ scratchKeep, const u64 *__restrict__ scratchOff /*[nLong] pow2-padded regions*/,
    unsigned long long *__restrict__ emitted) {
    extern __shared__ u32 shmem[];
    u16 *scoreDist = (u16 *) shmem;               // 65536 x u16 (the reference uses unsigned short too)
    u32 *hier = shmem + 32768;                    // 128
    __shared__ u32 sThreshold, sCnt, sCand, sWritten;
    __shared__ int sTooMuch;
    __shared__ u32 sConsidered;
    __shared__ unsigned long long sHash[EXL_BLOCK / 32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    for (u32 li = blockIdx.x; li < nLong; li += gridDim.x) {
        const u64 seq = longList[li];
        const u32 myRank = rankOf[seq];
        if (myRank < P.posLo || myRank >= P.posHi) continue;   // block-uniform
        const u32 L = len[seq];
        const u64 *w = words + wordOff[seq];
        const u32 *xm = xmask + wordOff[seq];
        const u32 W = L >= (u32) P.k ? L - P.k + 1 : 0;
        for (int i = tid; i < 32768 + 128; i += EXL_BLOCK) shmem[i] = 0;
        if (tid == 0) { sCnt = 0; sCand = 0; sWritten = 0; }
        __syncthreads();
        // identity hash
        u64 part = 0;
        for (u32 p = tid; p < L; p += EXL_BLOCK) part += (u64) base_code(w, xm, p) * pow31(L - 1 - p);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
        if (lane == 0) sHash[warp] = part;
        // pass 1: histograms
        u32 myCnt = 0;
        for (u32 p = tid; p < W; p += EXL_BLOCK) {
            WinRec r;
            if (make_window(w, xm, p, L, P, r)) {
                myCnt++;
                // 16-bit counters packed two per word
                atomicAdd(&shmem[r.score >> 1], (r.score & 1) ? 0x10000u : 1u);
                atomicAdd(&hier[r.score >> 9], 1u);
            }
        }
        atomicAdd(&sCnt, myCnt);
        __syncthreads();
        if (tid == 0) {
            u64 h = 0;
            for (int i = 0; i < EXL_BLOCK / 32; i++) h += sHash[i];
            if (P.identity) emit_identity(cdg_xxh64_u64(h, P.seed), seq, P, idKeys, idVals, mainKeys, mainVals, extraBase, extraCap, extraCount, overflow, myRank);
            u32 cnt = sCnt;
            float consF = (float) (P.kmersPerSeq - 1) + P.scale * (float) L;
            u64 cons64 = consF > 0.f ? (u64) consF : 0;
            u32 considered = (u32) (cons64 < cnt ? cons64 : cnt);
            u32 threshold = 0; u64 inBins = 0;
            if (cnt > 0) {                                            // kmermatcher.cpp:228-241
                u32 ht = 0;
                for (ht = 0; ht < 128 && inBins < considered; ht++) inBins += hier[ht];
                ht -= (ht > 0) ? 1 : 0;
                inBins -= hier[ht];
                for (threshold = ht * 512; threshold <= 65535 && inBins < considered; threshold++) inBins += scoreDist[threshold];
            }
            sThreshold = threshold; sTooMuch = (int) (inBins - considered); sConsidered = considered;
        }
        __syncthreads();
        const u32 threshold = sThreshold;
        WinRec *cand = scratch + scratchOff[li];
        u8 *keep = scratchKeep + scratchOff[li];
        // pass 2: compact candidates (score < threshold) — order irrelevant, sorted next
        for (u32 base = 0; base < W; base += EXL_BLOCK) {
            u32 p = base + tid;
            WinRec r; bool ok = false;
            if (p < W) ok = make_window(w, xm, p, L, P, r) && r.score < threshold;
            u32 b = __ballot_sync(0xffffffffu, ok);
            u32 wbase = 0;
            if (lane == 0 && b) wbase = atomicAdd(&sCand, __popc(b));
            wbase = __shfl_sync(0xffffffffu, wbase, 0);
            if (ok) cand[wbase + __popc(b & ((1u << lane) - 1))] = r;
        }
        __syncthreads();
        const u32 nc = sCand;
        u32 np2 = 1; while (np2 < nc) np2 <<= 1;
        for (u32 i = nc + tid; i < np2; i += EXL_BLOCK) { WinRec pad; pad.kmer = CDG_SENTINEL; pad.pos = 0xFFFFFFFFu; pad.score = 0x10000u; cand[i] = pad; }
        for (u32 i = tid; i < np2; i += EXL_BLOCK) keep[i] = 0;
        __syncthreads();
        for (u32 kk = 2; kk <= np2; kk <<= 1) {
            for (u32 j = kk >> 1; j > 0; j >>= 1) {
                for (u32 i = tid; i < np2; i += EXL_BLOCK) {
                    u32 ixj = i ^ j;
                    if (ixj > i) {
                        WinRec a = cand[i], b = cand[ixj];
                        bool up = (i & kk) == 0;
                        if (up ? win_less(b, a) : win_less(a, b)) { cand[i] = b; cand[ixj] = a; }
                    }
                }
                __syncthreads();
            }
        }
        if (tid == 0) select_sequential(cand, nc, sConsidered, threshold, sTooMuch, keep);
        __syncthreads();
        const u64 slot0 = slotOff[seq] - P.slotBase;
        const u32 cap = kmer_capacity(L, P.k, P.kmersPerSeq, P.scale);
        for (u32 base = 0; base < nc; base += EXL_BLOCK) {
            u32 i = base + tid;
            bool own = i < nc && keep[i] && cand[i].score >= P.hashLo && cand[i].score <= P.hashHi;
            u32 b = __ballot_sync(0xffffffffu, own);
            u32 wbase = 0;
            if (lane == 0 && b) wbase = atomicAdd(&sWritten, __popc(b));
            wbase = __shfl_sync(0xffffffffu, wbase, 0);
            if (own) {
                u64 dst = slot0 + wbase + __popc(b & ((1u << lane) - 1));
                mainKeys[dst] = cand[i].kmer;
                mainVals[dst] = cdg_pack_val((u32) seq, cand[i].pos, owner_of(cand[i].score, P.world));
            }
        }
        __syncthreads();
        const u32 written = sWritten;
        for (u32 s = written + tid; s < cap; s += EXL_BLOCK) { mainKeys[slot0 + s] = CDG_SENTINEL; mainVals[slot0 + s] = cdg_pack_val(0, 0, P.sentinelOwner); }
        if (tid == 0 && written) atomicAdd(emitted, (unsigned long long) written);
        __syncthreads();
    }
}

// routing of grouped records to the rank that owns the representative (rep index ranges of width repChunk):
// count, then unordered scatter (sort 2 follows on the receiver)
__global__ void __launch_bounds__(256) route_count_kernel(const u64 *__restrict__ key2, u64 n, int repShift, u32 repChunk, u32 world,
                                                          unsigned long long *__restrict__ counts) {
    __shared__ u32 sc[16];
    if (threadIdx.x < 16) sc[threadIdx.x] = 0;
    __syncthreads();
    for (u64 i = (u64) blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (u64) gridDim.x * blockDim.x) {
        u32 o = (u32) (key2[i] >> repShift) / repChunk; if (o >= world) o = world - 1;
        atomicAdd(&sc[o], 1u);
    }
    __syncthreads();
    if (threadIdx.x < world && sc[threadIdx.x]) atomicAdd(&counts[threadIdx.x], (unsigned long long) sc[threadIdx.x]);
}
// outTab == nullptr: one local output array, cursors start at the segment offsets.  outTab != nullptr: the send side of the
// grouped-record all-to-all — outTab[o] is this rank's segment of rank o's receive buffer (peer memory over NVLink, CUDA IPC),
// cursors start at 0; a block's records for one owner are appended as one burst.
__global__ void __launch_bounds__(256) route_scatter_kernel(const u64 *__restrict__ key2, u64 n, int repShift, u32 repChunk, u32 world,
                                                            unsigned long long *__restrict__ cursors /*[world]*/,
                                                            u64 *__restrict__ out, u64 *const *__restrict__ outTab = nullptr) {
    // block-aggregated append: per-owner counts of a 2048-record tile in shared memory, one global atomic per owner and block
    constexpr int ITEMS = 8;
    __shared__ u32 sCnt[16];
    __shared__ unsigned long long sBase[16];
    const u64 tiles = (n + 256 * ITEMS - 1) / (256 * ITEMS);
    for (u64 tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
        if (threadIdx.x < 16) sCnt[threadIdx.x] = 0;
        __syncthreads();
        u64 k[ITEMS]; u32 o[ITEMS], r[ITEMS];
#pragma unroll
        for (int j = 0; j < ITEMS; j++) {
            const u64 i = tile * (256 * ITEMS) + (u64) j * 256 + threadIdx.x;
            o[j] = 0xFFFFFFFFu; k[j] = 0; r[j] = 0;
            if (i < n) {
                k[j] = key2[i];
                o[j] = (u32) (k[j] >> repShift) / repChunk; if (o[j] >= world) o[j] = world - 1;
                // warp-aggregated shared atomic
                const u32 peers = __match_any_sync(__activemask(), o[j]);
                const int leader = __ffs(peers) - 1;
                u32 base = 0;
                if ((threadIdx.x & 31) == leader) base = atomicAdd(&sCnt[o[j]], (u32) __popc(peers));
                base = __shfl_sync(peers, base, leader);
                r[j] = base + __popc(peers & ((1u << (threadIdx.x & 31)) - 1));
            }
        }
        __syncthreads();
        if (threadIdx.x < world && sCnt[threadIdx.x]) sBase[threadIdx.x] = atomicAdd(&cursors[threadIdx.x], (unsigned long long) sCnt[threadIdx.x]);
        __syncthreads();
#pragma unroll
        for (int j = 0; j < ITEMS; j++) if (o[j] != 0xFFFFFFFFu) (outTab ? outTab[o[j]] : out)[sBase[o[j]] + r[j]] = k[j];
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------------------------------
// group assignment.  Input: records sorted by (key | bit63).  The thread of the first record of each run
// of equal k-mers finds the representative = (max seqLen, min id, min pos) — the first record of the run
// under compareRepSequenceAndIdAndPosReverse (kmermatcher.h:76-96) — and writes one compact record per
// member (the representative included) for sort 2:
//   key2 = rep << (idBits + diagBits + 1) | id << (diagBits + 1) | (diag + diagBias) << 1 | needRev
struct GroupParams {
    int idBits, diagBits; i32 diagBias; int onlyExtendable; int firstRunQuirk;   // :466 first run: repIsReverse=false
    float covThr; int covMode;   // Util::canBeCovered(covThr, covMode, repLen, memberLen), applied when !onlyExtendable (:526-529)
};
// Util::canBeCovered (lib/mmseqs/src/commons/Util.cpp:533-550)
CDG_D bool can_be_covered(float covThr, int covMode, float q, float t) {
    switch (covMode) {
        case 0: return (q / t >= covThr) && (t / q >= covThr);
        case 1: return (q / t) >= covThr;
        case 2: return (t / q) >= covThr;
        case 3: return ((t / q) >= covThr) && (t / q) <= 1.0f;
        case 4: return ((q / t) >= covThr) && (q / t) <= 1.0f;
        case 5: return (fminf(t, q) / fmaxf(t, q)) >= covThr;
        default: return true;
    }
}
// WIDE records (2*idBits + diagBits + 1 > 64, e.g. 5*10^7 sequences with contigs of 10 kb and more): two words per member,
//   lo = id << (diagBits + 1) | (diag + diagBias) << 1 | needRev,   hi = rep
// sorted as (lo bits, then hi bits) by two pair sorts; the (rep,id) pair = hi << idBits | lo >> (diagBits + 1) still fits 64 bits.
CDG_D bool rep_better(u32 lenA, u32 idA, u32 posA, u64 keyA, u32 lenB, u32 idB, u32 posB, u64 keyB) {   // A before B ?
    if (lenA != lenB) return lenA > lenB;
    if (idA != idB) return idA < idB;
    if (posA != posB) return posA < posB;
    return (keyA & CDG_B63) > (keyB & CDG_B63);
}
CDG_D void member_diag(bool repRev, u32 repLen, u32 repPos, bool tRev, u32 tLen, u32 tPos, i32 &diag, bool &need) {   // :489-521
    i32 q, t;
    if (repRev && !tRev) { q = (i32) repPos; t = (i32) tPos; need = true; }
    else if (repRev && tRev) { q = (i32) (repLen - 1) - (i32) repPos; t = (i32) (tLen - 1) - (i32) tPos; need = false; }
    else if (!repRev && tRev) { q = (i32) (repLen - 1) - (i32) repPos; t = (i32) (tLen - 1) - (i32) tPos; need = true; }
    else { q = (i32) repPos; t = (i32) tPos; need = false; }
    diag = q - t;
}
// Walk-free: k-mer slots are laid out in (sequence length desc, id asc) order and the radix sort is stable, so
// the first record of every run of equal k-mers already IS the representative (same-sequence repeats at the head
// are resolved by a short scan).  One thread per record: the index of its run head comes from a block-wide max-scan
// of head indices, the output position from a block-wide exclusive scan of keep flags.
static constexpr int GROUP_BLOCK = 256;
static constexpr int GROUP_ITEMS = 8;
static constexpr int GROUP_TILE = GROUP_BLOCK * GROUP_ITEMS;
#define GP(i) ((i) + ((i) >> 3))   // shared-memory index of tile record i (one pad word per 8)
// COV: assignGroup's canBeCovered filter is compiled in (-c > 0 or a length coverage mode: linclust); the assembler's calls (-c 0) run the
// instantiation without it (the two float divisions per member cost 0.5 ms per 10 M reads)
template <bool WIDE = false, bool COV = false>
__global__ void __launch_bounds__(GROUP_BLOCK) group_kernel(const u64 *__restrict__ keys, const u64 *__restrict__ vals, u64 n,
                                                            const u32 *__restrict__ len, GroupParams G, u64 *__restrict__ out,
                                                            unsigned long long *__restrict__ outCount, u64 *__restrict__ outHi = nullptr) {
    // blocked arrangement (thread t owns records [8t, 8t+8)) on 64-bit words: unpadded, the 32 lanes of a load hit 2 of the 16
    // 8-byte bank pairs (16-way conflict, 3.7 x the ideal wavefronts in profiles/r01e_src_group_kernel.csv); one pad word per 8 records
    // spreads them over all 16
    __shared__ u64 sKey[GP(GROUP_TILE) + 2];   // + the record after the tile (singleton test)
    __shared__ u64 sVal[GP(GROUP_TILE) + 2];
    __shared__ u32 shWarp[GROUP_BLOCK / 32 + 1];
    __shared__ u32 sWarpMax[GROUP_BLOCK / 32];
    __shared__ unsigned long long sBase;
    __shared__ u64 sTileHead;              // head index of the run that is open at the tile start
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const u64 tileBase = (u64) blockIdx.x * GROUP_TILE;
    const u32 tileN = (u32) (n - tileBase < (u64) GROUP_TILE ? n - tileBase : (u64) GROUP_TILE);
    for (u32 i = tid; i < tileN; i += GROUP_BLOCK) { sKey[GP(i)] = keys[tileBase + i]; sVal[GP(i)] = vals[tileBase + i]; }
    if (tid == 0) {
        sKey[GP(tileN)] = tileBase + tileN < n ? keys[tileBase + tileN] : CDG_SENTINEL;
        // run open at the tile start: walk back to its head (one thread per tile, runs are short)
        u64 h = tileBase;
        if (tileBase > 0 && keys[tileBase] != CDG_SENTINEL) {   // (a tile inside the sentinel tail has nothing to find)
            const u64 k0 = keys[tileBase] | CDG_B63;
            while (h > 0 && (keys[h - 1] | CDG_B63) == k0) h--;
        }
        sTileHead = h;
    }
    __syncthreads();
    // head flags and run-head index by max-scan (blocked arrangement: thread t owns records [t*ITEMS, (t+1)*ITEMS))
    const u32 first = tid * GROUP_ITEMS;
    u32 headOf[GROUP_ITEMS];   // tile-local index of the run head + 1, 0 = before the tile
    u32 runMax = 0;
#pragma unroll
    for (int j = 0; j < GROUP_ITEMS; j++) {
        const u32 li = first + j;
        if (li < tileN) {
            const u64 k = sKey[GP(li)] | CDG_B63;
            const bool head = li == 0 ? (sTileHead == tileBase) : ((sKey[GP(li - 1)] | CDG_B63) != k);
            if (head) runMax = li + 1;
        }
        headOf[j] = runMax;
    }
    // exclusive max-scan of runMax over threads
    u32 incl = runMax;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const u32 t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl = max(incl, t); }
    if (lane == 31) sWarpMax[warp] = incl;
    __syncthreads();
    u32 carry = 0;
#pragma unroll
    for (int w = 0; w < GROUP_BLOCK / 32; w++) if (w < warp) carry = max(carry, sWarpMax[w]);
    const u32 prevLane = __shfl_up_sync(0xffffffffu, incl, 1);
    if (lane > 0) carry = max(carry, prevLane);
    // per record: representative, diagonal, keep flag
    u64 outKey[GROUP_ITEMS];
    u32 outRep[WIDE ? GROUP_ITEMS : 1];
    u32 keep = 0;
    const int sh1 = G.diagBits + 1, sh2 = G.idBits + G.diagBits + 1;
    u64 cHead = ~0ull, cRepKey = 0; u32 cRepId = 0, cRepPos = 0, cRepLen = 0;   // representative of the run of the previous record
#pragma unroll
    for (int j = 0; j < GROUP_ITEMS; j++) {
        const u32 li = first + j;
        outKey[j] = 0;
        if (li >= tileN) continue;
        const u64 k = sKey[GP(li)];
        if (k == CDG_SENTINEL) continue;
        const u32 hl = max(headOf[j], carry);
        const u64 headIdx = hl ? tileBase + hl - 1 : sTileHead;
        const bool isHead = headIdx == tileBase + li;
        if (isHead && (sKey[GP(li + 1)] == CDG_SENTINEL || (sKey[GP(li + 1)] | CDG_B63) != (k | CDG_B63))) continue;   // singleton (:476)
        // representative = head record; repeats of the head's own sequence: smallest position (forward first)
        auto keyAt = [&](u64 x) -> u64 { return (x >= tileBase && x < tileBase + tileN) ? sKey[GP((u32) (x - tileBase))] : (x < n ? keys[x] : CDG_SENTINEL); };
        auto valAt = [&](u64 x) -> u64 { return (x >= tileBase && x < tileBase + tileN) ? sVal[GP((u32) (x - tileBase))] : vals[x]; };
        // consecutive records of a thread mostly belong to the same run (coverage): the representative is looked up once per run
        if (headIdx != cHead) {
            cHead = headIdx;
            cRepKey = keyAt(headIdx);
            const u64 repVal = valAt(headIdx);
            cRepId = cdg_val_id(repVal);
            cRepPos = cdg_val_pos(repVal);
            for (u64 x = headIdx + 1; x < n; x++) {
                const u64 kx = keyAt(x);
                if (kx == CDG_SENTINEL || (kx | CDG_B63) != (k | CDG_B63)) break;
                const u64 vx = valAt(x);
                if (cdg_val_id(vx) != cRepId) break;
                const u32 px = cdg_val_pos(vx);
                if (px < cRepPos || (px == cRepPos && (kx & CDG_B63) > (cRepKey & CDG_B63))) { cRepPos = px; cRepKey = kx; }
            }
            cRepLen = len[cRepId];
        }
        const u64 repKey = cRepKey; const u32 repId = cRepId, repPos = cRepPos, repLen = cRepLen;
        const bool repRev = (G.firstRunQuirk && headIdx == 0) ? false : ((repKey & CDG_B63) == 0);
        const u64 v = sVal[GP(li)];
        const u32 id = cdg_val_id(v), tl = len[id];
        i32 d; bool need;
        member_diag(repRev, repLen, repPos, (k & CDG_B63) == 0, tl, cdg_val_pos(v), d, need);
        if (G.onlyExtendable && !(d < 0 || d > (i32) repLen - (i32) tl)) continue;   // :526-536
        if (COV && !G.onlyExtendable && !can_be_covered(G.covThr, G.covMode, (float) repLen, (float) tl)) continue;   // :527-529
        outKey[j] = (WIDE ? 0ull : ((u64) repId << sh2)) | ((u64) id << sh1) | ((u64) (u32) (d + G.diagBias) << 1) | (need ? 1ull : 0ull);
        if (WIDE) outRep[j] = repId;
        keep |= 1u << j;
    }
    // compaction: block scan of keep counts, one atomic per block
    u32 blockTotal;
    const u32 myOff = cdgscan::block_excl_scan((u32) __popc(keep), &blockTotal, shWarp);
    if (tid == 0) sBase = blockTotal ? atomicAdd(outCount, (unsigned long long) blockTotal) : 0ull;
    // the kept records go through shared memory (every thread is past its last read of the tile: the scan has barriers) so that the
    // block writes one contiguous, coalesced run instead of 8-byte stores 64 bytes apart
    u32 o = myOff;
#pragma unroll
    for (int j = 0; j < GROUP_ITEMS; j++) if (keep & (1u << j)) { if (WIDE) sVal[GP(o)] = outRep[j]; sKey[GP(o)] = outKey[j]; o++; }
    __syncthreads();
    const u64 dst = sBase;
    for (u32 i = tid; i < blockTotal; i += GROUP_BLOCK) { out[dst + i] = sKey[GP(i)]; if (WIDE) outHi[dst + i] = sVal[GP(i)]; }
}

// ---------------------------------------------------------------------------------------------------
// hit reduction over key2 sorted ascending.  A pair head is the first record of a (rep,target) pair; its
// thread walks forward while the TARGET id stays equal — deliberately not checking the rep, which
// reproduces the reference's rep-boundary bleed (kmermatcher.cpp:886-888) — and picks the most frequent
// diagonal (ties: the larger one), the number of records and the strand of the last record on the best
// diagonal (:889-905).  Self pairs (target == rep) produce no hit (:906).
struct ReduceParams { int idBits, diagBits; i32 diagBias; };
// flag = last record of a (rep,target) pair with target != rep: one hit each, in sorted (= pref DB) order
struct PairEndFlag {
    const u64 *keys; u64 n; int idBits, diagBits; u64 nOwn;   // records [nOwn, n) are a neighbour's head run: walked, never emitted
    const u64 *hi = nullptr;                                  // WIDE records: the representatives
    CDG_D u64 pair_at(u64 i) const { return hi ? ((hi[i] << idBits) | (keys[i] >> (diagBits + 1))) : (keys[i] >> (diagBits + 1)); }
    CDG_D u32 operator()(u64 i) const {
        if (i >= nOwn) return 0;
        const u64 pair = pair_at(i);
        if (i + 1 < n && pair_at(i + 1) == pair) return 0;
        const u32 id = (u32) (pair & ((1ull << idBits) - 1)), rep = (u32) (pair >> idBits);
        return id != rep ? 1u : 0u;
    }
};

// with -c 0 canBeCovered holds for every pair except in the two length modes (Util.cpp:541-544)
inline bool group_cov_active(const GroupParams &G) { return !G.onlyExtendable && (G.covThr > 0.f || G.covMode == 3 || G.covMode == 4); }
inline void launch_group(cudaStream_t st, bool wide, unsigned grid, const u64 *keys, const u64 *vals, u64 n, const u32 *len, const GroupParams &G, u64 *out,
                         unsigned long long *outCount, u64 *outHi) {
    const bool cov = group_cov_active(G);
    if (wide) { if (cov) group_kernel<true, true><<<grid, GROUP_BLOCK, 0, st>>>(keys, vals, n, len, G, out, outCount, outHi); else group_kernel<true, false><<<grid, GROUP_BLOCK, 0, st>>>(keys, vals, n, len, G, out, outCount, outHi); }
    else { if (cov) group_kernel<false, true><<<grid, GROUP_BLOCK, 0, st>>>(keys, vals, n, len, G, out, outCount, nullptr); else group_kernel<false, false><<<grid, GROUP_BLOCK, 0, st>>>(keys, vals, n, len, G, out, outCount, nullptr); }
}

// tile sums of PairEndFlag for one-word records: 128-bit loads (two records per thread and load), the record after the pair by shuffle
// — the generic tile_sum_kernel evaluates the functor per record and loads every key twice
__global__ void __launch_bounds__(cdgscan::BLOCK) pair_end_count_kernel(const u64 *__restrict__ keys /*16-byte aligned*/, u64 n, u64 nOwn, int idBits, int diagBits,
                                                                        unsigned long long *__restrict__ tileSums) {
    __shared__ u32 shWarp[cdgscan::BLOCK / 32 + 1];
    const int lane = threadIdx.x & 31;
    const u64 base = (u64) blockIdx.x * cdgscan::TILE;
    const int sh = diagBits + 1;
    const u64 idMask = (1ull << idBits) - 1;
    u32 s = 0;
#pragma unroll
    for (int j = 0; j < cdgscan::ITEMS / 2; j++) {
        const u64 i = base + 2ull * (u64) (j * cdgscan::BLOCK + threadIdx.x);   // records i, i + 1
        u64 k0 = 0, k1 = 0;
        if (i + 1 < n) { const ulonglong2 t = *reinterpret_cast<const ulonglong2 *>(keys + i); k0 = t.x; k1 = t.y; }
        else if (i < n) k0 = keys[i];
        u64 k2 = __shfl_down_sync(0xffffffffu, k0, 1);
        if (lane == 31) k2 = i + 2 < n ? keys[i + 2] : 0;
        const u64 p0 = k0 >> sh, p1 = k1 >> sh, p2 = k2 >> sh;
        if (i < nOwn && !(i + 1 < n && p1 == p0) && (u32) (p0 & idMask) != (u32) (p0 >> idBits)) s++;
        if (i + 1 < nOwn && !(i + 2 < n && p2 == p1) && (u32) (p1 & idMask) != (u32) (p1 >> idBits)) s++;
    }
    u32 tot;
    cdgscan::block_excl_scan(s, &tot, shWarp);
    if (threadIdx.x == 0) tileSums[blockIdx.x] = tot;
}

// sequential replay of the reference's walk (kmermatcher.cpp:886-905) from the first record `s` of a pair; used for
// pairs that start before the tile and for pairs whose target-id run bleeds into the next representative
CDG_D void reduce_walk(const PairEndFlag &f, u64 n, u64 s, const ReduceParams &R, u32 &top, i32 &best, bool &brev) {
    const u64 *__restrict__ keys = f.keys;
    const u64 idMask = (1ull << R.idBits) - 1, diagMask = (1ull << R.diagBits) - 1;
    const u64 k = keys[s];
    const u64 p0 = f.pair_at(s);
    const u32 tgt = (u32) (p0 & idMask);
    u32 prevRep = (u32) (p0 >> R.idBits);
    i32 prev = (i32) ((k >> 1) & diagMask);
    best = prev; brev = false; top = 0;
    u32 dc = 0, mx = 0; bool runNeed = false;
    for (u64 e = s; e < n; e++) {
        const u64 ke = keys[e];
        const u64 pe = f.pair_at(e);
        if ((u32) (pe & idMask) != tgt) break;   // the rep is deliberately not checked
        const i32 d = (i32) ((ke >> 1) & diagMask);
        const u32 repE = (u32) (pe >> R.idBits);
        const bool sameRun = prev == d;
        dc = sameRun ? dc + 1 : 1;
        // equal (rep,id,diag) records are unordered in the reference; our order puts needRev last -> OR over the run,
        // restarted at a rep boundary inside a bleeding walk (only the later rep's records can be "last")
        runNeed = (sameRun && repE == prevRep) ? (runNeed || (ke & 1ull)) : ((ke & 1ull) != 0);
        if (dc >= mx) { best = d; mx = dc; brev = runNeed; }
        prev = d; prevRep = repE; top++;
    }
}

// Walk-free reduction: one thread per record, two block-wide segmented scans (diagonal-run length / strand OR, then
// best run per pair), results taken at the last record of each pair.
static constexpr int RED_BLOCK = cdgscan::BLOCK;
static constexpr int RED_ITEMS = cdgscan::ITEMS;
static constexpr int RED_TILE = cdgscan::TILE;

// inclusive segmented scan over the tile (blocked arrangement); heads: bit j = item j starts a segment
template <class Op>
CDG_D void block_seg_scan(u64 (&val)[RED_ITEMS], u32 heads, Op op, u64 *shVal /*[warps]*/, u32 *shFlag /*[warps]*/) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int j = 1; j < RED_ITEMS; j++) if (!(heads & (1u << j))) val[j] = op(val[j - 1], val[j]);
    u64 agg = val[RED_ITEMS - 1];
    u32 flag = heads != 0;
    // warp inclusive segmented scan of the thread aggregates
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const u64 v2 = __shfl_up_sync(0xffffffffu, agg, o);
        const u32 f2 = __shfl_up_sync(0xffffffffu, flag, o);
        if (lane >= o) { if (!flag) agg = op(v2, agg); flag |= f2; }
    }
    __syncthreads();
    if (lane == 31) { shVal[warp] = agg; shFlag[warp] = flag; }
    __syncthreads();
    // carry into this thread = segmented combination of everything before it in the tile
    u64 carry = 0; u32 have = 0;   // have: carry holds a value
    for (int w = 0; w < warp; w++) {
        const u64 v = shVal[w]; const u32 f = shFlag[w];
        carry = (f || !have) ? v : op(carry, v);
        have = 1;
    }
    const u64 pv = __shfl_up_sync(0xffffffffu, agg, 1);
    const u32 pf = __shfl_up_sync(0xffffffffu, flag, 1);
    if (lane > 0) { carry = (pf || !have) ? pv : op(carry, pv); have = 1; }
    if (have) {
#pragma unroll
        for (int j = 0; j < RED_ITEMS; j++) { if (heads & (1u << j)) break; val[j] = op(carry, val[j]); }
    }
}
struct OpLenNeed { CDG_D u64 operator()(u64 a, u64 b) const { return (((a >> 1) + (b >> 1)) << 1) | ((a | b) & 1ull); } };
struct OpMax { CDG_D u64 operator()(u64 a, u64 b) const { return a > b ? a : b; } };

template <bool WIDE = false>
__global__ void __launch_bounds__(RED_BLOCK) reduce_kernel(PairEndFlag f, ReduceParams R, const unsigned long long *__restrict__ tileSums,
                                                           u32 *__restrict__ hitRep, u32 *__restrict__ hitTgt,
                                                           i32 *__restrict__ hitScore, i32 *__restrict__ hitDiag) {
    __shared__ __align__(16) u64 sKeyRaw[RED_TILE + 4];
    __shared__ u64 sPairRaw[WIDE ? RED_TILE + 4 : 1];   // WIDE: the (rep,id) pair of every record of sKey (NONE outside the array)
    u64 *sKey = sKeyRaw + 1;             // [0] = record before the tile, [1..tileN] = the tile (16-byte aligned), [tileN+1] = record after it
    u64 *sPair = sPairRaw + (WIDE ? 1 : 0);
    __shared__ __align__(8) u64 sBar;
    __shared__ u64 shVal[RED_BLOCK / 32];
    __shared__ u32 shFlag[RED_BLOCK / 32];
    __shared__ u32 shWarp[RED_BLOCK / 32 + 1];
    __shared__ u32 sWarpMax[RED_BLOCK / 32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const u64 base = (u64) blockIdx.x * RED_TILE;
    const u32 tileN = (u32) (f.n - base < (u64) RED_TILE ? f.n - base : (u64) RED_TILE);
    const u64 NONE = ~0ull;   // pair field of "no record": real pairs have at most 2*idBits < 64 bits
    const bool tma = cdg_tma_ok(f.keys + base, tileN * 8);
    if (tma) {   // the tile is one contiguous run: one TMA bulk copy
        if (tid == 0) cdg_mbar_init(&sBar, 1);
        __syncthreads();
        if (tid == 0) { cdg_mbar_expect_tx(&sBar, tileN * 8); cdg_tma_load_1d(sKey + 1, f.keys + base, tileN * 8, &sBar); }
    } else {
        for (u32 i = tid; i < tileN; i += RED_BLOCK) sKey[i + 1] = f.keys[base + i];
    }
    if (tid == 0) { sKey[0] = base > 0 ? f.keys[base - 1] : NONE; sKey[tileN + 1] = base + tileN < f.n ? f.keys[base + tileN] : NONE; }
    if (WIDE) {
        for (u32 i = tid; i < tileN; i += RED_BLOCK) sPair[i + 1] = f.pair_at(base + i);
        if (tid == 0) { sPair[0] = base > 0 ? f.pair_at(base - 1) : NONE; sPair[tileN + 1] = base + tileN < f.n ? f.pair_at(base + tileN) : NONE; }
    }
    if (tma) cdg_mbar_wait(&sBar, 0);
    __syncthreads();
    const int pairShift = R.diagBits + 1;
    const u64 idMask = (1ull << R.idBits) - 1, diagMask = (1ull << R.diagBits) - 1;
    // pair of the record at shared index x (0 = before the tile)
    auto pairAt = [&](u32 x) -> u64 { if (WIDE) return sPair[x]; const u64 k = sKey[x]; return k == NONE ? NONE : (k >> pairShift); };
    const u32 first = tid * RED_ITEMS;
    u32 pairHeads = 0, runHeads = 0, pairEnds = 0, runEnds = 0;
    u64 v1[RED_ITEMS];
    u32 headOf[RED_ITEMS]; u32 runMax = 0;   // tile-local index + 1 of the pair head, 0 = before the tile
    // the thread's 8 consecutive records as four 128-bit loads (lane stride 64 B: 4-way conflicts per quarter warp, 16 wavefronts per
    // load) instead of eight 64-bit loads with a 16-way conflict each (32 wavefronts per load: profiles/r01e_src_reduce_kernel.csv had
    // 3 x the ideal shared-memory wavefronts); the neighbours' edge records come by shuffle
    static_assert(RED_ITEMS == 8, "reduce_kernel loads 8 records per thread as 4 x 16 bytes");
    u64 kk[RED_ITEMS + 2];   // kk[0] = record before the thread's first, kk[1..8] = its records, kk[9] = record after its last
    {
        const ulonglong2 *src = reinterpret_cast<const ulonglong2 *>(sKey + 1 + first);
#pragma unroll
        for (int j = 0; j < RED_ITEMS / 2; j++) { const ulonglong2 t2 = src[j]; kk[1 + 2 * j] = t2.x; kk[2 + 2 * j] = t2.y; }
        const u64 up = __shfl_up_sync(0xffffffffu, kk[RED_ITEMS], 1), down = __shfl_down_sync(0xffffffffu, kk[1], 1);
        kk[0] = lane > 0 ? up : sKey[first];
        kk[RED_ITEMS + 1] = lane < 31 ? down : sKey[first + RED_ITEMS + 1];
    }
    auto pairOf = [&](u64 k) -> u64 { return k == NONE ? NONE : (k >> pairShift); };
#pragma unroll
    for (int j = 0; j < RED_ITEMS; j++) {
        const u32 li = first + j;
        v1[j] = 0;
        if (li < tileN) {
            const u64 k = kk[j + 1], kp = kk[j], kn = kk[j + 2];
            const bool ph = WIDE ? pairAt(li) != pairAt(li + 1) : pairOf(kp) != pairOf(k), pe = WIDE ? pairAt(li + 2) != pairAt(li + 1) : pairOf(kn) != pairOf(k);
            const bool rh = ph || ((kp >> 1) & diagMask) != ((k >> 1) & diagMask);
            const bool re = pe || ((kn >> 1) & diagMask) != ((k >> 1) & diagMask);
            pairHeads |= ph ? 1u << j : 0; runHeads |= rh ? 1u << j : 0; pairEnds |= pe ? 1u << j : 0; runEnds |= re ? 1u << j : 0;
            v1[j] = 2ull | (k & 1ull);   // (length 1) << 1 | needRev
            if (ph) runMax = li + 1;
        } else { pairHeads |= 1u << j; runHeads |= 1u << j; }
        headOf[j] = runMax;
    }
    // scan 1: per diagonal run (length, OR of needRev)
    block_seg_scan(v1, runHeads, OpLenNeed(), shVal, shFlag);
    // scan 2: per pair, best run so far = max of (runLength, position, needRev) taken at run ends
    u64 v2[RED_ITEMS];
#pragma unroll
    for (int j = 0; j < RED_ITEMS; j++) v2[j] = (runEnds & (1u << j)) ? (((v1[j] >> 1) << 34) | ((u64) (first + j) << 1) | (v1[j] & 1ull)) : 0ull;
    block_seg_scan(v2, pairHeads, OpMax(), shVal, shFlag);
    // pair-head index by max-scan
    u32 incl = runMax;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const u32 t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl = max(incl, t); }
    __syncthreads();
    if (lane == 31) sWarpMax[warp] = incl;
    __syncthreads();
    u32 carry = 0;
#pragma unroll
    for (int w = 0; w < RED_BLOCK / 32; w++) if (w < warp) carry = max(carry, sWarpMax[w]);
    const u32 prevLane = __shfl_up_sync(0xffffffffu, incl, 1);
    if (lane > 0) carry = max(carry, prevLane);
    // hits at pair ends, compacted in order
    u32 flags = 0;
#pragma unroll
    for (int j = 0; j < RED_ITEMS; j++) {
        const u32 li = first + j;
        if (li < tileN && (pairEnds & (1u << j)) && base + li < f.nOwn) {
            const u64 pair = pairAt(li + 1);
            if ((u32) (pair & idMask) != (u32) (pair >> R.idBits)) flags |= 1u << j;
        }
    }
    u32 tot;
    const u32 ex = cdgscan::block_excl_scan((u32) __popc(flags), &tot, shWarp);
    u64 h = tileSums[blockIdx.x] + ex;
#pragma unroll
    for (int j = 0; j < RED_ITEMS; j++) {
        if (!(flags & (1u << j))) continue;
        const u32 li = first + j;
        const u64 pr0 = pairAt(li + 1);
        const u32 tgt = (u32) (pr0 & idMask), rep = (u32) (pr0 >> R.idBits);
        const u32 hl = max(headOf[j], carry);
        const u64 pn = pairAt(li + 2);
        const bool bleeds = pn != NONE && (u32) (pn & idMask) == tgt;   // next rep's first target is the same id
        u32 top; i32 best; bool brev;
        if (hl != 0 && !bleeds) {
            const u32 startLi = hl - 1;
            top = li - startLi + 1;
            const u64 b = v2[j];
            best = (i32) ((sKey[(u32) ((b >> 1) & 0xFFFFFFFFull) + 1] >> 1) & diagMask);
            brev = (b & 1ull) != 0;
        } else {
            u64 s = base + li;
            if (hl != 0) s = base + hl - 1;
            else { while (s > 0 && f.pair_at(s - 1) == pr0) s--; }
            reduce_walk(f, f.n, s, R, top, best, brev);
        }
        hitRep[h] = rep; hitTgt[h] = tgt;
        hitScore[h] = brev ? -(i32) top : (i32) top;
        hitDiag[h] = best - R.diagBias;
        h++;
    }
}

}   // namespace cdgkmer
