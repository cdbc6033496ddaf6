// What crosses PCIe (packed sequences in, alignment SoA + corrected sequences out):
//   unpack_kernel          2-bit words + N mask -> the DB byte layout on the device (cdg_upload_packed)
//   compact alignment export: only the alignments rescorediagonal writes (rescorediagonal.cpp:304-314), 14 bytes each
//   patch_kernel           corrected reads as (byte offset, new base) patches (correction.cpp:465-470 rewrites whole entries)
#pragma once
#include "cdg_common.cuh"
#include "cdg_kmer.cuh"
#include "cdg_sort.cuh"

namespace cdgexp {

// cdg_upload_packed, first step (one thread per sequence): the words / N mask are sanitised in place (N positions and the padding
// of the last word hold code 0, mask bits beyond the sequence are cleared) so that they are bit-identical to what pack_kernel
// derives from the same bytes; flags[seq] = the sequence holds an N; *flagged counts those sequences.  The DB byte layout is only
// rebuilt (unpack_kernel) when some sequence holds an N — the byte paths of rescore / correction are reached through flagged
// sequences only — or when a caller asks for bytes (ensure_bytes in cdg_api.cu).
__global__ void __launch_bounds__(256) sanitize_packed_kernel(u64 *__restrict__ words, u32 *__restrict__ xmask, const u32 *__restrict__ len,
                                                              const u64 *__restrict__ wordOff, u64 seqLo, u64 seqHi, u8 *__restrict__ flags,
                                                              unsigned long long *__restrict__ flagged) {
    const u64 seq = seqLo + (u64) blockIdx.x * blockDim.x + threadIdx.x;
    u32 any = 0;
    if (seq < seqHi) {
        const u32 L = len[seq];
        const u64 w0 = wordOff[seq];
        const u32 nw = (L + 31) >> 5;
        for (u32 w = 0; w < nw; w++) {
            const u32 c = L - (w << 5) < 32 ? L - (w << 5) : 32;
            const u64 word = words[w0 + w];
            const u32 xm0 = xmask[w0 + w];
            u32 xm = xm0;
            if (c < 32) xm &= ~(0xFFFFFFFFu >> c);
            const u64 valid = c == 32 ? ~0ull : ~(~0ull >> (2 * c));
            const u64 clean = word & valid & ~cdgkmer::interleave32(xm, xm);
            if (clean != word) words[w0 + w] = clean;
            if (xm != xm0) xmask[w0 + w] = xm;
            any |= xm;
        }
        flags[seq] = any ? 1 : 0;
    }
    const u32 b = __ballot_sync(0xffffffffu, any != 0);
    if ((threadIdx.x & 31) == 0 && b) atomicAdd(flagged, (unsigned long long) __popc(b));
}

// One warp per sequence: the DB byte layout from sanitised words.  onlyIfFlagged (device scalar, may be null): return at once when it
// is zero (no sequence holds an N: nothing reads the bytes).
__global__ void __launch_bounds__(256) unpack_kernel(const u64 *__restrict__ words, const u32 *__restrict__ xmask, const u64 *__restrict__ off,
                                                     const u32 *__restrict__ len, const u64 *__restrict__ wordOff, u64 seqLo, u64 seqHi,
                                                     u8 *__restrict__ data, const unsigned long long *__restrict__ onlyIfFlagged) {
    if (onlyIfFlagged && *onlyIfFlagged == 0) return;
    const int lane = threadIdx.x & 31;
    const u64 gwarp = ((u64) blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = ((u64) gridDim.x * blockDim.x) >> 5;
    for (u64 seq = seqLo + gwarp; seq < seqHi; seq += nwarps) {
        const u32 L = len[seq];
        const u64 w0 = wordOff[seq];
        u8 *d = data + off[seq];
        const u32 nw = (L + 31) >> 5;
        for (u32 w = 0; w < nw; w++) {
            const u32 c = L - (w << 5) < 32 ? L - (w << 5) : 32;
            const u64 word = words[w0 + w];
            const u32 xm = xmask[w0 + w];
            if ((u32) lane < c) {
                const u32 code = (u32) (word >> (62 - 2 * lane)) & 3u;
                d[(w << 5) + lane] = ((xm >> (31 - lane)) & 1u) ? (u8) 'N' : (u8) cdg_num2aa((int) code);
            }
        }
        if (lane == 0) { d[L] = (u8) '\n'; d[L + 1] = 0; }
    }
}

// ---- compact alignment export -------------------------------------------------------------------------------------------
struct LineArrays {
    const u32 *lineQ, *lineT; const i32 *linePref; const u8 *keep; const i32 *score; const u32 *idCnt; const float *seqId;
    const i32 *qS, *qE, *dS, *dE; const SeqMeta *meta;
};
// the self alignment every plain ACGT query has: implied, not exported
CDG_D bool implicit_self(const LineArrays &A, u64 l) {
    const u32 q = A.lineQ[l];
    if (q != A.lineT[l] || A.linePref[l] < 0) return false;
    const SeqMeta m = A.meta[q];
    if (m.lenFlags & META_EXC) return false;
    const i32 L = (i32) (m.lenFlags & META_LEN);
    return L > 0 && A.score[l] == 2 * L && A.idCnt[l] == (u32) L && A.qS[l] == 0 && A.qE[l] == L - 1 && A.dS[l] == 0 && A.dE[l] == L - 1;
}
struct ExportFlag {
    LineArrays A;
    CDG_D u32 operator()(u64 l) const { return (A.keep[l] && !implicit_self(A, l)) ? 1u : 0u; }
};

template <class T, class TS>
__global__ void __launch_bounds__(cdgscan::BLOCK) compact_write_kernel(ExportFlag f, u64 M, const unsigned long long *__restrict__ tileSums,
                                                                       u32 *__restrict__ target, T *__restrict__ score, T *__restrict__ idCnt,
                                                                       TS *__restrict__ qStart, TS *__restrict__ dbStart, T *__restrict__ alnLenRev) {
    __shared__ u32 shWarp[cdgscan::BLOCK / 32 + 1];
    u64 prefix[cdgscan::ITEMS]; u32 val[cdgscan::ITEMS];
    cdgscan::tile_prefix(f, M, tileSums, shWarp, prefix, val);
    const u64 base = (u64) blockIdx.x * cdgscan::TILE;
#pragma unroll
    for (int i = 0; i < cdgscan::ITEMS; i++) {
        const u64 l = base + (u64) i * cdgscan::BLOCK + threadIdx.x;
        if (l < M && val[i]) {
            const u64 o = prefix[i];
            const i32 qs = f.A.qS[l], qe = f.A.qE[l], ds = f.A.dS[l], de = f.A.dE[l];
            const bool rev = qs > qe;
            const u32 alnLen = (u32) (de - ds + 1);
            target[o] = f.A.lineT[l]; score[o] = (T) f.A.score[l]; idCnt[o] = (T) f.A.idCnt[l];
            qStart[o] = (TS) qs; dbStart[o] = (TS) ds; alnLenRev[o] = (T) ((alnLen << 1) | (rev ? 1u : 0u));
        }
    }
}
// per query: exported alignments | implicit-self bit 31
__global__ void __launch_bounds__(256) compact_count_kernel(ExportFlag f, const u64 *__restrict__ lineOff, u64 nQ, u32 *__restrict__ count) {
    const u64 q = (u64) blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= nQ) return;
    u32 c = 0, imp = 0;
    for (u64 l = lineOff[q]; l < lineOff[q + 1]; l++) {
        if (!f.A.keep[l]) continue;
        if (implicit_self(f.A, l)) imp = 0x80000000u; else c++;
    }
    count[q] = c | imp;
}

// ---- corrected reads -> patches ------------------------------------------------------------------------------------------
// bytes [b0, b1) of the DB layout (b0 a multiple of 16 below the range start is fine: both buffers are cudaMalloc-aligned)
__global__ void __launch_bounds__(256) patch_kernel(const u8 *__restrict__ in, const u8 *__restrict__ out, u64 b0, u64 b1, u64 *__restrict__ patches,
                                                    u64 cap, unsigned long long *__restrict__ count) {
    const u64 c0 = b0 >> 4, c1 = (b1 + 15) >> 4;
    for (u64 c = c0 + (u64) blockIdx.x * blockDim.x + threadIdx.x; c < c1; c += (u64) gridDim.x * blockDim.x) {
        const uint4 a = reinterpret_cast<const uint4 *>(in)[c], b = reinterpret_cast<const uint4 *>(out)[c];
        if (a.x == b.x && a.y == b.y && a.z == b.z && a.w == b.w) continue;
        const u32 aw[4] = {a.x, a.y, a.z, a.w}, bw[4] = {b.x, b.y, b.z, b.w};
        for (int w = 0; w < 4; w++) {
            u32 x = aw[w] ^ bw[w];
            while (x) {
                const int byte = (__ffs((int) x) - 1) >> 3;
                x &= ~(0xFFu << (8 * byte));
                const u64 pos = (c << 4) + (u64) (4 * w + byte);
                if (pos < b0 || pos >= b1) continue;
                const u64 slot = atomicAdd(count, 1ull);
                if (slot < cap) patches[slot] = (pos << 8) | (u64) ((bw[w] >> (8 * byte)) & 0xFFu);
            }
        }
    }
}

}   // namespace cdgexp
