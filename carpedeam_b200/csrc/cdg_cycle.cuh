// cyclecheck on the device (src/assembler/cyclecheck.cpp:31-269): which contigs wrap around (circular genomes /
// terminal redundancy) and at which diagonal.
// The reference sorts the k-mers of every sequence three times (front / middle / back third) and merges the sorted lists; here ALL k-mers
// of ALL sequences are one array of 16-byte records (key = k-mer index, value = sequence << 32 | position), sorted once by the k-mer with
// the in-house radix sort — the sort is stable and the records are emitted in (sequence, position) order, so every run of equal k-mers is
// ordered by sequence, then position: front, middle, back part of one sequence are adjacent and in that order.
//   cycle_emit_kernel   k-mer index = sum code_i * 4^i over the mapped letters (Indexer.h:21-84; an X adds 4 * 4^i: windows with X are not
//                       skipped, cyclecheck.cpp:104-128)
//   cycle_match_kernel  per record of the middle / back part: first occurrence of its k-mer in the front (and middle) part of the same
//                       sequence -> diagonal hit (:146-209); the part is decided with the iterator position before nextKmer() (:107-108)
//   cycle_match_head_kernel  the same for records deep inside very long sub-runs (a k-mer repeated > 4096 times in one sequence), linear
//   cycle_band_kernel   first diagonal whose 1 % band holds more than 0.24 hits per possible k-mer (:212-236)
#pragma once
#include "cdg_common.cuh"

namespace cdgcycle {

// part of the k-mer at position p: 0 front, 1 middle, 2 back (position 0 has iterator position UINT_MAX -> back)
CDG_HD int part_of(u32 p, u32 third) {
    const u32 it = p - 1u;
    return it < third + 1u ? 0 : (it < 2u * third + 1u ? 1 : 2);
}

// one warp per sequence; recOff[s] = first record of sequence s (prefix sum of the window counts), sequences >= maxSeqLen own no records
__global__ void __launch_bounds__(256) cycle_emit_kernel(const u8 *__restrict__ data, const u64 *__restrict__ off, const u32 *__restrict__ len, u64 n, int k,
                                                         u32 maxSeqLen, const u64 *__restrict__ recOff, u64 *__restrict__ keys, u64 *__restrict__ vals) {
    const int lane = threadIdx.x & 31;
    const u64 gwarp = ((u64) blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = ((u64) gridDim.x * blockDim.x) >> 5;
    for (u64 s = gwarp; s < n; s += nwarps) {
        const u32 L = len[s];
        if (L >= maxSeqLen || L < (u32) k) continue;
        const u8 *q = data + off[s];
        const u64 base = recOff[s];
        for (u32 p = lane; p + (u32) k <= L; p += 32) {
            u64 idx = 0;
            for (int i = k - 1; i >= 0; i--) idx = idx * 4ull + (u64) cdg_map_letter(q[p + i]);   // Horner: code_{k-1} * 4^{k-1} + ... + code_0
            keys[base + p] = idx;
            vals[base + p] = (s << 32) | (u64) p;
        }
    }
}

// ---- matching: host/device bodies (tests/cycle_host_harness.cu runs them on the CPU against the oracle) ---------------------------
CDG_HD void cycle_add_hit(u32 *p) {
#ifdef __CUDA_ARCH__
    atomicAdd(p, 1u);
#else
    ++*p;
#endif
}
// hits of the record at `pos` (part 1 or 2) against the first front / first middle occurrence of its k-mer in its sequence
CDG_HD void cycle_record_hits(int part, u32 pos, u32 third, bool haveF, u32 posF, bool haveM, u32 posM, u32 *hits) {
    if (haveF) { const i32 d = (i32) (pos - posF); if (d >= (i32) third) cycle_add_hit(&hits[(u32) d - third]); }
    if (part == 2 && haveM) { const i32 d = (i32) (pos - posM); if (d >= (i32) third) cycle_add_hit(&hits[(u32) d - third]); }
}
// Record r walks BACK over the records of the same k-mer and sequence (descending position): the last front record seen is the first front
// occurrence (the sub-run's head), the last middle record seen the first middle occurrence.  A record more than `cap` records behind its
// sub-run's head gives up (k-mers repeated thousands of times inside one sequence: low-complexity contigs) and is served by
// cycle_match_head, so that the work per sub-run stays linear beyond the cap.
CDG_HD void cycle_match_record(const u64 *__restrict__ keys, const u64 *__restrict__ vals, u64 r, const u32 *__restrict__ len, const u64 *__restrict__ diagOff,
                               u32 *__restrict__ diagHits, u32 cap) {
    const u64 key = keys[r], v = vals[r];
    const u32 s = (u32) (v >> 32), pos = (u32) v;
    const u32 third = len[s] / 3u;
    const int part = part_of(pos, third);
    if (part == 0) return;
    bool haveF = false, haveM = false; u32 posF = 0, posM = 0, steps = 0;
    for (u64 x = r; x > 0;) {
        x--;
        if (keys[x] != key) break;
        const u64 vx = vals[x];
        if ((u32) (vx >> 32) != s) break;
        if (++steps > cap) return;
        const u32 px = (u32) vx;
        const int pp = part_of(px, third);
        if (pp == 0) { haveF = true; posF = px; }
        else if (pp == 1) { haveM = true; posM = px; }
        // (a back record below a middle one cannot happen: position order; position 0 — back by the iterator quirk — comes first)
    }
    cycle_record_hits(part, pos, third, haveF, posF, haveM, posM, diagHits + diagOff[s]);
}
// Record r, if it is the head of its sub-run, walks FORWARD once and serves the records more than `cap` behind it (none for almost every k-mer)
CDG_HD void cycle_match_head(const u64 *__restrict__ keys, const u64 *__restrict__ vals, u64 R, u64 r, const u32 *__restrict__ len,
                             const u64 *__restrict__ diagOff, u32 *__restrict__ diagHits, u32 cap) {
    const u64 key = keys[r], v = vals[r];
    const u32 s = (u32) (v >> 32);
    if (r > 0 && keys[r - 1] == key && (u32) (vals[r - 1] >> 32) == s) return;   // not a head
    const u32 third = len[s] / 3u;
    u32 *hits = diagHits + diagOff[s];
    bool haveF = false, haveM = false; u32 posF = 0, posM = 0;
    for (u64 x = r; x < R; x++) {
        if (keys[x] != key) break;
        const u64 vx = vals[x];
        if ((u32) (vx >> 32) != s) break;
        const u32 px = (u32) vx;
        const int pp = part_of(px, third);
        if (pp != 0 && x - r > (u64) cap) cycle_record_hits(pp, px, third, haveF, posF, haveM, posM, hits);
        if (pp == 0 && !haveF) { haveF = true; posF = px; }
        else if (pp == 1 && !haveM) { haveM = true; posM = px; }
    }
}
static constexpr u32 CYCLE_WALK_CAP = 4096;

// records sorted by k-mer (stable): one thread per record
__global__ void __launch_bounds__(256) cycle_match_kernel(const u64 *__restrict__ keys, const u64 *__restrict__ vals, u64 R, const u32 *__restrict__ len,
                                                          const u64 *__restrict__ diagOff, u32 *__restrict__ diagHits, u32 cap) {
    const u64 r = (u64) blockIdx.x * blockDim.x + threadIdx.x;
    if (r < R) cycle_match_record(keys, vals, r, len, diagOff, diagHits, cap);
}
__global__ void __launch_bounds__(256) cycle_match_head_kernel(const u64 *__restrict__ keys, const u64 *__restrict__ vals, u64 R, const u32 *__restrict__ len,
                                                               const u64 *__restrict__ diagOff, u32 *__restrict__ diagHits, u32 cap) {
    const u64 r = (u64) blockIdx.x * blockDim.x + threadIdx.x;
    if (r < R) cycle_match_head(keys, vals, R, r, len, diagOff, diagHits, cap);
}

// one warp per sequence: lanes test 32 consecutive diagonals, the first passing one wins
__global__ void __launch_bounds__(256) cycle_band_kernel(const u32 *__restrict__ len, u64 n, int k, u32 maxSeqLen, const u64 *__restrict__ diagOff,
                                                         const u32 *__restrict__ diagHits, u32 *__restrict__ split) {
    const int lane = threadIdx.x & 31;
    const u64 gwarp = ((u64) blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = ((u64) gridDim.x * blockDim.x) >> 5;
    for (u64 s = gwarp; s < n; s += nwarps) {
        const u32 L = len[s], third = L / 3u;
        u32 found = 0;
        if (L < maxSeqLen) {
            const u32 *hits = diagHits + diagOff[s];
            for (u32 d0 = 0; d0 < 2u * third && !found; d0 += 32) {
                const u32 d = d0 + lane;
                bool ok = false;
                if (d < 2u * third) {
                    const u32 hd = hits[d];
                    if (hd != 0) {
                        const u32 diag = d + third, diaglen = L - diag;
                        const u32 gapwindow = (u32) ((double) diaglen * 0.01);
                        const u32 lower = d >= gapwindow ? d - gapwindow : 0u;          // max(0, (int) (d - gapwindow))
                        const u32 upper = min(d + gapwindow, 2u * third);
                        u32 band = 0;
                        for (u32 i = lower; i <= upper; i++) { const u32 h = hits[i]; if (h <= hd) band += h; }
                        const float rate = (float) band / (float) ((u64) diaglen - (u64) k + 1ull);   // size_t arithmetic (wraps for diaglen < k - 1)
                        ok = (double) rate > 0.24;   // HIT_RATE_THRESHOLD
                    }
                }
                const u32 b = __ballot_sync(0xffffffffu, ok);
                if (b) found = d0 + (u32) (__ffs((int) b) - 1) + third;
            }
        }
        if (lane == 0) split[s] = found;
    }
}

}   // namespace cdgcycle
