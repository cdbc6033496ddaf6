// In-house LSD radix sort (onesweep style: one global histogram pass, then per digit ONE read and ONE
// write of the records, chained-scan with decoupled look-back between tiles) and a 3-kernel exclusive
// scan.  Replaces the reference's ips4o SORT_PARALLEL calls (kmermatcher.cpp:412,431).  HBM-bound:
// algorithmic traffic per record and digit = sizeof(record) read + write.
//
// The sort is stable and sorts records by the bit range [bitLo, bitHi) of a "digit word":
// either the key word or the value word of each (key,val) pair, so a 128-bit ordering is done as
// passes over the value bits followed by passes over the key bits.
#pragma once
#include "cdg_common.cuh"

namespace cdgsort {

static constexpr int RADIX_BITS = 8;
static constexpr int RADIX = 1 << RADIX_BITS;
static constexpr int BLOCK = 256;                 // == RADIX: thread d owns digit d in the scan phases
static constexpr int WARPS = BLOCK / 32;
static constexpr int ITEMS = 16;
static constexpr int TILE = BLOCK * ITEMS;        // 4096 records per tile
static constexpr int MAX_PASSES = 16;
static constexpr u32 FLAG_AGG = 1u << 30, FLAG_INC = 2u << 30, FLAG_MASK = 3u << 30, VAL_MASK = ~FLAG_MASK;
static constexpr u64 MAX_N = (1ull << 30) - 1;    // 30-bit tile-state counters

struct Pass { int word; int shift; u32 mask; };   // word: 0 = key, 1 = val
struct Plan { int n; Pass p[MAX_PASSES]; };

CDG_D u32 digit_of(u64 w, int shift, u32 mask) { return (u32) (w >> shift) & mask; }

// ---- kernel 1: histograms of every pass in one read of the digit words --------------------------
__global__ void __launch_bounds__(BLOCK) hist_kernel(const u64 *__restrict__ keys, const u64 *__restrict__ vals, u64 n,
                                                     Plan plan, unsigned long long *__restrict__ hist /*[n_pass][RADIX]*/) {
    extern __shared__ u32 sh[];   // [plan.n][RADIX]
    for (int i = threadIdx.x; i < plan.n * RADIX; i += BLOCK) sh[i] = 0;
    __syncthreads();
    bool needKey = false, needVal = false;
    for (int p = 0; p < plan.n; p++) { if (plan.p[p].word == 0) needKey = true; else needVal = true; }
    u64 stride = (u64) gridDim.x * BLOCK;
    for (u64 i = (u64) blockIdx.x * BLOCK + threadIdx.x; i < n; i += stride) {
        u64 k = needKey ? keys[i] : 0, v = needVal ? vals[i] : 0;
        for (int p = 0; p < plan.n; p++) {
            u64 w = plan.p[p].word == 0 ? k : v;
            atomicAdd(&sh[p * RADIX + digit_of(w, plan.p[p].shift, plan.p[p].mask)], 1u);
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < plan.n * RADIX; i += BLOCK) {
        u32 c = sh[i];
        if (c) atomicAdd(&hist[i], (unsigned long long) c);
    }
}

// ---- kernel 2: exclusive scan of each pass's 256 bins -> global bin bases -------------------------
__global__ void __launch_bounds__(RADIX) binscan_kernel(unsigned long long *__restrict__ hist, int nPass) {
    __shared__ unsigned long long s[RADIX];
    for (int p = 0; p < nPass; p++) {
        unsigned long long c = hist[p * RADIX + threadIdx.x];
        s[threadIdx.x] = c;
        __syncthreads();
        for (int off = 1; off < RADIX; off <<= 1) {   // Hillis-Steele inclusive scan
            unsigned long long t = threadIdx.x >= off ? s[threadIdx.x - off] : 0;
            __syncthreads();
            s[threadIdx.x] += t;
            __syncthreads();
        }
        hist[p * RADIX + threadIdx.x] = s[threadIdx.x] - c;
        __syncthreads();
    }
}

// ---- kernel 3: one digit pass ----------------------------------------------------------------------
// Tile order = order of arrival (atomic ticket), so every tile a block waits for is already running.
template <bool HAS_VAL>
__global__ void __launch_bounds__(BLOCK) onesweep_kernel(const u64 *__restrict__ keysIn, u64 *__restrict__ keysOut,
                                                         const u64 *__restrict__ valsIn, u64 *__restrict__ valsOut, u64 n,
                                                         Pass pass, const unsigned long long *__restrict__ binBase,
                                                         u32 *tileState /*[tiles][RADIX]*/, u32 *ticket) {
    __shared__ u64 exch[TILE];
    __shared__ u32 warpHist[WARPS][RADIX];
    __shared__ u32 tileOff[RADIX];          // exclusive offset of digit d inside the sorted tile
    __shared__ unsigned long long gBase[RADIX];   // global position of the tile's first record with digit d
    __shared__ u8 digitAt[TILE];
    __shared__ u32 sTile;
    __shared__ u32 warpSum[WARPS];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) sTile = atomicAdd(ticket, 1u);
    for (int i = tid; i < WARPS * RADIX; i += BLOCK) (&warpHist[0][0])[i] = 0;
    __syncthreads();
    const u32 tile = sTile;
    const u64 tileBase = (u64) tile * TILE;
    const u32 tileValid = (u32) (n - tileBase < (u64) TILE ? n - tileBase : (u64) TILE);

    // load (warp-striped: warp w owns records [w*ITEMS*32, (w+1)*ITEMS*32) of the tile, item-major)
    u64 key[ITEMS];
    u32 dig[ITEMS];
    u32 rank[ITEMS];
    const u32 warpBase = warp * ITEMS * 32;
#pragma unroll
    for (int i = 0; i < ITEMS; i++) {
        u32 local = warpBase + i * 32 + lane;
        bool ok = local < tileValid;
        key[i] = ok ? keysIn[tileBase + local] : 0;
        u64 w = key[i];
        if (HAS_VAL && pass.word == 1) w = ok ? valsIn[tileBase + local] : 0;
        dig[i] = ok ? digit_of(w, pass.shift, pass.mask) : pass.mask;   // padding ranks last in the last bin
    }
    // stable rank inside the warp's chunk
#pragma unroll
    for (int i = 0; i < ITEMS; i++) {
        u32 d = dig[i];
        u32 peers = __match_any_sync(0xffffffffu, d);
        u32 pre = warpHist[warp][d];
        __syncwarp();
        if (lane == (__ffs(peers) - 1)) warpHist[warp][d] = pre + __popc(peers);
        __syncwarp();
        rank[i] = pre + __popc(peers & ((1u << lane) - 1));
    }
    __syncthreads();
    // thread d: scan digit d over the warps, publish the tile aggregate, look back
    {
        const int d = tid;
        u32 run = 0;
#pragma unroll
        for (int w = 0; w < WARPS; w++) { u32 c = warpHist[w][d]; warpHist[w][d] = run; run += c; }
        volatile u32 *state = tileState;
        u64 excl = 0;
        if (tile == 0) {
            state[d] = FLAG_INC | run;
        } else {
            state[(u64) tile * RADIX + d] = FLAG_AGG | run;
            i64 p = (i64) tile - 1;
            while (true) {
                u32 s = state[(u64) p * RADIX + d];
                if ((s & FLAG_MASK) == 0) continue;   // predecessor has not published yet
                excl += s & VAL_MASK;
                if ((s & FLAG_MASK) == FLAG_INC) break;
                p--;
            }
            state[(u64) tile * RADIX + d] = FLAG_INC | (u32) (excl + run);
        }
        gBase[d] = binBase[d] + excl;
        // exclusive scan of the tile's digit counts over d (block scan via warp shuffles)
        u32 incl = run;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) { u32 t = __shfl_up_sync(0xffffffffu, incl, off); if (lane >= off) incl += t; }
        if (lane == 31) warpSum[warp] = incl;
        __syncthreads();
        u32 wpre = 0;
#pragma unroll
        for (int w = 0; w < WARPS; w++) if (w < warp) wpre += warpSum[w];
        tileOff[d] = wpre + incl - run;
    }
    __syncthreads();
    // scatter into tile order through shared memory, then coalesced runs to global
    u32 pos[ITEMS];
#pragma unroll
    for (int i = 0; i < ITEMS; i++) {
        pos[i] = tileOff[dig[i]] + warpHist[warp][dig[i]] + rank[i];
        exch[pos[i]] = key[i];
        digitAt[pos[i]] = (u8) dig[i];
    }
    __syncthreads();
    for (u32 j = tid; j < tileValid; j += BLOCK) {
        u32 d = digitAt[j];
        keysOut[gBase[d] + (j - tileOff[d])] = exch[j];
    }
    if (HAS_VAL) {
        __syncthreads();
#pragma unroll
        for (int i = 0; i < ITEMS; i++) {
            u32 local = warpBase + i * 32 + lane;
            if (local < tileValid) exch[pos[i]] = valsIn[tileBase + local];
        }
        __syncthreads();
        for (u32 j = tid; j < tileValid; j += BLOCK) {
            u32 d = digitAt[j];
            valsOut[gBase[d] + (j - tileOff[d])] = exch[j];
        }
    }
}

// plan: passes of <= 8 bits over val bits [vLo,vHi) then key bits [kLo,kHi)
inline Plan make_plan(int vLo, int vHi, int kLo, int kHi) {
    Plan pl; pl.n = 0;
    for (int b = vLo; b < vHi; b += RADIX_BITS) { int w = (vHi - b < RADIX_BITS) ? vHi - b : RADIX_BITS; pl.p[pl.n++] = Pass{1, b, (1u << w) - 1}; }
    for (int b = kLo; b < kHi; b += RADIX_BITS) { int w = (kHi - b < RADIX_BITS) ? kHi - b : RADIX_BITS; pl.p[pl.n++] = Pass{0, b, (1u << w) - 1}; }
    return pl;
}

struct Workspace {   // device scratch owned by the caller
    unsigned long long *hist;   // [MAX_PASSES][RADIX]
    u32 *tileState;             // [tiles][RADIX]
    u32 *ticket;
    u64 tilesCap;
};
inline u64 tiles_for(u64 n) { return (n + TILE - 1) / TILE; }

// Sorts in place semantics: input in (keysA, valsA); result pointer pair returned through outKeys/outVals
// (either A or B, depending on the number of passes).  Returns number of kernels launched, <0 on error.
inline int radix_sort(cudaStream_t st, Workspace &ws, u64 *keysA, u64 *keysB, u64 *valsA, u64 *valsB, u64 n, const Plan &plan,
                      u64 **outKeys, u64 **outVals, int smCount) {
    *outKeys = keysA; *outVals = valsA;
    if (n == 0 || plan.n == 0) return 0;
    if (n > MAX_N || tiles_for(n) > ws.tilesCap || plan.n > MAX_PASSES) return -1;
    int launches = 0;
    cudaMemsetAsync(ws.hist, 0, sizeof(unsigned long long) * MAX_PASSES * RADIX, st);
    int hb = smCount * 8;
    u64 need = (n + BLOCK - 1) / BLOCK;
    if ((u64) hb > need) hb = (int) need;
    hist_kernel<<<hb, BLOCK, plan.n * RADIX * sizeof(u32), st>>>(keysA, valsA, n, plan, ws.hist);
    binscan_kernel<<<1, RADIX, 0, st>>>(ws.hist, plan.n);
    launches += 2;
    u64 tiles = tiles_for(n);
    u64 *ki = keysA, *ko = keysB, *vi = valsA, *vo = valsB;
    for (int p = 0; p < plan.n; p++) {
        cudaMemsetAsync(ws.tileState, 0, tiles * RADIX * sizeof(u32), st);
        cudaMemsetAsync(ws.ticket, 0, sizeof(u32), st);
        if (valsA) onesweep_kernel<true><<<(unsigned) tiles, BLOCK, 0, st>>>(ki, ko, vi, vo, n, plan.p[p], ws.hist + p * RADIX, ws.tileState, ws.ticket);
        else onesweep_kernel<false><<<(unsigned) tiles, BLOCK, 0, st>>>(ki, ko, nullptr, nullptr, n, plan.p[p], ws.hist + p * RADIX, ws.tileState, ws.ticket);
        launches++;
        u64 *t = ki; ki = ko; ko = t;
        t = vi; vi = vo; vo = t;
    }
    *outKeys = ki; *outVals = vi;
    return launches;
}

}   // namespace cdgsort

// ---- exclusive scan of f(i), i in [0,n), u32 -> u64 totals ------------------------------------------
namespace cdgscan {
static constexpr int BLOCK = 256;
static constexpr int ITEMS = 8;
static constexpr int TILE = BLOCK * ITEMS;

// block-wide exclusive scan of one u32 per thread; returns exclusive prefix, total in *total
CDG_D u32 block_excl_scan(u32 v, u32 *total, u32 *shWarp /*[BLOCK/32 + 1]*/) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    u32 incl = v;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) { u32 t = __shfl_up_sync(0xffffffffu, incl, off); if (lane >= off) incl += t; }
    __syncthreads();   // protect shWarp reuse between calls
    if (lane == 31) shWarp[warp] = incl;
    __syncthreads();
    u32 wpre = 0, tot = 0;
#pragma unroll
    for (int w = 0; w < BLOCK / 32; w++) { u32 s = shWarp[w]; if (w < warp) wpre += s; tot += s; }
    *total = tot;
    return wpre + incl - v;
}

// phase 1: per-tile sums of F(i)
template <class F>
__global__ void __launch_bounds__(BLOCK) tile_sum_kernel(F f, u64 n, unsigned long long *__restrict__ tileSums) {
    __shared__ u32 shWarp[BLOCK / 32 + 1];
    u64 base = (u64) blockIdx.x * TILE;
    u32 s = 0;
#pragma unroll
    for (int i = 0; i < ITEMS; i++) { u64 idx = base + (u64) i * BLOCK + threadIdx.x; if (idx < n) s += f(idx); }
    u32 tot;
    block_excl_scan(s, &tot, shWarp);
    if (threadIdx.x == 0) tileSums[blockIdx.x] = tot;
}
// phase 2: single-block exclusive scan of the tile sums (sequential over chunks of BLOCK)
__global__ void __launch_bounds__(1024) tile_scan_kernel(unsigned long long *__restrict__ tileSums, u64 nTiles, unsigned long long *__restrict__ total) {
    __shared__ unsigned long long sh[1024];
    __shared__ unsigned long long carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (u64 base = 0; base < nTiles; base += 1024) {
        u64 idx = base + threadIdx.x;
        unsigned long long v = idx < nTiles ? tileSums[idx] : 0;
        sh[threadIdx.x] = v;
        __syncthreads();
        for (int off = 1; off < 1024; off <<= 1) {
            unsigned long long t = threadIdx.x >= (unsigned) off ? sh[threadIdx.x - off] : 0;
            __syncthreads();
            sh[threadIdx.x] += t;
            __syncthreads();
        }
        unsigned long long c = carry;
        if (idx < nTiles) tileSums[idx] = c + sh[threadIdx.x] - v;
        __syncthreads();
        if (threadIdx.x == 1023) carry = c + sh[1023];
        __syncthreads();
    }
    if (threadIdx.x == 0 && total) *total = carry;
}
// phase 3 is fused into the consumer: consumer kernels call tile_prefix() to get each element's exclusive prefix.
// Element order inside a tile: idx = base + i*BLOCK + threadIdx.x (i-major), so the prefix is computed item by item.
template <class F>
CDG_D void tile_prefix(F f, u64 n, const unsigned long long *__restrict__ tileSums, u32 *shWarp, u64 outPrefix[ITEMS], u32 outVal[ITEMS]) {
    u64 base = (u64) blockIdx.x * TILE;
    u64 running = tileSums[blockIdx.x];
#pragma unroll
    for (int i = 0; i < ITEMS; i++) {
        u64 idx = base + (u64) i * BLOCK + threadIdx.x;
        u32 v = idx < n ? f(idx) : 0;
        u32 tot;
        u32 ex = block_excl_scan(v, &tot, shWarp);
        outPrefix[i] = running + ex;
        outVal[i] = v;
        running += tot;
    }
}
inline u64 tiles_for(u64 n) { return (n + TILE - 1) / TILE; }
}   // namespace cdgscan
