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
static constexpr int BLOCK = 256;                 // histogram kernel block
static constexpr int MAX_PASSES = 16;
// tile state = flag (2 bits) | digit count: one 32-bit word for passes below 2^30 records (the fast case: 64-bit states
// measured 64.7 % -> 53 % of the HBM peak on the 16-byte pass), one 64-bit word above
template <class ST> struct State;
template <> struct State<u32> { static constexpr u32 AGG = 1u << 30, INC = 2u << 30, FLAGS = 3u << 30, VAL = ~(3u << 30); };
template <> struct State<u64> { static constexpr u64 AGG = 1ull << 62, INC = 2ull << 62, FLAGS = 3ull << 62, VAL = ~(3ull << 62); };
static constexpr u64 SMALL_N = (1ull << 30) - 1;  // 30-bit tile-state counters
static constexpr u64 MAX_N = (1ull << 32) - 8192; // 32-bit tile ids and per-block counters
// resident blocks per SM the 256x16 digit pass is compiled for (64 registers, no spills, at 4; measured on B200)
#ifndef CDG_SWEEP_MB_PAIRS
#define CDG_SWEEP_MB_PAIRS 4
#endif
#ifndef CDG_SWEEP_MB_KEYS
#define CDG_SWEEP_MB_KEYS 4
#endif
static constexpr int ROUTE_MAX = 16;              // digits of a routing pass (= ranks)

struct Pass { int word; int shift; u32 mask; };   // word: 0 = key, 1 = val
struct Plan { int n; Pass p[MAX_PASSES]; };

CDG_D u32 digit_of(u64 w, int shift, u32 mask) { return (u32) (w >> shift) & mask; }
// tile-state words are single 32-bit values (flag | count): relaxed device-scope accesses are enough and much cheaper
// than the system-scope accesses `volatile` compiles to
CDG_D u32 ld_state(const u32 *p) { u32 v; asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v; }
CDG_D void st_state(u32 *p, u32 v) { asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" :: "l"(p), "r"(v) : "memory"); }
CDG_D u64 ld_state(const u64 *p) { u64 v; asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory"); return v; }
CDG_D void st_state(u64 *p, u64 v) { asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" :: "l"(p), "l"(v) : "memory"); }

// ---- kernel 1: histograms of every pass in one read of the digit words --------------------------
__global__ void __launch_bounds__(BLOCK) hist_kernel(const u64 *__restrict__ keys, const u64 *__restrict__ vals, u64 n,
                                                     Plan plan, unsigned long long *__restrict__ hist /*[n_pass][RADIX]*/) {
    extern __shared__ u32 sh[];   // [plan.n][RADIX]
    for (int i = threadIdx.x; i < plan.n * RADIX; i += BLOCK) sh[i] = 0;
    __syncthreads();
    bool needKey = false, needVal = false;
    for (int p = 0; p < plan.n; p++) { if (plan.p[p].word == 0) needKey = true; else needVal = true; }
    u64 stride = (u64) gridDim.x * BLOCK;
    if (needKey && !needVal && (((unsigned long long) keys) & 15ull) == 0) {
        // keys only (every caller today): 128-bit loads, two keys per thread and load
        const ulonglong2 *k2 = reinterpret_cast<const ulonglong2 *>(keys);
        const u64 pairs = n >> 1;
        for (u64 i = (u64) blockIdx.x * BLOCK + threadIdx.x; i < pairs; i += stride) {
            const ulonglong2 kk = k2[i];
            for (int p = 0; p < plan.n; p++) {
                atomicAdd(&sh[p * RADIX + digit_of(kk.x, plan.p[p].shift, plan.p[p].mask)], 1u);
                atomicAdd(&sh[p * RADIX + digit_of(kk.y, plan.p[p].shift, plan.p[p].mask)], 1u);
            }
        }
        if ((n & 1) && blockIdx.x == 0 && threadIdx.x == 0)
            for (int p = 0; p < plan.n; p++) atomicAdd(&sh[p * RADIX + digit_of(keys[n - 1], plan.p[p].shift, plan.p[p].mask)], 1u);
    } else {
        for (u64 i = (u64) blockIdx.x * BLOCK + threadIdx.x; i < n; i += stride) {
            u64 k = needKey ? keys[i] : 0, v = needVal ? vals[i] : 0;
            for (int p = 0; p < plan.n; p++) {
                u64 w = plan.p[p].word == 0 ? k : v;
                atomicAdd(&sh[p * RADIX + digit_of(w, plan.p[p].shift, plan.p[p].mask)], 1u);
            }
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
// Digits always come from the key word.  Shared memory: exchange buffer (tile of u64), per-warp digit
// counters (turned into per-warp scatter bases), per-digit `delta` = global base - tile offset.
template <int BLOCK_T, int ITEMS_T, bool HAS_VAL>
struct SweepSmem {
    static constexpr int TILE_T = BLOCK_T * ITEMS_T;
    static constexpr int WARPS_T = BLOCK_T / 32;
    u64 exch[TILE_T];
    u32 warpHist[WARPS_T][RADIX];
    unsigned long long delta[RADIX];
    u32 warpSum[8];
    u32 sTile;
    u64 *routeK[ROUTE_MAX]; u64 *routeV[ROUTE_MAX];   // ROUTE: per-digit output bases (peer-mapped receive buffers)
    u64 routeCap[ROUTE_MAX];                          // ... and how many records each may take (a record beyond is dropped and flagged)
    u8 digitAt[HAS_VAL ? TILE_T : 4];
};

// lanes of the warp that hold the same 8-bit digit: intersection of 8 ballots, 5 instructions per bit
// (MATCH.ANY iterates once per distinct value in the warp, ~30 for random digits; measured 25-30 % slower on B200)
CDG_D u32 digit_peers(u32 d) {
    u32 m = 0xffffffffu;
#pragma unroll
    for (int b = 0; b < RADIX_BITS; b++) {
        const u32 vote = __ballot_sync(0xffffffffu, (d >> b) & 1u);
        const u32 bitmask = (u32) ((i32) (d << (31 - b)) >> 31);   // all ones when bit b of d is set
        m &= ~(vote ^ bitmask);                                    // bit set: keep voters, else keep non-voters (one LOP3)
    }
    return m;
}

// ROUTE: the pass is the send side of an all-to-all.  Digit d (< ROUTE_MAX) is a destination rank and its records go, in
// stable order, to routeK[d] / routeV[d] — the rank's segment of peer d's receive buffer, mapped into this process over
// NVLink (CUDA IPC) — instead of one local output array: the partition and the transfer are one kernel, the tile's run
// of each digit is written as one coalesced burst into peer memory.  binBase must be all zero (offsets are per segment).
template <int BLOCK_T, int ITEMS_T, bool HAS_VAL, bool ROUTE = false, class ST = u32>
__global__ void __launch_bounds__(BLOCK_T, (BLOCK_T == 256 && ITEMS_T == 16) ? (HAS_VAL ? CDG_SWEEP_MB_PAIRS : CDG_SWEEP_MB_KEYS) : (BLOCK_T == 512 && ITEMS_T == 16) ? 2 : 1) onesweep_kernel(const u64 *__restrict__ keysIn, u64 *__restrict__ keysOut,
                                                           const u64 *__restrict__ valsIn, u64 *__restrict__ valsOut, u64 n,
                                                           Pass pass, const unsigned long long *__restrict__ binBase,
                                                           ST *tileState /*[tiles][RADIX]*/, u32 *ticket,
                                                           u64 *const *__restrict__ routeK = nullptr, u64 *const *__restrict__ routeV = nullptr,
                                                           const u64 *__restrict__ routeCap = nullptr /*ROUTE: records destination d can take*/, u32 *routeOverflow = nullptr) {
    constexpr int TILE_T = BLOCK_T * ITEMS_T, WARPS_T = BLOCK_T / 32;
    extern __shared__ __align__(16) unsigned char sweepRaw[];
    SweepSmem<BLOCK_T, ITEMS_T, HAS_VAL> &S = *reinterpret_cast<SweepSmem<BLOCK_T, ITEMS_T, HAS_VAL> *>(sweepRaw);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (ROUTE && tid < ROUTE_MAX) { S.routeK[tid] = routeK[tid]; if (HAS_VAL) S.routeV[tid] = routeV[tid]; S.routeCap[tid] = routeCap ? routeCap[tid] : ~0ull; }
    if (tid == 0) S.sTile = atomicAdd(ticket, 1u);
    for (int i = tid; i < WARPS_T * RADIX; i += BLOCK_T) (&S.warpHist[0][0])[i] = 0;
    __syncthreads();
    const u32 tile = S.sTile;
    const u64 tileBase = (u64) tile * TILE_T;
    const u32 tileValid = (u32) (n - tileBase < (u64) TILE_T ? n - tileBase : (u64) TILE_T);
    const bool full = tileValid == TILE_T;

    // load (warp-striped: warp w owns records [w*ITEMS*32, (w+1)*ITEMS*32) of the tile, item-major)
    u64 key[ITEMS_T];
    u32 dig[ITEMS_T];    // digit in the low byte, later rank << 8 | digit
    const u32 warpBase = warp * ITEMS_T * 32;
    const u64 *kin = keysIn + tileBase + warpBase + lane;
    if (full) {
#pragma unroll
        for (int i = 0; i < ITEMS_T; i++) key[i] = kin[i * 32];
    } else {
#pragma unroll
        for (int i = 0; i < ITEMS_T; i++) key[i] = warpBase + i * 32 + lane < tileValid ? kin[i * 32] : ~0ull;   // padding -> last bin, ranks last
    }
    // stable rank inside the warp's chunk: all peer masks first (independent), then the chain through the warp's counters
    u32 peers[ITEMS_T];
#pragma unroll
    for (int i = 0; i < ITEMS_T; i++) { dig[i] = digit_of(key[i], pass.shift, pass.mask); peers[i] = digit_peers(dig[i]); }
    u32 *myHist = S.warpHist[warp];
    const u32 ltMask = (1u << lane) - 1;
#pragma unroll
    for (int i = 0; i < ITEMS_T; i++) {
        const u32 d = dig[i];
        const u32 pre = myHist[d];
        const u32 lower = peers[i] & ltMask;
        __syncwarp();
        if (lower == 0) myHist[d] = pre + __popc(peers[i]);   // lowest lane of the peer group publishes
        __syncwarp();
        dig[i] = d | ((pre + __popc(lower)) << 8);
    }
    __syncthreads();
    // thread d: scan digit d over the warps, publish the tile aggregate, look back
    u32 run = 0, incl = 0;
    if (tid < RADIX) {
        const int d = tid;
#pragma unroll
        for (int w = 0; w < WARPS_T; w++) run += S.warpHist[w][d];
        ST *state = tileState;
        u64 excl = 0;
        if (tile == 0) {
            st_state(state + d, (ST) (State<ST>::INC | run));
        } else {
            st_state(state + (u64) tile * RADIX + d, (ST) (State<ST>::AGG | run));
            // decoupled look-back, LOOK predecessors per round trip (independent loads), consumed nearest first
            constexpr int LOOK = 4;
            i64 p = (i64) tile - 1;
            bool done = false;
            while (!done) {
                ST sv[LOOK];
#pragma unroll
                for (int j = 0; j < LOOK; j++) sv[j] = (p - j) >= 0 ? ld_state(state + (u64) (p - j) * RADIX + d) : State<ST>::INC /* value 0: nothing before tile 0 */;
                int used = 0;
#pragma unroll
                for (int j = 0; j < LOOK; j++) {
                    if (!done && used == j) {
                        const ST f = sv[j] & State<ST>::FLAGS;
                        if (f != 0) { excl += sv[j] & State<ST>::VAL; used = j + 1; if (f == State<ST>::INC) done = true; }
                    }
                }
                p -= used;   // first predecessor that had not published yet is polled again
            }
            st_state(state + (u64) tile * RADIX + d, (ST) (State<ST>::INC | (ST) (excl + run)));
        }
        // exclusive scan of the tile's digit counts over d (8 warps of 32 digits)
        incl = run;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) { const u32 t = __shfl_up_sync(0xffffffffu, incl, off); if (lane >= off) incl += t; }
        if (lane == 31) S.warpSum[warp] = incl;
        S.delta[d] = binBase[d] + excl;   // completed below
    }
    __syncthreads();
    if (tid < RADIX) {
        const int d = tid;
        u32 wpre = 0;
#pragma unroll
        for (int w = 0; w < 8; w++) if (w < warp) wpre += S.warpSum[w];
        const u32 tileOff = wpre + incl - run;   // first position of digit d in the sorted tile
        S.delta[d] -= tileOff;                    // out index = delta[d] + position in tile
        u32 base = tileOff;                       // per-warp scatter base = tile offset + counts of the warps before
#pragma unroll
        for (int w = 0; w < WARPS_T; w++) { const u32 c = S.warpHist[w][d]; S.warpHist[w][d] = base; base += c; }
    }
    __syncthreads();
    // scatter into tile order through shared memory, then coalesced runs to global
    u32 pos[ITEMS_T];
#pragma unroll
    for (int i = 0; i < ITEMS_T; i++) {
        const u32 d = dig[i] & 0xFFu;
        pos[i] = (dig[i] >> 8) + myHist[d];
        S.exch[pos[i]] = key[i];
        if (HAS_VAL) S.digitAt[pos[i]] = (u8) d;
    }
    u64 v[HAS_VAL ? ITEMS_T : 1];
    if (HAS_VAL) {   // values are requested now and arrive while the keys are written out
        const u64 *vin = valsIn + tileBase + warpBase + lane;
#pragma unroll
        for (int i = 0; i < ITEMS_T; i++) v[i] = (full || warpBase + i * 32 + lane < tileValid) ? vin[i * 32] : 0;
    }
    __syncthreads();
    if (full) {
#pragma unroll
        for (int i = 0; i < ITEMS_T; i++) {
            const u32 j = tid + i * BLOCK_T;
            const u64 k = S.exch[j];
            const u32 dg = digit_of(k, pass.shift, pass.mask);
            if (ROUTE) { const u64 idx = S.delta[dg] + j; if (idx < S.routeCap[dg & (ROUTE_MAX - 1)]) S.routeK[dg & (ROUTE_MAX - 1)][idx] = k; else *routeOverflow = 1u; }
            else keysOut[S.delta[dg] + j] = k;
        }
    } else {
#pragma unroll
        for (int i = 0; i < ITEMS_T; i++) {
            const u32 j = tid + i * BLOCK_T;
            if (j < tileValid) {
                const u64 k = S.exch[j];
                const u32 dg = digit_of(k, pass.shift, pass.mask);
                if (ROUTE) { const u64 idx = S.delta[dg] + j; if (idx < S.routeCap[dg & (ROUTE_MAX - 1)]) S.routeK[dg & (ROUTE_MAX - 1)][idx] = k; else *routeOverflow = 1u; }
                else keysOut[S.delta[dg] + j] = k;
            }
        }
    }
    if (HAS_VAL) {
        __syncthreads();
#pragma unroll
        for (int i = 0; i < ITEMS_T; i++) S.exch[pos[i]] = v[i];
        __syncthreads();
#pragma unroll
        for (int i = 0; i < ITEMS_T; i++) {
            const u32 j = tid + i * BLOCK_T;
            if (full || j < tileValid) {
                const u32 dg = S.digitAt[j];
                if (ROUTE) { const u64 idx = S.delta[dg] + j; if (idx < S.routeCap[dg & (ROUTE_MAX - 1)]) S.routeV[dg & (ROUTE_MAX - 1)][idx] = S.exch[j]; }
                else valsOut[S.delta[dg] + j] = S.exch[j];
            }
        }
    }
}

// ---- kernel 3b: persistent, software-pipelined digit pass --------------------------------------------------
// Same algorithm; block b of a grid that is fully resident processes tiles b, b+G, b+2G, ... (every tile's predecessor
// belongs to a resident block that reaches it in the same round, so the chained scan cannot stall on an unscheduled
// tile).  The keys of the block's next tile are requested right after the current tile has been scattered into shared
// memory and the look-back is moved behind the scatter, so the global round trips (key loads, predecessor states)
// overlap the shared-memory work instead of serialising the tile.
template <int BLOCK_T, int ITEMS_T, bool HAS_VAL>
__global__ void __launch_bounds__(BLOCK_T) onesweep_persistent_kernel(const u64 *__restrict__ keysIn, u64 *__restrict__ keysOut,
                                                                      const u64 *__restrict__ valsIn, u64 *__restrict__ valsOut, u64 n,
                                                                      Pass pass, const unsigned long long *__restrict__ binBase,
                                                                      u32 *tileState /*[tiles][RADIX]*/, u32 *ticket) {
    constexpr int TILE_T = BLOCK_T * ITEMS_T, WARPS_T = BLOCK_T / 32;
    extern __shared__ __align__(16) unsigned char sweepRaw[];
    SweepSmem<BLOCK_T, ITEMS_T, HAS_VAL> &S = *reinterpret_cast<SweepSmem<BLOCK_T, ITEMS_T, HAS_VAL> *>(sweepRaw);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const u32 numTiles = (u32) ((n + TILE_T - 1) / TILE_T);
    const u32 warpBase = warp * ITEMS_T * 32;

    (void) ticket;
    u32 tile = blockIdx.x;
    u64 key[ITEMS_T];
    if (tile < numTiles) {
        const u64 tb = (u64) tile * TILE_T;
        const u32 tv = (u32) (n - tb < (u64) TILE_T ? n - tb : (u64) TILE_T);
#pragma unroll
        for (int i = 0; i < ITEMS_T; i++) { const u32 local = warpBase + i * 32 + lane; key[i] = local < tv ? keysIn[tb + local] : ~0ull; }
    }
    while (tile < numTiles) {
        const u64 tileBase = (u64) tile * TILE_T;
        const u32 tileValid = (u32) (n - tileBase < (u64) TILE_T ? n - tileBase : (u64) TILE_T);
        for (int i = tid; i < WARPS_T * RADIX; i += BLOCK_T) (&S.warpHist[0][0])[i] = 0;
        __syncthreads();
        // rank
        u32 rank[ITEMS_T];
        {
            u32 peers[ITEMS_T];
#pragma unroll
            for (int i = 0; i < ITEMS_T; i++) {
                const u32 d = digit_of(key[i], pass.shift, pass.mask);
                u32 m = 0xffffffffu;
#pragma unroll
                for (int b = 0; b < RADIX_BITS; b++) {
                    const u32 vote = __ballot_sync(0xffffffffu, (d >> b) & 1u);
                    m &= ((d >> b) & 1u) ? vote : ~vote;
                }
                peers[i] = m;
            }
#pragma unroll
            for (int i = 0; i < ITEMS_T; i++) {
                const u32 d = digit_of(key[i], pass.shift, pass.mask);
                const u32 pre = S.warpHist[warp][d];
                const u32 lower = peers[i] & ((1u << lane) - 1);
                __syncwarp();
                if (lower == 0) S.warpHist[warp][d] = pre + __popc(peers[i]);
                __syncwarp();
                rank[i] = pre + __popc(lower);
            }
        }
        __syncthreads();
        // tile aggregate out as early as possible; tile-local offsets
        u32 run = 0, incl = 0;
        if (tid < RADIX) {
            const int d = tid;
#pragma unroll
            for (int w = 0; w < WARPS_T; w++) run += S.warpHist[w][d];
            st_state(tileState + (u64) tile * RADIX + d, (tile == 0 ? State<u32>::INC : State<u32>::AGG) | run);
            incl = run;
#pragma unroll
            for (int off = 1; off < 32; off <<= 1) { const u32 t = __shfl_up_sync(0xffffffffu, incl, off); if (lane >= off) incl += t; }
            if (lane == 31) S.warpSum[warp] = incl;
        }
        __syncthreads();
        u32 tileOff = 0;
        if (tid < RADIX) {
            const int d = tid;
            u32 wpre = 0;
#pragma unroll
            for (int w = 0; w < 8; w++) if (w < warp) wpre += S.warpSum[w];
            tileOff = wpre + incl - run;
            u32 base = tileOff;
#pragma unroll
            for (int w = 0; w < WARPS_T; w++) { const u32 c = S.warpHist[w][d]; S.warpHist[w][d] = base; base += c; }
        }
        __syncthreads();
        // scatter into tile order
#pragma unroll
        for (int i = 0; i < ITEMS_T; i++) {
            const u32 d = digit_of(key[i], pass.shift, pass.mask);
            rank[i] += S.warpHist[warp][d];
            S.exch[rank[i]] = key[i];
            if (HAS_VAL) S.digitAt[rank[i]] = (u8) d;
        }
        // requests that overlap the look-back: values of this tile, keys of the next tile, the ticket after that
        u64 v[HAS_VAL ? ITEMS_T : 1];
        if (HAS_VAL) {
#pragma unroll
            for (int i = 0; i < ITEMS_T; i++) { const u32 local = warpBase + i * 32 + lane; v[i] = local < tileValid ? valsIn[tileBase + local] : 0; }
        }
        const u32 nextTile = tile + gridDim.x;
        if (nextTile < numTiles) {
            const u64 tb = (u64) nextTile * TILE_T;
            const u32 tv = (u32) (n - tb < (u64) TILE_T ? n - tb : (u64) TILE_T);
#pragma unroll
            for (int i = 0; i < ITEMS_T; i++) { const u32 local = warpBase + i * 32 + lane; key[i] = local < tv ? keysIn[tb + local] : ~0ull; }
        }
        // look-back (predecessors published their aggregates a scatter phase ago)
        if (tid < RADIX) {
            const int d = tid;
            u32 *state = tileState;
            u64 excl = 0;
            if (tile != 0) {
                constexpr int LOOK = 8;
                i64 p = (i64) tile - 1;
                bool done = false;
                while (!done) {
                    u32 sv[LOOK];
#pragma unroll
                    for (int j = 0; j < LOOK; j++) sv[j] = (p - j) >= 0 ? ld_state(state + (u64) (p - j) * RADIX + d) : 0x80000000u;
                    int used = 0;
#pragma unroll
                    for (int j = 0; j < LOOK; j++) {
                        if (!done && used == j) {
                            const u32 f = sv[j] & State<u32>::FLAGS;
                            if (f != 0) { excl += sv[j] & State<u32>::VAL; used = j + 1; if (f == State<u32>::INC) done = true; }
                        }
                    }
                    p -= used;
                }
                st_state(state + (u64) tile * RADIX + d, State<u32>::INC | (u32) (excl + run));
            }
            S.delta[d] = binBase[d] + excl - tileOff;
        }
        __syncthreads();
#pragma unroll
        for (int i = 0; i < ITEMS_T; i++) {
            const u32 j = tid + i * BLOCK_T;
            if (j < tileValid) {
                const u64 k = S.exch[j];
                keysOut[S.delta[digit_of(k, pass.shift, pass.mask)] + j] = k;
            }
        }
        if (HAS_VAL) {
            __syncthreads();
#pragma unroll
            for (int i = 0; i < ITEMS_T; i++) S.exch[rank[i]] = v[i];
            __syncthreads();
#pragma unroll
            for (int i = 0; i < ITEMS_T; i++) {
                const u32 j = tid + i * BLOCK_T;
                if (j < tileValid) valsOut[S.delta[S.digitAt[j]] + j] = S.exch[j];
            }
        }
        __syncthreads();   // exch / delta / warpHist are reused by the next tile
        tile = nextTile;
    }
}

template <int BLOCK_T, int ITEMS_T, bool HAS_VAL>
inline void launch_persistent(cudaStream_t st, const u64 *ki, u64 *ko, const u64 *vi, u64 *vo, u64 n, const Pass &p,
                              const unsigned long long *binBase, u32 *tileState, u32 *ticket, int smCount) {
    typedef SweepSmem<BLOCK_T, ITEMS_T, HAS_VAL> SM;
    static int blocksPerSm = 0;
    if (!blocksPerSm) {
        cudaFuncSetAttribute(onesweep_persistent_kernel<BLOCK_T, ITEMS_T, HAS_VAL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) sizeof(SM));
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocksPerSm, onesweep_persistent_kernel<BLOCK_T, ITEMS_T, HAS_VAL>, BLOCK_T, sizeof(SM));
        if (blocksPerSm < 1) blocksPerSm = 1;
    }
    const u64 tiles = (n + SM::TILE_T - 1) / SM::TILE_T;
    u64 grid = (u64) smCount * blocksPerSm;
    if (grid > tiles) grid = tiles;
    onesweep_persistent_kernel<BLOCK_T, ITEMS_T, HAS_VAL><<<(unsigned) grid, BLOCK_T, sizeof(SM), st>>>(ki, ko, vi, vo, n, p, binBase, tileState, ticket);
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
    u64 tilesCap;               // in tiles of MIN_TILE records
};
static constexpr int MIN_TILE = 2048;
inline u64 tiles_for(u64 n) { return (n + MIN_TILE - 1) / MIN_TILE; }

template <int BLOCK_T, int ITEMS_T, bool HAS_VAL>
inline void launch_pass(cudaStream_t st, const u64 *ki, u64 *ko, const u64 *vi, u64 *vo, u64 n, const Pass &p, const unsigned long long *binBase,
                        u32 *tileState, u32 *ticket) {
    typedef SweepSmem<BLOCK_T, ITEMS_T, HAS_VAL> SM;
    static bool attrSet = false;
    if (!attrSet) { cudaFuncSetAttribute(onesweep_kernel<BLOCK_T, ITEMS_T, HAS_VAL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) sizeof(SM)); attrSet = true; }
    const u64 tiles = (n + SM::TILE_T - 1) / SM::TILE_T;
    onesweep_kernel<BLOCK_T, ITEMS_T, HAS_VAL><<<(unsigned) tiles, BLOCK_T, sizeof(SM), st>>>(ki, ko, vi, vo, n, p, binBase, tileState, ticket);
}
// same pass with 64-bit tile states: 2^30 records and more (the workspace holds twice the state bytes then)
template <int BLOCK_T, int ITEMS_T, bool HAS_VAL>
inline void launch_pass_wide(cudaStream_t st, const u64 *ki, u64 *ko, const u64 *vi, u64 *vo, u64 n, const Pass &p, const unsigned long long *binBase,
                             u32 *tileState, u32 *ticket) {
    typedef SweepSmem<BLOCK_T, ITEMS_T, HAS_VAL> SM;
    static bool attrSet = false;
    if (!attrSet) { cudaFuncSetAttribute(onesweep_kernel<BLOCK_T, ITEMS_T, HAS_VAL, false, u64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) sizeof(SM)); attrSet = true; }
    const u64 tiles = (n + SM::TILE_T - 1) / SM::TILE_T;
    onesweep_kernel<BLOCK_T, ITEMS_T, HAS_VAL, false, u64><<<(unsigned) tiles, BLOCK_T, sizeof(SM), st>>>(ki, ko, vi, vo, n, p, binBase, (u64 *) tileState, ticket);
}

// routing pass (send side of the all-to-all): digits = destination ranks, outputs = peer-mapped segment bases (device tables)
// base: per-destination index of the first record of this launch (all zero, or the running counts of a chunked exchange)
inline void launch_route_pass(cudaStream_t st, const u64 *ki, const u64 *vi, u64 n, const Pass &p, const unsigned long long *zeroBase,
                              u32 *tileState, u32 *ticket, u64 *const *routeK, u64 *const *routeV, const u64 *routeCap = nullptr, u32 *routeOverflow = nullptr) {
    typedef SweepSmem<256, 16, true> SM;
    static bool attrSet = false;
    if (!attrSet) { cudaFuncSetAttribute(onesweep_kernel<256, 16, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) sizeof(SM)); attrSet = true; }
    const u64 tiles = (n + SM::TILE_T - 1) / SM::TILE_T;
    onesweep_kernel<256, 16, true, true><<<(unsigned) tiles, 256, sizeof(SM), st>>>(ki, nullptr, vi, nullptr, n, p, zeroBase, tileState, ticket, routeK, routeV, routeCap, routeOverflow);
}
// after a routing pass over n records: base[d] += records of digit d (the inclusive state of the pass's last tile)
__global__ void route_advance_kernel(const u32 *__restrict__ tileState, u64 tiles, unsigned long long *__restrict__ base) {
    const int d = threadIdx.x;
    if (d < ROUTE_MAX && tiles) base[d] += (unsigned long long) (tileState[(tiles - 1) * RADIX + d] & State<u32>::VAL);
}
inline u64 route_tiles(u64 n) { return (n + SweepSmem<256, 16, true>::TILE_T - 1) / SweepSmem<256, 16, true>::TILE_T;
}

// variant: 0 = 256 threads x 16 keys, 1 = 256 x 8, 2 = 512 x 8, 3 = 384... (tuning knob; default chosen by measurements)
inline void launch_variant(int variant, bool hasVal, cudaStream_t st, const u64 *ki, u64 *ko, const u64 *vi, u64 *vo, u64 n, const Pass &p,
                           const unsigned long long *binBase, u32 *tileState, u32 *ticket, int smCount) {
    switch (variant * 2 + (hasVal ? 1 : 0)) {
        case 8: launch_persistent<256, 16, false>(st, ki, ko, vi, vo, n, p, binBase, tileState, ticket, smCount); break;
        case 9: launch_persistent<256, 16, true>(st, ki, ko, vi, vo, n, p, binBase, tileState, ticket, smCount); break;
        case 10: launch_persistent<256, 8, false>(st, ki, ko, vi, vo, n, p, binBase, tileState, ticket, smCount); break;
        case 11: launch_persistent<256, 8, true>(st, ki, ko, vi, vo, n, p, binBase, tileState, ticket, smCount); break;
        case 12: launch_persistent<512, 8, false>(st, ki, ko, vi, vo, n, p, binBase, tileState, ticket, smCount); break;
        case 13: launch_persistent<512, 8, true>(st, ki, ko, vi, vo, n, p, binBase, tileState, ticket, smCount); break;
        case 14: launch_persistent<384, 8, false>(st, ki, ko, vi, vo, n, p, binBase, tileState, ticket, smCount); break;
        case 15: launch_persistent<384, 8, true>(st, ki, ko, vi, vo, n, p, binBase, tileState, ticket, smCount); break;
        case 16: launch_pass<512, 16, false>(st, ki, ko, vi, vo, n, p, binBase, tileState, ticket); break;   // 8192-record tiles
        case 17: launch_pass<512, 16, true>(st, ki, ko, vi, vo, n, p, binBase, tileState, ticket); break;
        case 0: launch_pass<256, 16, false>(st, ki, ko, vi, vo, n, p, binBase, tileState, ticket); break;
        case 1: launch_pass<256, 16, true>(st, ki, ko, vi, vo, n, p, binBase, tileState, ticket); break;
        case 2: launch_pass<256, 8, false>(st, ki, ko, vi, vo, n, p, binBase, tileState, ticket); break;
        case 3: launch_pass<256, 8, true>(st, ki, ko, vi, vo, n, p, binBase, tileState, ticket); break;
        case 4: launch_pass<512, 8, false>(st, ki, ko, vi, vo, n, p, binBase, tileState, ticket); break;
        case 5: launch_pass<512, 8, true>(st, ki, ko, vi, vo, n, p, binBase, tileState, ticket); break;
        case 6: launch_pass<512, 4, false>(st, ki, ko, vi, vo, n, p, binBase, tileState, ticket); break;
        default: launch_pass<512, 4, true>(st, ki, ko, vi, vo, n, p, binBase, tileState, ticket); break;
    }
}
// records per tile of a (non-persistent) variant; 0 = unknown (persistent variants: clear the state for MIN_TILE-sized tiles)
inline u64 variant_tile(int variant, bool wide, bool hasVal) {
    if (wide) return hasVal ? 4096 : 8192;
    switch (variant) { case 0: return 4096; case 1: return 2048; case 2: return 4096; case 3: return 2048; case 8: return 8192; default: return 0; }
}
inline int &default_variant(bool hasVal) { static int v[2] = {8, 0}; return v[hasVal ? 1 : 0]; }   // measured on B200 (profiles/sortbench_tiles.py): keys only 512 x 16 (40.2 % of the HBM peak vs 38.3 %), pairs 256 x 16 (64.7 %)

// Input in (keysA, valsA); result pointer pair returned through outKeys/outVals (A or B, depending on the number
// of passes).  Returns number of kernels launched, <0 on error.  passMs (optional): per-pass event timing sink.
inline int radix_sort(cudaStream_t st, Workspace &ws, u64 *keysA, u64 *keysB, u64 *valsA, u64 *valsB, u64 n, const Plan &plan,
                      u64 **outKeys, u64 **outVals, int smCount, int variant = -1, cudaEvent_t *passEvents = nullptr /*[2*plan.n]*/) {
    *outKeys = keysA; *outVals = valsA;
    if (n == 0 || plan.n == 0) return 0;
    const bool wide = n > SMALL_N;   // 64-bit tile states: only the two default tile shapes are built for them
    if (n > MAX_N || tiles_for(n) * (wide ? 2 : 1) > ws.tilesCap || plan.n > MAX_PASSES) return -1;
    for (int p = 0; p < plan.n; p++) if (plan.p[p].word != 0) return -1;   // digits come from the key word
    if (variant < 0 || wide) variant = default_variant(valsA != nullptr);
    int launches = 0;
    cudaMemsetAsync(ws.hist, 0, sizeof(unsigned long long) * MAX_PASSES * RADIX, st);
    int hb = smCount * 8;
    u64 need = (n + BLOCK - 1) / BLOCK;
    if ((u64) hb > need) hb = (int) need;
    hist_kernel<<<hb, BLOCK, plan.n * RADIX * sizeof(u32), st>>>(keysA, valsA, n, plan, ws.hist);
    binscan_kernel<<<1, RADIX, 0, st>>>(ws.hist, plan.n);
    launches += 2;
    u64 *ki = keysA, *ko = keysB, *vi = valsA, *vo = valsB;
    // tile states to clear per pass: those of the tiles the pass really has (4096- or 8192-record tiles: a half or a quarter of the
    // MIN_TILE-based bound, 1.2 GB less memset traffic per 10 M-read iteration)
    const u64 vt = variant_tile(variant, wide, valsA != nullptr);
    const u64 stateTiles = vt ? (n + vt - 1) / vt : tiles_for(n);
    for (int p = 0; p < plan.n; p++) {
        cudaMemsetAsync(ws.tileState, 0, stateTiles * RADIX * sizeof(u32) * (wide ? 2 : 1), st);
        cudaMemsetAsync(ws.ticket, 0, sizeof(u32), st);
        if (passEvents) cudaEventRecord(passEvents[2 * p], st);
        if (!wide) launch_variant(variant, valsA != nullptr, st, ki, ko, vi, vo, n, plan.p[p], ws.hist + p * RADIX, ws.tileState, ws.ticket, smCount);
        else if (valsA) launch_pass_wide<256, 16, true>(st, ki, ko, vi, vo, n, plan.p[p], ws.hist + p * RADIX, ws.tileState, ws.ticket);
        else launch_pass_wide<512, 16, false>(st, ki, ko, vi, vo, n, plan.p[p], ws.hist + p * RADIX, ws.tileState, ws.ticket);
        if (passEvents) cudaEventRecord(passEvents[2 * p + 1], st);
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
// phase 2: single-block exclusive scan of the tile sums: 4096 sums per round staged through shared memory (coalesced global
// accesses), 4 consecutive sums per thread, warp-shuffle scans
__global__ void __launch_bounds__(1024) tile_scan_kernel(unsigned long long *__restrict__ tileSums, u64 nTiles, unsigned long long *__restrict__ total) {
    constexpr int PER = 4;
    __shared__ unsigned long long stage[1024 * PER];
    __shared__ unsigned long long shWarp[32];
    __shared__ unsigned long long carry;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (u64 base = 0; base < nTiles; base += 1024ull * PER) {
#pragma unroll
        for (int j = 0; j < PER; j++) { const u64 i = base + (u64) j * 1024 + threadIdx.x; stage[j * 1024 + threadIdx.x] = i < nTiles ? tileSums[i] : 0ull; }
        __syncthreads();
        unsigned long long v[PER], sum = 0;
#pragma unroll
        for (int j = 0; j < PER; j++) { v[j] = stage[threadIdx.x * PER + j]; sum += v[j]; }
        unsigned long long incl = sum;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const unsigned long long x = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += x; }
        if (lane == 31) shWarp[warp] = incl;
        __syncthreads();
        if (warp == 0) {
            unsigned long long w = shWarp[lane], wi = w;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const unsigned long long x = __shfl_up_sync(0xffffffffu, wi, o); if (lane >= o) wi += x; }
            shWarp[lane] = wi - w;   // exclusive over the warps
        }
        __syncthreads();
        unsigned long long run = carry + shWarp[warp] + incl - sum;
#pragma unroll
        for (int j = 0; j < PER; j++) { stage[threadIdx.x * PER + j] = run; run += v[j]; }
        __syncthreads();
        if (threadIdx.x == 1023) carry = run;
#pragma unroll
        for (int j = 0; j < PER; j++) { const u64 i = base + (u64) j * 1024 + threadIdx.x; if (i < nTiles) tileSums[i] = stage[j * 1024 + threadIdx.x]; }
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
