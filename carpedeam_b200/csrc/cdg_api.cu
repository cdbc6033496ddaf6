// libcarpedeam_gpu.so — C ABI (include/carpedeam_gpu.h) over the CUDA kernels.  sm_100a only.
#include <cuda_runtime.h>
#include <stdarg.h>
#include <algorithm>
#include <stdlib.h>
#include <string.h>
#include <string>
#include <thread>
#include <vector>

#include "../../include/carpedeam_gpu.h"
#include "cdg_align.cuh"
#include "cdg_common.cuh"
#include "cdg_cycle.cuh"
#include "cdg_export.cuh"
#include "cdg_ext.cuh"
#include "cdg_hamming.cuh"
#include "cdg_kmer.cuh"
#include "cdg_merge.cuh"
#include "cdg_pairlik.cuh"
#include "cdg_sort.cuh"

namespace {

struct DevBuf {
    void *p = nullptr; size_t cap = 0;
    cudaError_t reserve(size_t bytes) {
        if (bytes <= cap) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr; cap = 0;
        size_t want = bytes + (bytes >> 4) + 256;
        cudaError_t e = cudaMalloc(&p, want);
        if (e == cudaSuccess) cap = want;
        return e;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
    template <class T> T *as() const { return (T *) p; }
};
struct HostBuf {   // pinned
    void *p = nullptr; size_t cap = 0;
    cudaError_t reserve(size_t bytes) {
        if (bytes <= cap) return cudaSuccess;
        if (p) cudaFreeHost(p);
        p = nullptr; cap = 0;
        size_t want = bytes + (bytes >> 4) + 256;
        cudaError_t e = cudaMallocHost(&p, want);
        if (e == cudaSuccess) cap = want;
        return e;
    }
    void release() { if (p) cudaFreeHost(p); p = nullptr; cap = 0; }
    template <class T> T *as() const { return (T *) p; }
};

enum { EV_START, EV_EXTRACT, EV_SORT1, EV_GROUP, EV_SORT2, EV_REDUCE, EV_RS0, EV_RS1, EV_CO0, EV_CO1, EV_N };

}   // namespace

static constexpr int CDG_MAX_WORLD = 16;
static constexpr int CDG_UPLOAD_CHUNKS = 8;   // pieces of an asynchronous upload (copy of piece c+1 overlaps pack + extraction of piece c)
static constexpr int CDG_IPC_BUFFERS = 5;   // recvK, recvV, recvIdK, recvIdV, recv2
static constexpr int CDG_IPC_HANDLE_BYTES = 64 + 8;   // cudaIpcMemHandle_t + byte offset of the buffer inside its allocation
static constexpr u64 CDG_HEAD_RUN = 4096;   // records of the next rank's head run appended before the reduction
struct KmPlanFwd {
    bool wide = false;   // grouped records need more than 64 bits: two-word records (single GPU only)
    u64 totalSlots = 0;  // k-mer slots of the whole set (every rank computes the same number)
    int idBits, diagBits; u32 world, rank; u32 posLo, posHi; u64 slotBase, slots; u32 extraCap; u32 hashLo, hashHi; u32 repChunk; u64 qLo, nQ;
};
struct cdg_ctx {
    int device = 0, smCount = 148;
    cudaStream_t st = nullptr, copySt = nullptr;   // copySt: device-to-host exports that overlap the next stage
    cudaEvent_t evReady = nullptr, evCopied = nullptr; bool copyPending = false;
    std::string err;
    cudaEvent_t ev[EV_N];
    cudaEvent_t passEv[2 * cdgsort::MAX_PASSES]; int passEvN = 0;   // per-pass timing of sort 1 (roofline of the dominant kernel)
    bool evValid[EV_N];
    u64 launches = 0;
    u32 wideThreshold = 65535;   // test knob CDG_WIDE_THRESHOLD: route queries with more counted targets to the generic kernel
    // resident sequence set
    u64 n = 0, dataBytes = 0, totalBases = 0, totalWords = 0; u32 maxLen = 0;
    DevBuf dData, dOff, dLen, dExt, dWordOff, dWords, dXmask, dFlags, dLikOff, dMeta;
    // scan / sort scratch
    DevBuf dTileSums, dScalars, dHist, dTileState;
    // kmermatcher
    DevBuf dSlotOff, dKeysA, dKeysB, dValsA, dValsB, dIdKeysA, dIdKeysB, dIdValsA, dIdValsB, dKey2A, dKey2B;
    DevBuf dLongList, dScratchOff, dScratch, dScratchKeep, dShortScratch, dShortKeep, dOrdKA, dOrdKB, dOrdVA, dOrdVB, dRankOf, dSlotRank;
    DevBuf dHitRep, dHitTgt, dHitScore, dHitDiag;
    HostBuf hHitRep, hHitTgt, hHitScore, hHitDiag, hHitDiag16;
    u64 nHits = 0, nRecords = 0, nGrouped = 0; bool hitsValid = false;
    u32 sort1Passes = 0, sort2Passes = 0; u64 sort1Records = 0;
    u64 *sortedMainK = nullptr, *sortedMainV = nullptr, *sortedIdK = nullptr, *sortedIdV = nullptr, *sortedKey2 = nullptr;
    KmPlanFwd plan; u64 qLo = 0, nQ = 0; u64 mgRecv = 0, mgR2 = 0; u64 nIdent = 0;
    // chunked upload: the DB bytes arrive in chunks on copySt; the chunks that are not packed yet (sequence ranges + events)
    struct UploadChunk { u64 seqLo, seqHi; cudaEvent_t ev; };
    std::vector<UploadChunk> pending; cudaEvent_t chunkEv[CDG_UPLOAD_CHUNKS]; cudaEvent_t evUploadStart = nullptr;
    const u64 *ordV = nullptr;   // sequence index by (length desc, id asc) rank (km_prepare)
    u64 *idSrcK = nullptr, *idSrcV = nullptr, *idAltK = nullptr, *idAltV = nullptr;   // identity records to sort (nIdent) and their ping-pong partner
    DevBuf dRecvIdK, dRecvIdV, dRouteTab;
    // peer-mapped receive buffers (CUDA IPC): [which][peer] base pointers as opened in this process, and the handle they came from
    void *peerPtr[CDG_IPC_BUFFERS][CDG_MAX_WORLD] = {}; void *peerBase[CDG_IPC_BUFFERS][CDG_MAX_WORLD] = {};
    unsigned char peerHandle[CDG_IPC_BUFFERS][CDG_MAX_WORLD][CDG_IPC_HANDLE_BYTES] = {};
    u64 mgSlice = 0; u64 *mgMainK = nullptr, *mgMainV = nullptr, *mgIdK = nullptr, *mgIdV = nullptr;   // phase-1 arrays awaiting the scatter
    // lines / alignments
    DevBuf dLineOff, dLineQ, dLineT, dLinePref, dLineDiag, dKeep, dScore, dBits, dEvalue, dSeqId, dQS, dQE, dDS, dDE, dLineSel, dIdCnt;
    HostBuf hLineOff, hLineT, hKeep, hScore, hBits, hEvalue, hSeqId, hQS, hQE, hDS, hDE;
    // compact exports (cdg_alns_compact / cdg_patches): device staging + pinned host copies
    DevBuf dCmpCount, dCmpTarget, dCmpFields, dPatches;
    HostBuf hCmpCount, hCmpTarget, hCmpFields, hPatches;
    u64 nCompact = 0, nChangedLast = 0; bool correctedValid = false; double dbResUsed = 0; bool monotoneOffsets = true; bool packedUpload = false;
    // packed uploads rebuild the DB bytes only when something reads them (a sequence with N, or a byte-consuming entry point)
    bool planReady = false; std::vector<cudaEvent_t> chunkDone;
    // km_prepare's result ((length desc, id asc) order, slot prefix, slice bounds) is a property of the resident set and of
    // (k, kmers per sequence, world, rank): kept until the next upload
    u64 uploadGen = 0, prepGen = ~0ull; int prepK = 0, prepKps = 0; float prepScale = 0.f; u32 prepWorld = 0, prepRank = 0; KmPlanFwd prepPlan; u32 splitPasses = 1; DevBuf dHashHist, dChunkK, dChunkV, dRouteBase;
    DevBuf dExtI32, dExtF32, dExtF64, dExtU8, dExtU32, dExtU64, dExtCons, dExtKeys; HostBuf hExt;   // cdg_extension_scores
    DevBuf dSetFacts, dSlowList, dPrep, dAlnLen, dMism, dKey2HiA, dKey2HiB; u64 *sortedKey2Hi = nullptr; bool bytesValid = true; bool patchesDirect = false;
    u64 nLines = 0, nKept = 0; bool linesValid = false, alnsValid = false, alnsFromText = false;
    // correction
    DevBuf dTables, dOut, dLik, dWideList, dRecvK, dRecvV, dRecv2, dQList, dPow31, dKey2Big; bool damageSet = false;
    u64 headCap = 0;   // records of a neighbour's head run that fit behind sortedKey2 (cdg_mg_sort2 / cdg_mg_head_area)
    // pair likelihood (a19)
    DevBuf dPairModel, dPairIn, dPairCons, dPairOut; bool pairModelSet = false;
    // contig-merge scores (N3)
    DevBuf dMergeLog; cdgmerge::MergeModel mergeModel; bool mergeModelSet = false;
    DevBuf dCycleRecOff, dCycleDiagOff, dCycleHits, dCycleSplit;   // cdg_cyclecheck
    bool qListValid = false;  // dQList / SC_QLIST hold the multi-line queries of the current lines
    u64 hScalars[64] = {0};   // host copy of the scalar block taken after sort 1 (km_sort_group)
};

namespace {

int fail(cdg_ctx *c, int code, const char *fmt, ...) {
    char buf[1024];
    va_list ap; va_start(ap, fmt); vsnprintf(buf, sizeof buf, fmt, ap); va_end(ap);
    if (c) c->err = buf;
    return code;
}
#define CK(call)                                                                                             \
    do {                                                                                                     \
        cudaError_t e_ = (call);                                                                             \
        if (e_ != cudaSuccess) {                                                                             \
            (void) cudaGetLastError(); /* a recovered failure (IPC probe, OOM + retry) must not poison the next launch check */ \
            return fail(ctx, CDG_ERR_CUDA, "%s:%d %s: %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_)); \
        }                                                                                                    \
    } while (0)
// launch check; with CDG_SYNC_DEBUG=1 in the environment every check also waits for the stream, so that an asynchronous kernel
// fault is reported at the launch site that caused it (compute-sanitizer is not always available)
static cudaError_t cdg_launch_check(cdg_ctx *ctx) {
    static const bool dbg = getenv("CDG_SYNC_DEBUG") != nullptr;
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess && dbg && ctx) e = cudaStreamSynchronize(ctx->st);
    return e;
}
#define CKL() CK(cdg_launch_check(ctx))

// scalars block on the device: [0] scan total, [1] emitted records, [2] grouped count, [3] kept lines,
// [4] changed bases, [5] (u32) extraCount, [6] (u32) overflow, [7] (u32) long count, [8] long scratch total, [9] (u32) sort ticket
enum { SC_SCAN = 0, SC_EMIT, SC_GROUPED, SC_KEPT, SC_CHANGED, SC_EXTRA, SC_OVERFLOW, SC_NLONG, SC_LONGSCR, SC_TICKET, SC_WIDE, SC_QLIST, SC_PATCH, SC_SLOW, SC_RTOVF, SC_ROUTE /*16 counters*/, SC_BOUNDS = SC_ROUTE + 16 /*17 bounds*/, SC_N = SC_BOUNDS + 20 };
static_assert(SC_N <= 64, "cdg_ctx::hScalars holds the scalar block");

void rec(cdg_ctx *ctx, int e) { cudaEventRecord(ctx->ev[e], ctx->st); ctx->evValid[e] = true; }

struct LenWordsF { const u32 *len; __device__ u32 operator()(u64 i) const { return (len[i] + 31) >> 5; } };
struct LenBasesF { const u32 *len; __device__ u32 operator()(u64 i) const { return len[i]; } };
struct LenEntryF { const u32 *len; __device__ u32 operator()(u64 i) const { return len[i] + 2; } };   // BASES + "\n\0"
struct OrderedCapacityF { const u64 *order; const u32 *len; int k, kps; float scale;
    __device__ u32 operator()(u64 r) const { return cdgkmer::kmer_capacity(len[(u32) order[r]], k, kps, scale); } };
__global__ void __launch_bounds__(256) order_init_kernel(const u32 *__restrict__ len, u64 n, u32 maxLen, u64 *__restrict__ k, u64 *__restrict__ v) {
    u64 i = (u64) blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) { k[i] = (u64) (maxLen - len[i]); v[i] = i; }
}
__global__ void __launch_bounds__(256) order_scatter_kernel(const u64 *__restrict__ order, const u64 *__restrict__ prefix, u64 n,
                                                            u64 *__restrict__ slotOff, u32 *__restrict__ rankOf) {
    u64 r = (u64) blockIdx.x * blockDim.x + threadIdx.x;
    if (r < n) { const u32 seq = (u32) order[r]; slotOff[seq] = prefix[r]; rankOf[seq] = (u32) r; }
}
__global__ void __launch_bounds__(256) max_len_kernel(const u32 *__restrict__ len, u64 n, u32 *__restrict__ out) {
    u32 m = 0;
    for (u64 i = (u64) blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (u64) gridDim.x * blockDim.x) m = max(m, len[i]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0 && m) atomicMax(out, m);
}
// bit 0: an entry ends beyond the data, bit 1: entries are not laid out in index order
__global__ void __launch_bounds__(256) layout_check_kernel(const u64 *__restrict__ off, const u32 *__restrict__ len, u64 n, u64 dataBytes, u32 *__restrict__ out) {
    u32 f = 0;
    for (u64 i = (u64) blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (u64) gridDim.x * blockDim.x) {
        const u64 e = off[i] + len[i];
        if (e > dataBytes || e < off[i]) f |= 1u;
        if (i + 1 < n && off[i + 1] < e) f |= 2u;
    }
    f = __reduce_or_sync(0xffffffffu, f);
    if ((threadIdx.x & 31) == 0 && f) atomicOr(out, f);
}
template <class F>
__global__ void __launch_bounds__(cdgscan::BLOCK) prefix_write_kernel(F f, u64 n, const unsigned long long *__restrict__ tileSums, u64 *__restrict__ out) {
    __shared__ u32 shWarp[cdgscan::BLOCK / 32 + 1];
    u64 prefix[cdgscan::ITEMS]; u32 val[cdgscan::ITEMS];
    cdgscan::tile_prefix(f, n, tileSums, shWarp, prefix, val);
    u64 base = (u64) blockIdx.x * cdgscan::TILE;
#pragma unroll
    for (int i = 0; i < cdgscan::ITEMS; i++) { u64 idx = base + (u64) i * cdgscan::BLOCK + threadIdx.x; if (idx < n) out[idx] = prefix[i]; }
}
__global__ void set_u64_kernel(u64 *p, const unsigned long long *v) { *p = *v; }

// exclusive scan of f over [0,n) into out[0..n) (and out[n] = total when writeTotal); total also left in scalars[SC_SCAN]
template <class F>
int scan_to(cdg_ctx *ctx, F f, u64 n, u64 *out, bool writeTotal) {
    u64 tiles = cdgscan::tiles_for(n);
    CK(ctx->dTileSums.reserve((tiles + 1) * sizeof(unsigned long long)));
    unsigned long long *ts = ctx->dTileSums.as<unsigned long long>();
    unsigned long long *sc = ctx->dScalars.as<unsigned long long>();
    if (n == 0) { CK(cudaMemsetAsync(sc + SC_SCAN, 0, 8, ctx->st)); if (writeTotal) CK(cudaMemsetAsync(out, 0, 8, ctx->st)); return 0; }
    cdgscan::tile_sum_kernel<<<(unsigned) tiles, cdgscan::BLOCK, 0, ctx->st>>>(f, n, ts);
    cdgscan::tile_scan_kernel<<<1, 1024, 0, ctx->st>>>(ts, tiles, sc + SC_SCAN);
    prefix_write_kernel<<<(unsigned) tiles, cdgscan::BLOCK, 0, ctx->st>>>(f, n, ts, out);
    ctx->launches += 3;
    if (writeTotal) { set_u64_kernel<<<1, 1, 0, ctx->st>>>(out + n, sc + SC_SCAN); ctx->launches++; }
    CKL();
    return 0;
}

int read_scalar(cdg_ctx *ctx, int idx, u64 *out) {
    CK(cudaMemcpyAsync(out, ctx->dScalars.as<unsigned long long>() + idx, 8, cudaMemcpyDeviceToHost, ctx->st));
    CK(cudaStreamSynchronize(ctx->st));
    return 0;
}

__global__ void __launch_bounds__(256) long_list_kernel(const u32 *__restrict__ len, u64 n, int k, int kps, float scale, u32 *__restrict__ list,
                                                        u64 *__restrict__ scratchOff, u32 *nLong, unsigned long long *scratchTotal) {
    u64 i = (u64) blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    u32 L = len[i];
    if (cdgkmer::is_short_seq(L, k, kps, scale)) return;
    u32 W = L >= (u32) k ? L - k + 1 : 0;
    u32 np2 = 1; while (np2 < W) np2 <<= 1;
    u32 slot = atomicAdd(nLong, 1u);
    list[slot] = (u32) i;
    scratchOff[slot] = atomicAdd(scratchTotal, (unsigned long long) np2);
}

int ceil_log2(u64 x) { int b = 0; while ((1ull << b) < x) b++; return b; }

int sort_ws(cdg_ctx *ctx, u64 n, cdgsort::Workspace &ws) {
    u64 tiles = (cdgsort::tiles_for(n) + 1) * (n > cdgsort::SMALL_N ? 2 : 1);   // 64-bit tile states above 2^30 records
    CK(ctx->dHist.reserve(sizeof(unsigned long long) * cdgsort::MAX_PASSES * cdgsort::RADIX));
    CK(ctx->dTileState.reserve(tiles * cdgsort::RADIX * sizeof(u32)));
    ws.hist = ctx->dHist.as<unsigned long long>();
    ws.tileState = ctx->dTileState.as<u32>();
    ws.ticket = (u32 *) (ctx->dScalars.as<unsigned long long>() + SC_TICKET);
    ws.tilesCap = tiles;
    return 0;
}

}   // namespace

// ------------------------------------------------------------------------------------------------------
extern "C" int cdg_create(int device, cdg_ctx **out) {
    if (!out) return CDG_ERR_ARG;
    *out = nullptr;
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0 || device < 0 || device >= count) return CDG_ERR_CUDA;   // no CPU fallback: fail loudly
    cdg_ctx *ctx = new cdg_ctx();
    ctx->device = device;
    if (cudaSetDevice(device) != cudaSuccess) { delete ctx; return CDG_ERR_CUDA; }
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) { delete ctx; return CDG_ERR_CUDA; }
    ctx->smCount = prop.multiProcessorCount;
    if (cudaStreamCreateWithFlags(&ctx->st, cudaStreamNonBlocking) != cudaSuccess) { delete ctx; return CDG_ERR_CUDA; }
    if (cudaStreamCreateWithFlags(&ctx->copySt, cudaStreamNonBlocking) != cudaSuccess) { delete ctx; return CDG_ERR_CUDA; }
    cudaEventCreateWithFlags(&ctx->evReady, cudaEventDisableTiming); cudaEventCreateWithFlags(&ctx->evCopied, cudaEventDisableTiming);
    for (int i = 0; i < EV_N; i++) { cudaEventCreate(&ctx->ev[i]); ctx->evValid[i] = false; }
    for (int i = 0; i < 2 * cdgsort::MAX_PASSES; i++) cudaEventCreate(&ctx->passEv[i]);
    for (int i = 0; i < CDG_UPLOAD_CHUNKS; i++) cudaEventCreateWithFlags(&ctx->chunkEv[i], cudaEventDisableTiming);
    cudaEventCreateWithFlags(&ctx->evUploadStart, cudaEventDisableTiming);
    if (ctx->dScalars.reserve(SC_N * 8) != cudaSuccess) { delete ctx; return CDG_ERR_CUDA; }
    cudaFuncSetAttribute(cdgalign::correct_fast_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) sizeof(cdgalign::CorrSmem));
    cudaFuncSetAttribute(cdgkmer::extract_long_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (32768 + 128) * 4);
    if (ctx->dPow31.reserve(cdgkmer::POW31_TABLE * 8) != cudaSuccess) { delete ctx; return CDG_ERR_CUDA; }
    cdgkmer::pow31_table_kernel<<<1, 32, 0, ctx->st>>>(ctx->dPow31.as<u64>(), cdgkmer::POW31_TABLE);
    if (const char *e = getenv("CDG_WIDE_THRESHOLD")) { long v = atol(e); if (v >= 1 && v < 65535) ctx->wideThreshold = (u32) v; }
    *out = ctx;
    return CDG_OK;
}

extern "C" void cdg_destroy(cdg_ctx *ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->st);
    for (int w = 0; w < CDG_IPC_BUFFERS; w++) for (int pe = 0; pe < CDG_MAX_WORLD; pe++) if (ctx->peerBase[w][pe]) cudaIpcCloseMemHandle(ctx->peerBase[w][pe]);
    DevBuf *d[] = {&ctx->dMeta, &ctx->dData, &ctx->dOff, &ctx->dLen, &ctx->dExt, &ctx->dWordOff, &ctx->dWords, &ctx->dXmask, &ctx->dFlags, &ctx->dLikOff,
                   &ctx->dTileSums, &ctx->dScalars, &ctx->dHist, &ctx->dTileState, &ctx->dSlotOff, &ctx->dKeysA, &ctx->dKeysB, &ctx->dValsA,
                   &ctx->dValsB, &ctx->dIdKeysA, &ctx->dIdKeysB, &ctx->dIdValsA, &ctx->dIdValsB, &ctx->dKey2A, &ctx->dKey2B, &ctx->dLongList,
                   &ctx->dOrdKA, &ctx->dOrdKB, &ctx->dOrdVA, &ctx->dOrdVB, &ctx->dRankOf, &ctx->dSlotRank, &ctx->dScratchOff, &ctx->dScratch, &ctx->dScratchKeep, &ctx->dShortScratch, &ctx->dShortKeep, &ctx->dHitRep, &ctx->dHitTgt,
                   &ctx->dHitScore, &ctx->dHitDiag, &ctx->dLineOff, &ctx->dLineQ, &ctx->dLineT, &ctx->dLinePref, &ctx->dLineDiag, &ctx->dKeep,
                   &ctx->dScore, &ctx->dBits, &ctx->dEvalue, &ctx->dSeqId, &ctx->dQS, &ctx->dQE, &ctx->dDS, &ctx->dDE, &ctx->dLineSel,
                   &ctx->dTables, &ctx->dOut, &ctx->dLik, &ctx->dWideList, &ctx->dRecvK, &ctx->dRecvV, &ctx->dRecv2, &ctx->dQList, &ctx->dPow31,
                   &ctx->dPairModel, &ctx->dPairIn, &ctx->dPairCons, &ctx->dPairOut, &ctx->dRecvIdK, &ctx->dRecvIdV, &ctx->dRouteTab,
                   &ctx->dIdCnt, &ctx->dCmpCount, &ctx->dCmpTarget, &ctx->dCmpFields, &ctx->dPatches, &ctx->dKey2Big, &ctx->dSetFacts, &ctx->dSlowList, &ctx->dPrep, &ctx->dAlnLen, &ctx->dMism, &ctx->dKey2HiA, &ctx->dKey2HiB, &ctx->dHashHist, &ctx->dChunkK, &ctx->dChunkV, &ctx->dRouteBase, &ctx->dExtI32, &ctx->dExtF32, &ctx->dExtF64, &ctx->dExtU8, &ctx->dExtU32, &ctx->dExtU64,
                   &ctx->dExtCons, &ctx->dExtKeys};
    for (DevBuf *b : d) b->release();
    HostBuf *h[] = {&ctx->hHitRep, &ctx->hHitTgt, &ctx->hHitScore, &ctx->hHitDiag, &ctx->hHitDiag16, &ctx->hLineOff, &ctx->hLineT, &ctx->hKeep,
                    &ctx->hScore, &ctx->hBits, &ctx->hEvalue, &ctx->hSeqId, &ctx->hQS, &ctx->hQE, &ctx->hDS, &ctx->hDE,
                    &ctx->hCmpCount, &ctx->hCmpTarget, &ctx->hCmpFields, &ctx->hPatches, &ctx->hExt};
    for (HostBuf *b : h) b->release();
    for (int i = 0; i < EV_N; i++) cudaEventDestroy(ctx->ev[i]);
    for (cudaEvent_t e : ctx->chunkDone) cudaEventDestroy(e);
    for (int i = 0; i < 2 * cdgsort::MAX_PASSES; i++) cudaEventDestroy(ctx->passEv[i]);
    cudaStreamSynchronize(ctx->copySt);
    for (int i = 0; i < CDG_UPLOAD_CHUNKS; i++) cudaEventDestroy(ctx->chunkEv[i]);
    cudaEventDestroy(ctx->evUploadStart);
    cudaStreamDestroy(ctx->st); cudaStreamDestroy(ctx->copySt); cudaEventDestroy(ctx->evReady); cudaEventDestroy(ctx->evCopied);
    delete ctx;
}

extern "C" const char *cdg_last_error(const cdg_ctx *ctx) { return ctx ? ctx->err.c_str() : "no context (CUDA device unavailable?)"; }
extern "C" void *cdg_stream(cdg_ctx *ctx) { return ctx ? (void *) ctx->st : nullptr; }

// ------------------------------------------------------------------------------------------------------
// pack + meta of the chunks whose bytes have arrived but are not packed yet (every consumer of the packed set calls this;
// the k-mer extraction interleaves it with its own per-chunk launches instead, see km_extract)
static int pack_chunk(cdg_ctx *ctx, u64 seqLo, u64 seqHi) {
    if (seqHi <= seqLo) return 0;
    const u64 cnt = seqHi - seqLo;
    const unsigned pb = (unsigned) std::min<u64>((cnt * 32 + 255) / 256, (u64) ctx->smCount * 16);
    if (ctx->packedUpload) {   // cdg_upload_packed: the words are the input; the DB bytes are rebuilt only if a sequence holds an N
        unsigned long long *flagged = ctx->dSetFacts.as<unsigned long long>();
        cdgexp::sanitize_packed_kernel<<<(unsigned) ((cnt + 255) / 256), 256, 0, ctx->st>>>(ctx->dWords.as<u64>(), ctx->dXmask.as<u32>(), ctx->dLen.as<u32>(),
                                                                                           ctx->dWordOff.as<u64>(), seqLo, seqHi, ctx->dFlags.as<u8>(), flagged);
        cdgexp::unpack_kernel<<<pb, 256, 0, ctx->st>>>(ctx->dWords.as<u64>(), ctx->dXmask.as<u32>(), ctx->dOff.as<u64>(), ctx->dLen.as<u32>(), ctx->dWordOff.as<u64>(),
                                                       seqLo, seqHi, ctx->dData.as<u8>(), flagged);
        ctx->launches++;
    } else
    cdgkmer::pack_kernel<<<pb, 256, 0, ctx->st>>>(ctx->dData.as<u8>(), ctx->dOff.as<u64>(), ctx->dLen.as<u32>(), ctx->dWordOff.as<u64>(), seqLo, seqHi,
                                                  ctx->dWords.as<u64>(), ctx->dXmask.as<u32>(), ctx->dFlags.as<u8>());
    cdgalign::meta_kernel<<<(unsigned) ((cnt + 255) / 256), 256, 0, ctx->st>>>(ctx->dOff.as<u64>(), ctx->dWordOff.as<u64>(), ctx->dLen.as<u32>(), ctx->dFlags.as<u8>(),
                                                                           ctx->dExt.as<u8>(), seqLo, seqHi, ctx->dMeta.as<cdgalign::SeqMeta>());
    ctx->launches += 2;
    CKL();
    return 0;
}
static int ensure_packed(cdg_ctx *ctx) {
    for (const cdg_ctx::UploadChunk &c : ctx->pending) {
        CK(cudaStreamWaitEvent(ctx->st, c.ev, 0));
        if (int rc = pack_chunk(ctx, c.seqLo, c.seqHi)) return rc;
    }
    ctx->pending.clear();
    return 0;
}
// the DB byte layout of a packed upload, for the entry points that read or return bytes (no-op after cdg_upload_seqs)
static int ensure_bytes(cdg_ctx *ctx) {
    if (ctx->bytesValid) return 0;
    if (int rc = ensure_packed(ctx)) return rc;
    u64 flagged = 0;
    CK(cudaMemcpyAsync(&flagged, ctx->dSetFacts.p, 8, cudaMemcpyDeviceToHost, ctx->st));
    CK(cudaStreamSynchronize(ctx->st));
    if (flagged == 0 && ctx->n) {   // (otherwise pack_chunk's conditional unpack has already run)
        const unsigned pb = (unsigned) std::min<u64>((ctx->n * 32 + 255) / 256, (u64) ctx->smCount * 16);
        cdgexp::unpack_kernel<<<pb, 256, 0, ctx->st>>>(ctx->dWords.as<u64>(), ctx->dXmask.as<u32>(), ctx->dOff.as<u64>(), ctx->dLen.as<u32>(), ctx->dWordOff.as<u64>(),
                                                       0, ctx->n, ctx->dData.as<u8>(), nullptr);
        ctx->launches++;
        CKL();
    }
    ctx->bytesValid = true;
    return 0;
}

// chunks > 1: `data` and `offsets` are host memory; the DB bytes are copied in `chunks` pieces on the copy stream and the set is
// packed lazily (ensure_packed / km_extract), so that the copy of one piece overlaps pack + extraction of the previous one
static int upload_impl(cdg_ctx *ctx, const uint8_t *data, uint64_t data_bytes, const uint64_t *offsets, const uint32_t *lens, const uint8_t *ext,
                       uint64_t n, int chunks) {
    if (!ctx) return CDG_ERR_ARG;
    if (n == 0 || !data || !offsets || !lens) return fail(ctx, CDG_ERR_ARG, "cdg_upload_seqs: empty or NULL input");
    if (n >= (1ull << 31)) return fail(ctx, CDG_ERR_UNSUPPORTED, "cdg_upload_seqs: more than 2^31 sequences per context");
    CK(cudaSetDevice(ctx->device));
    if (!ctx->pending.empty()) { CK(cudaStreamSynchronize(ctx->copySt)); ctx->pending.clear(); }   // an upload nobody consumed
    ctx->hitsValid = ctx->linesValid = ctx->alnsValid = false; ctx->packedUpload = false; ctx->correctedValid = false; ctx->bytesValid = true;
    ctx->n = n; ctx->dataBytes = data_bytes; ctx->qLo = 0; ctx->nQ = n; ctx->uploadGen++;
    CK(ctx->dData.reserve(data_bytes + 64));
    CK(ctx->dOff.reserve(n * 8)); CK(ctx->dLen.reserve(n * 4)); CK(ctx->dExt.reserve(n)); CK(ctx->dFlags.reserve(n));
    CK(ctx->dWordOff.reserve((n + 1) * 8)); CK(ctx->dLikOff.reserve((n + 1) * 8));
    CK(cudaMemcpyAsync(ctx->dOff.p, offsets, n * 8, cudaMemcpyDefault, ctx->st));
    CK(cudaMemcpyAsync(ctx->dLen.p, lens, n * 4, cudaMemcpyDefault, ctx->st));
    if (ext) CK(cudaMemcpyAsync(ctx->dExt.p, ext, n, cudaMemcpyDefault, ctx->st));
    else CK(cudaMemsetAsync(ctx->dExt.p, 0, n, ctx->st));
    if (chunks > CDG_UPLOAD_CHUNKS) chunks = CDG_UPLOAD_CHUNKS;
    if (chunks <= 1 || n < 4096) {
        CK(cudaMemcpyAsync(ctx->dData.p, data, data_bytes, cudaMemcpyDefault, ctx->st));
        CK(cudaEventRecord(ctx->chunkEv[0], ctx->st));
        ctx->pending.push_back(cdg_ctx::UploadChunk{0, n, ctx->chunkEv[0]});
    } else {
        // the copy stream starts once everything queued on the main stream (which may still read the old set) is done
        CK(cudaEventRecord(ctx->evUploadStart, ctx->st));
        CK(cudaStreamWaitEvent(ctx->copySt, ctx->evUploadStart, 0));
        u64 s0 = 0;
        for (int c = 0; c < chunks; c++) {
            u64 s1 = n;
            if (c + 1 < chunks) {   // first sequence whose offset reaches the next byte target
                const u64 target = data_bytes / chunks * (c + 1);
                s1 = (u64) (std::lower_bound(offsets + s0, offsets + n, target) - offsets);
            }
            const u64 b0 = c == 0 ? 0 : offsets[s0], b1 = s1 < n ? offsets[s1] : data_bytes;
            if (b1 > b0) CK(cudaMemcpyAsync(ctx->dData.as<u8>() + b0, data + b0, b1 - b0, cudaMemcpyDefault, ctx->copySt));
            CK(cudaEventRecord(ctx->chunkEv[c], ctx->copySt));
            ctx->pending.push_back(cdg_ctx::UploadChunk{s0, s1, ctx->chunkEv[c]});
            s0 = s1;
        }
        // The pieces assume that the entries lie in the data file in index order without overlap (what createdb and the stages'
        // single-threaded writers produce).  Checked here, on the host, while the copies are already in flight; any other layout
        // (index sorted by key over a multi-threaded writer's data file) gets one plain copy behind the pieces instead.
        bool ordered = true;
        {
            const int T = 8;
            bool ok[T];
            std::vector<std::thread> th;
            for (int t = 0; t < T; t++)
                th.emplace_back([&, t]() {
                    const u64 lo = (n - 1) * t / T, hi = (n - 1) * (t + 1) / T;
                    bool o = true;
                    for (u64 i = lo; i < hi; i++) o &= offsets[i] + lens[i] <= offsets[i + 1];
                    ok[t] = o;
                });
            for (int t = 0; t < T; t++) { th[t].join(); ordered = ordered && ok[t]; }
        }
        if (!ordered || offsets[n - 1] + lens[n - 1] > data_bytes) {
            if (n && offsets[n - 1] + lens[n - 1] > data_bytes && ordered)
                return fail(ctx, CDG_ERR_ARG, "cdg_upload_seqs_async: the last entry ends beyond data_bytes");
            CK(cudaMemcpyAsync(ctx->dData.p, data, data_bytes, cudaMemcpyDefault, ctx->copySt));
            CK(cudaEventRecord(ctx->chunkEv[0], ctx->copySt));
            ctx->pending.clear();
            ctx->pending.push_back(cdg_ctx::UploadChunk{0, n, ctx->chunkEv[0]});
        }
    }
    // facts needed for sizing: total bases, longest sequence, packed words (all reduced on the device, from the lengths alone)
    unsigned long long *sc = ctx->dScalars.as<unsigned long long>();
    CK(cudaMemsetAsync(sc + SC_WIDE, 0, 8, ctx->st));
    CK(cudaMemsetAsync(sc + SC_QLIST, 0, 8, ctx->st));
    ctx->qListValid = false;
    max_len_kernel<<<ctx->smCount * 4, 256, 0, ctx->st>>>(ctx->dLen.as<u32>(), n, (u32 *) (sc + SC_WIDE));
    // entries must lie inside the data (the kernels index with them); entries that are not in index order are fine for
    // everything except the per-range exports of the multi-GPU phases (cdg_mg_finish)
    layout_check_kernel<<<ctx->smCount * 4, 256, 0, ctx->st>>>(ctx->dOff.as<u64>(), ctx->dLen.as<u32>(), n, data_bytes, (u32 *) (sc + SC_QLIST));
    ctx->launches += 2;
    int rc;
    if ((rc = scan_to(ctx, LenBasesF{ctx->dLen.as<u32>()}, n, ctx->dLikOff.as<u64>(), true))) return rc;
    u64 total = 0, mx64 = 0, layout = 0;
    if ((rc = read_scalar(ctx, SC_SCAN, &total))) return rc;
    if ((rc = read_scalar(ctx, SC_WIDE, &mx64))) return rc;
    if ((rc = read_scalar(ctx, SC_QLIST, &layout))) return rc;
    if ((u32) layout & 1u) { ctx->n = 0; ctx->pending.clear(); cudaStreamSynchronize(ctx->copySt); return fail(ctx, CDG_ERR_ARG, "cdg_upload_seqs: an entry (offset + length) ends beyond data_bytes"); }
    ctx->monotoneOffsets = ((u32) layout & 2u) == 0;
    ctx->totalBases = total; ctx->maxLen = (u32) mx64;
    if ((rc = scan_to(ctx, LenWordsF{ctx->dLen.as<u32>()}, n, ctx->dWordOff.as<u64>(), true))) return rc;
    u64 words = 0;
    if ((rc = read_scalar(ctx, SC_SCAN, &words))) return rc;
    ctx->totalWords = words;
    if (words >= (1ull << 32)) return fail(ctx, CDG_ERR_UNSUPPORTED, "cdg_upload_seqs: more than 2^32 packed words per context");
    CK(ctx->dWords.reserve((words + 2) * 8)); CK(ctx->dXmask.reserve((words + 2) * 4));
    CK(ctx->dMeta.reserve(n * sizeof(cdgalign::SeqMeta)));
    return CDG_OK;
}

extern "C" int cdg_upload_seqs(cdg_ctx *ctx, const uint8_t *data, uint64_t data_bytes, const uint64_t *offsets, const uint32_t *lens,
                               const uint8_t *ext, uint64_t n) {
    int rc = upload_impl(ctx, data, data_bytes, offsets, lens, ext, n, 1);
    if (rc) return rc;
    if ((rc = ensure_packed(ctx))) return rc;
    CK(cudaStreamSynchronize(ctx->st));
    return CDG_OK;
}

extern "C" int cdg_upload_seqs_async(cdg_ctx *ctx, const uint8_t *data, uint64_t data_bytes, const uint64_t *offsets, const uint32_t *lens,
                                     const uint8_t *ext, uint64_t n) {
    return upload_impl(ctx, data, data_bytes, offsets, lens, ext, n, CDG_UPLOAD_CHUNKS);
}

// 2-bit packed input: the words (and the N mask) travel on the copy stream while the layout scans run; the DB bytes are rebuilt
// on the device by unpack_kernel when the first consumer needs them (pack_chunk)
extern "C" int cdg_upload_packed(cdg_ctx *ctx, const uint64_t *words, const uint32_t *nmask, const uint32_t *lens, const uint8_t *ext, uint64_t n,
                                 uint64_t n_words) {
    if (!ctx) return CDG_ERR_ARG;
    if (n == 0 || !words || !lens) return fail(ctx, CDG_ERR_ARG, "cdg_upload_packed: empty or NULL input");
    if (n >= (1ull << 31)) return fail(ctx, CDG_ERR_UNSUPPORTED, "cdg_upload_packed: more than 2^31 sequences per context");
    if (n_words >= (1ull << 32)) return fail(ctx, CDG_ERR_UNSUPPORTED, "cdg_upload_packed: more than 2^32 packed words per context");
    CK(cudaSetDevice(ctx->device));
    if (!ctx->pending.empty()) { CK(cudaStreamSynchronize(ctx->copySt)); ctx->pending.clear(); }
    ctx->hitsValid = ctx->linesValid = ctx->alnsValid = false; ctx->packedUpload = true; ctx->monotoneOffsets = true; ctx->correctedValid = false; ctx->bytesValid = false;
    CK(ctx->dSetFacts.reserve(64)); CK(cudaMemsetAsync(ctx->dSetFacts.p, 0, 64, ctx->st));
    ctx->n = n; ctx->qLo = 0; ctx->nQ = n; ctx->uploadGen++;
    CK(ctx->dOff.reserve(n * 8)); CK(ctx->dLen.reserve(n * 4)); CK(ctx->dExt.reserve(n)); CK(ctx->dFlags.reserve(n));
    CK(ctx->dWordOff.reserve((n + 1) * 8)); CK(ctx->dLikOff.reserve((n + 1) * 8));
    CK(ctx->dWords.reserve((n_words + 2) * 8)); CK(ctx->dXmask.reserve((n_words + 2) * 4));
    CK(ctx->dMeta.reserve(n * sizeof(cdgalign::SeqMeta)));
    CK(cudaMemcpyAsync(ctx->dLen.p, lens, n * 4, cudaMemcpyDefault, ctx->st));
    if (ext) CK(cudaMemcpyAsync(ctx->dExt.p, ext, n, cudaMemcpyDefault, ctx->st));
    else CK(cudaMemsetAsync(ctx->dExt.p, 0, n, ctx->st));
    CK(cudaEventRecord(ctx->evUploadStart, ctx->st));   // work queued earlier may still read the old set
    CK(cudaStreamWaitEvent(ctx->copySt, ctx->evUploadStart, 0));
    CK(cudaMemcpyAsync(ctx->dWords.p, words, n_words * 8, cudaMemcpyDefault, ctx->copySt));
    if (nmask) CK(cudaMemcpyAsync(ctx->dXmask.p, nmask, n_words * 4, cudaMemcpyDefault, ctx->copySt));
    else CK(cudaMemsetAsync(ctx->dXmask.p, 0, n_words * 4, ctx->copySt));
    CK(cudaEventRecord(ctx->chunkEv[0], ctx->copySt));
    ctx->pending.push_back(cdg_ctx::UploadChunk{0, n, ctx->chunkEv[0]});
    unsigned long long *sc = ctx->dScalars.as<unsigned long long>();
    CK(cudaMemsetAsync(sc + SC_WIDE, 0, 8, ctx->st));
    max_len_kernel<<<ctx->smCount * 4, 256, 0, ctx->st>>>(ctx->dLen.as<u32>(), n, (u32 *) (sc + SC_WIDE));
    ctx->launches++;
    int rc;
    u64 total = 0, mx64 = 0, wordsNeeded = 0, bytes = 0;
    if ((rc = scan_to(ctx, LenBasesF{ctx->dLen.as<u32>()}, n, ctx->dLikOff.as<u64>(), true))) return rc;
    if ((rc = read_scalar(ctx, SC_SCAN, &total))) return rc;
    if ((rc = read_scalar(ctx, SC_WIDE, &mx64))) return rc;
    if ((rc = scan_to(ctx, LenWordsF{ctx->dLen.as<u32>()}, n, ctx->dWordOff.as<u64>(), true))) return rc;
    if ((rc = read_scalar(ctx, SC_SCAN, &wordsNeeded))) return rc;
    if ((rc = scan_to(ctx, LenEntryF{ctx->dLen.as<u32>()}, n, ctx->dOff.as<u64>(), false))) return rc;
    if ((rc = read_scalar(ctx, SC_SCAN, &bytes))) return rc;
    if (wordsNeeded != n_words) {
        cudaStreamSynchronize(ctx->copySt); ctx->pending.clear(); ctx->n = 0;
        return fail(ctx, CDG_ERR_ARG, "cdg_upload_packed: lens need %llu words, n_words = %llu", (unsigned long long) wordsNeeded, (unsigned long long) n_words);
    }
    ctx->totalBases = total; ctx->maxLen = (u32) mx64; ctx->totalWords = n_words; ctx->dataBytes = bytes;
    CK(ctx->dData.reserve(bytes + 64));
    return CDG_OK;
}

// ------------------------------------------------------------------------------------------------------
// ---- kmermatcher building blocks (shared by the single-GPU path and the multi-GPU phases) -----------------------------
// idBits/diagBits: key2 layout; world/rank; posLo/posHi: slice of the (length desc, id asc) order this context extracts;
// slotBase/slots: first slot / number of slots of the slice; hashLo/hashHi: hash range this context groups; repChunk: width
// of the representative (= query) index range per rank; qLo/nQ: query range this context rescores and corrects
typedef KmPlanFwd KmPlan;

__global__ void slice_bounds_kernel(const u64 *__restrict__ prefix, u64 n, u64 total, u32 world, u64 *__restrict__ bounds /*[world+1]*/) {
    const u32 j = threadIdx.x;
    if (j > world) return;
    if (j == world) { bounds[j] = n; return; }
    const u64 target = total / world * j;   // first rank position whose slot prefix reaches j/world of the slots
    u64 lo = 0, hi = n;
    while (lo < hi) { const u64 mid = (lo + hi) >> 1; if (prefix[mid] < target) lo = mid + 1; else hi = mid; }
    bounds[j] = lo;
}

static int km_bits(cdg_ctx *ctx, const cdg_kmer_params *p, KmPlan &K) {
    if (p->k < 8 || p->k > 31) return fail(ctx, CDG_ERR_UNSUPPORTED, "k=%d outside [8,31]", p->k);
    if (ctx->maxLen >= (1u << CDG_OWNER_SHIFT)) return fail(ctx, CDG_ERR_UNSUPPORTED, "sequences longer than 2^28 are not supported");
    K.idBits = std::max(1, ceil_log2(ctx->n));
    K.diagBits = std::max(1, ceil_log2(2ull * ctx->maxLen + 1));
    K.wide = 2 * K.idBits + K.diagBits + 1 > 64 || getenv("CDG_FORCE_WIDE_KEY2") != nullptr;   // (test knob: the two-word path on any set)
    if (K.wide && K.world > 1)
        return fail(ctx, CDG_ERR_UNSUPPORTED, "n=%llu sequences x max length %u needs %d grouped-record bits (>64): the two-word records are "
                    "implemented for the single-GPU path only", (unsigned long long) ctx->n, ctx->maxLen, 2 * K.idBits + K.diagBits + 1);
    return 0;
}

// (length desc, id asc) order, slot prefix in that order, rankOf, slice of this rank
static int km_prepare_compute(cdg_ctx *ctx, const cdg_kmer_params *p, KmPlanFwd &K);
static int km_prepare(cdg_ctx *ctx, const cdg_kmer_params *p, KmPlan &K) {
    if (ctx->prepGen == ctx->uploadGen && ctx->prepK == p->k && ctx->prepKps == p->kmers_per_seq && ctx->prepScale == p->kmers_per_seq_scale &&
        ctx->prepWorld == K.world && ctx->prepRank == K.rank) {   // same resident set, same parameters: nothing to recompute
        const bool wide = K.wide; const int idBits = K.idBits, diagBits = K.diagBits;
        K = ctx->prepPlan; K.wide = wide; K.idBits = idBits; K.diagBits = diagBits;
        return 0;
    }
    int rcPrep = km_prepare_compute(ctx, p, K);
    if (rcPrep) return rcPrep;
    ctx->prepGen = ctx->uploadGen; ctx->prepK = p->k; ctx->prepKps = p->kmers_per_seq; ctx->prepScale = p->kmers_per_seq_scale; ctx->prepWorld = K.world; ctx->prepRank = K.rank;
    ctx->prepPlan = K;
    return 0;
}
static int km_prepare_compute(cdg_ctx *ctx, const cdg_kmer_params *p, KmPlan &K) {
    const u64 n = ctx->n;
    unsigned long long *sc = ctx->dScalars.as<unsigned long long>();
    CK(ctx->dSlotOff.reserve((n + 1) * 8)); CK(ctx->dSlotRank.reserve((n + 1) * 8)); CK(ctx->dRankOf.reserve(n * 4));
    CK(ctx->dOrdKA.reserve(n * 8)); CK(ctx->dOrdKB.reserve(n * 8)); CK(ctx->dOrdVA.reserve(n * 8)); CK(ctx->dOrdVB.reserve(n * 8));
    int rc;
    cdgsort::Workspace ws;
    if ((rc = sort_ws(ctx, n, ws))) return rc;
    order_init_kernel<<<(unsigned) ((n + 255) / 256), 256, 0, ctx->st>>>(ctx->dLen.as<u32>(), n, ctx->maxLen, ctx->dOrdKA.as<u64>(), ctx->dOrdVA.as<u64>());
    u64 *ordK, *ordV;
    int lo = cdgsort::radix_sort(ctx->st, ws, ctx->dOrdKA.as<u64>(), ctx->dOrdKB.as<u64>(), ctx->dOrdVA.as<u64>(), ctx->dOrdVB.as<u64>(), n,
                                 cdgsort::make_plan(0, 0, 0, std::max(1, ceil_log2((u64) ctx->maxLen + 1))), &ordK, &ordV, ctx->smCount);
    if (lo < 0) return fail(ctx, CDG_ERR_UNSUPPORTED, "length-order sort failed");
    ctx->launches += lo + 1;
    ctx->ordV = ordV;
    if ((rc = scan_to(ctx, OrderedCapacityF{ordV, ctx->dLen.as<u32>(), p->k, p->kmers_per_seq, p->kmers_per_seq_scale}, n, ctx->dSlotRank.as<u64>(), true))) return rc;
    // per-sequence views of the order (slot offset, rank): the single-GPU extraction walks the sequences in index order, and the
    // block-per-sequence kernel for long sequences looks them up; a multi-GPU slice of short reads walks the order itself
    if (K.world == 1 || !cdgkmer::is_short_seq(ctx->maxLen, p->k, p->kmers_per_seq, p->kmers_per_seq_scale)) {
        order_scatter_kernel<<<(unsigned) ((n + 255) / 256), 256, 0, ctx->st>>>(ordV, ctx->dSlotRank.as<u64>(), n, ctx->dSlotOff.as<u64>(), ctx->dRankOf.as<u32>());
        ctx->launches++;
    }
    u64 total = 0;
    if ((rc = read_scalar(ctx, SC_SCAN, &total))) return rc;
    K.posLo = 0; K.posHi = (u32) n; K.slotBase = 0; K.slots = total; K.totalSlots = total;
    if (K.world > 1) {   // slot-balanced slices of the order
        u64 *bounds = (u64 *) (sc + SC_BOUNDS);
        slice_bounds_kernel<<<1, 32, 0, ctx->st>>>(ctx->dSlotRank.as<u64>(), n, total, K.world, bounds);
        ctx->launches++;
        u64 hb[CDG_MAX_WORLD + 1];
        CK(cudaMemcpyAsync(hb, bounds, (K.world + 1) * 8, cudaMemcpyDeviceToHost, ctx->st));
        CK(cudaStreamSynchronize(ctx->st));
        K.posLo = (u32) hb[K.rank]; K.posHi = (u32) hb[K.rank + 1];
        u64 sb[2] = {0, total};
        if (K.posLo < n) CK(cudaMemcpyAsync(&sb[0], ctx->dSlotRank.as<u64>() + K.posLo, 8, cudaMemcpyDeviceToHost, ctx->st));
        if (K.posHi < n) CK(cudaMemcpyAsync(&sb[1], ctx->dSlotRank.as<u64>() + K.posHi, 8, cudaMemcpyDeviceToHost, ctx->st));
        CK(cudaStreamSynchronize(ctx->st));
        if (K.posLo >= n) sb[0] = total;
        K.slotBase = sb[0]; K.slots = sb[1] - sb[0];
    }
    K.extraCap = (u32) (1024 + (n >> 10));
    K.repChunk = (u32) ((n + K.world - 1) / K.world);
    K.qLo = (u64) K.repChunk * K.rank; if (K.qLo > n) K.qLo = n;
    K.nQ = std::min<u64>(K.repChunk, n - K.qLo);
    K.hashLo = (u32) (((u64) K.rank * 65536 + K.world - 1) / K.world);
    K.hashHi = (u32) (((u64) (K.rank + 1) * 65536 + K.world - 1) / K.world) - 1;
    return 0;
}

// extraction of the slice into (keys, vals)[0, K.slots); extras (identity hashes inside the k-mer key range) at extraBase
static int km_extract(cdg_ctx *ctx, const cdg_kmer_params *p, const KmPlan &K, u64 *keys, u64 *vals, u64 *idKeys, u64 *idVals, u64 extraBase,
                      int identityMode /*ExtractParams::identity*/, u32 hashLo, u32 hashHi, u32 sentinelOwner = 0, bool byRank = false) {
    byRank = byRank || identityMode == 2;
    using namespace cdgkmer;
    const u64 n = ctx->n;
    unsigned long long *sc = ctx->dScalars.as<unsigned long long>();
    ExtractParams E; E.k = p->k; E.kmersPerSeq = p->kmers_per_seq; E.scale = p->kmers_per_seq_scale; E.seed = p->hash_shift;
    E.hashLo = hashLo; E.hashHi = hashHi; E.world = K.world; E.posLo = K.posLo; E.posHi = K.posHi; E.slotBase = K.slotBase; E.identity = identityMode; E.sentinelOwner = sentinelOwner;
    int rc;
    u64 nLong64 = 0, longScr = 0;
    if (!is_short_seq(ctx->maxLen, p->k, p->kmers_per_seq, p->kmers_per_seq_scale)) {   // (monotonic in the length: the longest sequence decides)
        CK(ctx->dLongList.reserve(n * 4)); CK(ctx->dScratchOff.reserve(n * 8));
        CK(cudaMemsetAsync(sc + SC_NLONG, 0, 8, ctx->st)); CK(cudaMemsetAsync(sc + SC_LONGSCR, 0, 8, ctx->st));   // (km_extract runs once per chunk / hash range)
        long_list_kernel<<<(unsigned) ((n + 255) / 256), 256, 0, ctx->st>>>(ctx->dLen.as<u32>(), n, p->k, p->kmers_per_seq, p->kmers_per_seq_scale,
                                                                            ctx->dLongList.as<u32>(), ctx->dScratchOff.as<u64>(), (u32 *) (sc + SC_NLONG),
                                                                            sc + SC_LONGSCR);
        ctx->launches++;
        if ((rc = read_scalar(ctx, SC_NLONG, &nLong64))) return rc;
        if ((rc = read_scalar(ctx, SC_LONGSCR, &longScr))) return rc;
    }
    const u32 nLong = (u32) nLong64;
    int perSm = 3;   // persistent grid: as many resident blocks as the registers allow (asked once per kernel variant)
    {
        static int shortPerSm[3] = {0, 0, 0};
        const u32 maxW0 = ctx->maxLen >= (u32) p->k ? ctx->maxLen - p->k + 1 : 0;
        const int v = maxW0 <= 64 ? 0 : (maxW0 <= 128 ? 1 : 2);
        if (!shortPerSm[v]) {
            int q = 0;
            if (v == 0) cudaOccupancyMaxActiveBlocksPerMultiprocessor(&q, extract_short_kernel<2>, EX_WARPS * 32, 0);
            else if (v == 1) cudaOccupancyMaxActiveBlocksPerMultiprocessor(&q, extract_short_kernel<4>, EX_WARPS * 32, 0);
            else cudaOccupancyMaxActiveBlocksPerMultiprocessor(&q, extract_short_kernel<8>, EX_WARPS * 32, 0);
            shortPerSm[v] = q < 1 ? 1 : q;
        }
        perSm = shortPerSm[v];
    }
    const int shortBlocks = ctx->smCount * perSm;
    static int groupPerSm[3] = {0, 0, 0};
    if (!groupPerSm[0]) {
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&groupPerSm[0], extract_group_kernel<8, 8, true>, EX_WARPS * 32, 0);
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&groupPerSm[1], extract_group_kernel<16, 8, true>, EX_WARPS * 32, 0);
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&groupPerSm[2], extract_group_kernel<32, 8, true>, EX_WARPS * 32, 0);
        for (int &g : groupPerSm) if (g < 1) g = 1;
    }
    const int groupBlocks[3] = {ctx->smCount * groupPerSm[0], ctx->smCount * groupPerSm[1], ctx->smCount * groupPerSm[2]};
    CK(ctx->dSlowList.reserve((n + 1) * 4));
    CK(cudaMemsetAsync(sc + SC_SLOW, 0, 8, ctx->st));
    const u64 gridWarps = (u64) shortBlocks * EX_WARPS;
    CK(ctx->dShortScratch.reserve(gridWarps * 2 * EX_MAXW * sizeof(WinRec)));
    CK(ctx->dShortKeep.reserve(gridWarps * EX_MAXW));
    // Index-order extraction of a set whose bytes are still arriving (asynchronous upload): per chunk wait for its bytes, pack it,
    // extract it — the copy of the next chunk runs meanwhile.  Otherwise (resident set, or a rank-order slice) one launch.
    std::vector<cdg_ctx::UploadChunk> pieces;
    if (!byRank && ctx->pending.size() > 1) { pieces = ctx->pending; ctx->pending.clear(); }
    else { if ((rc = ensure_packed(ctx))) return rc; pieces.push_back(cdg_ctx::UploadChunk{0, n, nullptr}); }
    for (const cdg_ctx::UploadChunk &piece : pieces) {
        if (piece.ev) { CK(cudaStreamWaitEvent(ctx->st, piece.ev, 0)); if ((rc = pack_chunk(ctx, piece.seqLo, piece.seqHi))) return rc; }
        E.seqLo = piece.seqLo; E.seqHi = piece.seqHi;
        const u32 maxW = ctx->maxLen >= (u32) p->k ? ctx->maxLen - p->k + 1 : 0;   // short sequences have <= 256 windows
#define CDG_EXTRACT_ARGS_(LIST, LISTN) (ctx->dWords.as<u64>(), ctx->dXmask.as<u32>(), ctx->dMeta.as<SeqMeta>(), n, E, ctx->dSlotOff.as<u64>(), keys, vals, idKeys, idVals, extraBase, K.extraCap, (u32 *) (sc + SC_EXTRA), (u32 *) (sc + SC_OVERFLOW), ctx->dRankOf.as<u32>(), ctx->dPow31.as<u64>(), ctx->dShortScratch.as<WinRec>(), ctx->dShortKeep.as<u8>(), sc + SC_EMIT, byRank ? ctx->ordV : (const u64 *) nullptr, ctx->dSlotRank.as<u64>(), LIST, LISTN)
        // the 16-bit XXH64 score of a window only matters for ownership (hash range, owner rank) and inside the exact replay
        const bool noScore = hashLo == 0 && hashHi >= 65535 && K.world == 1;
        // fast path: GROUP lanes per sequence, possible repeats go to the slow list; then the exact warp-per-sequence kernel on the list
        const u32 *slowList = ctx->dSlowList.as<u32>(); const u32 *slowCount = (const u32 *) (sc + SC_SLOW);
#define CDG_GROUP_ARGS (ctx->dWords.as<u64>(), ctx->dXmask.as<u32>(), ctx->dMeta.as<SeqMeta>(), n, E, ctx->dSlotOff.as<u64>(), keys, vals, idKeys, idVals, extraBase, K.extraCap, (u32 *) (sc + SC_EXTRA), (u32 *) (sc + SC_OVERFLOW), ctx->dRankOf.as<u32>(), ctx->dPow31.as<u64>(), sc + SC_EMIT, byRank ? ctx->ordV : (const u64 *) nullptr, ctx->dSlotRank.as<u64>(), ctx->dSlowList.as<u32>(), (u32 *) (sc + SC_SLOW))
        // group size from the mean length: 8 lanes cover reads of up to 64 windows (L <= 63 + k), longer short sequences are rare in read
        // sets (config 3: < 0.1 %) and go through the slow list; sets of mostly longer sequences take 16 or 32 lanes
        const u32 meanW = (u32) (ctx->totalBases / std::max<u64>(1, n)) >= (u32) p->k ? (u32) (ctx->totalBases / std::max<u64>(1, n)) - p->k + 1 : 0;
        int gsel = (maxW <= 64 || meanW <= 40) ? 0 : ((maxW <= 128 || meanW <= 80) ? 1 : 2);
        if (const char *e = getenv("CDG_EXTRACT_GROUP")) { const int v = atoi(e); if (v >= 0 && v <= 2) gsel = v; }   // test knob
        if (gsel == 0) { if (noScore) extract_group_kernel<8, 8, false><<<groupBlocks[0], EX_WARPS * 32, 0, ctx->st>>> CDG_GROUP_ARGS; else extract_group_kernel<8, 8, true><<<groupBlocks[0], EX_WARPS * 32, 0, ctx->st>>> CDG_GROUP_ARGS; }
        else if (gsel == 1) { if (noScore) extract_group_kernel<16, 8, false><<<groupBlocks[1], EX_WARPS * 32, 0, ctx->st>>> CDG_GROUP_ARGS; else extract_group_kernel<16, 8, true><<<groupBlocks[1], EX_WARPS * 32, 0, ctx->st>>> CDG_GROUP_ARGS; }
        else { if (noScore) extract_group_kernel<32, 8, false><<<groupBlocks[2], EX_WARPS * 32, 0, ctx->st>>> CDG_GROUP_ARGS; else extract_group_kernel<32, 8, true><<<groupBlocks[2], EX_WARPS * 32, 0, ctx->st>>> CDG_GROUP_ARGS; }
#undef CDG_GROUP_ARGS
        ctx->launches++;
        CKL();
        if (getenv("CDG_DEBUG_EXTRACT")) {
            u64 a = 0, b = 0;
            read_scalar(ctx, SC_SLOW, &a); read_scalar(ctx, SC_EMIT, &b);
            fprintf(stderr, "[extract] gsel %d n %llu pos [%u,%u) after group kernel: slow list %u, emitted %llu\n", gsel, (unsigned long long) n, E.posLo, E.posHi, (u32) a, (unsigned long long) b);
        }
        {
            ExtractParams keepE = E;
            E.identity = 0;   // the identity records of the listed sequences are already out
#define CDG_EXTRACT_LIST_ARGS CDG_EXTRACT_ARGS_(slowList, slowCount)
            if (maxW <= 64) extract_short_kernel<2><<<shortBlocks, EX_WARPS * 32, 0, ctx->st>>> CDG_EXTRACT_LIST_ARGS;
            else if (maxW <= 128) extract_short_kernel<4><<<shortBlocks, EX_WARPS * 32, 0, ctx->st>>> CDG_EXTRACT_LIST_ARGS;
            else extract_short_kernel<8><<<shortBlocks, EX_WARPS * 32, 0, ctx->st>>> CDG_EXTRACT_LIST_ARGS;
#undef CDG_EXTRACT_LIST_ARGS
            E = keepE;
            if (getenv("CDG_DEBUG_EXTRACT")) { u64 b = 0; read_scalar(ctx, SC_EMIT, &b); fprintf(stderr, "[extract] after the list kernel: emitted %llu\n", (unsigned long long) b); }
            CK(cudaMemsetAsync(sc + SC_SLOW, 0, 8, ctx->st));   // the next piece starts an empty list
        }
#undef CDG_EXTRACT_ARGS_
        ctx->launches++;
        CKL();
    }
    if (nLong) {
        CK(ctx->dScratch.reserve((longScr + 1) * sizeof(WinRec))); CK(ctx->dScratchKeep.reserve(longScr + 1));
        unsigned lb = nLong < (u32) ctx->smCount ? nLong : (u32) ctx->smCount;
        extract_long_kernel<<<lb, EXL_BLOCK, (32768 + 128) * 4, ctx->st>>>(
            ctx->dWords.as<u64>(), ctx->dXmask.as<u32>(), ctx->dWordOff.as<u64>(), ctx->dLen.as<u32>(), ctx->dLongList.as<u32>(), nLong, E,
            ctx->dSlotOff.as<u64>(), keys, vals, idKeys, idVals, extraBase, K.extraCap, (u32 *) (sc + SC_EXTRA), (u32 *) (sc + SC_OVERFLOW),
            ctx->dRankOf.as<u32>(), ctx->dScratch.as<WinRec>(), ctx->dScratchKeep.as<u8>(), ctx->dScratchOff.as<u64>(), sc + SC_EMIT);
        ctx->launches++;
        CKL();
    }
    return 0;
}

// The identity array is sorted by key bits [ID_SORT_LO, 63) only; this orders what is left: one thread per run of records that agree in
// those bits.  Runs of true duplicates (equal keys) are only read; a run that mixes keys (a 40-bit collision of two 64-bit hashes) is
// insertion-sorted in place by the low bits — stable, like the radix passes, so the result is the full 63-bit stable sort.
static constexpr int ID_SORT_LO = 23;
__global__ void __launch_bounds__(256) id_fixup_kernel(u64 *__restrict__ keys, u64 *__restrict__ vals, u64 n, int lo) {
    const u64 i = (u64) blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const u64 lowMask = (1ull << lo) - 1, hiMask = ~lowMask & ~CDG_B63;
    const u64 k = keys[i];
    if (k == CDG_SENTINEL) return;   // the unowned tail (hash-range calls): one long run of equal keys
    if (i > 0 && ((keys[i - 1] ^ k) & hiMask) == 0) return;   // not the head of its run
    u64 j = i + 1; bool mixed = false;
    while (j < n && ((keys[j] ^ k) & hiMask) == 0) { mixed = mixed || ((keys[j] ^ k) & lowMask) != 0; j++; }
    if (!mixed) return;
    for (u64 a = i + 1; a < j; a++) {
        const u64 ka = keys[a], va = vals[a];
        u64 b = a;
        while (b > i && (keys[b - 1] & lowMask) > (ka & lowMask)) { keys[b] = keys[b - 1]; vals[b] = vals[b - 1]; b--; }
        keys[b] = ka; vals[b] = va;
    }
}

// sort 1 (main array of Rtot records in (ka,va) with partner (kb,vb); identity array of n records) + group -> key2 in dKey2A
static int km_sort_group(cdg_ctx *ctx, const cdg_kmer_params *p, const KmPlan &K, u64 *ka, u64 *kb, u64 *va, u64 *vb, u64 Rtot, int quirkMode,
                         u64 *minKeyOut, u64 *R2out) {
    using namespace cdgkmer;
    const u64 n = ctx->n;
    unsigned long long *sc = ctx->dScalars.as<unsigned long long>();
    int rc;
    cdgsort::Workspace ws;
    if ((rc = sort_ws(ctx, Rtot > n ? Rtot : n, ws))) return rc;
    cdgsort::Plan plan1 = cdgsort::make_plan(0, 0, 0, 2 * p->k);
    int l1 = cdgsort::radix_sort(ctx->st, ws, ka, kb, va, vb, Rtot, plan1, &ctx->sortedMainK, &ctx->sortedMainV, ctx->smCount, -1, ctx->passEv);
    ctx->passEvN = plan1.n; ctx->sort1Records = Rtot;
    // identity records (64-bit hashes, one per sequence): five passes over key bits [23, 63) and a fix-up of the (expected: n^2 / 2^41)
    // runs that agree in those 40 bits but not below, instead of eight passes over all 63 bits
    int idLo = ID_SORT_LO;
    if (const char *e = getenv("CDG_ID_SORT_LO")) { const int v = atoi(e); if (v >= 0 && v < 63) idLo = v; }   // test knob: a high value sends every run through the fix-up
    cdgsort::Plan planI = cdgsort::make_plan(0, 0, idLo, 63);
    int l2 = cdgsort::radix_sort(ctx->st, ws, ctx->idSrcK, ctx->idAltK, ctx->idSrcV, ctx->idAltV, ctx->nIdent, planI, &ctx->sortedIdK, &ctx->sortedIdV, ctx->smCount);
    if (l1 < 0 || l2 < 0) return fail(ctx, CDG_ERR_UNSUPPORTED, "radix sort: input too large for one pass set");
    if (ctx->nIdent) { id_fixup_kernel<<<(unsigned) ((ctx->nIdent + 255) / 256), 256, 0, ctx->st>>>(ctx->sortedIdK, ctx->sortedIdV, ctx->nIdent, idLo); l2++; }
    ctx->launches += l1 + l2; ctx->sort1Passes = plan1.n;
    CKL();
    rec(ctx, EV_SORT1);
    // one host round trip for everything the host needs after sort 1: the scalar block (overflow flag, emitted / extra counts) and the smallest key
    u64 firstKey = CDG_SENTINEL;
    CK(cudaMemcpyAsync(ctx->hScalars, sc, SC_N * 8, cudaMemcpyDeviceToHost, ctx->st));
    if (Rtot) CK(cudaMemcpyAsync(&firstKey, ctx->sortedMainK, 8, cudaMemcpyDeviceToHost, ctx->st));
    CK(cudaStreamSynchronize(ctx->st));
    if ((u32) ctx->hScalars[SC_OVERFLOW]) return fail(ctx, CDG_ERR_UNSUPPORTED, "identity-record overflow area exhausted");
    if (minKeyOut) *minKeyOut = firstKey == CDG_SENTINEL ? CDG_SENTINEL : (firstKey | CDG_B63);
    if (quirkMode < 0) return 0;   // multi-GPU: the caller decides the first-run quirk after comparing the ranks' minima
    return 0;
}

// append: the grouped records go behind the ones already counted in SC_GROUPED (hash-range passes); mainSkip: leading records of
// the sorted main array that are grouped separately (the first run of a hash range, see run_kmermatch_split)
static int km_group(cdg_ctx *ctx, const cdg_kmer_params *p, const KmPlan &K, u64 Rtot, bool mainHasGlobalFirstRun, bool idHasGlobalFirstRun, u64 *R2out,
                    bool append = false, u64 mainSkip = 0, u64 key2CapOverride = 0) {
    using namespace cdgkmer;
    const u64 n = ctx->n;
    unsigned long long *sc = ctx->dScalars.as<unsigned long long>();
    int rc;
    const u64 key2Cap = key2CapOverride ? key2CapOverride : Rtot + n + CDG_HEAD_RUN;
    CK(ctx->dKey2A.reserve(key2Cap * 8)); CK(ctx->dKey2B.reserve(key2Cap * 8));
    if (K.wide) { CK(ctx->dKey2HiA.reserve(key2Cap * 8)); CK(ctx->dKey2HiB.reserve(key2Cap * 8)); }
    if (!append) CK(cudaMemsetAsync(sc + SC_GROUPED, 0, 8, ctx->st));
    const u64 *mainK = ctx->sortedMainK + mainSkip, *mainV = ctx->sortedMainV + mainSkip;
    Rtot -= mainSkip;
    GroupParams G; G.idBits = K.idBits; G.diagBits = K.diagBits; G.diagBias = (i32) ctx->maxLen; G.onlyExtendable = p->include_only_extendable; G.covThr = p->cov_thr; G.covMode = p->cov_mode;
    G.firstRunQuirk = mainHasGlobalFirstRun ? 1 : 0;   // kmermatcher.cpp:466: only the globally first run
    if (Rtot) launch_group(ctx->st, K.wide, (unsigned) ((Rtot + GROUP_TILE - 1) / GROUP_TILE), mainK, mainV, Rtot, ctx->dLen.as<u32>(), G, ctx->dKey2A.as<u64>(), sc + SC_GROUPED,
                           K.wide ? ctx->dKey2HiA.as<u64>() : nullptr);
    GroupParams GI = G; GI.firstRunQuirk = idHasGlobalFirstRun ? 1 : 0;
    const u64 nId = ctx->nIdent;
    if (nId) launch_group(ctx->st, K.wide, (unsigned) ((nId + GROUP_TILE - 1) / GROUP_TILE), ctx->sortedIdK, ctx->sortedIdV, nId, ctx->dLen.as<u32>(), GI, ctx->dKey2A.as<u64>(),
                          sc + SC_GROUPED, K.wide ? ctx->dKey2HiA.as<u64>() : nullptr);
    ctx->launches += 2;
    CKL();
    if ((rc = read_scalar(ctx, SC_GROUPED, R2out))) return rc;
    rec(ctx, EV_GROUP);
    return 0;
}

// sort 2 of R2 records in dKey2A (partner dKey2B); result pointer in ctx->sortedKey2
static int km_sort2(cdg_ctx *ctx, const KmPlan &K, u64 R2, u64 *src = nullptr, u64 *alt = nullptr) {
    int rc;
    cdgsort::Workspace ws;
    if (R2 > cdgsort::MAX_N) return fail(ctx, CDG_ERR_UNSUPPORTED, "grouped records exceed one sort");
    if ((rc = sort_ws(ctx, std::max<u64>(R2, ctx->n), ws))) return rc;
    u64 *dummy = nullptr;
    ctx->sortedKey2Hi = nullptr;
    int l3;
    if (K.wide) {   // two-word records: pair sort by (id, diag) on the lo words carrying hi, then by rep on the hi words carrying lo (LSD, stable)
        cdgsort::Plan planLo = cdgsort::make_plan(0, 0, 1, K.idBits + K.diagBits + 1), planHi = cdgsort::make_plan(0, 0, 0, K.idBits);
        u64 *lo1, *hi1;
        int la = cdgsort::radix_sort(ctx->st, ws, ctx->dKey2A.as<u64>(), ctx->dKey2B.as<u64>(), ctx->dKey2HiA.as<u64>(), ctx->dKey2HiB.as<u64>(), R2, planLo, &lo1, &hi1, ctx->smCount);
        if (la < 0) return fail(ctx, CDG_ERR_UNSUPPORTED, "radix sort 2 failed");
        u64 *loAlt = lo1 == ctx->dKey2A.as<u64>() ? ctx->dKey2B.as<u64>() : ctx->dKey2A.as<u64>();
        u64 *hiAlt = hi1 == ctx->dKey2HiA.as<u64>() ? ctx->dKey2HiB.as<u64>() : ctx->dKey2HiA.as<u64>();
        int lb = cdgsort::radix_sort(ctx->st, ws, hi1, hiAlt, lo1, loAlt, R2, planHi, &ctx->sortedKey2Hi, &ctx->sortedKey2, ctx->smCount);
        if (lb < 0) return fail(ctx, CDG_ERR_UNSUPPORTED, "radix sort 2 failed");
        l3 = la + lb; ctx->sort2Passes = planLo.n + planHi.n;
    } else {
    cdgsort::Plan plan2 = cdgsort::make_plan(0, 0, 1, 2 * K.idBits + K.diagBits + 1);   // bit 0 (needRev) is not an ordering key
    if (!src) { src = ctx->dKey2A.as<u64>(); alt = ctx->dKey2B.as<u64>(); }
    l3 = cdgsort::radix_sort(ctx->st, ws, src, alt, nullptr, nullptr, R2, plan2, &ctx->sortedKey2, &dummy, ctx->smCount);
    if (l3 < 0) return fail(ctx, CDG_ERR_UNSUPPORTED, "radix sort 2 failed");
    ctx->sort2Passes = plan2.n;
    }
    ctx->launches += l3;
    CKL();
    rec(ctx, EV_SORT2);
    return 0;
}

// best diagonal per pair over sortedKey2[0, nTotal); pairs ending at or after nOwn are walked but not emitted
static int km_reduce(cdg_ctx *ctx, const KmPlan &K, u64 nOwn, u64 nTotal) {
    using namespace cdgkmer;
    unsigned long long *sc = ctx->dScalars.as<unsigned long long>();
    int rc;
    u64 H = 0;
    if (nOwn > 0) {
        PairEndFlag f; f.keys = ctx->sortedKey2; f.n = nTotal; f.idBits = K.idBits; f.diagBits = K.diagBits; f.nOwn = nOwn; f.hi = K.wide ? ctx->sortedKey2Hi : nullptr;
        u64 tiles = cdgscan::tiles_for(nTotal);
        CK(ctx->dTileSums.reserve((tiles + 1) * 8));
        unsigned long long *ts = ctx->dTileSums.as<unsigned long long>();
        if (!K.wide && (((unsigned long long) f.keys) & 15ull) == 0)
            pair_end_count_kernel<<<(unsigned) tiles, cdgscan::BLOCK, 0, ctx->st>>>(f.keys, nTotal, nOwn, K.idBits, K.diagBits, ts);
        else cdgscan::tile_sum_kernel<<<(unsigned) tiles, cdgscan::BLOCK, 0, ctx->st>>>(f, nTotal, ts);
        cdgscan::tile_scan_kernel<<<1, 1024, 0, ctx->st>>>(ts, tiles, sc + SC_SCAN);
        ctx->launches += 2;
        if ((rc = read_scalar(ctx, SC_SCAN, &H))) return rc;
        CK(ctx->dHitRep.reserve((H + 1) * 4)); CK(ctx->dHitTgt.reserve((H + 1) * 4)); CK(ctx->dHitScore.reserve((H + 1) * 4)); CK(ctx->dHitDiag.reserve((H + 1) * 4));
        ReduceParams R; R.idBits = K.idBits; R.diagBits = K.diagBits; R.diagBias = (i32) ctx->maxLen;
        if (K.wide) reduce_kernel<true><<<(unsigned) tiles, cdgscan::BLOCK, 0, ctx->st>>>(f, R, ts, ctx->dHitRep.as<u32>(), ctx->dHitTgt.as<u32>(), ctx->dHitScore.as<i32>(),
                                                                                      ctx->dHitDiag.as<i32>());
        else reduce_kernel<<<(unsigned) tiles, cdgscan::BLOCK, 0, ctx->st>>>(f, R, ts, ctx->dHitRep.as<u32>(), ctx->dHitTgt.as<u32>(), ctx->dHitScore.as<i32>(),
                                                                      ctx->dHitDiag.as<i32>());
        ctx->launches++;
        CKL();
    } else {
        CK(ctx->dHitRep.reserve(4)); CK(ctx->dHitTgt.reserve(4)); CK(ctx->dHitScore.reserve(4)); CK(ctx->dHitDiag.reserve(4));
    }
    rec(ctx, EV_REDUCE);
    ctx->nHits = H; ctx->hitsValid = true; ctx->linesValid = false; ctx->alnsValid = false;
    return 0;
}

// ---- hash-range passes on one context (kmermatcher.cpp:615-627 memory sizing, :742-784 setupKmerSplits, :665-675 the split loop) ----
// The k-mer records of the set do not fit (or `passes` is forced): the 16-bit hash space is cut into `M` contiguous ranges of about
// equal record count from a 65536-bin histogram, and per range the sequences are extracted chunk by chunk with the range as hash
// filter, compacted into the range's record array, sorted and grouped; the grouped records of all ranges accumulate and the second
// sort + reduction run once over all of them, so the result is the UNSPLIT one (the reference's split mode applies the first-run
// quirk per split and merges unsigned-char scores on the host, :953-1111).  One resident copy of the sequences; peak memory =
// 4 arrays of the largest range + a chunk buffer + the grouped records.  The first run of every range is grouped at the end,
// when the range holding the globally smallest k-mer (first-run quirk, :466) is known.
static int run_kmermatch_split(cdg_ctx *ctx, const cdg_kmer_params *p, const KmPlan &K, int M) {
    using namespace cdgkmer;
    const u64 n = ctx->n;
    unsigned long long *sc = ctx->dScalars.as<unsigned long long>();
    int rc;
    if ((rc = ensure_packed(ctx))) return rc;
    // 1. histogram of the window hashes -> ranges
    CK(ctx->dHashHist.reserve(65536 * 8));
    CK(cudaMemsetAsync(ctx->dHashHist.p, 0, 65536 * 8, ctx->st));
    ExtractParams E0; E0.k = p->k; E0.kmersPerSeq = p->kmers_per_seq; E0.scale = p->kmers_per_seq_scale; E0.seed = p->hash_shift;
    E0.hashLo = 0; E0.hashHi = 65535; E0.world = 1; E0.posLo = 0; E0.posHi = (u32) n; E0.slotBase = 0; E0.identity = 0; E0.seqLo = 0; E0.seqHi = n;
    hash_hist_kernel<<<ctx->smCount * 8, 256, 0, ctx->st>>>(ctx->dWords.as<u64>(), ctx->dXmask.as<u32>(), ctx->dMeta.as<SeqMeta>(), n, E0, ctx->dHashHist.as<unsigned long long>());
    ctx->launches++;
    CKL();
    std::vector<u64> hist(65536);
    CK(cudaMemcpyAsync(hist.data(), ctx->dHashHist.p, 65536 * 8, cudaMemcpyDeviceToHost, ctx->st));
    CK(cudaStreamSynchronize(ctx->st));
    u64 totalWin = 0;
    for (u64 v : hist) totalWin += v;
    if (totalWin == 0) return -100;   // no k-mer at all: the caller runs the unsplit path (nothing to split)
    std::vector<u32> lo, hi; std::vector<u64> cnt;
    {
        const u64 target = (totalWin + (u64) M - 1) / (u64) M;
        u64 acc = 0, accAll = 0; u32 start = 0;
        for (u32 b = 0; b < 65536; b++) {
            acc += hist[b]; accAll += hist[b];
            if (b == 65535 || (accAll >= target * (lo.size() + 1) && (int) lo.size() < M - 1)) { lo.push_back(start); hi.push_back(b); cnt.push_back(acc); acc = 0; start = b + 1; }
        }
    }
    M = (int) lo.size();
    u64 rangeCap = 0;
    for (u64 c : cnt) rangeCap = std::max(rangeCap, c);
    rangeCap += K.extraCap + 1;
    if (rangeCap > cdgsort::MAX_N) return fail(ctx, CDG_ERR_UNSUPPORTED, "a hash range still holds %llu records (> 2^32): more passes needed", (unsigned long long) rangeCap);
    // 2. chunks of the (length desc, id asc) order with about half a range of slots each
    const u64 chunkTarget = std::max<u64>(K.slots / (2ull * (u64) M) + 1, 1ull << 20);
    const int C = (int) std::min<u64>((K.slots + chunkTarget - 1) / chunkTarget, 4096);
    std::vector<u32> cpos(C + 1); std::vector<u64> cslot(C + 1);
    for (int c = 0; c <= C; c++) cpos[c] = (u32) (n * (u64) c / (u64) C);
    for (int c = 0; c < C; c++) CK(cudaMemcpyAsync(&cslot[c], ctx->dSlotRank.as<u64>() + cpos[c], 8, cudaMemcpyDeviceToHost, ctx->st));
    CK(cudaStreamSynchronize(ctx->st));
    cslot[C] = K.slots;
    u64 chunkCap = 0;
    for (int c = 0; c < C; c++) chunkCap = std::max(chunkCap, cslot[c + 1] - cslot[c]);
    CK(ctx->dChunkK.reserve((chunkCap + K.extraCap + 1) * 8)); CK(ctx->dChunkV.reserve((chunkCap + K.extraCap + 1) * 8));
    CK(ctx->dKeysA.reserve(rangeCap * 8)); CK(ctx->dKeysB.reserve(rangeCap * 8)); CK(ctx->dValsA.reserve(rangeCap * 8)); CK(ctx->dValsB.reserve(rangeCap * 8));
    CK(ctx->dIdKeysA.reserve(n * 8)); CK(ctx->dIdKeysB.reserve(n * 8)); CK(ctx->dIdValsA.reserve(n * 8)); CK(ctx->dIdValsB.reserve(n * 8));
    const u64 key2Cap = std::min<u64>(totalWin, K.slots) + n + CDG_HEAD_RUN;
    CK(cudaMemsetAsync(sc + SC_GROUPED, 0, 8, ctx->st));
    unsigned long long *cursor = sc + SC_ROUTE;   // records of the current range so far
    std::vector<DevBuf> side(M); std::vector<u64> runLen(M, 0), minKey(M, CDG_SENTINEL);
    auto releaseSide = [&]() { for (DevBuf &b : side) b.release(); };
    auto compactInto = [&](const u64 *ck, const u64 *cv, u64 cntSlots) -> int {   // chunk buffer -> behind the range's records (stable)
        if (!cntSlots) return 0;
        const u64 tiles = cdgscan::tiles_for(cntSlots);
        CK(ctx->dTileSums.reserve((tiles + 1) * 8));
        unsigned long long *ts = ctx->dTileSums.as<unsigned long long>();
        NotSentinelF f{ck};
        cdgscan::tile_sum_kernel<<<(unsigned) tiles, cdgscan::BLOCK, 0, ctx->st>>>(f, cntSlots, ts);
        cdgscan::tile_scan_kernel<<<1, 1024, 0, ctx->st>>>(ts, tiles, sc + SC_SCAN);
        compact_records_kernel<<<(unsigned) tiles, cdgscan::BLOCK, 0, ctx->st>>>(f, cv, cntSlots, ts, cursor, ctx->dKeysA.as<u64>(), ctx->dValsA.as<u64>());
        add_u64_kernel<<<1, 1, 0, ctx->st>>>(cursor, sc + SC_SCAN);
        ctx->launches += 4;
        return 0;
    };
    ctx->nIdent = n;
    for (int r = 0; r < M; r++) {
        CK(cudaMemsetAsync(cursor, 0, 8, ctx->st));
        CK(cudaMemsetAsync(sc + SC_EXTRA, 0, 8, ctx->st));
        CK(cudaMemsetAsync(ctx->dChunkK.as<u64>() + chunkCap, 0xFF, (size_t) K.extraCap * 8, ctx->st));
        CK(cudaMemsetAsync(ctx->dChunkV.as<u64>() + chunkCap, 0, (size_t) K.extraCap * 8, ctx->st));
        for (int c = 0; c < C; c++) {
            KmPlan Kc = K; Kc.posLo = cpos[c]; Kc.posHi = cpos[c + 1]; Kc.slotBase = cslot[c]; Kc.slots = cslot[c + 1] - cslot[c];
            if (Kc.posHi <= Kc.posLo) continue;
            if ((rc = km_extract(ctx, p, Kc, ctx->dChunkK.as<u64>(), ctx->dChunkV.as<u64>(), ctx->dIdKeysA.as<u64>(), ctx->dIdValsA.as<u64>(), chunkCap, 1, lo[r], hi[r], 0,
                                 true))) { releaseSide(); return rc; }
            if ((rc = compactInto(ctx->dChunkK.as<u64>(), ctx->dChunkV.as<u64>(), Kc.slots))) { releaseSide(); return rc; }
        }
        if ((rc = compactInto(ctx->dChunkK.as<u64>() + chunkCap, ctx->dChunkV.as<u64>() + chunkCap, K.extraCap))) { releaseSide(); return rc; }
        CKL();
        if (r == 0) rec(ctx, EV_EXTRACT);
        u64 Rr = 0;
        if ((rc = read_scalar(ctx, SC_ROUTE, &Rr))) { releaseSide(); return rc; }
        if (Rr > rangeCap) { releaseSide(); return fail(ctx, CDG_ERR_CUDA, "hash range %d holds %llu records, the histogram bound was %llu", r, (unsigned long long) Rr, (unsigned long long) rangeCap); }
        ctx->idSrcK = ctx->dIdKeysA.as<u64>(); ctx->idSrcV = ctx->dIdValsA.as<u64>(); ctx->idAltK = ctx->dIdKeysB.as<u64>(); ctx->idAltV = ctx->dIdValsB.as<u64>();
        if ((rc = km_sort_group(ctx, p, K, ctx->dKeysA.as<u64>(), ctx->dKeysB.as<u64>(), ctx->dValsA.as<u64>(), ctx->dValsB.as<u64>(), Rr, 0, &minKey[r], nullptr))) { releaseSide(); return rc; }
        if (Rr) {
            first_run_kernel<<<1, 1024, 0, ctx->st>>>(ctx->sortedMainK, Rr, sc + SC_SCAN);
            ctx->launches++;
            if ((rc = read_scalar(ctx, SC_SCAN, &runLen[r]))) { releaseSide(); return rc; }
        }
        if (runLen[r]) {   // the first run waits in a side buffer until the range with the globally smallest k-mer is known
            if (side[r].reserve(runLen[r] * 16) != cudaSuccess) { releaseSide(); return fail(ctx, CDG_ERR_CUDA, "out of memory (first run of a hash range)"); }
            CK(cudaMemcpyAsync(side[r].p, ctx->sortedMainK, runLen[r] * 8, cudaMemcpyDeviceToDevice, ctx->st));
            CK(cudaMemcpyAsync(side[r].as<u64>() + runLen[r], ctx->sortedMainV, runLen[r] * 8, cudaMemcpyDeviceToDevice, ctx->st));
        }
        u64 r2 = 0;
        if ((rc = km_group(ctx, p, K, Rr, false, false, &r2, true, runLen[r], key2Cap))) { releaseSide(); return rc; }
    }
    // 3. the first runs; the quirk goes to the range with the smallest key (ties cannot happen: a k-mer has one hash)
    int rStar = -1;
    for (int r = 0; r < M; r++) if (minKey[r] != CDG_SENTINEL && (rStar < 0 || minKey[r] < minKey[rStar])) rStar = r;
    GroupParams G; G.idBits = K.idBits; G.diagBits = K.diagBits; G.diagBias = (i32) ctx->maxLen; G.onlyExtendable = p->include_only_extendable; G.covThr = p->cov_thr; G.covMode = p->cov_mode;
    for (int r = 0; r < M; r++) {
        if (!runLen[r]) continue;
        G.firstRunQuirk = r == rStar ? 1 : 0;
        const unsigned gb = (unsigned) ((runLen[r] + GROUP_TILE - 1) / GROUP_TILE);
        launch_group(ctx->st, K.wide, gb, side[r].as<u64>(), side[r].as<u64>() + runLen[r], runLen[r], ctx->dLen.as<u32>(), G, ctx->dKey2A.as<u64>(), sc + SC_GROUPED,
                     K.wide ? ctx->dKey2HiA.as<u64>() : nullptr);
        ctx->launches++;
    }
    CKL();
    u64 R2 = 0, emitted = 0;
    rc = read_scalar(ctx, SC_GROUPED, &R2);
    releaseSide();
    if (rc) return rc;
    if ((rc = read_scalar(ctx, SC_EMIT, &emitted))) return rc;
    rec(ctx, EV_GROUP);
    ctx->nRecords = emitted + n; ctx->nGrouped = R2; ctx->splitPasses = (u32) M;
    if ((rc = km_sort2(ctx, K, R2))) return rc;
    return km_reduce(ctx, K, R2, R2);
}

static int run_kmermatch(cdg_ctx *ctx, const cdg_kmer_params *p) {
    const u64 n = ctx->n;
    KmPlan K; K.world = 1; K.rank = 0;
    int rc = km_bits(ctx, p, K);
    if (rc) return rc;
    unsigned long long *sc = ctx->dScalars.as<unsigned long long>();
    CK(cudaMemsetAsync(sc, 0, SC_N * 8, ctx->st));
    rec(ctx, EV_START);
    if ((rc = km_prepare(ctx, p, K))) return rc;
    ctx->plan = K; ctx->qLo = 0; ctx->nQ = n; ctx->nIdent = n; ctx->splitPasses = 1;
    const u64 Rtot = K.slots + K.extraCap;
    {   // number of hash-range passes: forced, or from the memory the unsplit arrays would need (kmermatcher.cpp:615-627)
        int M = p->split_passes;
        if (const char *e = getenv("CDG_FORCE_SPLITS")) { const int v = atoi(e); if (v > 0) M = v; }   // test knob
        const u64 key2Each = (Rtot + n + CDG_HEAD_RUN) * 8;
        // steady state (the same set again): every array the unsplit run needs is already allocated — no cudaMemGetInfo, which costs
        // from a fraction of a millisecond to several milliseconds of host time per call depending on the box
        if (M <= 0 && Rtot <= cdgsort::MAX_N && ctx->dKeysA.cap >= Rtot * 8 && ctx->dKeysB.cap >= Rtot * 8 && ctx->dValsA.cap >= Rtot * 8 &&
            ctx->dValsB.cap >= Rtot * 8 && ctx->dKey2A.cap >= key2Each && ctx->dKey2B.cap >= key2Each &&
            (!K.wide || (ctx->dKey2HiA.cap >= key2Each && ctx->dKey2HiB.cap >= key2Each)))
            M = 1;
        if (M <= 0) {
            size_t freeB = 0, totalB = 0;
            CK(cudaMemGetInfo(&freeB, &totalB));
            const u64 have = ctx->dKeysA.cap + ctx->dKeysB.cap + ctx->dValsA.cap + ctx->dValsB.cap + ctx->dKey2A.cap + ctx->dKey2B.cap;
            const u64 needMain = 4 * Rtot * 8, needKey2 = 2 * (Rtot + n + CDG_HEAD_RUN) * 8 * (K.wide ? 2 : 1);
            const u64 budget = (u64) ((double) (freeB + have) * 0.85);
            if (Rtot <= cdgsort::MAX_N && needMain + needKey2 <= budget) M = 1;
            else if (needKey2 >= budget) return fail(ctx, CDG_ERR_UNSUPPORTED, "the grouped records alone (%llu bytes) exceed the free device memory: shard the set over more GPUs", (unsigned long long) needKey2);
            else M = (int) std::min<u64>(64, std::max<u64>(2, (needMain + needMain / 2 + (budget - needKey2) - 1) / (budget - needKey2)));
            if (Rtot > cdgsort::MAX_N) M = std::max<int>(M, (int) (Rtot / (cdgsort::MAX_N / 2)) + 1);
        }
        if (M > 1 && p->hash_lo == 0 && p->hash_hi >= 65535) {
            rc = run_kmermatch_split(ctx, p, K, M);
            if (rc != -100) return rc;
        }
    }
    if (Rtot > cdgsort::MAX_N) return fail(ctx, CDG_ERR_UNSUPPORTED, "%llu k-mer records exceed one sort (2^32) and no hash-range split applies", (unsigned long long) Rtot);
    CK(ctx->dKeysA.reserve(Rtot * 8)); CK(ctx->dKeysB.reserve(Rtot * 8)); CK(ctx->dValsA.reserve(Rtot * 8)); CK(ctx->dValsB.reserve(Rtot * 8));
    CK(ctx->dIdKeysA.reserve(n * 8)); CK(ctx->dIdKeysB.reserve(n * 8)); CK(ctx->dIdValsA.reserve(n * 8)); CK(ctx->dIdValsB.reserve(n * 8));
    CK(cudaMemsetAsync(ctx->dKeysA.as<u64>() + K.slots, 0xFF, (size_t) K.extraCap * 8, ctx->st));
    CK(cudaMemsetAsync(ctx->dValsA.as<u64>() + K.slots, 0, (size_t) K.extraCap * 8, ctx->st));
    if ((rc = km_extract(ctx, p, K, ctx->dKeysA.as<u64>(), ctx->dValsA.as<u64>(), ctx->dIdKeysA.as<u64>(), ctx->dIdValsA.as<u64>(), K.slots, 1, p->hash_lo,
                         p->hash_hi))) return rc;
    rec(ctx, EV_EXTRACT);
    ctx->idSrcK = ctx->dIdKeysA.as<u64>(); ctx->idSrcV = ctx->dIdValsA.as<u64>(); ctx->idAltK = ctx->dIdKeysB.as<u64>(); ctx->idAltV = ctx->dIdValsB.as<u64>();
    u64 minKey = 0;
    if ((rc = km_sort_group(ctx, p, K, ctx->dKeysA.as<u64>(), ctx->dKeysB.as<u64>(), ctx->dValsA.as<u64>(), ctx->dValsB.as<u64>(), Rtot, 0, &minKey, nullptr))) return rc;
    u64 R2 = 0;
    const u64 emitted = ctx->hScalars[SC_EMIT], extra = ctx->hScalars[SC_EXTRA];   // (fetched by km_sort_group with the overflow flag)
    const bool mainNonEmpty = minKey != CDG_SENTINEL;
    if ((rc = km_group(ctx, p, K, Rtot, mainNonEmpty, !mainNonEmpty, &R2))) return rc;
    u64 ownedIds = n;   // every sequence owns one identity record when the hash range is full
    if (!(p->hash_lo == 0 && p->hash_hi >= 65535) && n) {   // a sub-range call: the identity records whose hash lies in the range (:244)
        const u64 tiles = cdgscan::tiles_for(n);
        CK(ctx->dTileSums.reserve((tiles + 1) * 8));
        cdgkmer::NotSentinelF f{ctx->sortedIdK};
        cdgscan::tile_sum_kernel<<<(unsigned) tiles, cdgscan::BLOCK, 0, ctx->st>>>(f, n, ctx->dTileSums.as<unsigned long long>());
        cdgscan::tile_scan_kernel<<<1, 1024, 0, ctx->st>>>(ctx->dTileSums.as<unsigned long long>(), tiles, sc + SC_SCAN);
        ctx->launches += 2;
        if ((rc = read_scalar(ctx, SC_SCAN, &ownedIds))) return rc;
        ownedIds += (u32) extra;   // identity hashes inside the k-mer key range live in the main array
    }
    ctx->nRecords = emitted + ownedIds;
    ctx->nGrouped = R2;
    if ((rc = km_sort2(ctx, K, R2))) return rc;
    return km_reduce(ctx, K, R2, R2);
}

__global__ void __launch_bounds__(256) diag16_kernel(const i32 *__restrict__ d, u64 n, short *__restrict__ o) {
    u64 i = (u64) blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) o[i] = (short) d[i];
}

static int export_hits(cdg_ctx *ctx, cdg_hits *out) {
    const u64 H = ctx->nHits;
    CK(ctx->hHitRep.reserve((H + 1) * 4)); CK(ctx->hHitTgt.reserve((H + 1) * 4)); CK(ctx->hHitScore.reserve((H + 1) * 4)); CK(ctx->hHitDiag16.reserve((H + 1) * 2));
    if (H) {
        CK(ctx->dLineDiag.reserve((H + 1) * 2));   // scratch for the int16 view (rebuilt by the line builder later)
        diag16_kernel<<<(unsigned) ((H + 255) / 256), 256, 0, ctx->st>>>(ctx->dHitDiag.as<i32>(), H, (short *) ctx->dLineDiag.p);
        ctx->launches++;
        CK(cudaMemcpyAsync(ctx->hHitRep.p, ctx->dHitRep.p, H * 4, cudaMemcpyDeviceToHost, ctx->st));
        CK(cudaMemcpyAsync(ctx->hHitTgt.p, ctx->dHitTgt.p, H * 4, cudaMemcpyDeviceToHost, ctx->st));
        CK(cudaMemcpyAsync(ctx->hHitScore.p, ctx->dHitScore.p, H * 4, cudaMemcpyDeviceToHost, ctx->st));
        CK(cudaMemcpyAsync(ctx->hHitDiag16.p, ctx->dLineDiag.p, H * 2, cudaMemcpyDeviceToHost, ctx->st));
    }
    CK(cudaStreamSynchronize(ctx->st));
    out->n_hits = H; out->rep = ctx->hHitRep.as<u32>(); out->target = ctx->hHitTgt.as<u32>(); out->score = ctx->hHitScore.as<i32>();
    out->diag = ctx->hHitDiag16.as<int16_t>(); out->n_records = ctx->nRecords; out->n_grouped = ctx->nGrouped;
    return CDG_OK;
}

extern "C" int cdg_kmermatch(cdg_ctx *ctx, const cdg_kmer_params *p, cdg_hits *out) {
    if (!ctx || !p) return CDG_ERR_ARG;
    if (ctx->n == 0) return fail(ctx, CDG_ERR_ARG, "cdg_kmermatch: no sequence set uploaded");
    CK(cudaSetDevice(ctx->device));
    int rc = run_kmermatch(ctx, p);
    if (rc) return rc;
    if (out) return export_hits(ctx, out);
    CK(cudaStreamSynchronize(ctx->st));
    return CDG_OK;
}

// ------------------------------------------------------------------------------------------------------
static int reserve_lines(cdg_ctx *ctx, u64 M) {
    const u64 n = ctx->n;
    CK(ctx->dLineOff.reserve((n + 2) * 8)); CK(ctx->dLineQ.reserve((M + 1) * 4)); CK(ctx->dLineT.reserve((M + 1) * 4));
    CK(ctx->dLinePref.reserve((M + 1) * 4)); CK(ctx->dLineDiag.reserve((M + 1) * 2));
    CK(ctx->dKeep.reserve(M + 1)); CK(ctx->dScore.reserve((M + 1) * 4)); CK(ctx->dBits.reserve((M + 1) * 4)); CK(ctx->dEvalue.reserve((M + 1) * 8));
    CK(ctx->dSeqId.reserve((M + 1) * 4)); CK(ctx->dQS.reserve((M + 1) * 4)); CK(ctx->dQE.reserve((M + 1) * 4)); CK(ctx->dDS.reserve((M + 1) * 4));
    CK(ctx->dDE.reserve((M + 1) * 4)); CK(ctx->dLineSel.reserve(M + 1)); CK(ctx->dIdCnt.reserve((M + 1) * 4));
    return 0;
}

static int build_lines_from_hits(cdg_ctx *ctx) {
    const u64 nQ = ctx->nQ, qLo = ctx->qLo, H = ctx->nHits, M = nQ + H;
    int rc = reserve_lines(ctx, M);
    if (rc) return rc;
    unsigned long long *sc = ctx->dScalars.as<unsigned long long>();
    CK(ctx->dQList.reserve((nQ + 1) * 4));
    CK(cudaMemsetAsync(sc + SC_QLIST, 0, 8, ctx->st));
    cdgalign::lines_self_kernel<<<(unsigned) ((nQ + 1 + 255) / 256), 256, 0, ctx->st>>>(ctx->dHitRep.as<u32>(), H, qLo, nQ, ctx->dLineOff.as<u64>(),
                                                                                        ctx->dLineQ.as<u32>(), ctx->dLineT.as<u32>(),
                                                                                        ctx->dLinePref.as<i32>(), ctx->dLineDiag.as<u16>(),
                                                                                        ctx->dQList.as<u32>(), (u32 *) (sc + SC_QLIST));
    ctx->qListValid = true;   // run_correct takes the list of multi-line queries from here
    ctx->launches++;
    if (H) {
        cdgalign::lines_hits_kernel<<<(unsigned) ((H + 255) / 256), 256, 0, ctx->st>>>(ctx->dHitRep.as<u32>(), ctx->dHitTgt.as<u32>(), ctx->dHitScore.as<i32>(),
                                                                                     ctx->dHitDiag.as<i32>(), H, qLo, ctx->dLineQ.as<u32>(), ctx->dLineT.as<u32>(),
                                                                                     ctx->dLinePref.as<i32>(), ctx->dLineDiag.as<u16>());
        ctx->launches++;
    }
    CKL();
    ctx->nLines = M; ctx->linesValid = true;
    return 0;
}

// caller-provided CSR lists (parsed from a pref / aln DB): offsets must be monotone and end at n_lines, targets must be
// indices of the resident set — the kernels index with them
static int validate_csr(cdg_ctx *ctx, const char *what, const uint64_t *off, uint64_t nq, const uint32_t *target, uint64_t n_lines) {
    if (off[0] != 0 || off[nq] != n_lines) return fail(ctx, CDG_ERR_ARG, "%s: offsets must start at 0 and end at the number of lines", what);
    for (uint64_t q = 0; q < nq; q++)
        if (off[q + 1] < off[q]) return fail(ctx, CDG_ERR_ARG, "%s: offsets decrease at query %llu", what, (unsigned long long) q);
    if (n_lines && !target) return fail(ctx, CDG_ERR_ARG, "%s: no target array", what);
    for (uint64_t l = 0; l < n_lines; l++)
        if (target[l] >= ctx->n) return fail(ctx, CDG_ERR_ARG, "%s: line %llu: target index %u outside the resident set (%llu sequences)", what,
                                             (unsigned long long) l, target[l], (unsigned long long) ctx->n);
    return 0;
}

extern "C" int cdg_set_pref_lines(cdg_ctx *ctx, const uint64_t *line_off, uint64_t n_queries, const uint32_t *target, const int32_t *pref_score,
                                  const uint16_t *diag, uint64_t n_lines) {
    if (!ctx || !line_off) return CDG_ERR_ARG;
    if (n_queries != ctx->n) return fail(ctx, CDG_ERR_ARG, "cdg_set_pref_lines: %llu queries, resident set has %llu", (unsigned long long) n_queries,
                                         (unsigned long long) ctx->n);
    CK(cudaSetDevice(ctx->device));
    if (int v = validate_csr(ctx, "cdg_set_pref_lines", line_off, n_queries, target, n_lines)) return v;
    int rc = reserve_lines(ctx, n_lines);
    if (rc) return rc;
    CK(cudaMemcpyAsync(ctx->dLineOff.p, line_off, (n_queries + 1) * 8, cudaMemcpyHostToDevice, ctx->st));
    if (n_lines) {
        CK(cudaMemcpyAsync(ctx->dLineT.p, target, n_lines * 4, cudaMemcpyHostToDevice, ctx->st));
        CK(cudaMemcpyAsync(ctx->dLinePref.p, pref_score, n_lines * 4, cudaMemcpyHostToDevice, ctx->st));
        CK(cudaMemcpyAsync(ctx->dLineDiag.p, diag, n_lines * 2, cudaMemcpyHostToDevice, ctx->st));
    }
    cdgalign::lines_fillq_kernel<<<(unsigned) ((n_queries + 255) / 256), 256, 0, ctx->st>>>(ctx->dLineOff.as<u64>(), n_queries, ctx->dLineQ.as<u32>());
    ctx->launches++;
    CKL();
    CK(cudaStreamSynchronize(ctx->st));
    ctx->nLines = n_lines; ctx->linesValid = true; ctx->alnsValid = false; ctx->qLo = 0; ctx->nQ = ctx->n; ctx->qListValid = false;
    return CDG_OK;
}

static int run_rescore(cdg_ctx *ctx, const cdg_rescore_params *p) {
    if (int rcp = ensure_packed(ctx)) return rcp;
    if (!ctx->linesValid) {
        if (!ctx->hitsValid) return fail(ctx, CDG_ERR_ARG, "cdg_rescore: no prefilter lines (run cdg_kmermatch or cdg_set_pref_lines)");
        int rc = build_lines_from_hits(ctx);
        if (rc) return rc;
    }
    const u64 M = ctx->nLines;
    unsigned long long *sc = ctx->dScalars.as<unsigned long long>();
    CK(cudaMemsetAsync(sc + SC_KEPT, 0, 8, ctx->st));
    rec(ctx, EV_RS0);
    cdgalign::RescoreParams R; R.seqIdThr = p->min_seq_id; R.evalThr = p->evalue_thr;
    R.dbRes = (double) (p->db_residues ? p->db_residues : ctx->totalBases);
    if (M) {
        cdgalign::rescore_kernel<<<(unsigned) ((M + 255) / 256), 256, 0, ctx->st>>>(
            ctx->dData.as<u8>(), ctx->dMeta.as<cdgalign::SeqMeta>(), ctx->dWords.as<u64>(),
            ctx->dLineQ.as<u32>(), ctx->dLineT.as<u32>(), ctx->dLinePref.as<i32>(), ctx->dLineDiag.as<u16>(), M, R, ctx->dKeep.as<u8>(),
            ctx->dScore.as<i32>(), ctx->dBits.as<i32>(), ctx->dEvalue.as<double>(), ctx->dSeqId.as<float>(), ctx->dQS.as<i32>(), ctx->dQE.as<i32>(),
            ctx->dDS.as<i32>(), ctx->dDE.as<i32>(), ctx->dIdCnt.as<u32>(), sc + SC_KEPT);
        ctx->launches++;
        CKL();
    }
    ctx->dbResUsed = R.dbRes;
    rec(ctx, EV_RS1);
    ctx->alnsValid = true; ctx->alnsFromText = false;
    return CDG_OK;
}

// overlap = true: the copies run on the copy stream behind the rescore kernel; the caller finishes with wait_exports()
static int export_alns(cdg_ctx *ctx, cdg_alns *out, bool overlap = false) {
    const u64 M = ctx->nLines, n = ctx->nQ;
    cudaStream_t cs = ctx->st;
    CK(ctx->hLineOff.reserve((n + 1) * 8)); CK(ctx->hLineT.reserve((M + 1) * 4)); CK(ctx->hKeep.reserve(M + 1)); CK(ctx->hScore.reserve((M + 1) * 4));
    CK(ctx->hBits.reserve((M + 1) * 4)); CK(ctx->hEvalue.reserve((M + 1) * 8)); CK(ctx->hSeqId.reserve((M + 1) * 4)); CK(ctx->hQS.reserve((M + 1) * 4));
    CK(ctx->hQE.reserve((M + 1) * 4)); CK(ctx->hDS.reserve((M + 1) * 4)); CK(ctx->hDE.reserve((M + 1) * 4));
    u64 kept = 0;
    int rc = read_scalar(ctx, SC_KEPT, &kept);   // synchronises the compute stream: rescore has finished
    if (rc) return rc;
    ctx->nKept = kept;
    if (overlap) cs = ctx->copySt;
    CK(cudaMemcpyAsync(ctx->hLineOff.p, ctx->dLineOff.p, (n + 1) * 8, cudaMemcpyDeviceToHost, cs));
    if (M) {
        CK(cudaMemcpyAsync(ctx->hLineT.p, ctx->dLineT.p, M * 4, cudaMemcpyDeviceToHost, cs));
        CK(cudaMemcpyAsync(ctx->hKeep.p, ctx->dKeep.p, M, cudaMemcpyDeviceToHost, cs));
        CK(cudaMemcpyAsync(ctx->hScore.p, ctx->dScore.p, M * 4, cudaMemcpyDeviceToHost, cs));
        CK(cudaMemcpyAsync(ctx->hBits.p, ctx->dBits.p, M * 4, cudaMemcpyDeviceToHost, cs));
        CK(cudaMemcpyAsync(ctx->hEvalue.p, ctx->dEvalue.p, M * 8, cudaMemcpyDeviceToHost, cs));
        CK(cudaMemcpyAsync(ctx->hSeqId.p, ctx->dSeqId.p, M * 4, cudaMemcpyDeviceToHost, cs));
        CK(cudaMemcpyAsync(ctx->hQS.p, ctx->dQS.p, M * 4, cudaMemcpyDeviceToHost, cs));
        CK(cudaMemcpyAsync(ctx->hQE.p, ctx->dQE.p, M * 4, cudaMemcpyDeviceToHost, cs));
        CK(cudaMemcpyAsync(ctx->hDS.p, ctx->dDS.p, M * 4, cudaMemcpyDeviceToHost, cs));
        CK(cudaMemcpyAsync(ctx->hDE.p, ctx->dDE.p, M * 4, cudaMemcpyDeviceToHost, cs));
    }
    if (overlap) { CK(cudaEventRecord(ctx->evCopied, cs)); ctx->copyPending = true; } else CK(cudaStreamSynchronize(cs));
    out->n_lines = M; out->n_queries = n; out->line_off = ctx->hLineOff.as<u64>(); out->target = ctx->hLineT.as<u32>(); out->keep = ctx->hKeep.as<u8>();
    out->score = ctx->hScore.as<i32>(); out->bits = ctx->hBits.as<i32>(); out->evalue = ctx->hEvalue.as<double>(); out->seq_id = ctx->hSeqId.as<float>();
    out->q_start = ctx->hQS.as<i32>(); out->q_end = ctx->hQE.as<i32>(); out->db_start = ctx->hDS.as<i32>(); out->db_end = ctx->hDE.as<i32>();
    return CDG_OK;
}


// ---- compact exports -------------------------------------------------------------------------------------------------------
// Alignments the reference writes (keep == 1), 14 bytes each when every sequence is shorter than 32768, implicit self lines
// dropped.  overlap: the device-to-host copies run on the copy stream (the caller finishes with wait_exports()).
static int export_alns_compact(cdg_ctx *ctx, cdg_alns_compact *out, bool overlap) {
    if (!ctx->alnsValid || ctx->alnsFromText) return fail(ctx, CDG_ERR_ARG, "compact alignment export needs the alignments of cdg_rescore on this context");
    const u64 M = ctx->nLines, nQ = ctx->nQ;
    unsigned long long *sc = ctx->dScalars.as<unsigned long long>();
    cdgexp::ExportFlag f;
    f.A = cdgexp::LineArrays{ctx->dLineQ.as<u32>(), ctx->dLineT.as<u32>(), ctx->dLinePref.as<i32>(), ctx->dKeep.as<u8>(), ctx->dScore.as<i32>(), ctx->dIdCnt.as<u32>(),
                             ctx->dSeqId.as<float>(), ctx->dQS.as<i32>(), ctx->dQE.as<i32>(), ctx->dDS.as<i32>(), ctx->dDE.as<i32>(), ctx->dMeta.as<SeqMeta>()};
    const u64 tiles = cdgscan::tiles_for(M);
    CK(ctx->dTileSums.reserve((tiles + 1) * 8));
    unsigned long long *ts = ctx->dTileSums.as<unsigned long long>();
    u64 A = 0, kept = 0;
    if (M) {
        cdgscan::tile_sum_kernel<<<(unsigned) tiles, cdgscan::BLOCK, 0, ctx->st>>>(f, M, ts);
        cdgscan::tile_scan_kernel<<<1, 1024, 0, ctx->st>>>(ts, tiles, sc + SC_SCAN);
        ctx->launches += 2;
        CKL();
        u64 h[2];
        CK(cudaMemcpyAsync(&h[0], sc + SC_SCAN, 8, cudaMemcpyDeviceToHost, ctx->st));
        CK(cudaMemcpyAsync(&h[1], sc + SC_KEPT, 8, cudaMemcpyDeviceToHost, ctx->st));
        CK(cudaStreamSynchronize(ctx->st));
        A = h[0]; kept = h[1];
    }
    ctx->nKept = kept; ctx->nCompact = A;
    const int width = ctx->maxLen < 32768 ? 2 : 4;
    const u64 fieldBytes = (A + 8) * (u64) width;   // 5 arrays of A elements, each padded to 16 bytes
    const u64 stride = (fieldBytes + 15) & ~15ull;
    CK(ctx->dCmpCount.reserve((nQ + 1) * 4)); CK(ctx->dCmpTarget.reserve((A + 1) * 4)); CK(ctx->dCmpFields.reserve(5 * stride));
    CK(ctx->hCmpCount.reserve((nQ + 1) * 4)); CK(ctx->hCmpTarget.reserve((A + 1) * 4)); CK(ctx->hCmpFields.reserve(5 * stride));
    u8 *df = ctx->dCmpFields.as<u8>();
    if (M) {
        if (width == 2)
            cdgexp::compact_write_kernel<u16, short><<<(unsigned) tiles, cdgscan::BLOCK, 0, ctx->st>>>(f, M, ts, ctx->dCmpTarget.as<u32>(), (u16 *) df, (u16 *) (df + stride),
                                                                                                     (short *) (df + 2 * stride), (short *) (df + 3 * stride), (u16 *) (df + 4 * stride));
        else
            cdgexp::compact_write_kernel<u32, i32><<<(unsigned) tiles, cdgscan::BLOCK, 0, ctx->st>>>(f, M, ts, ctx->dCmpTarget.as<u32>(), (u32 *) df, (u32 *) (df + stride),
                                                                                                   (i32 *) (df + 2 * stride), (i32 *) (df + 3 * stride), (u32 *) (df + 4 * stride));
        ctx->launches++;
    }
    if (nQ) { cdgexp::compact_count_kernel<<<(unsigned) ((nQ + 255) / 256), 256, 0, ctx->st>>>(f, ctx->dLineOff.as<u64>(), nQ, ctx->dCmpCount.as<u32>()); ctx->launches++; }
    CKL();
    cudaStream_t cs = ctx->st;
    if (overlap) { CK(cudaEventRecord(ctx->evReady, ctx->st)); CK(cudaStreamWaitEvent(ctx->copySt, ctx->evReady, 0)); cs = ctx->copySt; }
    if (nQ) CK(cudaMemcpyAsync(ctx->hCmpCount.p, ctx->dCmpCount.p, nQ * 4, cudaMemcpyDeviceToHost, cs));
    if (A) {
        CK(cudaMemcpyAsync(ctx->hCmpTarget.p, ctx->dCmpTarget.p, A * 4, cudaMemcpyDeviceToHost, cs));
        CK(cudaMemcpyAsync(ctx->hCmpFields.p, ctx->dCmpFields.p, 4 * stride + A * (u64) width, cudaMemcpyDeviceToHost, cs));
    }
    if (overlap) { CK(cudaEventRecord(ctx->evCopied, cs)); ctx->copyPending = true; } else CK(cudaStreamSynchronize(cs));
    const u8 *hf = ctx->hCmpFields.as<u8>();
    out->n_queries = nQ; out->n_alns = A; out->first_query = ctx->qLo; out->count = ctx->hCmpCount.as<u32>(); out->target = ctx->hCmpTarget.as<u32>();
    out->width = width; out->score = hf; out->id_cnt = hf + stride; out->q_start = hf + 2 * stride; out->db_start = hf + 3 * stride;
    out->aln_len_rev = hf + 4 * stride; out->db_residues = (u64) ctx->dbResUsed;
    return CDG_OK;
}

// corrected reads of the byte range [b0, b1) as patches (the last cdg_correct's output against the resident input)
static int export_patches(cdg_ctx *ctx, u64 b0, u64 b1, u64 nChanged, cdg_patches *out) {
    unsigned long long *sc = ctx->dScalars.as<unsigned long long>();
    CK(ctx->dPatches.reserve((nChanged + 1) * 8)); CK(ctx->hPatches.reserve((nChanged + 1) * 8));
    u64 got = 0;
    if (ctx->patchesDirect) {   // the correction kernel wrote the list itself (all of it lies in the context's query range)
        if (nChanged) CK(cudaMemcpyAsync(ctx->hPatches.p, ctx->dPatches.p, nChanged * 8, cudaMemcpyDeviceToHost, ctx->st));
        CK(cudaStreamSynchronize(ctx->st));
        got = nChanged;
    } else if (nChanged && b1 > b0) {
        CK(cudaMemsetAsync(sc + SC_SCAN, 0, 8, ctx->st));
        const u64 chunks = ((b1 + 15) >> 4) - (b0 >> 4);
        const unsigned blocks = (unsigned) std::min<u64>((chunks + 255) / 256, (u64) ctx->smCount * 16);
        cdgexp::patch_kernel<<<blocks, 256, 0, ctx->st>>>(ctx->dData.as<u8>(), ctx->dOut.as<u8>(), b0, b1, ctx->dPatches.as<u64>(), nChanged, sc + SC_SCAN);
        ctx->launches++;
        CKL();
        CK(cudaMemcpyAsync(ctx->hPatches.p, ctx->dPatches.p, nChanged * 8, cudaMemcpyDeviceToHost, ctx->st));
        int rc = read_scalar(ctx, SC_SCAN, &got);
        if (rc) return rc;
        if (got != nChanged) return fail(ctx, CDG_ERR_CUDA, "patch export: %llu differing bytes, the correction counted %llu", (unsigned long long) got, (unsigned long long) nChanged);
    }
    out->n = got; out->entry = ctx->hPatches.as<u64>();
    return CDG_OK;
}

extern "C" int cdg_export_alns_compact(cdg_ctx *ctx, cdg_alns_compact *out) {
    if (!ctx || !out) return CDG_ERR_ARG;
    CK(cudaSetDevice(ctx->device));
    return export_alns_compact(ctx, out, false);
}

extern "C" int cdg_rescore(cdg_ctx *ctx, const cdg_rescore_params *p, cdg_alns *out) {
    if (!ctx || !p) return CDG_ERR_ARG;
    if (ctx->n == 0) return fail(ctx, CDG_ERR_ARG, "cdg_rescore: no sequence set uploaded");
    CK(cudaSetDevice(ctx->device));
    int rc = run_rescore(ctx, p);
    if (rc) return rc;
    if (out) return export_alns(ctx, out);
    u64 kept = 0;
    rc = read_scalar(ctx, SC_KEPT, &kept);
    ctx->nKept = kept;
    return rc;
}

// ---- N3: cyclecheck (cdg_cycle.cuh) ------------------------------------------------------------------------------------------
struct CycleWindowsF { const u32 *len; int k; u32 maxSeqLen; __device__ u32 operator()(u64 i) const { const u32 L = len[i]; return (L < maxSeqLen && L >= (u32) k) ? L - (u32) k + 1u : 0u; } };
struct CycleDiagsF { const u32 *len; __device__ u32 operator()(u64 i) const { return 2u * (len[i] / 3u) + 1u; } };

extern "C" int cdg_cyclecheck(cdg_ctx *ctx, const cdg_cycle_params *p, cdg_cycle_out *out) {
    if (!ctx || !p || !out) return CDG_ERR_ARG;
    if (ctx->n == 0) return fail(ctx, CDG_ERR_ARG, "cdg_cyclecheck: no resident sequence set");
    if (p->k < 1 || p->k > 31) return fail(ctx, CDG_ERR_ARG, "cdg_cyclecheck: k = %d (1..31; the reference pins 22)", p->k);
    CK(cudaSetDevice(ctx->device));
    if (int rcb = ensure_bytes(ctx)) return rcb;
    const u64 n = ctx->n;
    unsigned long long *sc = ctx->dScalars.as<unsigned long long>();
    int rc;
    // record and diagonal-counter offsets per sequence
    CK(ctx->dCycleRecOff.reserve((n + 1) * 8)); CK(ctx->dCycleDiagOff.reserve((n + 1) * 8)); CK(ctx->dCycleSplit.reserve((n + 1) * 4));
    u64 R = 0, D = 0;
    if ((rc = scan_to(ctx, CycleWindowsF{ctx->dLen.as<u32>(), p->k, p->max_seq_len}, n, ctx->dCycleRecOff.as<u64>(), true))) return rc;
    if ((rc = read_scalar(ctx, SC_SCAN, &R))) return rc;
    if ((rc = scan_to(ctx, CycleDiagsF{ctx->dLen.as<u32>()}, n, ctx->dCycleDiagOff.as<u64>(), true))) return rc;
    if ((rc = read_scalar(ctx, SC_SCAN, &D))) return rc;
    if (R > cdgsort::MAX_N) return fail(ctx, CDG_ERR_UNSUPPORTED, "cdg_cyclecheck: %llu k-mers exceed one sort (2^32): call it per block of sequences", (unsigned long long) R);
    CK(ctx->dKeysA.reserve((R + 1) * 8)); CK(ctx->dKeysB.reserve((R + 1) * 8)); CK(ctx->dValsA.reserve((R + 1) * 8)); CK(ctx->dValsB.reserve((R + 1) * 8));
    CK(ctx->dCycleHits.reserve((D + 1) * 4));
    CK(cudaMemsetAsync(ctx->dCycleHits.p, 0, (D + 1) * 4, ctx->st));
    const unsigned wgrid = (unsigned) std::min<u64>((n * 32 + 255) / 256, (u64) ctx->smCount * 16);
    if (R) {
        cdgcycle::cycle_emit_kernel<<<wgrid, 256, 0, ctx->st>>>(ctx->dData.as<u8>(), ctx->dOff.as<u64>(), ctx->dLen.as<u32>(), n, p->k, p->max_seq_len,
                                                               ctx->dCycleRecOff.as<u64>(), ctx->dKeysA.as<u64>(), ctx->dValsA.as<u64>());
        ctx->launches++;
        CKL();
        cdgsort::Workspace ws;
        if ((rc = sort_ws(ctx, R, ws))) return rc;
        // k-mer index < 5 * 4^k / 3 (letters 0..4): 2k + 1 key bits
        cdgsort::Plan plan = cdgsort::make_plan(0, 0, 0, 2 * p->k + 1);
        u64 *sk = nullptr, *sv = nullptr;
        const int l = cdgsort::radix_sort(ctx->st, ws, ctx->dKeysA.as<u64>(), ctx->dKeysB.as<u64>(), ctx->dValsA.as<u64>(), ctx->dValsB.as<u64>(), R, plan, &sk, &sv, ctx->smCount);
        if (l < 0) return fail(ctx, CDG_ERR_UNSUPPORTED, "cdg_cyclecheck: radix sort: input too large for one pass set");
        ctx->launches += l;
        cdgcycle::cycle_match_kernel<<<(unsigned) ((R + 255) / 256), 256, 0, ctx->st>>>(sk, sv, R, ctx->dLen.as<u32>(), ctx->dCycleDiagOff.as<u64>(), ctx->dCycleHits.as<u32>(),
                                                                                       cdgcycle::CYCLE_WALK_CAP);
        cdgcycle::cycle_match_head_kernel<<<(unsigned) ((R + 255) / 256), 256, 0, ctx->st>>>(sk, sv, R, ctx->dLen.as<u32>(), ctx->dCycleDiagOff.as<u64>(),
                                                                                            ctx->dCycleHits.as<u32>(), cdgcycle::CYCLE_WALK_CAP);
        ctx->launches += 2;
        CKL();
    }
    cdgcycle::cycle_band_kernel<<<wgrid, 256, 0, ctx->st>>>(ctx->dLen.as<u32>(), n, p->k, p->max_seq_len, ctx->dCycleDiagOff.as<u64>(), ctx->dCycleHits.as<u32>(),
                                                           ctx->dCycleSplit.as<u32>());
    ctx->launches++;
    CKL();
    CK(ctx->hExt.reserve(n * 4 + 16));
    CK(cudaMemcpyAsync(ctx->hExt.p, ctx->dCycleSplit.p, n * 4, cudaMemcpyDeviceToHost, ctx->st));
    CK(cudaStreamSynchronize(ctx->st));
    ctx->hitsValid = false; ctx->linesValid = false;   // the record arrays of the k-mer matcher were reused
    out->n = n; out->n_kmers = R; out->split_diagonal = ctx->hExt.as<u32>();
    u64 nc = 0;
    for (u64 i = 0; i < n; i++) nc += out->split_diagonal[i] != 0;
    out->n_cyclic = nc;
    return CDG_OK;
}

// ---- N4: rescorediagonal --rescore-mode 0 (linclust's Hamming pre-clustering), cdg_hamming.cuh ------------------------------------
extern "C" int cdg_rescore_hamming(cdg_ctx *ctx, const cdg_hamming_params *p, cdg_hamming_out *out) {
    if (!ctx || !p || !out) return CDG_ERR_ARG;
    if (ctx->n == 0) return fail(ctx, CDG_ERR_ARG, "cdg_rescore_hamming: no sequence set uploaded");
    if (p->cov_mode < 0 || p->cov_mode > 5 || p->seq_id_mode < 0 || p->seq_id_mode > 2)
        return fail(ctx, CDG_ERR_ARG, "cdg_rescore_hamming: --cov-mode %d / --seq-id-mode %d unknown", p->cov_mode, p->seq_id_mode);
    CK(cudaSetDevice(ctx->device));
    if (int rcp = ensure_packed(ctx)) return rcp;
    if (!ctx->linesValid) {
        if (!ctx->hitsValid) return fail(ctx, CDG_ERR_ARG, "cdg_rescore_hamming: no prefilter lines (run cdg_kmermatch or cdg_set_pref_lines)");
        if (int rc = build_lines_from_hits(ctx)) return rc;
    }
    const u64 M = ctx->nLines, n = ctx->nQ;
    unsigned long long *sc = ctx->dScalars.as<unsigned long long>();
    CK(cudaMemsetAsync(sc + SC_KEPT, 0, 8, ctx->st));
    cdgham::HammingParams H; H.covThr = p->cov_thr; H.seqIdThr = p->min_seq_id; H.covMode = p->cov_mode; H.seqIdMode = p->seq_id_mode;
    H.alnLenThr = p->min_aln_len; H.wrapped = p->wrapped_scoring;
    if (M) {
        cdgham::hamming_kernel<<<(unsigned) ((M + 255) / 256), 256, 0, ctx->st>>>(
            ctx->dData.as<u8>(), ctx->dMeta.as<cdgalign::SeqMeta>(), ctx->dWords.as<u64>(), ctx->dLineQ.as<u32>(), ctx->dLineT.as<u32>(), ctx->dLinePref.as<i32>(),
            ctx->dLineDiag.as<u16>(), M, H, ctx->dKeep.as<u8>(), ctx->dScore.as<i32>(), ctx->dQS.as<i32>(), ctx->dSeqId.as<float>(), sc + SC_KEPT);
        ctx->launches++;
        CKL();
    }
    ctx->alnsValid = false;   // the alignment arrays now hold Hamming results
    CK(ctx->hLineOff.reserve((n + 1) * 8)); CK(ctx->hLineT.reserve((M + 1) * 4)); CK(ctx->hKeep.reserve(M + 1)); CK(ctx->hScore.reserve((M + 1) * 4));
    CK(ctx->hQS.reserve((M + 1) * 4)); CK(ctx->hSeqId.reserve((M + 1) * 4));
    u64 kept = 0;
    if (int rc = read_scalar(ctx, SC_KEPT, &kept)) return rc;
    CK(cudaMemcpyAsync(ctx->hLineOff.p, ctx->dLineOff.p, (n + 1) * 8, cudaMemcpyDeviceToHost, ctx->st));
    if (M) {
        CK(cudaMemcpyAsync(ctx->hLineT.p, ctx->dLineT.p, M * 4, cudaMemcpyDeviceToHost, ctx->st));
        CK(cudaMemcpyAsync(ctx->hKeep.p, ctx->dKeep.p, M, cudaMemcpyDeviceToHost, ctx->st));
        CK(cudaMemcpyAsync(ctx->hScore.p, ctx->dScore.p, M * 4, cudaMemcpyDeviceToHost, ctx->st));
        CK(cudaMemcpyAsync(ctx->hQS.p, ctx->dQS.p, M * 4, cudaMemcpyDeviceToHost, ctx->st));
        CK(cudaMemcpyAsync(ctx->hSeqId.p, ctx->dSeqId.p, M * 4, cudaMemcpyDeviceToHost, ctx->st));
    }
    CK(cudaStreamSynchronize(ctx->st));
    out->n_lines = M; out->n_queries = n; out->n_kept = kept; out->line_off = ctx->hLineOff.as<u64>(); out->target = ctx->hLineT.as<u32>();
    out->keep = ctx->hKeep.as<u8>(); out->pref_score = ctx->hScore.as<i32>(); out->diag = ctx->hQS.as<i32>(); out->seq_id = ctx->hSeqId.as<float>();
    return CDG_OK;
}

// ------------------------------------------------------------------------------------------------------
extern "C" int cdg_set_damage(cdg_ctx *ctx, const cdg_damage *d) {
    if (!ctx || !d) return CDG_ERR_ARG;
    CK(cudaSetDevice(ctx->device));
    static_assert(sizeof(cdg_damage) == sizeof(cdgalign::DamageTables), "damage table layout");
    CK(ctx->dTables.reserve(sizeof(cdgalign::DamageTables)));
    CK(cudaMemcpyAsync(ctx->dTables.p, d, sizeof(cdg_damage), cudaMemcpyHostToDevice, ctx->st));
    CK(cudaStreamSynchronize(ctx->st));
    ctx->damageSet = true;
    return CDG_OK;
}

extern "C" int cdg_set_alignments(cdg_ctx *ctx, const uint64_t *aln_off, uint64_t n_queries, const uint32_t *target, const float *seq_id,
                                  const int32_t *q_start, const int32_t *q_end, const int32_t *db_start, const int32_t *db_end, uint64_t n_alns) {
    if (!ctx || !aln_off) return CDG_ERR_ARG;
    if (n_queries != ctx->n) return fail(ctx, CDG_ERR_ARG, "cdg_set_alignments: %llu queries, resident set has %llu", (unsigned long long) n_queries,
                                         (unsigned long long) ctx->n);
    CK(cudaSetDevice(ctx->device));
    if (int v = validate_csr(ctx, "cdg_set_alignments", aln_off, n_queries, target, n_alns)) return v;
    int rc = reserve_lines(ctx, n_alns);
    if (rc) return rc;
    CK(cudaMemcpyAsync(ctx->dLineOff.p, aln_off, (n_queries + 1) * 8, cudaMemcpyHostToDevice, ctx->st));
    if (n_alns) {
        CK(cudaMemcpyAsync(ctx->dLineT.p, target, n_alns * 4, cudaMemcpyHostToDevice, ctx->st));
        CK(cudaMemcpyAsync(ctx->dSeqId.p, seq_id, n_alns * 4, cudaMemcpyHostToDevice, ctx->st));
        CK(cudaMemcpyAsync(ctx->dQS.p, q_start, n_alns * 4, cudaMemcpyHostToDevice, ctx->st));
        CK(cudaMemcpyAsync(ctx->dQE.p, q_end, n_alns * 4, cudaMemcpyHostToDevice, ctx->st));
        CK(cudaMemcpyAsync(ctx->dDS.p, db_start, n_alns * 4, cudaMemcpyHostToDevice, ctx->st));
        CK(cudaMemcpyAsync(ctx->dDE.p, db_end, n_alns * 4, cudaMemcpyHostToDevice, ctx->st));
        CK(cudaMemsetAsync(ctx->dKeep.p, 1, n_alns, ctx->st));
        cdgalign::lines_fillq_kernel<<<(unsigned) ((n_queries + 255) / 256), 256, 0, ctx->st>>>(ctx->dLineOff.as<u64>(), n_queries, ctx->dLineQ.as<u32>());
        ctx->launches++;
        CKL();
    }
    CK(cudaStreamSynchronize(ctx->st));
    ctx->nLines = n_alns; ctx->linesValid = false; ctx->hitsValid = false; ctx->alnsValid = true; ctx->alnsFromText = true; ctx->qLo = 0; ctx->nQ = ctx->n;
    ctx->qListValid = false;
    return CDG_OK;
}

static int wait_exports(cdg_ctx *ctx) {
    if (ctx->copyPending) { CK(cudaEventSynchronize(ctx->evCopied)); ctx->copyPending = false; }
    return 0;
}

// patchMode: corrected bases are emitted as patches (byte offset << 8 | letter) straight from the kernel, no corrected copy of the
// DB bytes exists (cdg_iteration_compact, cdg_mg_finish without a byte output); a packed upload then never materialises bytes.
// Falls back to the byte mode when the generic wide-counter kernel is needed or the patch list overflows.
static int run_correct(cdg_ctx *ctx, const cdg_corr_params *p, uint8_t *out_data, uint64_t *n_changed, double *lik, bool patchMode = false) {
    if (int rcp = ensure_packed(ctx)) return rcp;
    if (!ctx->alnsValid) return fail(ctx, CDG_ERR_ARG, "cdg_correct: no alignments (run cdg_rescore or cdg_set_alignments)");
    if (!ctx->damageSet) return fail(ctx, CDG_ERR_ARG, "cdg_correct: cdg_set_damage not called");
    if (out_data) patchMode = false;
    if (!patchMode) if (int rcb = ensure_bytes(ctx)) return rcb;
    const u64 n = ctx->nQ, qLo = ctx->qLo;
    unsigned long long *sc = ctx->dScalars.as<unsigned long long>();
    const u64 patchCap = patchMode ? std::max<u64>(1ull << 20, ctx->totalBases / 32) : 0;
    if (patchMode) { CK(ctx->dPatches.reserve((patchCap + 1) * 8)); CK(cudaMemsetAsync(sc + SC_PATCH, 0, 8, ctx->st)); }
    else CK(ctx->dOut.reserve(ctx->dataBytes + 64));
    CK(cudaMemsetAsync(sc + SC_CHANGED, 0, 8, ctx->st));
    rec(ctx, EV_CO0);
    if (!patchMode) CK(cudaMemcpyAsync(ctx->dOut.p, ctx->dData.p, ctx->dataBytes, cudaMemcpyDeviceToDevice, ctx->st));
    double *dLik = nullptr;
    if (lik) {
        CK(ctx->dLik.reserve((ctx->totalBases + 1) * 4 * 8));
        CK(cudaMemsetAsync(ctx->dLik.p, 0xFF, ctx->totalBases * 4 * 8, ctx->st));   // all-ones = NaN
        dLik = ctx->dLik.as<double>();
    }
    cdgalign::CorrParams C; C.seqIdThr = p->min_seqid_corr_reads; C.ryThr = p->min_ryseq_id_corr_reads; C.roundSeqId = ctx->alnsFromText ? 0 : 1; C.wideThreshold = ctx->wideThreshold;
    const u64 wideCap = ctx->wideThreshold < 65535 ? ctx->n + 16 : (ctx->nLines >> 16) + 16;
    CK(ctx->dWideList.reserve(wideCap * 4));
    CK(cudaMemsetAsync(sc + SC_WIDE, 0, 8, ctx->st));
    if (!ctx->qListValid) {   // (lines built from the hit list come with their list: build_lines_from_hits)
        CK(ctx->dQList.reserve((n + 1) * 4));
        CK(cudaMemsetAsync(sc + SC_QLIST, 0, 8, ctx->st));
        if (n) cdgalign::multi_line_queries_kernel<<<(unsigned) ((n + 255) / 256), 256, 0, ctx->st>>>(ctx->dLineOff.as<u64>(), n, ctx->dQList.as<u32>(), (u32 *) (sc + SC_QLIST));
        ctx->launches++;
        CKL();
    }
    ctx->qListValid = false;   // consumed: SC_QLIST is shared scratch
    // per-line part of pass 2 (RY identity, thresholds, mismatch columns): one thread per line
    const u64 M = ctx->nLines;
    CK(ctx->dPrep.reserve(M + 1)); CK(ctx->dAlnLen.reserve((M + 1) * 4)); CK(ctx->dMism.reserve((M + 1) * 24));
    if (M) {
        cdgalign::corr_prep_kernel<<<(unsigned) ((M + 255) / 256), 256, 0, ctx->st>>>(
            ctx->dData.as<u8>(), ctx->dMeta.as<cdgalign::SeqMeta>(), ctx->dWords.as<u64>(), qLo, ctx->dLineOff.as<u64>(), ctx->dLineQ.as<u32>(), ctx->dLineT.as<u32>(),
            ctx->dKeep.as<u8>(), ctx->dSeqId.as<float>(), ctx->dQS.as<i32>(), ctx->dQE.as<i32>(), ctx->dDS.as<i32>(), ctx->dDE.as<i32>(), M, C,
            ctx->dPrep.as<u8>(), ctx->dAlnLen.as<u32>(), ctx->dMism.as<u64>());
        ctx->launches++;
        CKL();
    }
    const int blocks = ctx->smCount * 4;
    cdgalign::correct_fast_kernel<<<blocks, cdgalign::CF_WARPS * 32, sizeof(cdgalign::CorrSmem), ctx->st>>>(
        ctx->dQList.as<u32>(), (const u32 *) (sc + SC_QLIST), ctx->dData.as<u8>(), ctx->dMeta.as<cdgalign::SeqMeta>(), ctx->dWords.as<u64>(), qLo, n, ctx->dLineOff.as<u64>(), ctx->dLineT.as<u32>(), ctx->dKeep.as<u8>(), ctx->dSeqId.as<float>(), ctx->dQS.as<i32>(),
        ctx->dQE.as<i32>(), ctx->dDS.as<i32>(), ctx->dDE.as<i32>(), C, ctx->dTables.as<cdgalign::DamageTables>(), ctx->dLineSel.as<u8>(),
        patchMode ? nullptr : ctx->dOut.as<u8>(), sc + SC_CHANGED, ctx->dLikOff.as<u64>(), dLik, ctx->dWideList.as<u32>(), (u32 *) (sc + SC_WIDE),
        patchMode ? ctx->dPatches.as<u64>() : nullptr, patchCap, sc + SC_PATCH, ctx->dPrep.as<u8>(), ctx->dAlnLen.as<u32>(), ctx->dMism.as<u64>());
    ctx->launches++;
    CKL();
    if (patchMode) {
        u64 h[3];
        CK(cudaMemcpyAsync(&h[0], sc + SC_WIDE, 8, cudaMemcpyDeviceToHost, ctx->st));
        CK(cudaMemcpyAsync(&h[1], sc + SC_PATCH, 8, cudaMemcpyDeviceToHost, ctx->st));
        CK(cudaMemcpyAsync(&h[2], sc + SC_CHANGED, 8, cudaMemcpyDeviceToHost, ctx->st));
        rec(ctx, EV_CO1);
        CK(cudaStreamSynchronize(ctx->st));
        if ((u32) h[0] != 0 || h[1] > patchCap) return run_correct(ctx, p, nullptr, n_changed, lik, false);   // deep pile-ups / very many corrections: byte mode
        if (lik) { CK(cudaMemcpyAsync(lik, ctx->dLik.p, ctx->totalBases * 4 * 8, cudaMemcpyDeviceToHost, ctx->st)); CK(cudaStreamSynchronize(ctx->st)); }
        if (n_changed) *n_changed = h[2];
        ctx->nChangedLast = h[2]; ctx->correctedValid = true; ctx->patchesDirect = true;
        return CDG_OK;
    }
    // queries with more than 65535 counted targets (16-bit counters): generic kernel, usually an empty list
    cdgalign::correct_wide_kernel<<<ctx->smCount, cdgalign::CORR_WARPS * 32, 0, ctx->st>>>(
        ctx->dWideList.as<u32>(), (const u32 *) (sc + SC_WIDE),
        ctx->dData.as<u8>(), ctx->dOff.as<u64>(), ctx->dLen.as<u32>(), ctx->dExt.as<u8>(), qLo, n, ctx->dLineOff.as<u64>(), ctx->dLineT.as<u32>(),
        ctx->dKeep.as<u8>(), ctx->dSeqId.as<float>(), ctx->dQS.as<i32>(), ctx->dQE.as<i32>(), ctx->dDS.as<i32>(), ctx->dDE.as<i32>(), C,
        ctx->dTables.as<cdgalign::DamageTables>(), ctx->dLineSel.as<u8>(), ctx->dOut.as<u8>(), sc + SC_CHANGED, ctx->dLikOff.as<u64>(), dLik);
    ctx->launches++;
    CKL();
    rec(ctx, EV_CO1);
    if (out_data) CK(cudaMemcpyAsync(out_data, ctx->dOut.p, ctx->dataBytes, cudaMemcpyDeviceToHost, ctx->st));
    if (lik) CK(cudaMemcpyAsync(lik, ctx->dLik.p, ctx->totalBases * 4 * 8, cudaMemcpyDeviceToHost, ctx->st));
    u64 ch = 0;
    int rc = read_scalar(ctx, SC_CHANGED, &ch);
    if (rc) return rc;
    if (n_changed) *n_changed = ch;
    ctx->nChangedLast = ch; ctx->correctedValid = true; ctx->patchesDirect = false;
    return CDG_OK;
}

extern "C" int cdg_export_patches(cdg_ctx *ctx, cdg_patches *out) {
    if (!ctx || !out) return CDG_ERR_ARG;
    CK(cudaSetDevice(ctx->device));
    if (!ctx->correctedValid) return fail(ctx, CDG_ERR_ARG, "cdg_export_patches: no cdg_correct result on this context");
    return export_patches(ctx, 0, ctx->dataBytes, ctx->nChangedLast, out);
}

extern "C" int cdg_iteration_compact(cdg_ctx *ctx, const cdg_iter_params *p, cdg_alns_compact *alns, cdg_patches *patches, uint64_t *n_changed) {
    if (!ctx || !p) return CDG_ERR_ARG;
    if (ctx->n == 0) return fail(ctx, CDG_ERR_ARG, "cdg_iteration_compact: no sequence set uploaded");
    CK(cudaSetDevice(ctx->device));
    int rc = run_kmermatch(ctx, &p->kmer);
    if (rc) return rc;
    if ((rc = run_rescore(ctx, &p->rescore))) return rc;
    if (alns && (rc = export_alns_compact(ctx, alns, true))) return rc;   // the export overlaps the correction kernel
    if (!alns) { u64 kept = 0; if ((rc = read_scalar(ctx, SC_KEPT, &kept))) return rc; ctx->nKept = kept; }
    u64 ch = 0;
    rc = run_correct(ctx, &p->corr, nullptr, &ch, nullptr, true);
    if (!rc && patches) rc = export_patches(ctx, 0, ctx->dataBytes, ch, patches);
    if (n_changed) *n_changed = ch;
    int rc2 = wait_exports(ctx);
    return rc ? rc : rc2;
}

extern "C" int cdg_correct(cdg_ctx *ctx, const cdg_corr_params *p, uint8_t *out_data, uint64_t *n_changed) {
    if (!ctx || !p) return CDG_ERR_ARG;
    CK(cudaSetDevice(ctx->device));
    return run_correct(ctx, p, out_data, n_changed, nullptr);
}
extern "C" int cdg_correct_likelihoods(cdg_ctx *ctx, const cdg_corr_params *p, double *lik) {
    if (!ctx || !p || !lik) return CDG_ERR_ARG;
    CK(cudaSetDevice(ctx->device));
    return run_correct(ctx, p, nullptr, nullptr, lik);
}

extern "C" int cdg_iteration(cdg_ctx *ctx, const cdg_iter_params *p, cdg_hits *hits, cdg_alns *alns, uint8_t *out_data, uint64_t *n_changed) {
    if (!ctx || !p) return CDG_ERR_ARG;
    if (ctx->n == 0) return fail(ctx, CDG_ERR_ARG, "cdg_iteration: no sequence set uploaded");
    CK(cudaSetDevice(ctx->device));
    int rc = run_kmermatch(ctx, &p->kmer);
    if (rc) return rc;
    if (hits && (rc = export_hits(ctx, hits))) return rc;
    if ((rc = run_rescore(ctx, &p->rescore))) return rc;
    if (alns && (rc = export_alns(ctx, alns, true))) return rc;   // alignment export overlaps the correction kernel
    if (!alns) { u64 kept = 0; if ((rc = read_scalar(ctx, SC_KEPT, &kept))) return rc; ctx->nKept = kept; }
    rc = run_correct(ctx, &p->corr, out_data, n_changed, nullptr);
    int rc2 = wait_exports(ctx);
    return rc ? rc : rc2;
}

// ------------------------------------------------------------------------------------------------------
// a19: damage-aware pair likelihood
extern "C" int cdg_set_pair_model(cdg_ctx *ctx, const cdg_pair_model *m) {
    if (!ctx || !m) return CDG_ERR_ARG;
    CK(cudaSetDevice(ctx->device));
    static_assert(sizeof(cdg_pair_model) == sizeof(cdgpair::PairModel), "pair model layout");
    CK(ctx->dPairModel.reserve(sizeof(cdgpair::PairModel)));
    CK(cudaMemcpyAsync(ctx->dPairModel.p, m, sizeof(cdg_pair_model), cudaMemcpyHostToDevice, ctx->st));
    CK(cudaStreamSynchronize(ctx->st));
    ctx->pairModelSet = true;
    return CDG_OK;
}

static int run_pair_loglik(cdg_ctx *ctx, const cdg_pairs *P, const uint8_t *cons, uint64_t consBytes, const uint64_t *consOff, double *sLen,
                           double *sRatio) {
    if (!ctx || !P || !sLen || !sRatio) return CDG_ERR_ARG;
    if (!ctx->pairModelSet) return fail(ctx, CDG_ERR_ARG, "cdg_pair_loglik: cdg_set_pair_model not called");
    if (ctx->n == 0) return fail(ctx, CDG_ERR_ARG, "cdg_pair_loglik: no resident sequence set");
    const u64 n = P->n;
    if (n == 0) return CDG_OK;
    CK(cudaSetDevice(ctx->device));
    if (int rcp = ensure_bytes(ctx)) return rcp;   // (packs first) the pair kernels read the DB bytes
    // validate on the host what would otherwise be an out-of-bounds device read
    for (u64 i = 0; i < n; i++) {
        if (P->query[i] >= ctx->n || P->target[i] >= ctx->n)
            return fail(ctx, CDG_ERR_ARG, "cdg_pair_loglik: pair %llu: sequence index out of range", (unsigned long long) i);
        if (P->q_start[i] < 0 || P->db_start[i] < 0 || P->q_end[i] < P->q_start[i] || P->db_end[i] - P->db_start[i] != P->q_end[i] - P->q_start[i])
            return fail(ctx, CDG_ERR_ARG, "cdg_pair_loglik: pair %llu: coordinates are not an oriented ungapped alignment", (unsigned long long) i);
        if (cons && consOff[i] > consBytes) return fail(ctx, CDG_ERR_ARG, "cdg_pair_loglik_consensus: pair %llu: consensus offset out of range", (unsigned long long) i);
    }
    // one staging buffer: query, target, qS, qE, dS, dE, maxAln (u32 each) | rev (u8) | consOff (u64)
    const u64 words = 7 * n;
    const u64 consOffAt = (words * 4 + n + 7) & ~7ull;
    CK(ctx->dPairIn.reserve(consOffAt + n * 8));
    u8 *base = ctx->dPairIn.as<u8>();
    const void *src[7] = {P->query, P->target, P->q_start, P->q_end, P->db_start, P->db_end, P->max_aln};
    for (int a = 0; a < 7; a++) CK(cudaMemcpyAsync(base + (u64) a * n * 4, src[a], n * 4, cudaMemcpyHostToDevice, ctx->st));
    CK(cudaMemcpyAsync(base + words * 4, P->rev, n, cudaMemcpyHostToDevice, ctx->st));
    if (cons) {
        CK(ctx->dPairCons.reserve(consBytes + 16));
        CK(cudaMemcpyAsync(ctx->dPairCons.p, cons, consBytes, cudaMemcpyHostToDevice, ctx->st));
        CK(cudaMemcpyAsync(base + consOffAt, consOff, n * 8, cudaMemcpyHostToDevice, ctx->st));
    }
    CK(ctx->dPairOut.reserve(2 * n * 8));
    double *dLen = ctx->dPairOut.as<double>(), *dRatio = dLen + n;
    const u32 *w = (const u32 *) base;
    const unsigned grid = (unsigned) std::min<u64>((n + cdgpair::PL_WARPS - 1) / cdgpair::PL_WARPS, (u64) ctx->smCount * 8);
#define CDG_PAIR_ARGS ctx->dData.as<u8>(), ctx->dMeta.as<SeqMeta>(), n, w, w + n, base + words * 4, (const i32 *) (w + 2 * n), (const i32 *) (w + 3 * n), \
                      (const i32 *) (w + 4 * n), (const i32 *) (w + 5 * n), w + 6 * n, ctx->dPairCons.as<u8>(), (const u64 *) (base + consOffAt),                 \
                      ctx->dPairModel.as<cdgpair::PairModel>(), dLen, dRatio
    if (cons) cdgpair::pair_loglik_kernel<1><<<grid, cdgpair::PL_WARPS * 32, 0, ctx->st>>>(CDG_PAIR_ARGS);
    else cdgpair::pair_loglik_kernel<0><<<grid, cdgpair::PL_WARPS * 32, 0, ctx->st>>>(CDG_PAIR_ARGS);
#undef CDG_PAIR_ARGS
    ctx->launches++;
    CKL();
    CK(cudaMemcpyAsync(sLen, dLen, n * 8, cudaMemcpyDeviceToHost, ctx->st));
    CK(cudaMemcpyAsync(sRatio, dRatio, n * 8, cudaMemcpyDeviceToHost, ctx->st));
    CK(cudaStreamSynchronize(ctx->st));
    return CDG_OK;
}

extern "C" int cdg_pair_loglik(cdg_ctx *ctx, const cdg_pairs *pairs, double *s_len_norm, double *s_ratio) {
    return run_pair_loglik(ctx, pairs, nullptr, 0, nullptr, s_len_norm, s_ratio);
}

extern "C" int cdg_pair_loglik_consensus(cdg_ctx *ctx, const cdg_pairs *pairs, const uint8_t *consensus, uint64_t consensus_bytes,
                                         const uint64_t *cons_off, double *s_len_norm, double *s_ratio) {
    if (!consensus || !cons_off) return CDG_ERR_ARG;
    return run_pair_loglik(ctx, pairs, consensus, consensus_bytes, cons_off, s_len_norm, s_ratio);
}

// ---- scoring front half of ancient_read_assemble on the alignments of the context ----------------------------
extern "C" int cdg_extension_scores(cdg_ctx *ctx, const cdg_ext_params *p, cdg_ext_scores *out) {
    if (!ctx || !p || !out) return CDG_ERR_ARG;
    if (ctx->n == 0) return fail(ctx, CDG_ERR_ARG, "cdg_extension_scores: no resident sequence set");
    if (!ctx->alnsValid) return fail(ctx, CDG_ERR_ARG, "cdg_extension_scores: no alignments (cdg_rescore or cdg_set_alignments)");
    if (!ctx->pairModelSet) return fail(ctx, CDG_ERR_ARG, "cdg_extension_scores: cdg_set_pair_model not called");
    if (ctx->qLo != 0 || ctx->nQ != ctx->n) return fail(ctx, CDG_ERR_UNSUPPORTED, "cdg_extension_scores: the alignments must cover all queries of the context");
    CK(cudaSetDevice(ctx->device));
    if (int rcb = ensure_bytes(ctx)) return rcb;
    using namespace cdgext;
    const u64 M = ctx->nLines, n = ctx->n, consBytes = 3 * ctx->totalBases;
    // device arrays: 4 x i32 coordinates | 2 x f32 identities | 2 x f64 scores | flags + rev (u8) | maxAln per line + maxLeft/Right per query (u32) | consOff (u64)
    CK(ctx->dExtI32.reserve((4 * M + 4) * 4)); CK(ctx->dExtF32.reserve((2 * M + 2) * 4)); CK(ctx->dExtF64.reserve((2 * M + 2) * 8)); CK(ctx->dExtU8.reserve(2 * M + 2));
    CK(ctx->dExtU32.reserve((M + 2 * n + 4) * 4)); CK(ctx->dExtU64.reserve((M + 1) * 8)); CK(ctx->dExtCons.reserve(consBytes + 16));
    i32 *oqS = ctx->dExtI32.as<i32>(), *oqE = oqS + M, *odS = oqE + M, *odE = odS + M;
    float *sid = ctx->dExtF32.as<float>(), *ry = sid + M;
    double *sLen = ctx->dExtF64.as<double>(), *sRatio = sLen + M;
    u8 *flags = ctx->dExtU8.as<u8>(), *revLine = flags + M;
    u32 *maxAlnLine = ctx->dExtU32.as<u32>(), *maxLeft = maxAlnLine + M, *maxRight = maxLeft + n;
    u64 *consOffLine = ctx->dExtU64.as<u64>();
    const u32 *dKeys = nullptr;
    if (p->keys) { CK(ctx->dExtKeys.reserve(n * 4)); CK(cudaMemcpyAsync(ctx->dExtKeys.p, p->keys, n * 4, cudaMemcpyHostToDevice, ctx->st)); dKeys = ctx->dExtKeys.as<u32>(); }
    CK(cudaMemsetAsync(maxLeft, 0, 2 * n * 4, ctx->st));
    ExtParams E; E.seqIdThr = p->min_seq_id; E.ryThr = p->min_ryseq_id; E.unsafe = p->unsafe; E.minCovSafe = p->min_cov_safe; E.roundSeqId = ctx->alnsFromText ? 0 : 1;
    const SeqMeta *meta = ctx->dMeta.as<SeqMeta>();
    const unsigned lb = (unsigned) ((M + 255) / 256);
    if (M) {
        ext_lines_kernel<<<lb, 256, 0, ctx->st>>>(ctx->dData.as<u8>(), meta, dKeys, ctx->dLikOff.as<u64>(), ctx->dLineQ.as<u32>(), ctx->dLineT.as<u32>(), ctx->dKeep.as<u8>(),
                                                 ctx->dSeqId.as<float>(), ctx->dQS.as<i32>(), ctx->dQE.as<i32>(), ctx->dDS.as<i32>(), ctx->dDE.as<i32>(), M, E, flags, oqS, oqE,
                                                 odS, odE, sid, ry, consOffLine);
        ctx->launches++;
    }
    ext_consensus_kernel<<<(unsigned) std::min<u64>((n * 32 + 255) / 256, (u64) ctx->smCount * 16), 256, 0, ctx->st>>>(
        ctx->dData.as<u8>(), meta, dKeys, ctx->dLikOff.as<u64>(), ctx->dLineOff.as<u64>(), n, ctx->dLineT.as<u32>(), flags, oqS, oqE, odS, odE, E, ctx->dExtCons.as<u8>());
    ctx->launches++;
    if (M) {
        ext_update_kernel<<<lb, 256, 0, ctx->st>>>(ctx->dData.as<u8>(), meta, ctx->dLineQ.as<u32>(), ctx->dLineT.as<u32>(), flags, oqS, oqE, odS, odE, consOffLine,
                                                  ctx->dExtCons.as<u8>(), M, sid, ry, maxLeft, maxRight);
        ext_select_kernel<<<lb, 256, 0, ctx->st>>>(meta, dKeys, ctx->dLineQ.as<u32>(), ctx->dLineT.as<u32>(), flags, oqS, odS, odE, sid, ry, maxLeft, maxRight, M, E, maxAlnLine,
                                                  revLine);
        const unsigned grid = (unsigned) std::min<u64>((M + cdgpair::PL_WARPS - 1) / cdgpair::PL_WARPS, (u64) ctx->smCount * 8);
        cdgpair::pair_loglik_kernel<1><<<grid, cdgpair::PL_WARPS * 32, 0, ctx->st>>>(
            ctx->dData.as<u8>(), meta, M, ctx->dLineQ.as<u32>(), ctx->dLineT.as<u32>(), revLine, oqS, oqE, odS, odE, maxAlnLine, ctx->dExtCons.as<u8>(), consOffLine,
            ctx->dPairModel.as<cdgpair::PairModel>(), sLen, sRatio, flags, EXT_SCORED);
        ctx->launches += 3;
    }
    CKL();
    // one pinned block: flags | 4 coordinates | 2 identities | 2 scores | maxLeft | maxRight | consensus | consensus offsets per query
    auto al = [](u64 x) { return (x + 15) & ~15ull; };
    const u64 oFlags = 0, oI = al(M), oF = oI + al(16 * M), oD = oF + al(8 * M), oML = oD + al(16 * M), oMR = oML + al(4 * n), oC = oMR + al(4 * n), oCO = oC + al(consBytes),
              total = oCO + al(8 * (n + 1));
    CK(ctx->hExt.reserve(total + 16));
    u8 *h = ctx->hExt.as<u8>();
    if (M) {
        CK(cudaMemcpyAsync(h + oFlags, flags, M, cudaMemcpyDeviceToHost, ctx->st));
        CK(cudaMemcpyAsync(h + oI, oqS, 16 * M, cudaMemcpyDeviceToHost, ctx->st));
        CK(cudaMemcpyAsync(h + oF, sid, 8 * M, cudaMemcpyDeviceToHost, ctx->st));
        CK(cudaMemcpyAsync(h + oD, sLen, 16 * M, cudaMemcpyDeviceToHost, ctx->st));
    }
    CK(cudaMemcpyAsync(h + oML, maxLeft, 4 * n, cudaMemcpyDeviceToHost, ctx->st));
    CK(cudaMemcpyAsync(h + oMR, maxRight, 4 * n, cudaMemcpyDeviceToHost, ctx->st));
    if (consBytes) CK(cudaMemcpyAsync(h + oC, ctx->dExtCons.p, consBytes, cudaMemcpyDeviceToHost, ctx->st));
    CK(cudaMemcpyAsync(h + oCO, ctx->dLikOff.p, 8 * (n + 1), cudaMemcpyDeviceToHost, ctx->st));
    CK(cudaStreamSynchronize(ctx->st));
    u64 *co = (u64 *) (h + oCO);
    for (u64 i = 0; i <= n; i++) co[i] *= 3;
    out->n_lines = M; out->n_queries = n; out->flags = h + oFlags;
    out->q_start = (const int32_t *) (h + oI); out->q_end = out->q_start + M; out->db_start = out->q_end + M; out->db_end = out->db_start + M;
    out->seq_id = (const float *) (h + oF); out->ry_seq_id = out->seq_id + M;
    out->s_len_norm = (const double *) (h + oD); out->s_ratio = out->s_len_norm + M;
    out->max_aln_left = (const uint32_t *) (h + oML); out->max_aln_right = (const uint32_t *) (h + oMR);
    out->consensus = h + oC; out->consensus_off = co;
    return CDG_OK;
}

// ---- N3: scoring front half of ancient_contig_merge (cdg_merge.cuh) ----------------------------------------------------------
extern "C" int cdg_set_merge_model(cdg_ctx *ctx, const cdg_merge_model *m) {
    if (!ctx || !m) return CDG_ERR_ARG;
    CK(cudaSetDevice(ctx->device));
    // deamMatches' length prior (nuclassembleUtil.cpp:1017-1030) needs log(): tabulated here with the host libm the reference uses
    const u32 nTab = cdgmerge::MRG_MAX_LOG_LEN + 1;
    std::vector<double> tab(nTab);
    const double logAdj = std::log(1.4e-9);
    for (u32 len = 0; len < nTab; len++) tab[len] = logAdj - 3.0 * std::log(len);
    CK(ctx->dMergeLog.reserve(nTab * 8));
    CK(cudaMemcpyAsync(ctx->dMergeLog.p, tab.data(), nTab * 8, cudaMemcpyHostToDevice, ctx->st));
    CK(cudaStreamSynchronize(ctx->st));
    for (int r = 0; r < 2; r++) { ctx->mergeModel.matchLik[r][0] = m->match_lik[r][1][3]; ctx->mergeModel.matchLik[r][1] = m->match_lik[r][2][0]; }
    ctx->mergeModel.absLogMin = std::abs(tab[10]); ctx->mergeModel.absLogMax = std::abs(tab[cdgmerge::MRG_MAX_LOG_LEN]);
    ctx->mergeModelSet = true;
    return CDG_OK;
}

extern "C" int cdg_merge_scores(cdg_ctx *ctx, const cdg_merge_params *p, cdg_merge_out *out) {
    if (!ctx || !p || !out) return CDG_ERR_ARG;
    if (ctx->n == 0) return fail(ctx, CDG_ERR_ARG, "cdg_merge_scores: no resident sequence set");
    if (!ctx->alnsValid) return fail(ctx, CDG_ERR_ARG, "cdg_merge_scores: no alignments (cdg_rescore or cdg_set_alignments)");
    if (!ctx->mergeModelSet) return fail(ctx, CDG_ERR_ARG, "cdg_merge_scores: cdg_set_merge_model not called");
    if (ctx->qLo != 0 || ctx->nQ != ctx->n) return fail(ctx, CDG_ERR_UNSUPPORTED, "cdg_merge_scores: the alignments must cover all queries of the context");
    CK(cudaSetDevice(ctx->device));
    if (int rcb = ensure_bytes(ctx)) return rcb;
    using namespace cdgmerge;
    const u64 M = ctx->nLines, n = ctx->n, consBytes = p->unsafe ? 3 * ctx->totalBases : 0;
    // device arrays: 4 x i32 coordinates | 5 x f32 (seqId0, ry0, seqId, ry, deamMatch) | flags (u8) | alnLengthCons (u32)
    CK(ctx->dExtI32.reserve((4 * M + 4) * 4)); CK(ctx->dExtF32.reserve((5 * M + 5) * 4)); CK(ctx->dExtU8.reserve(M + 2)); CK(ctx->dExtU32.reserve((M + 4) * 4));
    if (consBytes) CK(ctx->dExtCons.reserve(consBytes + 16));
    i32 *oqS = ctx->dExtI32.as<i32>(), *oqE = oqS + M, *odS = oqE + M, *odE = odS + M;
    float *sid0 = ctx->dExtF32.as<float>(), *ry0 = sid0 + M, *sid = ry0 + M, *ry = sid + M, *deam = ry + M;
    u8 *flags = ctx->dExtU8.as<u8>();
    u32 *alc = ctx->dExtU32.as<u32>();
    const u32 *dKeys = nullptr;
    if (p->keys) { CK(ctx->dExtKeys.reserve(n * 4)); CK(cudaMemcpyAsync(ctx->dExtKeys.p, p->keys, n * 4, cudaMemcpyHostToDevice, ctx->st)); dKeys = ctx->dExtKeys.as<u32>(); }
    MergeParams P; P.mergeThr = p->min_merge_seq_id; P.ryThr = p->min_ryseq_id; P.unsafe = p->unsafe;
    const SeqMeta *meta = ctx->dMeta.as<SeqMeta>();
    if (M) {
        const unsigned wgrid = (unsigned) std::min<u64>((M * 32 + 255) / 256, (u64) ctx->smCount * 16);
        merge_lines_kernel<<<wgrid, 256, 0, ctx->st>>>(ctx->dData.as<u8>(), meta, dKeys, ctx->dLineQ.as<u32>(), ctx->dLineT.as<u32>(), ctx->dKeep.as<u8>(), ctx->dQS.as<i32>(),
                                                      ctx->dQE.as<i32>(), ctx->dDS.as<i32>(), ctx->dDE.as<i32>(), M, P, flags, oqS, oqE, odS, odE, sid0, ry0);
        ctx->launches++;
        if (p->unsafe) {
            ExtParams E; E.seqIdThr = p->min_merge_seq_id; E.ryThr = p->min_ryseq_id; E.unsafe = 1; E.minCovSafe = p->min_cov_safe; E.roundSeqId = 0;
            ext_consensus_kernel<<<(unsigned) std::min<u64>((n * 32 + 255) / 256, (u64) ctx->smCount * 16), 256, 0, ctx->st>>>(
                ctx->dData.as<u8>(), meta, dKeys, ctx->dLikOff.as<u64>(), ctx->dLineOff.as<u64>(), n, ctx->dLineT.as<u32>(), flags, oqS, oqE, odS, odE, E, ctx->dExtCons.as<u8>());
            ctx->launches++;
        }
        merge_score_kernel<<<wgrid, 256, 0, ctx->st>>>(ctx->dData.as<u8>(), meta, ctx->dLikOff.as<u64>(), ctx->dLineQ.as<u32>(), ctx->dLineT.as<u32>(), flags, oqS, oqE, odS, odE,
                                                      sid0, ry0, p->unsafe ? ctx->dExtCons.as<u8>() : nullptr, ctx->dMergeLog.as<double>(), ctx->mergeModel, M, P, sid, ry,
                                                      alc, deam);
        ctx->launches++;
    }
    CKL();
    // one pinned block: flags | 4 coordinates | 5 floats | alnLengthCons
    auto al = [](u64 x) { return (x + 15) & ~15ull; };
    const u64 oFlags = 0, oI = al(M), oF = oI + al(16 * M), oU = oF + al(20 * M), total = oU + al(4 * M);
    CK(ctx->hExt.reserve(total + 16));
    u8 *h = ctx->hExt.as<u8>();
    if (M) {
        CK(cudaMemcpyAsync(h + oFlags, flags, M, cudaMemcpyDeviceToHost, ctx->st));
        CK(cudaMemcpyAsync(h + oI, oqS, 16 * M, cudaMemcpyDeviceToHost, ctx->st));
        CK(cudaMemcpyAsync(h + oF, sid0, 20 * M, cudaMemcpyDeviceToHost, ctx->st));
        CK(cudaMemcpyAsync(h + oU, alc, 4 * M, cudaMemcpyDeviceToHost, ctx->st));
    }
    CK(cudaStreamSynchronize(ctx->st));
    out->n_lines = M; out->flags = h + oFlags;
    out->q_start = (const int32_t *) (h + oI); out->q_end = out->q_start + M; out->db_start = out->q_end + M; out->db_end = out->db_start + M;
    out->seq_id_query = (const float *) (h + oF); out->ry_seq_id_query = out->seq_id_query + M; out->seq_id = out->ry_seq_id_query + M; out->ry_seq_id = out->seq_id + M;
    out->deam_match = out->ry_seq_id + M;
    out->aln_len_cons = (const uint32_t *) (h + oU);
    return CDG_OK;
}

extern "C" int cdg_get_stats(cdg_ctx *ctx, cdg_stats *out) {
    if (!ctx || !out) return CDG_ERR_ARG;
    memset(out, 0, sizeof *out);
    out->n_seqs = ctx->n; out->n_bases = ctx->totalBases; out->n_records = ctx->nRecords; out->n_grouped = ctx->nGrouped; out->n_hits = ctx->nHits;
    out->n_alns = ctx->nKept; out->sort1_passes = ctx->sort1Passes; out->sort2_passes = ctx->sort2Passes; out->kernel_launches = ctx->launches;
    auto el = [&](int a, int b) -> float {
        float ms = 0.f;
        if (ctx->evValid[a] && ctx->evValid[b] && cudaEventElapsedTime(&ms, ctx->ev[a], ctx->ev[b]) == cudaSuccess) return ms;
        return 0.f;
    };
    cudaStreamSynchronize(ctx->st);
    out->ms_extract = el(EV_START, EV_EXTRACT); out->ms_sort1 = el(EV_EXTRACT, EV_SORT1); out->ms_group = el(EV_SORT1, EV_GROUP);
    out->ms_sort2 = el(EV_GROUP, EV_SORT2);
    out->ms_sort1_passes = 0.f; out->n_sort1_pass_launches = 0; out->sort1_records = ctx->sort1Records; out->split_passes = ctx->splitPasses;
    for (int i = 0; i < ctx->passEvN; i++) { float ms = 0.f; if (cudaEventElapsedTime(&ms, ctx->passEv[2 * i], ctx->passEv[2 * i + 1]) == cudaSuccess) { out->ms_sort1_passes += ms; out->n_sort1_pass_launches++; } } out->ms_reduce = el(EV_SORT2, EV_REDUCE); out->ms_rescore = el(EV_RS0, EV_RS1); out->ms_correct = el(EV_CO0, EV_CO1);
    return CDG_OK;
}

// ------------------------------------------------------------------------------------------------------
// Multi-GPU phases (one context per rank, the whole sequence set resident on every rank).  The caller moves the
// buffers between ranks (NCCL all-to-all over NVLink through torch.distributed, see carpedeam_b200/multigpu.py):
//   phase 1  cdg_mg_extract    k-mers of this rank's slice of the sequences, partitioned by owner of the hash range
//            -> all-to-all of 16-byte records (the single data-carrying collective north_star names)
//   phase 2  cdg_mg_sort       identity records of the owned hash range (local), sort 1; returns the smallest key
//            -> the ranks compare their minima (8 bytes each): only the global first run gets the kmermatcher.cpp:466 quirk
//   phase 3  cdg_mg_group      group assignment, grouped records routed to the owner of the representative's index range
//            -> all-to-all of 8-byte grouped records (replaces the reference's host k-way merge of split files)
//   phase 4  cdg_mg_sort2      sort 2; the first records (same target id) go to the previous rank for the rep-boundary bleed
//   phase 5  cdg_mg_finish     reduce, rescore, correct for the rank's query range
extern "C" int cdg_mg_extract(cdg_ctx *ctx, const cdg_kmer_params *p, int world, int rank, int partition, uint64_t *send_counts, void **send_keys,
                              void **send_vals, uint64_t *id_send_counts, void **id_send_keys, void **id_send_vals) {
    if (!ctx || !p || !send_counts || !id_send_counts || world < 1 || world > CDG_MAX_WORLD || rank < 0 || rank >= world) return CDG_ERR_ARG;
    if (ctx->n == 0) return fail(ctx, CDG_ERR_ARG, "cdg_mg_extract: no sequence set uploaded");
    CK(cudaSetDevice(ctx->device));
    KmPlan K; K.world = (u32) world; K.rank = (u32) rank;
    int rc = km_bits(ctx, p, K);
    if (rc) return rc;
    unsigned long long *sc = ctx->dScalars.as<unsigned long long>();
    CK(cudaMemsetAsync(sc, 0, SC_N * 8, ctx->st));
    if ((rc = km_prepare(ctx, p, K))) return rc;
    ctx->plan = K; ctx->qLo = K.qLo; ctx->nQ = K.nQ;
    const u64 R = K.slots + K.extraCap;          // slice slots + spare slots for identity hashes inside the k-mer key range
    const u64 nSlice = (u64) K.posHi - K.posLo;  // sequences of the slice = identity records produced here
    if (R > cdgsort::SMALL_N) return fail(ctx, CDG_ERR_UNSUPPORTED, "slice has too many k-mer records for one routing pass (2^30)");
    CK(ctx->dKeysA.reserve((R + 1) * 8)); CK(ctx->dKeysB.reserve((R + 1) * 8)); CK(ctx->dValsA.reserve((R + 1) * 8)); CK(ctx->dValsB.reserve((R + 1) * 8));
    CK(ctx->dIdKeysA.reserve((nSlice + 1) * 8)); CK(ctx->dIdKeysB.reserve((nSlice + 1) * 8)); CK(ctx->dIdValsA.reserve((nSlice + 1) * 8)); CK(ctx->dIdValsB.reserve((nSlice + 1) * 8));
    CK(cudaMemsetAsync(ctx->dKeysA.as<u64>() + K.slots, 0xFF, (size_t) K.extraCap * 8, ctx->st));
    CK(cudaMemsetAsync(ctx->dValsA.as<u64>() + K.slots, 0, (size_t) K.extraCap * 8, ctx->st));
    // all hash values of the slice's sequences and one identity record per sequence of the slice, each tagged with its owner
    if ((rc = km_extract(ctx, p, K, ctx->dKeysA.as<u64>(), ctx->dValsA.as<u64>(), ctx->dIdKeysA.as<u64>(), ctx->dIdValsA.as<u64>(), K.slots, 2, 0, 65535))) return rc;
    // partition by owner: one stable radix pass over the owner bits of the VALUE word (roles of the arrays swapped)
    u64 *pv = ctx->dValsA.as<u64>(), *pk = ctx->dKeysA.as<u64>(), *iv = ctx->dIdValsA.as<u64>(), *ik = ctx->dIdKeysA.as<u64>();
    for (int o = 0; o < world; o++) { send_counts[o] = 0; id_send_counts[o] = 0; }
    cdgsort::Workspace ws;
    if ((rc = sort_ws(ctx, std::max<u64>(R, ctx->n), ws))) return rc;
    const int ownerBits = std::max(1, ceil_log2((u64) world));
    cdgsort::Plan plan = cdgsort::make_plan(0, 0, CDG_OWNER_SHIFT, CDG_OWNER_SHIFT + ownerBits);
    auto partitionPass = [&](u64 *va, u64 *vb, u64 *ka, u64 *kb, u64 cnt, u64 **outV, u64 **outK, uint64_t *counts) -> int {
        if (!cnt) return 0;
        if (!partition) {   // peer-memory exchange: only the per-owner counts now, cdg_mg_scatter_records writes straight into the peers
            CK(cudaMemsetAsync(ws.hist, 0, sizeof(unsigned long long) * cdgsort::RADIX, ctx->st));
            int hb = (int) std::min<u64>((cnt + cdgsort::BLOCK - 1) / cdgsort::BLOCK, (u64) ctx->smCount * 8);
            cdgsort::hist_kernel<<<hb, cdgsort::BLOCK, cdgsort::RADIX * sizeof(u32), ctx->st>>>(va, nullptr, cnt, plan, ws.hist);
            ctx->launches++;
            unsigned long long hc[cdgsort::RADIX];
            CK(cudaMemcpyAsync(hc, ws.hist, sizeof hc, cudaMemcpyDeviceToHost, ctx->st));
            CK(cudaStreamSynchronize(ctx->st));
            for (int o = 0; o < world; o++) counts[o] = hc[o];
            return 0;
        }
        int l = cdgsort::radix_sort(ctx->st, ws, va, vb, ka, kb, cnt, plan, outV, outK, ctx->smCount);
        if (l < 0) return fail(ctx, CDG_ERR_UNSUPPORTED, "owner partition failed");
        ctx->launches += l;
        unsigned long long hoff[cdgsort::RADIX];
        CK(cudaMemcpyAsync(hoff, ws.hist, sizeof hoff, cudaMemcpyDeviceToHost, ctx->st));   // exclusive bin offsets of the single pass
        CK(cudaStreamSynchronize(ctx->st));
        u64 acc = 0;
        for (int o = 0; o < world - 1; o++) { counts[o] = hoff[o + 1] - hoff[o]; acc += counts[o]; }
        counts[world - 1] = cnt - acc;
        return 0;
    };
    ctx->mgSlice = nSlice; ctx->mgMainK = ctx->dKeysA.as<u64>(); ctx->mgMainV = ctx->dValsA.as<u64>(); ctx->mgIdK = ctx->dIdKeysA.as<u64>(); ctx->mgIdV = ctx->dIdValsA.as<u64>();
    if ((rc = partitionPass(ctx->dValsA.as<u64>(), ctx->dValsB.as<u64>(), ctx->dKeysA.as<u64>(), ctx->dKeysB.as<u64>(), R, &pv, &pk, send_counts))) return rc;
    if ((rc = partitionPass(ctx->dIdValsA.as<u64>(), ctx->dIdValsB.as<u64>(), ctx->dIdKeysA.as<u64>(), ctx->dIdKeysB.as<u64>(), nSlice, &iv, &ik, id_send_counts))) return rc;
    CKL();
    if (send_keys) *send_keys = pk;
    if (send_vals) *send_vals = pv;
    if (id_send_keys) *id_send_keys = ik;
    if (id_send_vals) *id_send_vals = iv;
    return CDG_OK;
}

// ---- streamed record exchange: extraction and routing overlap ------------------------------------------------------------
// Every owner's receive area is split into `world` regions of cap_per_source records, one per source rank (in rank order, so the
// receiver's stable sort still sees the sources in (length desc, id asc) order).  A source extracts its slice in chunks and, while
// chunk c+1 is being extracted, the routing pass of chunk c writes each owner's records behind the ones already sent (running
// per-owner counts on the device) — no count matrix is needed before the transfer starts.  The unused tail of every region is
// filled with sentinels by the receiver (cdg_mg_recv_fill).  Hash ownership is uniform, so a region holds 1/world^2 of the set's
// k-mer slots plus a few per cent; a source that would overrun a region (degenerate sets) reports *overflow and the caller falls
// back to the counted exchange (cdg_mg_extract / cdg_mg_scatter_records).
extern "C" int cdg_mg_plan(cdg_ctx *ctx, const cdg_kmer_params *p, int world, int rank, uint64_t *cap_per_source) {
    if (!ctx || !p || !cap_per_source || world < 1 || world > CDG_MAX_WORLD || rank < 0 || rank >= world) return CDG_ERR_ARG;
    if (ctx->n == 0) return fail(ctx, CDG_ERR_ARG, "cdg_mg_plan: no sequence set uploaded");
    CK(cudaSetDevice(ctx->device));
    KmPlan K; K.world = (u32) world; K.rank = (u32) rank;
    int rc = km_bits(ctx, p, K);
    if (rc) return rc;
    unsigned long long *sc = ctx->dScalars.as<unsigned long long>();
    CK(cudaMemsetAsync(sc, 0, SC_N * 8, ctx->st));
    if ((rc = km_prepare(ctx, p, K))) return rc;
    ctx->plan = K; ctx->qLo = K.qLo; ctx->nQ = K.nQ; ctx->planReady = true;
    const u64 perPair = K.totalSlots / ((u64) world * (u64) world);
    *cap_per_source = ((perPair + perPair / 64 + K.extraCap + 65536 + 4095) / 4096) * 4096;   // + 1.5 % and the spare identity slots
    return CDG_OK;
}

extern "C" int cdg_mg_extract_stream(cdg_ctx *ctx, const cdg_kmer_params *p, int chunks, uint64_t cap_per_source, uint64_t *send_counts,
                                     uint64_t *id_send_counts, int *overflow) {
    if (!ctx || !p || !send_counts || !id_send_counts || !overflow || chunks < 1) return CDG_ERR_ARG;
    if (!ctx->planReady) return fail(ctx, CDG_ERR_ARG, "cdg_mg_extract_stream: call cdg_mg_plan first");
    CK(cudaSetDevice(ctx->device));
    ctx->planReady = false;
    const KmPlan K = ctx->plan;
    const u32 world = K.world, rank = K.rank;
    unsigned long long *sc = ctx->dScalars.as<unsigned long long>();
    int rc;
    const u64 R = K.slots + K.extraCap;
    const u64 nSlice = (u64) K.posHi - K.posLo;
    if (R > cdgsort::SMALL_N) return fail(ctx, CDG_ERR_UNSUPPORTED, "slice has too many k-mer records for one routing pass (2^30)");
    CK(ctx->dKeysA.reserve((R + 1) * 8)); CK(ctx->dValsA.reserve((R + 1) * 8)); CK(ctx->dKeysB.reserve((R + 1) * 8)); CK(ctx->dValsB.reserve((R + 1) * 8));
    CK(ctx->dIdKeysA.reserve((nSlice + 1) * 8)); CK(ctx->dIdKeysB.reserve((nSlice + 1) * 8)); CK(ctx->dIdValsA.reserve((nSlice + 1) * 8)); CK(ctx->dIdValsB.reserve((nSlice + 1) * 8));
    const u32 discard = world < (u32) cdgsort::ROUTE_MAX ? (u32) cdgsort::ROUTE_MAX - 1 : 0u;   // unused slots: a local trash area, not rank 0
    CK(cudaMemsetAsync(ctx->dKeysA.as<u64>() + K.slots, 0xFF, (size_t) K.extraCap * 8, ctx->st));
    {   // (value words of the spare slots carry the discard tag)
        std::vector<u64> fillv(K.extraCap, cdg_pack_val(0, 0, discard));
        CK(cudaMemcpyAsync(ctx->dValsA.as<u64>() + K.slots, fillv.data(), (size_t) K.extraCap * 8, cudaMemcpyHostToDevice, ctx->st));
        CK(cudaStreamSynchronize(ctx->st));
    }
    // routing tables: region `rank` of every owner's receive area; the discard destination is the local partner buffer
    cdgsort::Workspace ws;
    if ((rc = sort_ws(ctx, std::max<u64>(R, ctx->n), ws))) return rc;
    CK(ctx->dRouteTab.reserve(6 * cdgsort::ROUTE_MAX * sizeof(void *)));
    struct { u64 *v[cdgsort::ROUTE_MAX]; u64 *k[cdgsort::ROUTE_MAX]; u64 cap[cdgsort::ROUTE_MAX]; } tab = {};
    for (u32 o = 0; o < world; o++) {
        u64 *bk = o == rank ? ctx->dRecvK.as<u64>() : (u64 *) ctx->peerPtr[0][o], *bv = o == rank ? ctx->dRecvV.as<u64>() : (u64 *) ctx->peerPtr[1][o];
        if (!bk || !bv) return fail(ctx, CDG_ERR_ARG, "cdg_mg_extract_stream: receive buffers of rank %u are not mapped (cdg_mg_ipc_open)", o);
        tab.k[o] = bk + (u64) rank * cap_per_source; tab.v[o] = bv + (u64) rank * cap_per_source; tab.cap[o] = cap_per_source;
    }
    if (discard) { tab.k[discard] = ctx->dKeysB.as<u64>(); tab.v[discard] = ctx->dValsB.as<u64>(); tab.cap[discard] = R + 1; }
    CK(cudaMemcpyAsync(ctx->dRouteTab.p, &tab, sizeof tab, cudaMemcpyHostToDevice, ctx->st));
    u64 *const *dV = ctx->dRouteTab.as<u64 *>(), *const *dK = dV + cdgsort::ROUTE_MAX;
    const u64 *dCap = (const u64 *) (dK + cdgsort::ROUTE_MAX);
    CK(ctx->dRouteBase.reserve(cdgsort::RADIX * 8));
    unsigned long long *cursor = ctx->dRouteBase.as<unsigned long long>();      // records already sent per destination (one entry per digit of the pass)
    CK(cudaMemsetAsync(cursor, 0, cdgsort::RADIX * 8, ctx->st));
    CK(cudaMemsetAsync(sc + SC_RTOVF, 0, 8, ctx->st));
    // chunk boundaries: equal shares of the slice's ranks; their first slots from the prefix array
    std::vector<u32> cpos(chunks + 1); std::vector<u64> cslot(chunks + 1);
    for (int c = 0; c <= chunks; c++) cpos[c] = (u32) (K.posLo + nSlice * (u64) c / (u64) chunks);
    for (int c = 0; c <= chunks; c++) {
        if (cpos[c] >= ctx->n) cslot[c] = K.slotBase + K.slots;
        else CK(cudaMemcpyAsync(&cslot[c], ctx->dSlotRank.as<u64>() + cpos[c], 8, cudaMemcpyDeviceToHost, ctx->st));
    }
    CK(cudaStreamSynchronize(ctx->st));
    cslot[chunks] = K.slotBase + K.slots;
    const int ownerBits = 4;   // all 16 destinations: the ranks and the discard bin
    const cdgsort::Pass pass{0, CDG_OWNER_SHIFT, (1u << ownerBits) - 1};
    if ((int) ctx->chunkDone.size() < chunks + 1) {
        const size_t old = ctx->chunkDone.size();
        ctx->chunkDone.resize(chunks + 1);
        for (size_t i = old; i < ctx->chunkDone.size(); i++) CK(cudaEventCreateWithFlags(&ctx->chunkDone[i], cudaEventDisableTiming));
    }
    auto routeRange = [&](u64 slotLo, u64 slotHi) -> int {   // on the copy stream: digit words = values (owner tag), payload = keys
        const u64 cnt = slotHi - slotLo;
        if (!cnt) return 0;
        const u64 tiles = cdgsort::route_tiles(cnt);
        CK(cudaMemsetAsync(ws.tileState, 0, tiles * cdgsort::RADIX * sizeof(u32), ctx->copySt));
        CK(cudaMemsetAsync(ws.ticket, 0, sizeof(u32), ctx->copySt));
        cdgsort::launch_route_pass(ctx->copySt, ctx->dValsA.as<u64>() + slotLo, ctx->dKeysA.as<u64>() + slotLo, cnt, pass, cursor, ws.tileState, ws.ticket, dV, dK, dCap,
                                   (u32 *) (sc + SC_RTOVF));
        cdgsort::route_advance_kernel<<<1, 32, 0, ctx->copySt>>>(ws.tileState, tiles, cursor);
        ctx->launches += 2;
        return 0;
    };
    CK(cudaEventRecord(ctx->evReady, ctx->st));
    CK(cudaStreamWaitEvent(ctx->copySt, ctx->evReady, 0));   // tables, cursors and the spare slots are in place
    for (int c = 0; c < chunks; c++) {
        KmPlan Kc = K; Kc.posLo = cpos[c]; Kc.posHi = cpos[c + 1];
        if (Kc.posHi > Kc.posLo) {
            // identity records of the chunk land at their slice position (the kernels index them relative to the chunk's first rank)
            if ((rc = km_extract(ctx, p, Kc, ctx->dKeysA.as<u64>(), ctx->dValsA.as<u64>(), ctx->dIdKeysA.as<u64>() + (cpos[c] - K.posLo),
                                 ctx->dIdValsA.as<u64>() + (cpos[c] - K.posLo), K.slots, 2, 0, 65535, discard))) return rc;
        }
        CK(cudaEventRecord(ctx->chunkDone[c], ctx->st));
        CK(cudaStreamWaitEvent(ctx->copySt, ctx->chunkDone[c], 0));
        if ((rc = routeRange(cslot[c] - K.slotBase, cslot[c + 1] - K.slotBase))) return rc;
    }
    if ((rc = routeRange(K.slots, R))) return rc;     // the spare slots (identity hashes inside the k-mer key range)
    CKL();
    unsigned long long hc[17];
    CK(cudaMemcpyAsync(hc, cursor, 16 * 8, cudaMemcpyDeviceToHost, ctx->copySt));
    CK(cudaMemcpyAsync(hc + 16, sc + SC_RTOVF, 8, cudaMemcpyDeviceToHost, ctx->copySt));
    CK(cudaStreamSynchronize(ctx->copySt));
    CK(cudaStreamSynchronize(ctx->st));
    for (u32 o = 0; o < world; o++) send_counts[o] = hc[o];
    *overflow = (u32) hc[16] ? 1 : 0;
    // identity records: counted here, sent by cdg_mg_scatter_records(NULL, id offsets) once the ranks know each other's counts
    for (u32 o = 0; o < world; o++) id_send_counts[o] = 0;
    if (nSlice) {
        const int ob = std::max(1, ceil_log2((u64) world));
        cdgsort::Plan plan = cdgsort::make_plan(0, 0, CDG_OWNER_SHIFT, CDG_OWNER_SHIFT + ob);
        CK(cudaMemsetAsync(ws.hist, 0, sizeof(unsigned long long) * cdgsort::RADIX, ctx->st));
        int hb = (int) std::min<u64>((nSlice + cdgsort::BLOCK - 1) / cdgsort::BLOCK, (u64) ctx->smCount * 8);
        cdgsort::hist_kernel<<<hb, cdgsort::BLOCK, cdgsort::RADIX * sizeof(u32), ctx->st>>>(ctx->dIdValsA.as<u64>(), nullptr, nSlice, plan, ws.hist);
        ctx->launches++;
        unsigned long long hh[cdgsort::RADIX];
        CK(cudaMemcpyAsync(hh, ws.hist, sizeof hh, cudaMemcpyDeviceToHost, ctx->st));
        CK(cudaStreamSynchronize(ctx->st));
        for (u32 o = 0; o < world; o++) id_send_counts[o] = hh[o];
    }
    ctx->mgSlice = nSlice; ctx->mgMainK = ctx->dKeysA.as<u64>(); ctx->mgMainV = ctx->dValsA.as<u64>(); ctx->mgIdK = ctx->dIdKeysA.as<u64>(); ctx->mgIdV = ctx->dIdValsA.as<u64>();
    return CDG_OK;
}

// receiver side of the streamed exchange: region r of the receive area holds counts[r] records, the rest becomes sentinels
__global__ void __launch_bounds__(256) region_fill_kernel(u64 *__restrict__ keys, u64 *__restrict__ vals, const unsigned long long *__restrict__ counts, u32 world,
                                                          u64 cap) {
    const u32 r = blockIdx.y;
    if (r >= world) return;
    for (u64 i = counts[r] + (u64) blockIdx.x * blockDim.x + threadIdx.x; i < cap; i += (u64) gridDim.x * blockDim.x) {
        keys[(u64) r * cap + i] = CDG_SENTINEL; vals[(u64) r * cap + i] = 0;
    }
}
extern "C" int cdg_mg_recv_fill(cdg_ctx *ctx, const uint64_t *counts, uint64_t cap_per_source) {
    if (!ctx || !counts) return CDG_ERR_ARG;
    CK(cudaSetDevice(ctx->device));
    const u32 world = ctx->plan.world;
    unsigned long long *sc = ctx->dScalars.as<unsigned long long>();
    unsigned long long hc[16] = {};
    for (u32 r = 0; r < world; r++) { if (counts[r] > cap_per_source) return fail(ctx, CDG_ERR_ARG, "cdg_mg_recv_fill: a source sent more than a region holds"); hc[r] = counts[r]; }
    CK(cudaMemcpyAsync(sc + SC_ROUTE, hc, sizeof hc, cudaMemcpyHostToDevice, ctx->st));
    region_fill_kernel<<<dim3(64, world), 256, 0, ctx->st>>>(ctx->dRecvK.as<u64>(), ctx->dRecvV.as<u64>(), sc + SC_ROUTE, world, cap_per_source);
    ctx->launches++;
    CKL();
    CK(cudaStreamSynchronize(ctx->st));   // (hc is a stack array)
    ctx->mgRecv = (u64) world * cap_per_source;
    return CDG_OK;
}

// receive areas for `total` routed k-mer records and `id_total` identity records; returns device pointers the exchange writes into
extern "C" int cdg_mg_recv_records(cdg_ctx *ctx, uint64_t total, uint64_t id_total, void **recv_keys, void **recv_vals, void **id_recv_keys,
                                   void **id_recv_vals) {
    if (!ctx || !recv_keys || !recv_vals || !id_recv_keys || !id_recv_vals) return CDG_ERR_ARG;
    CK(cudaSetDevice(ctx->device));
    const u64 minBytes = 8ull << 20;   // own allocation block each (IPC maps whole blocks)
    CK(ctx->dRecvK.reserve(std::max<u64>((total + 1) * 8, minBytes))); CK(ctx->dRecvV.reserve(std::max<u64>((total + 1) * 8, minBytes)));
    CK(ctx->dRecvIdK.reserve(std::max<u64>((id_total + 1) * 8, minBytes))); CK(ctx->dRecvIdV.reserve(std::max<u64>((id_total + 1) * 8, minBytes)));
    ctx->mgRecv = total; ctx->nIdent = id_total;
    *recv_keys = ctx->dRecvK.p; *recv_vals = ctx->dRecvV.p; *id_recv_keys = ctx->dRecvIdK.p; *id_recv_vals = ctx->dRecvIdV.p;
    return CDG_OK;
}

extern "C" int cdg_mg_sort(cdg_ctx *ctx, const cdg_kmer_params *p, uint64_t *min_key) {
    if (!ctx || !p || !min_key) return CDG_ERR_ARG;
    CK(cudaSetDevice(ctx->device));
    using namespace cdgkmer;
    const KmPlan &K = ctx->plan;
    const u64 Rtot = ctx->mgRecv;
    if (Rtot > cdgsort::MAX_N) return fail(ctx, CDG_ERR_UNSUPPORTED, "too many routed records for one sort");
    int rc;
    // the send buffers of phase 1 are free now: they become the ping-pong partners of the receive areas
    CK(ctx->dKeysB.reserve((Rtot + 1) * 8)); CK(ctx->dValsB.reserve((Rtot + 1) * 8));
    CK(ctx->dIdKeysA.reserve((ctx->nIdent + 1) * 8)); CK(ctx->dIdValsA.reserve((ctx->nIdent + 1) * 8));
    // identity records arrived ordered by source rank = (length desc, id asc) order: the stable sort keeps the representative first
    ctx->idSrcK = ctx->dRecvIdK.as<u64>(); ctx->idSrcV = ctx->dRecvIdV.as<u64>(); ctx->idAltK = ctx->dIdKeysA.as<u64>(); ctx->idAltV = ctx->dIdValsA.as<u64>();
    u64 mk = 0;
    if ((rc = km_sort_group(ctx, p, K, ctx->dRecvK.as<u64>(), ctx->dKeysB.as<u64>(), ctx->dRecvV.as<u64>(), ctx->dValsB.as<u64>(), Rtot, -1, &mk, nullptr))) return rc;
    *min_key = mk;
    return CDG_OK;
}

// has_global_first_run: this rank's smallest key is the smallest of all ranks (ties: lowest rank)
extern "C" int cdg_mg_group(cdg_ctx *ctx, const cdg_kmer_params *p, int has_global_first_run, int partition, uint64_t *send_counts, void **send_key2) {
    if (!ctx || !p || !send_counts) return CDG_ERR_ARG;
    CK(cudaSetDevice(ctx->device));
    using namespace cdgkmer;
    const KmPlan &K = ctx->plan;
    const u64 Rtot = ctx->mgRecv;
    unsigned long long *sc = ctx->dScalars.as<unsigned long long>();
    int rc;
    u64 R2 = 0;
    // identity hashes are >= 4^k: the identity array can only hold the global first run if no rank has any k-mer record
    if ((rc = km_group(ctx, p, K, Rtot, has_global_first_run != 0, false, &R2))) return rc;
    ctx->nGrouped = R2;
    // route to the owner of the representative's index range
    const int repShift = K.idBits + K.diagBits + 1;
    unsigned long long *counts = sc + SC_ROUTE;
    CK(cudaMemsetAsync(counts, 0, 16 * 8, ctx->st));
    unsigned long long hc[16] = {0};
    if (R2) {
        const int blocks = (int) std::min<u64>((R2 + 255) / 256, (u64) ctx->smCount * 16);
        route_count_kernel<<<blocks, 256, 0, ctx->st>>>(ctx->dKey2A.as<u64>(), R2, repShift, K.repChunk, K.world, counts);
        CK(cudaMemcpyAsync(hc, counts, 16 * 8, cudaMemcpyDeviceToHost, ctx->st));
        CK(cudaStreamSynchronize(ctx->st));
        ctx->launches++;
        if (partition) {
            unsigned long long cur[16], acc = 0;
            for (u32 o = 0; o < 16; o++) { cur[o] = acc; acc += o < K.world ? hc[o] : 0; }
            CK(cudaMemcpyAsync(counts, cur, 16 * 8, cudaMemcpyHostToDevice, ctx->st));
            route_scatter_kernel<<<blocks, 256, 0, ctx->st>>>(ctx->dKey2A.as<u64>(), R2, repShift, K.repChunk, K.world, counts, ctx->dKey2B.as<u64>());
            ctx->launches++;
        }
        CKL();
    }
    for (u32 o = 0; o < K.world; o++) send_counts[o] = hc[o];
    CK(cudaStreamSynchronize(ctx->st));
    if (send_key2) *send_key2 = ctx->dKey2B.p;
    return CDG_OK;
}

extern "C" int cdg_mg_recv_grouped(cdg_ctx *ctx, uint64_t total, void **recv_key2) {
    if (!ctx || !recv_key2) return CDG_ERR_ARG;
    CK(cudaSetDevice(ctx->device));
    // dKey2A (group output, already scattered to dKey2B) is reused as the receive area; the exchange reads dKey2B
    CK(ctx->dRecv2.reserve(std::max<u64>((total + CDG_HEAD_RUN + 1) * 8, 8ull << 20)));
    ctx->mgR2 = total;
    *recv_key2 = ctx->dRecv2.p;
    return CDG_OK;
}

// ---- peer-memory exchange (NVLink, CUDA IPC): the partition pass of the sender writes straight into the receivers ----------
static DevBuf *ipc_buffer(cdg_ctx *ctx, int which) {
    switch (which) {
        case 0: return &ctx->dRecvK; case 1: return &ctx->dRecvV; case 2: return &ctx->dRecvIdK; case 3: return &ctx->dRecvIdV; case 4: return &ctx->dRecv2;
        default: return nullptr;
    }
}
// handle of receive buffer `which` (after cdg_mg_recv_records / cdg_mg_recv_grouped reserved it): 64-byte cudaIpcMemHandle_t
// followed by the buffer's byte offset inside its allocation
extern "C" int cdg_mg_ipc_export(cdg_ctx *ctx, int which, void *handle_out) {
    DevBuf *b = ctx ? ipc_buffer(ctx, which) : nullptr;
    if (!b || !handle_out || !b->p) return CDG_ERR_ARG;
    CK(cudaSetDevice(ctx->device));
    cudaIpcMemHandle_t h;
    CK(cudaIpcGetMemHandle(&h, b->p));
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    u64 offset = 0;
    {   // allocation base through the driver entry point (no link-time dependency on libcuda)
        typedef int (*GetRange)(unsigned long long *, size_t *, unsigned long long);
        void *fn = nullptr; cudaDriverEntryPointQueryResult qr;
        if (cudaGetDriverEntryPoint("cuMemGetAddressRange", &fn, cudaEnableDefault, &qr) == cudaSuccess && fn) {
            unsigned long long base = 0; size_t sz = 0;
            if (((GetRange) fn)(&base, &sz, (unsigned long long) (uintptr_t) b->p) == 0 && base) offset = (u64) (uintptr_t) b->p - base;
        }
    }
    memcpy(handle_out, &h, 64); memcpy((unsigned char *) handle_out + 64, &offset, 8);
    return CDG_OK;
}
// bytes receive buffer `which` can hold without being reallocated (0 = not allocated yet).  A buffer must not be freed while a
// peer still maps it: before a rank grows a buffer every peer drops its mapping (cdg_mg_ipc_close) — see multigpu.py.
extern "C" int cdg_mg_ipc_capacity(cdg_ctx *ctx, int which, uint64_t *bytes) {
    DevBuf *b = ctx ? ipc_buffer(ctx, which) : nullptr;
    if (!b || !bytes) return CDG_ERR_ARG;
    *bytes = b->cap;
    return CDG_OK;
}
extern "C" int cdg_mg_ipc_close(cdg_ctx *ctx, int which, int peer) {
    if (!ctx || which < 0 || which >= CDG_IPC_BUFFERS || peer < 0 || peer >= CDG_MAX_WORLD) return CDG_ERR_ARG;
    CK(cudaSetDevice(ctx->device));
    if (ctx->peerBase[which][peer]) {
        CK(cudaStreamSynchronize(ctx->st));
        CK(cudaIpcCloseMemHandle(ctx->peerBase[which][peer]));
        ctx->peerBase[which][peer] = nullptr; ctx->peerPtr[which][peer] = nullptr;
        memset(ctx->peerHandle[which][peer], 0, CDG_IPC_HANDLE_BYTES);
    }
    return CDG_OK;
}
// maps peer `peer`'s receive buffer `which` into this process (cached while the peer's handle stays the same)
extern "C" int cdg_mg_ipc_open(cdg_ctx *ctx, int which, int peer, const void *handle) {
    if (!ctx || !handle || which < 0 || which >= CDG_IPC_BUFFERS || peer < 0 || peer >= CDG_MAX_WORLD) return CDG_ERR_ARG;
    CK(cudaSetDevice(ctx->device));
    if (ctx->peerBase[which][peer] && memcmp(ctx->peerHandle[which][peer], handle, CDG_IPC_HANDLE_BYTES) == 0) return CDG_OK;
    if (ctx->peerBase[which][peer]) { CK(cudaIpcCloseMemHandle(ctx->peerBase[which][peer])); ctx->peerBase[which][peer] = nullptr; ctx->peerPtr[which][peer] = nullptr; }
    cudaIpcMemHandle_t h; u64 offset = 0;
    memcpy(&h, handle, 64); memcpy(&offset, (const unsigned char *) handle + 64, 8);
    void *base = nullptr;
    CK(cudaIpcOpenMemHandle(&base, h, cudaIpcMemLazyEnablePeerAccess));
    ctx->peerBase[which][peer] = base; ctx->peerPtr[which][peer] = (unsigned char *) base + offset;
    memcpy(ctx->peerHandle[which][peer], handle, CDG_IPC_HANDLE_BYTES);
    return CDG_OK;
}

// Send side of the record all-to-all: ONE stable partition pass per array pair whose per-owner outputs are the peers' receive
// buffers.  dst_off[o] / id_dst_off[o] = first record index of this rank's segment in rank o's k-mer / identity receive area
// (= records the lower ranks send to o).  Own segment: the local buffer.  The caller synchronises the ranks afterwards.
extern "C" int cdg_mg_scatter_records(cdg_ctx *ctx, const uint64_t *dst_off, const uint64_t *id_dst_off) {
    if (!ctx || !id_dst_off) return CDG_ERR_ARG;   // dst_off == NULL: identity records only (the k-mer records went with cdg_mg_extract_stream)
    CK(cudaSetDevice(ctx->device));
    const KmPlan &K = ctx->plan;
    const u32 world = K.world, rank = K.rank;
    const u64 R = K.slots + K.extraCap, nSlice = ctx->mgSlice;
    int rc;
    cdgsort::Workspace ws;
    if ((rc = sort_ws(ctx, std::max<u64>(R, ctx->n), ws))) return rc;
    CK(ctx->dRouteTab.reserve(4 * cdgsort::ROUTE_MAX * sizeof(void *)));
    u64 *tab[4][cdgsort::ROUTE_MAX] = {};
    for (u32 o = 0; o < world; o++) {
        DevBuf *own[4] = {&ctx->dRecvV, &ctx->dRecvK, &ctx->dRecvIdV, &ctx->dRecvIdK};   // digit word = value (owner bits), payload = key
        const int which[4] = {1, 0, 3, 2};
        for (int a = 0; a < 4; a++) {
            if (a < 2 && !dst_off) continue;
            u64 *base = o == rank ? own[a]->as<u64>() : (u64 *) ctx->peerPtr[which[a]][o];
            if (!base) return fail(ctx, CDG_ERR_ARG, "cdg_mg_scatter_records: receive buffer %d of rank %u is not mapped (cdg_mg_ipc_open)", which[a], o);
            tab[a][o] = base + (a < 2 ? dst_off[o] : id_dst_off[o]);
        }
    }
    CK(cudaMemcpyAsync(ctx->dRouteTab.p, tab, sizeof tab, cudaMemcpyHostToDevice, ctx->st));
    u64 *const *dtab = ctx->dRouteTab.as<u64 *>();
    const int ownerBits = std::max(1, ceil_log2((u64) world));
    const cdgsort::Pass pass{0, CDG_OWNER_SHIFT, (1u << ownerBits) - 1};
    auto route = [&](const u64 *digitWords, const u64 *payload, u64 cnt, u64 *const *tV, u64 *const *tK) -> int {
        if (!cnt) return 0;
        CK(cudaMemsetAsync(ws.hist, 0, sizeof(unsigned long long) * cdgsort::RADIX, ctx->st));
        CK(cudaMemsetAsync(ws.tileState, 0, cdgsort::tiles_for(cnt) * cdgsort::RADIX * sizeof(u32), ctx->st));
        CK(cudaMemsetAsync(ws.ticket, 0, sizeof(u32), ctx->st));
        cdgsort::launch_route_pass(ctx->st, digitWords, payload, cnt, pass, ws.hist, ws.tileState, ws.ticket, tV, tK);
        ctx->launches++;
        return 0;
    };
    if (dst_off && (rc = route(ctx->mgMainV, ctx->mgMainK, R, dtab, dtab + cdgsort::ROUTE_MAX))) return rc;
    if ((rc = route(ctx->mgIdV, ctx->mgIdK, nSlice, dtab + 2 * cdgsort::ROUTE_MAX, dtab + 3 * cdgsort::ROUTE_MAX))) return rc;
    CKL();
    CK(cudaStreamSynchronize(ctx->st));
    return CDG_OK;
}

// Send side of the grouped-record all-to-all: block-aggregated append into the peers' receive buffers (order is irrelevant,
// sort 2 follows).  dst_off[o] = first record index of this rank's segment in rank o's receive area.
extern "C" int cdg_mg_scatter_grouped(cdg_ctx *ctx, const uint64_t *dst_off) {
    if (!ctx || !dst_off) return CDG_ERR_ARG;
    CK(cudaSetDevice(ctx->device));
    using namespace cdgkmer;
    const KmPlan &K = ctx->plan;
    const u64 R2 = ctx->nGrouped;
    unsigned long long *sc = ctx->dScalars.as<unsigned long long>();
    CK(ctx->dRouteTab.reserve(4 * cdgsort::ROUTE_MAX * sizeof(void *)));
    u64 *tab[cdgsort::ROUTE_MAX] = {};
    for (u32 o = 0; o < K.world; o++) {
        u64 *base = o == K.rank ? ctx->dRecv2.as<u64>() : (u64 *) ctx->peerPtr[4][o];
        if (!base) return fail(ctx, CDG_ERR_ARG, "cdg_mg_scatter_grouped: receive buffer of rank %u is not mapped (cdg_mg_ipc_open)", o);
        tab[o] = base + dst_off[o];
    }
    CK(cudaMemcpyAsync(ctx->dRouteTab.p, tab, sizeof tab, cudaMemcpyHostToDevice, ctx->st));
    if (R2) {
        unsigned long long *cursors = sc + SC_ROUTE;
        CK(cudaMemsetAsync(cursors, 0, 16 * 8, ctx->st));
        const int blocks = (int) std::min<u64>((R2 + 255) / 256, (u64) ctx->smCount * 16);
        route_scatter_kernel<<<blocks, 256, 0, ctx->st>>>(ctx->dKey2A.as<u64>(), R2, K.idBits + K.diagBits + 1, K.repChunk, K.world, cursors, nullptr,
                                                          ctx->dRouteTab.as<u64 *>());
        ctx->launches++;
        CKL();
    }
    CK(cudaStreamSynchronize(ctx->st));
    return CDG_OK;
}

// leading records of the sorted grouped array that share the first record's target id (exact, any length: one block gallops
// over the array until the first record with another id)
__global__ void __launch_bounds__(1024) head_run_kernel(const u64 *__restrict__ keys, u64 n, int idShift, u64 idMask, unsigned long long *__restrict__ out) {
    __shared__ unsigned long long firstOther;
    if (threadIdx.x == 0) firstOther = ~0ull;
    __syncthreads();
    if (n == 0) { if (threadIdx.x == 0) *out = 0; return; }
    const u64 id0 = (keys[0] >> idShift) & idMask;
    for (u64 base = 0; base < n; base += 1024) {
        const u64 i = base + threadIdx.x;
        if (i < n && ((keys[i] >> idShift) & idMask) != id0) atomicMin(&firstOther, (unsigned long long) i);
        __syncthreads();
        if (firstOther != ~0ull) break;
    }
    if (threadIdx.x == 0) *out = firstOther == ~0ull ? n : firstOther;
}

// sorts the received grouped records; *sorted points at the sorted array; *head_count = number of leading records that share
// the first record's target id (they go to the previous rank: rep-boundary bleed, kmermatcher.cpp:886-888); *head_capacity =
// records of a neighbour's head run that fit behind the sorted array without moving it (cdg_mg_head_area handles more)
extern "C" int cdg_mg_sort2(cdg_ctx *ctx, void **sorted, uint64_t *head_count, uint64_t *head_capacity) {
    if (!ctx || !sorted || !head_count) return CDG_ERR_ARG;
    CK(cudaSetDevice(ctx->device));
    const KmPlan &K = ctx->plan;
    const u64 R2 = ctx->mgR2;
    int rc;
    CK(ctx->dKey2A.reserve((R2 + CDG_HEAD_RUN + 1) * 8));
    if ((rc = km_sort2(ctx, K, R2, ctx->dRecv2.as<u64>(), ctx->dKey2A.as<u64>()))) return rc;   // sorted in place of the receive area and its partner
    unsigned long long *sc = ctx->dScalars.as<unsigned long long>();
    head_run_kernel<<<1, 1024, 0, ctx->st>>>(ctx->sortedKey2, R2, K.diagBits + 1, (1ull << K.idBits) - 1, sc + SC_SCAN);
    ctx->launches++;
    CKL();
    u64 hcnt = 0;
    if ((rc = read_scalar(ctx, SC_SCAN, &hcnt))) return rc;
    ctx->headCap = CDG_HEAD_RUN;
    *sorted = ctx->sortedKey2; *head_count = hcnt;
    if (head_capacity) *head_capacity = CDG_HEAD_RUN;
    return CDG_OK;
}

// where the next rank's head run of n_incoming records goes: directly behind the sorted records.  Runs longer than the spare
// area (deep pile-ups, overlapping contigs) move the sorted array into a larger buffer first — no run is ever cut.
extern "C" int cdg_mg_head_area(cdg_ctx *ctx, uint64_t n_incoming, void **dst) {
    if (!ctx || !dst) return CDG_ERR_ARG;
    CK(cudaSetDevice(ctx->device));
    const u64 R2 = ctx->mgR2;
    if (n_incoming > ctx->headCap) {
        CK(ctx->dKey2Big.reserve((R2 + n_incoming + 1) * 8));
        CK(cudaMemcpyAsync(ctx->dKey2Big.p, ctx->sortedKey2, R2 * 8, cudaMemcpyDeviceToDevice, ctx->st));
        CK(cudaStreamSynchronize(ctx->st));
        ctx->sortedKey2 = ctx->dKey2Big.as<u64>();
        ctx->headCap = n_incoming;
    }
    *dst = ctx->sortedKey2 + R2;
    return CDG_OK;
}

// n_appended records of the next rank's head run have been written behind the rank's own sorted records
static int mg_finish_core(cdg_ctx *ctx, const cdg_rescore_params *rp, const cdg_corr_params *cp, uint64_t n_appended, cdg_hits *hits, cdg_alns *alns,
                          cdg_alns_compact *compact, cdg_patches *patches, uint8_t *out_range, uint64_t *range_offset, uint64_t *range_bytes,
                          uint64_t *n_changed) {
    if (!ctx || !rp || !cp) return CDG_ERR_ARG;
    CK(cudaSetDevice(ctx->device));
    const KmPlan &K = ctx->plan;
    const u64 R2 = ctx->mgR2;
    if (n_appended > ctx->headCap) return fail(ctx, CDG_ERR_ARG, "cdg_mg_finish: head run longer than the reserved area (call cdg_mg_head_area first)");
    if (!ctx->monotoneOffsets && (out_range || patches))
        return fail(ctx, CDG_ERR_UNSUPPORTED, "cdg_mg_finish: the DB entries are not laid out in index order, so a query range is not a byte range; "
                                              "rewrite the DB in index order (single-threaded writer) for the multi-GPU phases");
    int rc;
    if ((rc = km_reduce(ctx, K, R2, R2 + n_appended))) return rc;
    if (hits && (rc = export_hits(ctx, hits))) return rc;
    if ((rc = run_rescore(ctx, rp))) return rc;
    if (alns && (rc = export_alns(ctx, alns, true))) return rc;
    else if (compact && (rc = export_alns_compact(ctx, compact, true))) return rc;
    if (!alns && !compact) { u64 kept = 0; if ((rc = read_scalar(ctx, SC_KEPT, &kept))) return rc; ctx->nKept = kept; }
    u64 ch = 0;
    if ((rc = run_correct(ctx, cp, nullptr, &ch, nullptr, out_range == nullptr))) return rc;
    if (n_changed) *n_changed = ch;
    // corrected bytes of the rank's query range
    u64 b0 = 0, b1 = ctx->dataBytes;
    if (K.nQ) {
        CK(cudaMemcpyAsync(&b0, ctx->dOff.as<u64>() + K.qLo, 8, cudaMemcpyDeviceToHost, ctx->st));
        if (K.qLo + K.nQ < ctx->n) CK(cudaMemcpyAsync(&b1, ctx->dOff.as<u64>() + K.qLo + K.nQ, 8, cudaMemcpyDeviceToHost, ctx->st));
        CK(cudaStreamSynchronize(ctx->st));
    } else b0 = b1 = 0;
    if (range_offset) *range_offset = b0;
    if (range_bytes) *range_bytes = b1 - b0;
    if (patches && (rc = export_patches(ctx, b0, b1, ch, patches))) return rc;
    if (out_range && b1 > b0) CK(cudaMemcpyAsync(out_range, ctx->dOut.as<u8>() + b0, b1 - b0, cudaMemcpyDeviceToHost, ctx->st));
    CK(cudaStreamSynchronize(ctx->st));
    return wait_exports(ctx);
}
extern "C" int cdg_mg_finish(cdg_ctx *ctx, const cdg_rescore_params *rp, const cdg_corr_params *cp, uint64_t n_appended, cdg_hits *hits, cdg_alns *alns,
                             uint8_t *out_range, uint64_t *range_offset, uint64_t *range_bytes, uint64_t *n_changed) {
    return mg_finish_core(ctx, rp, cp, n_appended, hits, alns, nullptr, nullptr, out_range, range_offset, range_bytes, n_changed);
}
// same with the compact outputs of cdg_iteration_compact for the rank's query range (patch offsets are offsets into the whole DB)
extern "C" int cdg_mg_finish_compact(cdg_ctx *ctx, const cdg_rescore_params *rp, const cdg_corr_params *cp, uint64_t n_appended, cdg_alns_compact *alns,
                                     cdg_patches *patches, uint64_t *range_offset, uint64_t *range_bytes, uint64_t *n_changed) {
    return mg_finish_core(ctx, rp, cp, n_appended, nullptr, nullptr, alns, patches, nullptr, range_offset, range_bytes, n_changed);
}

// ------------------------------------------------------------------------------------------------------
extern "C" int cdg_test_radix_sort(cdg_ctx *ctx, uint64_t *keys, uint64_t *vals, uint64_t n, int bit_lo, int bit_hi) {
    if (!ctx || !keys) return CDG_ERR_ARG;
    CK(cudaSetDevice(ctx->device));
    DevBuf ka, kb, va, vb;
    CK(ka.reserve(n * 8 + 8)); CK(kb.reserve(n * 8 + 8));
    if (vals) { CK(va.reserve(n * 8 + 8)); CK(vb.reserve(n * 8 + 8)); }
    CK(cudaMemcpyAsync(ka.p, keys, n * 8, cudaMemcpyHostToDevice, ctx->st));
    if (vals) CK(cudaMemcpyAsync(va.p, vals, n * 8, cudaMemcpyHostToDevice, ctx->st));
    cdgsort::Workspace ws;
    int rc = sort_ws(ctx, n, ws);
    if (rc) return rc;
    cdgsort::Plan plan = cdgsort::make_plan(0, 0, bit_lo, bit_hi);
    u64 *ok, *ov;
    int l = cdgsort::radix_sort(ctx->st, ws, ka.as<u64>(), kb.as<u64>(), vals ? va.as<u64>() : nullptr, vals ? vb.as<u64>() : nullptr, n, plan, &ok, &ov,
                                ctx->smCount);
    if (l < 0) return fail(ctx, CDG_ERR_UNSUPPORTED, "cdg_test_radix_sort: too large");
    ctx->launches += l;
    CKL();
    CK(cudaMemcpyAsync(keys, ok, n * 8, cudaMemcpyDeviceToHost, ctx->st));
    if (vals) CK(cudaMemcpyAsync(vals, ov, n * 8, cudaMemcpyDeviceToHost, ctx->st));
    CK(cudaStreamSynchronize(ctx->st));
    ka.release(); kb.release(); va.release(); vb.release();
    return CDG_OK;
}

__global__ void hash_rc_kernel(const u64 *in, u64 n, int k, u64 seed, u64 *h, u64 *rc) {
    u64 i = (u64) blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) { h[i] = cdg_xxh64_u64(in[i], seed); rc[i] = cdgkmer::revcomp_kmer_dev(in[i], k); }   // the kernels' BREV form (oracle: the reference's)
}
extern "C" int cdg_test_hash_revcomp(cdg_ctx *ctx, const uint64_t *in, uint64_t n, int k, uint64_t seed, uint64_t *hash_out, uint64_t *rc_out) {
    if (!ctx || !in || !hash_out || !rc_out) return CDG_ERR_ARG;
    CK(cudaSetDevice(ctx->device));
    DevBuf a, b, c;
    CK(a.reserve(n * 8 + 8)); CK(b.reserve(n * 8 + 8)); CK(c.reserve(n * 8 + 8));
    CK(cudaMemcpyAsync(a.p, in, n * 8, cudaMemcpyHostToDevice, ctx->st));
    hash_rc_kernel<<<(unsigned) ((n + 255) / 256), 256, 0, ctx->st>>>(a.as<u64>(), n, k, seed, b.as<u64>(), c.as<u64>());
    ctx->launches++;
    CKL();
    CK(cudaMemcpyAsync(hash_out, b.p, n * 8, cudaMemcpyDeviceToHost, ctx->st));
    CK(cudaMemcpyAsync(rc_out, c.p, n * 8, cudaMemcpyDeviceToHost, ctx->st));
    CK(cudaStreamSynchronize(ctx->st));
    a.release(); b.release(); c.release();
    return CDG_OK;
}

__global__ void fill_random_kernel(u64 *k, u64 *v, u64 n, u64 seed) {
    u64 i = (u64) blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) { k[i] = cdg_xxh64_u64(i, seed); if (v) v[i] = i; }
}
// order check of a sort result: adjacent pairs out of order on the sorted bits (and, with values = original indices, unstable
// ties), plus wrapping sums of keys and values to compare with the input (keys xxh64(i), values i: sums are known in closed form
// only for the values, the key sum is taken before the sort)
__global__ void sort_check_kernel(const u64 *__restrict__ k, const u64 *__restrict__ v, u64 n, u64 mask, unsigned long long *out /*[3]: bad, sumK, sumV*/) {
    unsigned long long bad = 0, sk = 0, sv = 0;
    for (u64 i = (u64) blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (u64) gridDim.x * blockDim.x) {
        const u64 a = k[i];
        sk += a;
        if (v) sv += v[i];
        if (i + 1 < n) {
            const u64 b = k[i + 1];
            if ((a & mask) > (b & mask)) bad++;
            else if (v && (a & mask) == (b & mask) && v[i] > v[i + 1]) bad++;
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { bad += __shfl_xor_sync(0xffffffffu, bad, o); sk += __shfl_xor_sync(0xffffffffu, sk, o); sv += __shfl_xor_sync(0xffffffffu, sv, o); }
    if ((threadIdx.x & 31) == 0) { if (bad) atomicAdd(out, bad); atomicAdd(out + 1, sk); atomicAdd(out + 2, sv); }
}
// sort micro-benchmark on device-generated random keys: average milliseconds per full sort (histogram + all passes)
extern "C" int cdg_test_sort_bench(cdg_ctx *ctx, uint64_t n, int bit_lo, int bit_hi, int has_val, int variant, int reps, float *ms_per_sort,
                                   int *passes) {
    if (!ctx || !ms_per_sort) return CDG_ERR_ARG;
    CK(cudaSetDevice(ctx->device));
    DevBuf ka, kb, va, vb;
    CK(ka.reserve(n * 8 + 8)); CK(kb.reserve(n * 8 + 8));
    if (has_val) { CK(va.reserve(n * 8 + 8)); CK(vb.reserve(n * 8 + 8)); }
    cdgsort::Workspace ws;
    int rc = sort_ws(ctx, n, ws);
    if (rc) return rc;
    cdgsort::Plan plan = cdgsort::make_plan(0, 0, bit_lo, bit_hi);
    if (passes) *passes = plan.n;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    float total = 0.f;
    for (int r = 0; r < reps + 1; r++) {
        fill_random_kernel<<<(unsigned) ((n + 255) / 256), 256, 0, ctx->st>>>(ka.as<u64>(), has_val ? va.as<u64>() : nullptr, n, 1234 + r);
        unsigned long long *chk = ctx->dScalars.as<unsigned long long>() + SC_ROUTE;   // 6 scratch words: input sums, output check
        const u64 sortMask = (bit_hi >= 64 ? ~0ull : ((1ull << bit_hi) - 1)) & ~((1ull << bit_lo) - 1);
        if (r == reps) {
            CK(cudaMemsetAsync(chk, 0, 6 * 8, ctx->st));
            sort_check_kernel<<<ctx->smCount * 8, 256, 0, ctx->st>>>(ka.as<u64>(), has_val ? va.as<u64>() : nullptr, n, 0, chk);
        }
        cudaEventRecord(e0, ctx->st);
        u64 *ok, *ov;
        int l = cdgsort::radix_sort(ctx->st, ws, ka.as<u64>(), kb.as<u64>(), has_val ? va.as<u64>() : nullptr, has_val ? vb.as<u64>() : nullptr, n, plan,
                                    &ok, &ov, ctx->smCount, variant);
        if (l < 0) return fail(ctx, CDG_ERR_UNSUPPORTED, "cdg_test_sort_bench: sort refused");
        cudaEventRecord(e1, ctx->st);
        if (r == reps) {   // last repetition: the result must be a stable sorted permutation of the input
            sort_check_kernel<<<ctx->smCount * 8, 256, 0, ctx->st>>>(ok, has_val ? ov : nullptr, n, sortMask, chk + 3);
            unsigned long long h[6];
            CK(cudaMemcpyAsync(h, chk, sizeof h, cudaMemcpyDeviceToHost, ctx->st));
            CK(cudaStreamSynchronize(ctx->st));
            if (h[3] != 0 || h[4] != h[1] || h[5] != h[2])
                return fail(ctx, CDG_ERR_CUDA, "cdg_test_sort_bench: result is not a stable sorted permutation (%llu inversions, key sum %s, value sum %s)", h[3],
                            h[4] == h[1] ? "ok" : "differs", h[5] == h[2] ? "ok" : "differs");
        }
        CK(cudaStreamSynchronize(ctx->st));
        CKL();
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e0, e1);
        if (r > 0) total += ms;   // first repetition is warm-up
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    *ms_per_sort = total / reps;
    ka.release(); kb.release(); va.release(); vb.release();
    return CDG_OK;
}
extern "C" int cdg_test_set_sort_variant(int keys_only_variant, int pairs_variant) {
    if (keys_only_variant >= 0) cdgsort::default_variant(false) = keys_only_variant;
    if (pairs_variant >= 0) cdgsort::default_variant(true) = pairs_variant;
    return 0;
}
