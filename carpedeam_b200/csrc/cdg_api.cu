// libcarpedeam_gpu.so — C ABI (include/carpedeam_gpu.h) over the CUDA kernels.  sm_100a only.
#include <cuda_runtime.h>
#include <stdarg.h>
#include <algorithm>
#include <stdlib.h>
#include <string.h>
#include <string>
#include <vector>

#include "../../include/carpedeam_gpu.h"
#include "cdg_align.cuh"
#include "cdg_common.cuh"
#include "cdg_kmer.cuh"
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

struct cdg_ctx {
    int device = 0, smCount = 148;
    cudaStream_t st = nullptr;
    std::string err;
    cudaEvent_t ev[EV_N];
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
    u32 sort1Passes = 0, sort2Passes = 0;
    // lines / alignments
    DevBuf dLineOff, dLineQ, dLineT, dLinePref, dLineDiag, dKeep, dScore, dBits, dEvalue, dSeqId, dQS, dQE, dDS, dDE, dLineSel;
    HostBuf hLineOff, hLineT, hKeep, hScore, hBits, hEvalue, hSeqId, hQS, hQE, hDS, hDE;
    u64 nLines = 0, nKept = 0; bool linesValid = false, alnsValid = false, alnsFromText = false;
    // correction
    DevBuf dTables, dOut, dLik, dWideList; bool damageSet = false;
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
        if (e_ != cudaSuccess) return fail(ctx, CDG_ERR_CUDA, "%s:%d %s: %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_)); \
    } while (0)
#define CKL() CK(cudaGetLastError())

// scalars block on the device: [0] scan total, [1] emitted records, [2] grouped count, [3] kept lines,
// [4] changed bases, [5] (u32) extraCount, [6] (u32) overflow, [7] (u32) long count, [8] long scratch total, [9] (u32) sort ticket
enum { SC_SCAN = 0, SC_EMIT, SC_GROUPED, SC_KEPT, SC_CHANGED, SC_EXTRA, SC_OVERFLOW, SC_NLONG, SC_LONGSCR, SC_TICKET, SC_WIDE, SC_N = 16 };

void rec(cdg_ctx *ctx, int e) { cudaEventRecord(ctx->ev[e], ctx->st); ctx->evValid[e] = true; }

struct LenWordsF { const u32 *len; __device__ u32 operator()(u64 i) const { return (len[i] + 31) >> 5; } };
struct LenBasesF { const u32 *len; __device__ u32 operator()(u64 i) const { return len[i]; } };
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
struct CapacityF { const u32 *len; int k, kps; float scale; __device__ u32 operator()(u64 i) const { return cdgkmer::kmer_capacity(len[i], k, kps, scale); } };

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
    u64 tiles = cdgsort::tiles_for(n) + 1;
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
    for (int i = 0; i < EV_N; i++) { cudaEventCreate(&ctx->ev[i]); ctx->evValid[i] = false; }
    if (ctx->dScalars.reserve(SC_N * 8) != cudaSuccess) { delete ctx; return CDG_ERR_CUDA; }
    cudaFuncSetAttribute(cdgkmer::extract_short_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, cdgkmer::EX_WARPS * 2048 * 4);
    cudaFuncSetAttribute(cdgkmer::extract_short_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, cdgkmer::EX_WARPS * 2048 * 4);
    cudaFuncSetAttribute(cdgkmer::extract_short_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, cdgkmer::EX_WARPS * 2048 * 4);
    cudaFuncSetAttribute(cdgalign::correct_fast_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) sizeof(cdgalign::CorrSmem));
    cudaFuncSetAttribute(cdgkmer::extract_long_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (32768 + 128) * 4);
    if (const char *e = getenv("CDG_WIDE_THRESHOLD")) { long v = atol(e); if (v >= 1 && v < 65535) ctx->wideThreshold = (u32) v; }
    *out = ctx;
    return CDG_OK;
}

extern "C" void cdg_destroy(cdg_ctx *ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->st);
    DevBuf *d[] = {&ctx->dMeta, &ctx->dData, &ctx->dOff, &ctx->dLen, &ctx->dExt, &ctx->dWordOff, &ctx->dWords, &ctx->dXmask, &ctx->dFlags, &ctx->dLikOff,
                   &ctx->dTileSums, &ctx->dScalars, &ctx->dHist, &ctx->dTileState, &ctx->dSlotOff, &ctx->dKeysA, &ctx->dKeysB, &ctx->dValsA,
                   &ctx->dValsB, &ctx->dIdKeysA, &ctx->dIdKeysB, &ctx->dIdValsA, &ctx->dIdValsB, &ctx->dKey2A, &ctx->dKey2B, &ctx->dLongList,
                   &ctx->dOrdKA, &ctx->dOrdKB, &ctx->dOrdVA, &ctx->dOrdVB, &ctx->dRankOf, &ctx->dSlotRank, &ctx->dScratchOff, &ctx->dScratch, &ctx->dScratchKeep, &ctx->dShortScratch, &ctx->dShortKeep, &ctx->dHitRep, &ctx->dHitTgt,
                   &ctx->dHitScore, &ctx->dHitDiag, &ctx->dLineOff, &ctx->dLineQ, &ctx->dLineT, &ctx->dLinePref, &ctx->dLineDiag, &ctx->dKeep,
                   &ctx->dScore, &ctx->dBits, &ctx->dEvalue, &ctx->dSeqId, &ctx->dQS, &ctx->dQE, &ctx->dDS, &ctx->dDE, &ctx->dLineSel,
                   &ctx->dTables, &ctx->dOut, &ctx->dLik, &ctx->dWideList};
    for (DevBuf *b : d) b->release();
    HostBuf *h[] = {&ctx->hHitRep, &ctx->hHitTgt, &ctx->hHitScore, &ctx->hHitDiag, &ctx->hHitDiag16, &ctx->hLineOff, &ctx->hLineT, &ctx->hKeep,
                    &ctx->hScore, &ctx->hBits, &ctx->hEvalue, &ctx->hSeqId, &ctx->hQS, &ctx->hQE, &ctx->hDS, &ctx->hDE};
    for (HostBuf *b : h) b->release();
    for (int i = 0; i < EV_N; i++) cudaEventDestroy(ctx->ev[i]);
    cudaStreamDestroy(ctx->st);
    delete ctx;
}

extern "C" const char *cdg_last_error(const cdg_ctx *ctx) { return ctx ? ctx->err.c_str() : "no context (CUDA device unavailable?)"; }
extern "C" void *cdg_stream(cdg_ctx *ctx) { return ctx ? (void *) ctx->st : nullptr; }

// ------------------------------------------------------------------------------------------------------
extern "C" int cdg_upload_seqs(cdg_ctx *ctx, const uint8_t *data, uint64_t data_bytes, const uint64_t *offsets, const uint32_t *lens,
                               const uint8_t *ext, uint64_t n) {
    if (!ctx) return CDG_ERR_ARG;
    if (n == 0 || !data || !offsets || !lens) return fail(ctx, CDG_ERR_ARG, "cdg_upload_seqs: empty or NULL input");
    if (n >= (1ull << 31)) return fail(ctx, CDG_ERR_UNSUPPORTED, "cdg_upload_seqs: more than 2^31 sequences per context");
    CK(cudaSetDevice(ctx->device));
    ctx->hitsValid = ctx->linesValid = ctx->alnsValid = false;
    ctx->n = n; ctx->dataBytes = data_bytes;
    CK(ctx->dData.reserve(data_bytes + 64));
    CK(ctx->dOff.reserve(n * 8)); CK(ctx->dLen.reserve(n * 4)); CK(ctx->dExt.reserve(n)); CK(ctx->dFlags.reserve(n));
    CK(ctx->dWordOff.reserve((n + 1) * 8)); CK(ctx->dLikOff.reserve((n + 1) * 8));
    CK(cudaMemcpyAsync(ctx->dData.p, data, data_bytes, cudaMemcpyHostToDevice, ctx->st));
    CK(cudaMemcpyAsync(ctx->dOff.p, offsets, n * 8, cudaMemcpyHostToDevice, ctx->st));
    CK(cudaMemcpyAsync(ctx->dLen.p, lens, n * 4, cudaMemcpyHostToDevice, ctx->st));
    if (ext) CK(cudaMemcpyAsync(ctx->dExt.p, ext, n, cudaMemcpyHostToDevice, ctx->st));
    else CK(cudaMemsetAsync(ctx->dExt.p, 0, n, ctx->st));
    // host-side facts needed for sizing (one pass over lens)
    u64 total = 0; u32 mx = 0;
    for (u64 i = 0; i < n; i++) { total += lens[i]; if (lens[i] > mx) mx = lens[i]; }
    ctx->totalBases = total; ctx->maxLen = mx;
    int rc;
    if ((rc = scan_to(ctx, LenWordsF{ctx->dLen.as<u32>()}, n, ctx->dWordOff.as<u64>(), true))) return rc;
    u64 words = 0;
    if ((rc = read_scalar(ctx, SC_SCAN, &words))) return rc;
    ctx->totalWords = words;
    if ((rc = scan_to(ctx, LenBasesF{ctx->dLen.as<u32>()}, n, ctx->dLikOff.as<u64>(), true))) return rc;
    CK(ctx->dWords.reserve((words + 2) * 8)); CK(ctx->dXmask.reserve((words + 2) * 4));
    cdgkmer::pack_kernel<<<ctx->smCount * 16, 256, 0, ctx->st>>>(ctx->dData.as<u8>(), ctx->dOff.as<u64>(), ctx->dLen.as<u32>(),
                                                                                    ctx->dWordOff.as<u64>(), n, ctx->dWords.as<u64>(),
                                                                                    ctx->dXmask.as<u32>(), ctx->dFlags.as<u8>());
    if (words >= (1ull << 32)) return fail(ctx, CDG_ERR_UNSUPPORTED, "cdg_upload_seqs: more than 2^32 packed words per context");
    CK(ctx->dMeta.reserve(n * sizeof(cdgalign::SeqMeta)));
    cdgalign::meta_kernel<<<(unsigned) ((n + 255) / 256), 256, 0, ctx->st>>>(ctx->dOff.as<u64>(), ctx->dWordOff.as<u64>(), ctx->dLen.as<u32>(),
                                                                           ctx->dFlags.as<u8>(), ctx->dExt.as<u8>(), n, ctx->dMeta.as<cdgalign::SeqMeta>());
    ctx->launches += 2;
    CKL();
    CK(cudaStreamSynchronize(ctx->st));
    return CDG_OK;
}

// ------------------------------------------------------------------------------------------------------
static int run_kmermatch(cdg_ctx *ctx, const cdg_kmer_params *p) {
    using namespace cdgkmer;
    const u64 n = ctx->n;
    if (p->k < 8 || p->k > 31) return fail(ctx, CDG_ERR_UNSUPPORTED, "k=%d outside [8,31]", p->k);
    const int idBits = std::max(1, ceil_log2(n));
    const int diagBits = std::max(1, ceil_log2(2ull * ctx->maxLen + 1));
    if (2 * idBits + diagBits + 1 > 64)
        return fail(ctx, CDG_ERR_UNSUPPORTED, "n=%llu sequences x max length %u needs %d key bits (>64): split the set", (unsigned long long) n,
                    ctx->maxLen, 2 * idBits + diagBits + 1);
    unsigned long long *sc = ctx->dScalars.as<unsigned long long>();
    CK(cudaMemsetAsync(sc, 0, SC_N * 8, ctx->st));
    rec(ctx, EV_START);
    // slots, laid out in (sequence length desc, id asc) order: after the stable k-mer sort the first record of every
    // run is the representative of compareRepSequenceAndIdAndPosReverse (kmermatcher.h:76-96)
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
    if ((rc = scan_to(ctx, OrderedCapacityF{ordV, ctx->dLen.as<u32>(), p->k, p->kmers_per_seq, p->kmers_per_seq_scale}, n, ctx->dSlotRank.as<u64>(), false))) return rc;
    order_scatter_kernel<<<(unsigned) ((n + 255) / 256), 256, 0, ctx->st>>>(ordV, ctx->dSlotRank.as<u64>(), n, ctx->dSlotOff.as<u64>(), ctx->dRankOf.as<u32>());
    ctx->launches++;
    u64 slots = 0;
    if ((rc = read_scalar(ctx, SC_SCAN, &slots))) return rc;
    const u32 extraCap = (u32) (1024 + (n >> 10));
    const u64 Rtot = slots + extraCap;
    if (Rtot > cdgsort::MAX_N) return fail(ctx, CDG_ERR_UNSUPPORTED, "%llu k-mer records exceed one sort (2^30): use hash ranges", (unsigned long long) Rtot);
    CK(ctx->dKeysA.reserve(Rtot * 8)); CK(ctx->dKeysB.reserve(Rtot * 8)); CK(ctx->dValsA.reserve(Rtot * 8)); CK(ctx->dValsB.reserve(Rtot * 8));
    CK(ctx->dIdKeysA.reserve(n * 8)); CK(ctx->dIdKeysB.reserve(n * 8)); CK(ctx->dIdValsA.reserve(n * 8)); CK(ctx->dIdValsB.reserve(n * 8));
    CK(cudaMemsetAsync(ctx->dKeysA.as<u64>() + slots, 0xFF, (size_t) extraCap * 8, ctx->st));
    CK(cudaMemsetAsync(ctx->dValsA.as<u64>() + slots, 0, (size_t) extraCap * 8, ctx->st));
    ExtractParams E; E.k = p->k; E.kmersPerSeq = p->kmers_per_seq; E.scale = p->kmers_per_seq_scale; E.seed = p->hash_shift;
    E.hashLo = p->hash_lo; E.hashHi = p->hash_hi;
    // long sequences
    CK(ctx->dLongList.reserve(n * 4)); CK(ctx->dScratchOff.reserve(n * 8));
    long_list_kernel<<<(unsigned) ((n + 255) / 256), 256, 0, ctx->st>>>(ctx->dLen.as<u32>(), n, p->k, p->kmers_per_seq, p->kmers_per_seq_scale,
                                                                        ctx->dLongList.as<u32>(), ctx->dScratchOff.as<u64>(), (u32 *) (sc + SC_NLONG),
                                                                        sc + SC_LONGSCR);
    ctx->launches++;
    u64 nLong64 = 0, longScr = 0;
    if ((rc = read_scalar(ctx, SC_NLONG, &nLong64))) return rc;
    if ((rc = read_scalar(ctx, SC_LONGSCR, &longScr))) return rc;
    const u32 nLong = (u32) nLong64;
    // short sequences
    const int shortBlocks = ctx->smCount * 3;
    const u64 gridWarps = (u64) shortBlocks * EX_WARPS;
    CK(ctx->dShortScratch.reserve(gridWarps * 2 * EX_MAXW * sizeof(WinRec)));
    CK(ctx->dShortKeep.reserve(gridWarps * EX_MAXW));
    {
        const u32 maxW = ctx->maxLen >= (u32) p->k ? ctx->maxLen - p->k + 1 : 0;   // short sequences have <= 256 windows
#define CDG_EXTRACT_ARGS (ctx->dWords.as<u64>(), ctx->dXmask.as<u32>(), ctx->dWordOff.as<u64>(), ctx->dLen.as<u32>(), n, E, ctx->dSlotOff.as<u64>(), ctx->dKeysA.as<u64>(), ctx->dValsA.as<u64>(), ctx->dIdKeysA.as<u64>(), ctx->dIdValsA.as<u64>(), slots, extraCap, (u32 *) (sc + SC_EXTRA), (u32 *) (sc + SC_OVERFLOW), ctx->dRankOf.as<u32>(), ctx->dShortScratch.as<WinRec>(), ctx->dShortKeep.as<u8>(), sc + SC_EMIT)
        if (maxW <= 64) extract_short_kernel<2><<<shortBlocks, EX_WARPS * 32, EX_WARPS * 2048 * 4, ctx->st>>> CDG_EXTRACT_ARGS;
        else if (maxW <= 128) extract_short_kernel<4><<<shortBlocks, EX_WARPS * 32, EX_WARPS * 2048 * 4, ctx->st>>> CDG_EXTRACT_ARGS;
        else extract_short_kernel<8><<<shortBlocks, EX_WARPS * 32, EX_WARPS * 2048 * 4, ctx->st>>> CDG_EXTRACT_ARGS;
#undef CDG_EXTRACT_ARGS
    }
    ctx->launches++;
    CKL();
    if (nLong) {
        CK(ctx->dScratch.reserve((longScr + 1) * sizeof(WinRec))); CK(ctx->dScratchKeep.reserve(longScr + 1));
        unsigned lb = nLong < (u32) ctx->smCount ? nLong : (u32) ctx->smCount;
        extract_long_kernel<<<lb, EXL_BLOCK, (32768 + 128) * 4, ctx->st>>>(
            ctx->dWords.as<u64>(), ctx->dXmask.as<u32>(), ctx->dWordOff.as<u64>(), ctx->dLen.as<u32>(), ctx->dLongList.as<u32>(), nLong, E,
            ctx->dSlotOff.as<u64>(), ctx->dKeysA.as<u64>(), ctx->dValsA.as<u64>(), ctx->dIdKeysA.as<u64>(), ctx->dIdValsA.as<u64>(), slots, extraCap,
            (u32 *) (sc + SC_EXTRA), (u32 *) (sc + SC_OVERFLOW), ctx->dRankOf.as<u32>(), ctx->dScratch.as<WinRec>(), ctx->dScratchKeep.as<u8>(), ctx->dScratchOff.as<u64>(),
            sc + SC_EMIT);
        ctx->launches++;
        CKL();
    }
    rec(ctx, EV_EXTRACT);
    // sort 1: k-mer records by the 2k k-mer bits; identity records by their 63 hash bits
    if ((rc = sort_ws(ctx, Rtot > n ? Rtot : n, ws))) return rc;
    u64 *mk, *mv, *ik, *iv;
    cdgsort::Plan plan1 = cdgsort::make_plan(0, 0, 0, 2 * p->k);
    int l1 = cdgsort::radix_sort(ctx->st, ws, ctx->dKeysA.as<u64>(), ctx->dKeysB.as<u64>(), ctx->dValsA.as<u64>(), ctx->dValsB.as<u64>(), Rtot, plan1, &mk,
                                 &mv, ctx->smCount);
    cdgsort::Plan planI = cdgsort::make_plan(0, 0, 0, 63);
    int l2 = cdgsort::radix_sort(ctx->st, ws, ctx->dIdKeysA.as<u64>(), ctx->dIdKeysB.as<u64>(), ctx->dIdValsA.as<u64>(), ctx->dIdValsB.as<u64>(), n, planI,
                                 &ik, &iv, ctx->smCount);
    if (l1 < 0 || l2 < 0) return fail(ctx, CDG_ERR_UNSUPPORTED, "radix sort: input too large for one pass set");
    ctx->launches += l1 + l2; ctx->sort1Passes = plan1.n;
    CKL();
    rec(ctx, EV_SORT1);
    // overflow / counters
    u64 overflow = 0, emitted = 0, extra = 0;
    if ((rc = read_scalar(ctx, SC_OVERFLOW, &overflow))) return rc;
    if ((u32) overflow) return fail(ctx, CDG_ERR_UNSUPPORTED, "identity-record overflow area exhausted");
    if ((rc = read_scalar(ctx, SC_EMIT, &emitted))) return rc;
    if ((rc = read_scalar(ctx, SC_EXTRA, &extra))) return rc;
    // group assignment (main array, then identity array)
    const u64 key2Cap = Rtot + n;
    CK(ctx->dKey2A.reserve(key2Cap * 8)); CK(ctx->dKey2B.reserve(key2Cap * 8));
    GroupParams G; G.idBits = idBits; G.diagBits = diagBits; G.diagBias = (i32) ctx->maxLen; G.onlyExtendable = p->include_only_extendable;
    G.firstRunQuirk = (emitted + (u32) extra) > 0 ? 1 : 0;
    group_kernel<<<(unsigned) ((Rtot + GROUP_TILE - 1) / GROUP_TILE), GROUP_BLOCK, 0, ctx->st>>>(mk, mv, Rtot, ctx->dLen.as<u32>(), G, ctx->dKey2A.as<u64>(), sc + SC_GROUPED);
    GroupParams GI = G; GI.firstRunQuirk = G.firstRunQuirk ? 0 : 1;
    group_kernel<<<(unsigned) ((n + GROUP_TILE - 1) / GROUP_TILE), GROUP_BLOCK, 0, ctx->st>>>(ik, iv, n, ctx->dLen.as<u32>(), GI, ctx->dKey2A.as<u64>(), sc + SC_GROUPED);
    ctx->launches += 2;
    CKL();
    u64 R2 = 0;
    if ((rc = read_scalar(ctx, SC_GROUPED, &R2))) return rc;
    rec(ctx, EV_GROUP);
    // identity records actually emitted: count of non-sentinel id keys is not needed for parity; R1 = k-mers + extras + identities
    ctx->nRecords = emitted + n;   // every sequence owns one identity record when the hash range is full
    ctx->nGrouped = R2;
    // sort 2
    u64 *k2 = ctx->dKey2A.as<u64>(), *dummy = nullptr;
    cdgsort::Plan plan2 = cdgsort::make_plan(0, 0, 1, 2 * idBits + diagBits + 1);   // bit 0 (needRev) is not an ordering key
    if (R2 > cdgsort::MAX_N) return fail(ctx, CDG_ERR_UNSUPPORTED, "grouped records exceed one sort");
    if ((rc = sort_ws(ctx, std::max(R2, std::max(Rtot, n)), ws))) return rc;
    int l3 = cdgsort::radix_sort(ctx->st, ws, ctx->dKey2A.as<u64>(), ctx->dKey2B.as<u64>(), nullptr, nullptr, R2, plan2, &k2, &dummy, ctx->smCount);
    if (l3 < 0) return fail(ctx, CDG_ERR_UNSUPPORTED, "radix sort 2 failed");
    ctx->launches += l3; ctx->sort2Passes = plan2.n;
    CKL();
    rec(ctx, EV_SORT2);
    // reduce
    u64 H = 0;
    if (R2 > 0) {
        PairEndFlag f; f.keys = k2; f.n = R2; f.idBits = idBits; f.diagBits = diagBits;
        u64 tiles = cdgscan::tiles_for(R2);
        CK(ctx->dTileSums.reserve((tiles + 1) * 8));
        unsigned long long *ts = ctx->dTileSums.as<unsigned long long>();
        cdgscan::tile_sum_kernel<<<(unsigned) tiles, cdgscan::BLOCK, 0, ctx->st>>>(f, R2, ts);
        cdgscan::tile_scan_kernel<<<1, 1024, 0, ctx->st>>>(ts, tiles, sc + SC_SCAN);
        ctx->launches += 2;
        if ((rc = read_scalar(ctx, SC_SCAN, &H))) return rc;
        CK(ctx->dHitRep.reserve((H + 1) * 4)); CK(ctx->dHitTgt.reserve((H + 1) * 4)); CK(ctx->dHitScore.reserve((H + 1) * 4)); CK(ctx->dHitDiag.reserve((H + 1) * 4));
        ReduceParams R; R.idBits = idBits; R.diagBits = diagBits; R.diagBias = (i32) ctx->maxLen;
        reduce_kernel<<<(unsigned) tiles, cdgscan::BLOCK, 0, ctx->st>>>(f, R, ts, ctx->dHitRep.as<u32>(), ctx->dHitTgt.as<u32>(), ctx->dHitScore.as<i32>(),
                                                                      ctx->dHitDiag.as<i32>());
        ctx->launches++;
        CKL();
    } else {
        CK(ctx->dHitRep.reserve(4)); CK(ctx->dHitTgt.reserve(4)); CK(ctx->dHitScore.reserve(4)); CK(ctx->dHitDiag.reserve(4));
    }
    rec(ctx, EV_REDUCE);
    ctx->nHits = H; ctx->hitsValid = true; ctx->linesValid = false; ctx->alnsValid = false;
    return CDG_OK;
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
    CK(ctx->dDE.reserve((M + 1) * 4)); CK(ctx->dLineSel.reserve(M + 1));
    return 0;
}

static int build_lines_from_hits(cdg_ctx *ctx) {
    const u64 n = ctx->n, H = ctx->nHits, M = n + H;
    int rc = reserve_lines(ctx, M);
    if (rc) return rc;
    cdgalign::lines_self_kernel<<<(unsigned) ((n + 1 + 255) / 256), 256, 0, ctx->st>>>(ctx->dHitRep.as<u32>(), H, n, ctx->dLineOff.as<u64>(),
                                                                                       ctx->dLineQ.as<u32>(), ctx->dLineT.as<u32>(),
                                                                                       ctx->dLinePref.as<i32>(), ctx->dLineDiag.as<u16>());
    ctx->launches++;
    if (H) {
        cdgalign::lines_hits_kernel<<<(unsigned) ((H + 255) / 256), 256, 0, ctx->st>>>(ctx->dHitRep.as<u32>(), ctx->dHitTgt.as<u32>(), ctx->dHitScore.as<i32>(),
                                                                                     ctx->dHitDiag.as<i32>(), H, ctx->dLineQ.as<u32>(), ctx->dLineT.as<u32>(),
                                                                                     ctx->dLinePref.as<i32>(), ctx->dLineDiag.as<u16>());
        ctx->launches++;
    }
    CKL();
    ctx->nLines = M; ctx->linesValid = true;
    return 0;
}

extern "C" int cdg_set_pref_lines(cdg_ctx *ctx, const uint64_t *line_off, uint64_t n_queries, const uint32_t *target, const int32_t *pref_score,
                                  const uint16_t *diag, uint64_t n_lines) {
    if (!ctx || !line_off) return CDG_ERR_ARG;
    if (n_queries != ctx->n) return fail(ctx, CDG_ERR_ARG, "cdg_set_pref_lines: %llu queries, resident set has %llu", (unsigned long long) n_queries,
                                         (unsigned long long) ctx->n);
    CK(cudaSetDevice(ctx->device));
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
    ctx->nLines = n_lines; ctx->linesValid = true; ctx->alnsValid = false;
    return CDG_OK;
}

static int run_rescore(cdg_ctx *ctx, const cdg_rescore_params *p) {
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
            ctx->dData.as<u8>(), ctx->dOff.as<u64>(), ctx->dLen.as<u32>(), ctx->dWords.as<u64>(), ctx->dWordOff.as<u64>(), ctx->dFlags.as<u8>(),
            ctx->dLineQ.as<u32>(), ctx->dLineT.as<u32>(), ctx->dLinePref.as<i32>(), ctx->dLineDiag.as<u16>(), M, R, ctx->dKeep.as<u8>(),
            ctx->dScore.as<i32>(), ctx->dBits.as<i32>(), ctx->dEvalue.as<double>(), ctx->dSeqId.as<float>(), ctx->dQS.as<i32>(), ctx->dQE.as<i32>(),
            ctx->dDS.as<i32>(), ctx->dDE.as<i32>(), sc + SC_KEPT);
        ctx->launches++;
        CKL();
    }
    rec(ctx, EV_RS1);
    ctx->alnsValid = true; ctx->alnsFromText = false;
    return CDG_OK;
}

static int export_alns(cdg_ctx *ctx, cdg_alns *out) {
    const u64 M = ctx->nLines, n = ctx->n;
    CK(ctx->hLineOff.reserve((n + 1) * 8)); CK(ctx->hLineT.reserve((M + 1) * 4)); CK(ctx->hKeep.reserve(M + 1)); CK(ctx->hScore.reserve((M + 1) * 4));
    CK(ctx->hBits.reserve((M + 1) * 4)); CK(ctx->hEvalue.reserve((M + 1) * 8)); CK(ctx->hSeqId.reserve((M + 1) * 4)); CK(ctx->hQS.reserve((M + 1) * 4));
    CK(ctx->hQE.reserve((M + 1) * 4)); CK(ctx->hDS.reserve((M + 1) * 4)); CK(ctx->hDE.reserve((M + 1) * 4));
    CK(cudaMemcpyAsync(ctx->hLineOff.p, ctx->dLineOff.p, (n + 1) * 8, cudaMemcpyDeviceToHost, ctx->st));
    if (M) {
        CK(cudaMemcpyAsync(ctx->hLineT.p, ctx->dLineT.p, M * 4, cudaMemcpyDeviceToHost, ctx->st));
        CK(cudaMemcpyAsync(ctx->hKeep.p, ctx->dKeep.p, M, cudaMemcpyDeviceToHost, ctx->st));
        CK(cudaMemcpyAsync(ctx->hScore.p, ctx->dScore.p, M * 4, cudaMemcpyDeviceToHost, ctx->st));
        CK(cudaMemcpyAsync(ctx->hBits.p, ctx->dBits.p, M * 4, cudaMemcpyDeviceToHost, ctx->st));
        CK(cudaMemcpyAsync(ctx->hEvalue.p, ctx->dEvalue.p, M * 8, cudaMemcpyDeviceToHost, ctx->st));
        CK(cudaMemcpyAsync(ctx->hSeqId.p, ctx->dSeqId.p, M * 4, cudaMemcpyDeviceToHost, ctx->st));
        CK(cudaMemcpyAsync(ctx->hQS.p, ctx->dQS.p, M * 4, cudaMemcpyDeviceToHost, ctx->st));
        CK(cudaMemcpyAsync(ctx->hQE.p, ctx->dQE.p, M * 4, cudaMemcpyDeviceToHost, ctx->st));
        CK(cudaMemcpyAsync(ctx->hDS.p, ctx->dDS.p, M * 4, cudaMemcpyDeviceToHost, ctx->st));
        CK(cudaMemcpyAsync(ctx->hDE.p, ctx->dDE.p, M * 4, cudaMemcpyDeviceToHost, ctx->st));
    }
    u64 kept = 0;
    int rc = read_scalar(ctx, SC_KEPT, &kept);
    if (rc) return rc;
    ctx->nKept = kept;
    out->n_lines = M; out->n_queries = n; out->line_off = ctx->hLineOff.as<u64>(); out->target = ctx->hLineT.as<u32>(); out->keep = ctx->hKeep.as<u8>();
    out->score = ctx->hScore.as<i32>(); out->bits = ctx->hBits.as<i32>(); out->evalue = ctx->hEvalue.as<double>(); out->seq_id = ctx->hSeqId.as<float>();
    out->q_start = ctx->hQS.as<i32>(); out->q_end = ctx->hQE.as<i32>(); out->db_start = ctx->hDS.as<i32>(); out->db_end = ctx->hDE.as<i32>();
    return CDG_OK;
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
    }
    CK(cudaStreamSynchronize(ctx->st));
    ctx->nLines = n_alns; ctx->linesValid = false; ctx->hitsValid = false; ctx->alnsValid = true; ctx->alnsFromText = true;
    return CDG_OK;
}

static int run_correct(cdg_ctx *ctx, const cdg_corr_params *p, uint8_t *out_data, uint64_t *n_changed, double *lik) {
    if (!ctx->alnsValid) return fail(ctx, CDG_ERR_ARG, "cdg_correct: no alignments (run cdg_rescore or cdg_set_alignments)");
    if (!ctx->damageSet) return fail(ctx, CDG_ERR_ARG, "cdg_correct: cdg_set_damage not called");
    const u64 n = ctx->n;
    unsigned long long *sc = ctx->dScalars.as<unsigned long long>();
    CK(ctx->dOut.reserve(ctx->dataBytes + 64));
    CK(cudaMemsetAsync(sc + SC_CHANGED, 0, 8, ctx->st));
    rec(ctx, EV_CO0);
    CK(cudaMemcpyAsync(ctx->dOut.p, ctx->dData.p, ctx->dataBytes, cudaMemcpyDeviceToDevice, ctx->st));
    double *dLik = nullptr;
    if (lik) {
        CK(ctx->dLik.reserve((ctx->totalBases + 1) * 4 * 8));
        CK(cudaMemsetAsync(ctx->dLik.p, 0xFF, ctx->totalBases * 4 * 8, ctx->st));   // all-ones = NaN
        dLik = ctx->dLik.as<double>();
    }
    cdgalign::CorrParams C; C.seqIdThr = p->min_seqid_corr_reads; C.ryThr = p->min_ryseq_id_corr_reads; C.roundSeqId = ctx->alnsFromText ? 0 : 1; C.wideThreshold = ctx->wideThreshold;
    const u64 wideCap = ctx->wideThreshold < 65535 ? n + 16 : (ctx->nLines >> 16) + 16;
    CK(ctx->dWideList.reserve(wideCap * 4));
    CK(cudaMemsetAsync(sc + SC_WIDE, 0, 8, ctx->st));
    const int blocks = ctx->smCount * 4;
    cdgalign::correct_fast_kernel<<<blocks, cdgalign::CF_WARPS * 32, sizeof(cdgalign::CorrSmem), ctx->st>>>(
        ctx->dData.as<u8>(), ctx->dMeta.as<cdgalign::SeqMeta>(), ctx->dWords.as<u64>(), n, ctx->dLineOff.as<u64>(), ctx->dLineT.as<u32>(), ctx->dKeep.as<u8>(), ctx->dSeqId.as<float>(), ctx->dQS.as<i32>(),
        ctx->dQE.as<i32>(), ctx->dDS.as<i32>(), ctx->dDE.as<i32>(), C, ctx->dTables.as<cdgalign::DamageTables>(), ctx->dLineSel.as<u8>(),
        ctx->dOut.as<u8>(), sc + SC_CHANGED, ctx->dLikOff.as<u64>(), dLik, ctx->dWideList.as<u32>(), (u32 *) (sc + SC_WIDE));
    ctx->launches++;
    // queries with more than 65535 counted targets (16-bit counters): generic kernel, usually an empty list
    cdgalign::correct_wide_kernel<<<ctx->smCount, cdgalign::CORR_WARPS * 32, 0, ctx->st>>>(
        ctx->dWideList.as<u32>(), (const u32 *) (sc + SC_WIDE),
        ctx->dData.as<u8>(), ctx->dOff.as<u64>(), ctx->dLen.as<u32>(), ctx->dExt.as<u8>(), n, ctx->dLineOff.as<u64>(), ctx->dLineT.as<u32>(),
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
    return CDG_OK;
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
    if (alns && (rc = export_alns(ctx, alns))) return rc;
    if (!alns) { u64 kept = 0; if ((rc = read_scalar(ctx, SC_KEPT, &kept))) return rc; ctx->nKept = kept; }
    return run_correct(ctx, &p->corr, out_data, n_changed, nullptr);
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
    out->ms_sort2 = el(EV_GROUP, EV_SORT2); out->ms_reduce = el(EV_SORT2, EV_REDUCE); out->ms_rescore = el(EV_RS0, EV_RS1); out->ms_correct = el(EV_CO0, EV_CO1);
    return CDG_OK;
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
    if (i < n) { h[i] = cdg_xxh64_u64(in[i], seed); rc[i] = cdg_revcomp_kmer(in[i], k); }
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
        cudaEventRecord(e0, ctx->st);
        u64 *ok, *ov;
        int l = cdgsort::radix_sort(ctx->st, ws, ka.as<u64>(), kb.as<u64>(), has_val ? va.as<u64>() : nullptr, has_val ? vb.as<u64>() : nullptr, n, plan,
                                    &ok, &ov, ctx->smCount, variant);
        if (l < 0) return fail(ctx, CDG_ERR_UNSUPPORTED, "cdg_test_sort_bench: sort refused");
        cudaEventRecord(e1, ctx->st);
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
