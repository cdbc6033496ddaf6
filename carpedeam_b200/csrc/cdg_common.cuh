// Shared device/host helpers for libcarpedeam_gpu (sm_100a).  Scalar arithmetic that must be
// bit-identical to the reference lives here as __host__ __device__ functions.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

typedef uint64_t u64;
typedef uint32_t u32;
typedef uint16_t u16;
typedef uint8_t u8;
typedef int32_t i32;
typedef int64_t i64;

#define CDG_HD __host__ __device__ __forceinline__
#define CDG_D __device__ __forceinline__

static constexpr u64 CDG_B63 = 1ull << 63;
static constexpr u64 CDG_SENTINEL = ~0ull;   // unused record slot, sorts last (kmermatcher.cpp:40 memset 0xFF)

// ---- a1: letter mapping (NucleotideMatrix.cpp:17-61; alphabet order nucleotide.out:4) --------
// codes A0 C1 T2 G3 X4.  U,W->T  K,B,D,V,R,S->G  M,Y,H->C, anything else X; case-insensitive.
CDG_HD int cdg_map_letter(unsigned char c) {
    unsigned char u = (c >= 'a' && c <= 'z') ? (unsigned char) (c - 32) : c;
    switch (u) {
        case 'A': return 0;
        case 'C': case 'M': case 'Y': case 'H': return 1;
        case 'T': case 'U': case 'W': return 2;
        case 'G': case 'K': case 'B': case 'D': case 'V': case 'R': case 'S': return 3;
        default: return 4;
    }
}
// exact upper-case A/C/G/T -> code, everything else (lower case, IUPAC, N, ...) -> 4 ("exception"):
// only sequences without exceptions take the packed fast paths of rescore/correction.
CDG_HD int cdg_plain_code(unsigned char c) { return c == 'A' ? 0 : c == 'C' ? 1 : c == 'T' ? 2 : c == 'G' ? 3 : 4; }
CDG_HD char cdg_num2aa(int code) { return code == 0 ? 'A' : code == 1 ? 'C' : code == 2 ? 'T' : code == 3 ? 'G' : 'X'; }
CDG_HD int cdg_rev_code(int code) { return code == 4 ? 4 : (code ^ 2); }   // NucleotideMatrix::reverseResidue

// ---- a3: XXH64 of one little-endian u64 (xxhash v0.8.0; hashUInt64, kmermatcher.cpp:33-38) ----
CDG_HD u64 cdg_rotl64(u64 x, int r) { return (x << r) | (x >> (64 - r)); }
CDG_HD u64 cdg_xxh64_u64(u64 v, u64 seed) {
    const u64 P1 = 11400714785074694791ull, P2 = 14029467366897019727ull, P3 = 1609587929392839161ull,
              P4 = 9650029242287828579ull, P5 = 2870177450012600261ull;
    u64 h = seed + P5 + 8;
    u64 k = cdg_rotl64(v * P2, 31) * P1;
    h ^= k;
    h = cdg_rotl64(h, 27) * P1 + P4;
    h ^= h >> 33; h *= P2; h ^= h >> 29; h *= P3; h ^= h >> 32;
    return h;
}

// ---- a2: reverse complement of a 2-bit k-mer index (Util::revComplement, Util.cpp:601-638) ----
// complement = xor 0b10 per base (A0<->T2, C1<->G3); reverse the k 2-bit groups.
CDG_HD u64 cdg_rev2bit64(u64 x) {   // reverse all 32 2-bit groups of a word
    x = ((x >> 2) & 0x3333333333333333ull) | ((x & 0x3333333333333333ull) << 2);
    x = ((x >> 4) & 0x0F0F0F0F0F0F0F0Full) | ((x & 0x0F0F0F0F0F0F0F0Full) << 4);
    x = ((x >> 8) & 0x00FF00FF00FF00FFull) | ((x & 0x00FF00FF00FF00FFull) << 8);
    x = ((x >> 16) & 0x0000FFFF0000FFFFull) | ((x & 0x0000FFFF0000FFFFull) << 16);
    return (x >> 32) | (x << 32);
}
CDG_HD u64 cdg_revcomp_kmer(u64 idx, int k) {
    return cdg_rev2bit64(idx ^ 0xAAAAAAAAAAAAAAAAull) >> (64 - 2 * k);
}

#include "cdg_alp.h"   // E-values: Karlin-Altschul + ALP finite size correction (shared with the host code)

// ---- record value packing -----------------------------------------------------------------------
// k-mer record = {key u64: canonical k-mer | bit63 (1 = forward, 0 = reverse complement picked),
//                 val u64: id:32 | pos:32}  (KmerPosition, kmermatcher.h:49-54; seqLen is looked up by id)
// (multi-GPU: bits 28-31 of the value carry the owner rank of the k-mer's hash range until the records are routed)
static constexpr int CDG_OWNER_SHIFT = 28;
CDG_HD u64 cdg_pack_val(u32 id, u32 pos, u32 owner = 0) { return ((u64) id << 32) | ((u64) owner << CDG_OWNER_SHIFT) | pos; }
CDG_HD u32 cdg_val_id(u64 v) { return (u32) (v >> 32); }
CDG_HD u32 cdg_val_pos(u64 v) { return (u32) v & ((1u << CDG_OWNER_SHIFT) - 1); }

// per-sequence metadata in one 16-byte load: byte offset in the DB data, first packed word, length and flags
struct SeqMeta { u64 off; u32 wordOff; u32 lenFlags; };   // lenFlags: len | exception << 30 | ext << 31
static constexpr u32 META_EXC = 1u << 30, META_EXT = 1u << 31, META_LEN = (1u << 30) - 1;

// ---- TMA (bulk async copy engine): contiguous global -> shared tile loads ------------------------------------------
// One thread arms an mbarrier with the byte count and issues cp.async.bulk; the copy engine moves the tile while no LSU
// instruction is spent on it, every thread then waits on the barrier phase.  Source, destination and size must be
// multiples of 16 bytes (callers fall back to a plain loop otherwise).
CDG_D u32 cdg_smem_u32(const void *p) { return (u32) __cvta_generic_to_shared(p); }
CDG_D void cdg_mbar_init(u64 *bar, u32 arrivals) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(cdg_smem_u32(bar)), "r"(arrivals) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
CDG_D void cdg_mbar_expect_tx(u64 *bar, u32 bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(cdg_smem_u32(bar)), "r"(bytes) : "memory");
}
CDG_D void cdg_tma_load_1d(void *dstSmem, const void *srcGlobal, u32 bytes, u64 *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"(cdg_smem_u32(dstSmem)), "l"(srcGlobal), "r"(bytes), "r"(cdg_smem_u32(bar)) : "memory");
}
CDG_D void cdg_mbar_wait(u64 *bar, u32 parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "CDG_MBAR_WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra CDG_MBAR_DONE_%=;\n"
        "bra CDG_MBAR_WAIT_%=;\n"
        "CDG_MBAR_DONE_%=:\n"
        "}\n" :: "r"(cdg_smem_u32(bar)), "r"(parity) : "memory");
}
CDG_D bool cdg_tma_ok(const void *src, u32 bytes) { return (((unsigned long long) src) & 15ull) == 0 && (bytes & 15u) == 0 && bytes > 0; }

