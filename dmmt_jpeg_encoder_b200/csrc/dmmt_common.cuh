// dmmt_common.cuh -- shared device/host definitions of the sm_100a encode path.
// Geometry, per-image metadata, decoupled look-back helpers, block scans.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/dmmt_cuda.h"

namespace dmmt {

// ------------------------------------------------------------------------------------------
// Geometry of one image (or one MCU-row shard of an image) -- SURVEY.md Appendix A steps 2,5,8
struct Geom {
    int W, H;          // original pixel size of this image / shard rows
    int hr, vr;        // chroma rates (subsampling.rs:32-46)
    int mcus_x, mcus_y;
    int ypm;           // Y blocks per MCU = hr*vr
    int bpm;           // blocks per MCU = ypm + 2   (block_fold_iterator.rs:75-148)
    uint32_t n_mcus;
    uint32_t n_blocks; // stream blocks = n_mcus * bpm
};

// Per-image device metadata written by K2b/K4 (mirrors dmmt_image_meta + internals)
struct ImgMeta {
    unsigned long long scan_bits;   // accumulated by the 4 table CTAs of K2b
    unsigned long long out_len;     // whole file length, written by K4
    uint32_t header_len;
    uint32_t n_symbols[4];
    int32_t error;                  // first device-side error (DMMT_E_*)
    uint32_t n_stream_blocks;
    uint32_t reserved;
};
static_assert(sizeof(ImgMeta) == sizeof(dmmt_image_meta), "meta layout is part of the C ABI");

// Table indices
enum { T_YDC = 0, T_YAC = 1, T_CDC = 2, T_CAC = 3 };

// Symbol field of a TOKEN (and index of K3's shared-memory code LUT).  DC symbols are stored as they are.  An AC
// symbol (run << 4 | cat) is stored with its low nibble XOR-ed with the run nibble (and with 8 for the chroma
// table): the LUT entries are 8 bytes, so the bank pair of an entry is its index mod 16, and the frequent symbols
// of one category (0x01, 0x11, 0x21 ... and their chroma twins) would all sit in the SAME bank pair -- 44 % of K3's
// shared-memory wavefronts were bank-conflict replays.  The map is an involution per table; histograms, code
// tables and everything outside the token stream use the plain symbol.
__host__ __device__ constexpr uint32_t tok_swz(uint32_t table, uint32_t sym) {
    return (table & 1u) ? sym ^ (((sym >> 4) ^ ((table & 2u) << 2)) & 15u) : sym;
}

constexpr bool tok_swz_is_an_involution() {
    for (uint32_t t = 0; t < 4; t++)
        for (uint32_t s = 0; s < 256; s++)
            if (tok_swz(t, tok_swz(t, s)) != s || (tok_swz(t, s) >> 4) != (s >> 4) || tok_swz(t, s) > 255u) return false;
    return tok_swz(T_YDC, 0x0B) == 0x0B && tok_swz(T_YAC, 0xF0) == 0xFF && tok_swz(T_CAC, 0xF0) == 0xF7 &&
           tok_swz(T_CAC, 0x00) == 0x08 && tok_swz(T_YAC, 0x21) == 0x23;
}
static_assert(tok_swz_is_an_involution(), "tok_swz must be a bijection per table that keeps the run nibble");

// Encoder LUT entry: (len << 16) | right-aligned code.  len == 0 => symbol absent.
struct EncTables {
    uint32_t e[4][256];
};
// Length tables in the reference's Vec<SymbolCodeLength> order (ascending frequency), for DHT + tests
struct LenTables {
    uint8_t sym[4][256];
    uint8_t len[4][256];
};

// quantisation divisors as f32, natural (row-major) order: [0] luma, [1] chroma
struct QuantF {
    float q[2][64];
};

// zig-zag (frequency_block.rs:1-5): ZZ[i] = natural index of the i-th zig-zag coefficient
__host__ __device__ constexpr int zz_at(int i) {
    constexpr int t[64] = {0,  1,  8,  16, 9,  2,  3,  10, 17, 24, 32, 25, 18, 11, 4,  5,
                           12, 19, 26, 33, 40, 48, 41, 34, 27, 20, 13, 6,  7,  14, 21, 28,
                           35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23, 30, 37, 44, 51,
                           58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63};
    return t[i];
}

// ------------------------------------------------------------------------------------------
// Decoupled look-back state: one u64 per chunk = (flag << 62) | value.
// flag 0 = not ready, 1 = chunk aggregate available, 2 = inclusive prefix available.
constexpr unsigned long long LB_AGG = 1ull << 62;
constexpr unsigned long long LB_INC = 2ull << 62;
constexpr unsigned long long LB_VAL = (1ull << 62) - 1;

__device__ __forceinline__ unsigned long long ld_relaxed_u64(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_relaxed_u64(unsigned long long* p, unsigned long long v) {
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

// Called by ONE thread (thread 0) of the CTA that owns `chunk`.  Publishes the chunk aggregate,
// walks predecessors until an inclusive prefix is found, publishes the inclusive prefix and
// returns the EXCLUSIVE prefix.  Value and flag travel in one 64-bit word, so relaxed accesses
// are sufficient (no payload outside the word).
__device__ __forceinline__ unsigned long long lookback_exclusive(unsigned long long* state,
                                                                 int chunk,
                                                                 unsigned long long aggregate) {
    if (chunk == 0) {
        st_relaxed_u64(&state[0], LB_INC | aggregate);
        return 0ull;
    }
    st_relaxed_u64(&state[chunk], LB_AGG | aggregate);
    unsigned long long excl = 0ull;
    int j = chunk - 1;
    while (true) {
        unsigned long long s = ld_relaxed_u64(&state[j]);
        unsigned long long flag = s & ~LB_VAL;
        if (flag == 0ull) {
            __nanosleep(20);
            continue;
        }
        excl += (s & LB_VAL);
        if (flag == LB_INC) break;
        --j;  // flag == LB_AGG: keep walking (j >= 0 is guaranteed: chunk 0 always publishes INC)
    }
    st_relaxed_u64(&state[chunk], LB_INC | (excl + aggregate));
    return excl;
}

// The same in two halves, for a warp: lookback_publish_aggregate (ONE lane) makes the chunk's aggregate visible
// without waiting for anybody; lookback_resolve_warp (ALL 32 lanes of one warp, any time later) walks the
// predecessors 32 at a time -- one memory round trip per 32 chunks instead of one per chunk --, publishes the
// inclusive prefix and returns the exclusive prefix in every lane.  A CTA that publishes chunk i, goes on with
// other work and resolves chunk i later never waits for predecessors that started before it.
__device__ __forceinline__ void lookback_publish_aggregate(unsigned long long* state, int chunk,
                                                           unsigned long long aggregate) {
    st_relaxed_u64(&state[chunk], (chunk == 0 ? LB_INC : LB_AGG) | aggregate);
}
#ifndef LB_WARP_SLEEP_NS
#define LB_WARP_SLEEP_NS 20
#endif
__device__ __forceinline__ unsigned long long lookback_resolve_warp(unsigned long long* state, int chunk,
                                                                    unsigned long long aggregate) {
    const int lane = threadIdx.x & 31;
    if (chunk == 0) return 0ull;  // published as inclusive already
    unsigned long long excl = 0ull;
    int base = chunk - 1;
    while (true) {
        const int j = base - lane;
        const unsigned long long s = j >= 0 ? ld_relaxed_u64(&state[j]) : LB_INC;  // before chunk 0: inclusive 0
        const unsigned long long flag = s & ~LB_VAL;
        const unsigned int inc_mask = __ballot_sync(0xffffffffu, flag == LB_INC);
        const unsigned int wait_mask = __ballot_sync(0xffffffffu, flag == 0ull);
        const int f = inc_mask ? __ffs(inc_mask) - 1 : 32;                // nearest inclusive prefix in this window
        const unsigned int needed = f >= 31 ? 0xffffffffu : ((2u << f) - 1u);  // lanes 0..f (all when there is none)
        if (wait_mask & needed) {
            __nanosleep(LB_WARP_SLEEP_NS);
            continue;
        }
        unsigned long long v = lane <= f ? (s & LB_VAL) : 0ull;
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
        excl += v;
        if (inc_mask) break;
        base -= 32;
    }
    if (lane == 0) st_relaxed_u64(&state[chunk], LB_INC | (excl + aggregate));
    return excl;
}

// ------------------------------------------------------------------------------------------
// Programmatic dependent launch (PDL).  The kernels of the launch chain are stream-ordered, but each of them may be
// SCHEDULED while its predecessor still drains: pdl_wait() at its top blocks (in hardware, not by spinning) until the
// predecessor grid has completed and its writes are visible; pdl_trigger() says "every CTA of this grid that has reached
// this point no longer minds the next kernel's CTAs becoming resident".  Both are no-ops for a normal launch.
// Every kernel launched through launch_pdl() calls pdl_wait() in ALL its CTAs before anything else (ordering is then
// transitive along the chain).
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
// Measured on the B200 (tools/ab_pdl.sh): with the early trigger the waiting CTAs of the next kernel cost more than the
// overlap wins -- 1024 frames 7.23 -> 7.35 ms, a lone 4K frame 82.5 -> 93.5 us -- so the trigger is compiled out and a
// dependent kernel is only pre-staged (its launch latency hides behind its predecessor: 82.5 -> 81.3 us, batch unchanged).
#ifndef DMMT_PDL_TRIGGER
#define DMMT_PDL_TRIGGER 0
#endif
__device__ __forceinline__ void pdl_trigger() {
#if DMMT_PDL_TRIGGER
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
#endif
}

// ------------------------------------------------------------------------------------------
// Block-wide exclusive scan of one u32 per thread (blockDim.x == NT, multiple of 32, <= 1024).
// Returns the exclusive prefix; *total receives the block sum.  `warp_sums` = NT/32 + 1 u32 of smem.
template <int NT>
__device__ __forceinline__ uint32_t block_exclusive_scan(uint32_t v, uint32_t* warp_sums,
                                                         uint32_t* total) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    uint32_t inc = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        uint32_t t = __shfl_up_sync(0xffffffffu, inc, d);
        if (lane >= d) inc += t;
    }
    if (lane == 31) warp_sums[wid] = inc;
    __syncthreads();
    if (wid == 0) {
        uint32_t w = (lane < NT / 32) ? warp_sums[lane] : 0u;
        uint32_t winc = w;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            uint32_t t = __shfl_up_sync(0xffffffffu, winc, d);
            if (lane >= d) winc += t;
        }
        if (lane < NT / 32) warp_sums[lane] = winc - w;  // exclusive warp offsets
        if (lane == NT / 32 - 1) warp_sums[NT / 32] = winc;
    }
    __syncthreads();
    uint32_t excl = warp_sums[wid] + inc - v;
    *total = warp_sums[NT / 32];
    __syncthreads();  // warp_sums may be reused by the caller
    return excl;
}

__device__ __forceinline__ uint32_t bswap32(uint32_t x) { return __byte_perm(x, 0, 0x0123); }

}  // namespace dmmt
