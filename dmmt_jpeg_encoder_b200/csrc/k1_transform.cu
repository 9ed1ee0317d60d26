// k1_transform.cu -- K1: fused normalise -> black pad -> RGB->YCbCr -> chroma average ->
// Arai 8x8 DCT -> quantise -> zig-zag, one launch per batch of equally sized images.
//
// Replaces (file:line under /root/reference/src):
//   color.rs:45-53 (v/max), image/writer/jpeg/padder.rs:12-42 (black pad to MCU),
//   color.rs:75-100 (YCbCr, level shift inside Y), image/subsampling.rs:102-122,231-236,286-309
//   (window average + 8x8 re-tiling), cosine_transform/arai.rs:29-104 (rows then columns),
//   image/writer/jpeg/transformer/quantizer.rs:53-62, frequency_block.rs:1-5 (zig-zag),
//   block_entangler.rs:69-77 + encoder/block_fold_iterator.rs:75-148 (MCU stream order).
//
// Numerics: every f32 operation is an explicit round-to-nearest intrinsic (__fmul_rn/__fadd_rn/
// __fsub_rn/__fdiv_rn) in the reference's operation order, so nothing contracts to FMA and the
// coefficients are bit-identical to the CPU path (the file is also compiled with -fmad=false).
//
// Two kernel families.  k1_transform_p420<FMT, FUSED, VEC> (4:2:0 with the proven fast divisions; the path
// bench.py measures) is described at its definition below: 96-thread CTAs, 8-pixel strips, packed FP32, fused
// tokeniser.  The generic kernel k1_transform<HR, VR, FMT, DBG, EXACT> (4:4:4, 4:2:2, debug fetches, exact
// divisions) maps like this:
// Mapping (B200): one CTA = 128 threads = one tile of 256 x (8*VR) pixels of one MCU row.
//   phase A: thread = one 16 x VR pixel strip.  128-bit loads of the interleaved samples (scalar,
//            bounds-checked loads at ragged edges / unaligned pitches), colour conversion and the
//            chroma window sum in registers, planes staged in shared memory as
//            [row][16-byte chunk][strip] so both the phase-A stores and the phase-B loads are
//            conflict-free 128-bit accesses.
//   phase B: thread = one 8x8 block, all 64 samples in registers: 8 row passes + 8 column passes of
//            the Arai flow graph without any transpose, IEEE division by the quantiser, round half
//            away from zero, zig-zag by register renaming, 8 x 16-byte stores of the block.
// Output: int16 [n_mcus][blocks_per_mcu][64] in MCU-interleaved STREAM order, zig-zag inside a block.
#include <cstdlib>

#include "dmmt_kernels.h"

namespace dmmt {

namespace {

constexpr int K1_THREADS = 128;
constexpr int TILE_W = 256;

// arai.rs:7-26 -- the decimal literals of the reference (S0 != S4 on purpose)
constexpr float kA1 = 0.70710678118654752440f;  // FRAC_1_SQRT_2
constexpr float kA2 = 0.5411961f;
constexpr float kA3 = kA1;
constexpr float kA4 = 1.3065629f;
constexpr float kA5 = 0.3826834f;
constexpr float kS0 = 0.3535533f;
constexpr float kS1 = 0.2548978f;
constexpr float kS2 = 0.27059805f;
constexpr float kS3 = 0.30067244f;
constexpr float kS4 = 0.35355338f;
constexpr float kS5 = 0.4499881f;
constexpr float kS6 = 0.6532815f;
constexpr float kS7 = 1.2814577f;

// arai.rs:29-92, one 1-D pass on 8 registers
__device__ __forceinline__ void fast_arai(float& x0, float& x1, float& x2, float& x3, float& x4,
                                          float& x5, float& x6, float& x7) {
    const float v10 = __fadd_rn(x0, x7), v11 = __fadd_rn(x1, x6), v12 = __fadd_rn(x2, x5),
                v13 = __fadd_rn(x3, x4);
    const float v14 = __fsub_rn(x3, x4), v15 = __fsub_rn(x2, x5), v16 = __fsub_rn(x1, x6),
                v17 = __fsub_rn(x0, x7);

    const float v20 = __fadd_rn(v10, v13), v21 = __fadd_rn(v11, v12), v22 = __fsub_rn(v11, v12),
                v23 = __fsub_rn(v10, v13);
    const float v24 = __fsub_rn(-v14, v15), v25 = __fadd_rn(v15, v16), v26 = __fadd_rn(v16, v17);

    const float v30 = __fadd_rn(v20, v21), v31 = __fsub_rn(v20, v21), v32 = __fadd_rn(v22, v23);

    const float v42 = __fmul_rn(v32, kA1);
    const float t5 = __fmul_rn(__fadd_rn(v24, v26), kA5);  // (v24+v26)*A5 == (v26+v24)*A5
    const float v44 = __fsub_rn(__fmul_rn(-v24, kA2), t5);
    const float v45 = __fmul_rn(v25, kA3);
    const float v46 = __fsub_rn(__fmul_rn(v26, kA4), t5);

    const float v52 = __fadd_rn(v42, v23), v53 = __fsub_rn(v23, v42), v55 = __fadd_rn(v45, v17),
                v57 = __fsub_rn(v17, v45);

    const float v64 = __fadd_rn(v44, v57), v65 = __fadd_rn(v55, v46), v66 = __fsub_rn(v55, v46),
                v67 = __fsub_rn(v57, v44);

    x0 = __fmul_rn(v30, kS0);
    x4 = __fmul_rn(v31, kS4);
    x2 = __fmul_rn(v52, kS2);
    x6 = __fmul_rn(v53, kS6);
    x5 = __fmul_rn(v64, kS5);
    x1 = __fmul_rn(v65, kS1);
    x7 = __fmul_rn(v66, kS7);
    x3 = __fmul_rn(v67, kS3);
}

// color.rs:75-100
__device__ __forceinline__ void rgb_to_ycbcr(float r, float g, float b, float& y, float& cb,
                                             float& cr) {
    constexpr float kShift = 128.0f / 255.0f;  // folded in f32 like `128_f32 / 255_f32`
    y = __fmul_rn(__fsub_rn(__fadd_rn(__fadd_rn(__fmul_rn(r, 0.299f), __fmul_rn(g, 0.587f)),
                                      __fmul_rn(b, 0.114f)),
                            kShift),
                  255.0f);
    cb = __fmul_rn(__fadd_rn(__fadd_rn(__fmul_rn(r, -0.1687f), __fmul_rn(g, -0.3312f)),
                             __fmul_rn(b, 0.5f)),
                   255.0f);
    cr = __fmul_rn(__fadd_rn(__fadd_rn(__fmul_rn(r, 0.5f), __fmul_rn(g, -0.4186f)),
                             __fmul_rn(b, -0.0813f)),
                   255.0f);
}

// quantizer.rs:60 : (d / q as f32).round() as i16  (IEEE divide, half away from zero, saturate).
// Fast exact form: d / q == fma(d, rq_hi, d * rq_lo) with rq_hi = fl(1/q), rq_lo = fl(1/q - rq_hi):
// tools/exhaustive_div.cu proves on all 2^32 bit patterns of d and every q in 1..255 that the
// rounded+saturated i16 is identical (the only exceptions are d = +-inf, which need EXACT).
template <bool EXACT>
__device__ __forceinline__ uint32_t quantize(float d, float q, float rq_hi, float rq_lo) {
    float x;
    if constexpr (EXACT) x = __fdiv_rn(d, q);
    else x = __fmaf_rn(d, rq_hi, __fmul_rn(d, rq_lo));
    // round half away from zero: trunc(x + copysign(pred(0.5), x)); cvt.rzi.s16 saturates, NaN -> 0
    const float h = __int_as_float((__float_as_int(x) & 0x80000000) | 0x3EFFFFFF);
    unsigned short r;
    asm("cvt.rzi.s16.f32 %0, %1;" : "=h"(r) : "f"(__fadd_rn(x, h)));
    return (uint32_t)r;
}

template <int FMT>
struct Px;
template <>
struct Px<DMMT_RGB_U8> {
    static constexpr int kBytes = 3, kWords = 12;   // words per 16-pixel strip row
};
template <>
struct Px<DMMT_RGB_U16> {
    static constexpr int kBytes = 6, kWords = 24;
};
template <>
struct Px<DMMT_RGB_F32_NORM> {
    static constexpr int kBytes = 12, kWords = 48;
};

// Raw samples of one 16-pixel strip row.  Three cases per thread: fully inside + 16-byte aligned
// (128-bit loads), fully outside the image (black: padder.rs:18,27-38), ragged (bounds-checked
// scalar loads).  Samples outside the image are 0, which normalises to exactly 0.0.
template <int FMT>
__device__ __forceinline__ void load_strip_row(const uint8_t* __restrict__ row_base, int x0, int W,
                                               bool row_valid, bool vec_ok,
                                               uint32_t (&w)[Px<FMT>::kWords]) {
    constexpr int PB = Px<FMT>::kBytes, NW = Px<FMT>::kWords;
    if (row_valid && vec_ok && x0 + 16 <= W) {
        const uint4* p = reinterpret_cast<const uint4*>(row_base + (size_t)x0 * PB);
#pragma unroll
        for (int i = 0; i < NW / 4; i++) {
            const uint4 v = __ldg(p + i);
            w[4 * i] = v.x, w[4 * i + 1] = v.y, w[4 * i + 2] = v.z, w[4 * i + 3] = v.w;
        }
        return;
    }
#pragma unroll
    for (int i = 0; i < NW; i++) w[i] = 0u;
    if (!row_valid || x0 >= W) return;
    // ragged edge / unaligned pitch
    const int nval = min(16, W - x0) * 3;  // samples available
    if constexpr (FMT == DMMT_RGB_U8) {
        const uint8_t* q = row_base + (size_t)x0 * 3;
#pragma unroll
        for (int i = 0; i < 48; i++)
            if (i < nval) w[i >> 2] |= (uint32_t)q[i] << (8 * (i & 3));
    } else if constexpr (FMT == DMMT_RGB_U16) {
        const uint16_t* q = reinterpret_cast<const uint16_t*>(row_base) + (size_t)x0 * 3;
#pragma unroll
        for (int i = 0; i < 48; i++)
            if (i < nval) w[i >> 1] |= (uint32_t)q[i] << (16 * (i & 1));
    } else {
        const uint32_t* q = reinterpret_cast<const uint32_t*>(row_base) + (size_t)x0 * 3;
#pragma unroll
        for (int i = 0; i < 48; i++)
            if (i < nval) w[i] = q[i];
    }
}

// Raw samples of one row of an 8-pixel strip (the P420 kernels): 24 / 48 / 96 bytes, same three
// cases as load_strip_row.  The strip starts at a multiple of 8 pixels, so it is 8-byte (u8) or
// 16-byte (u16, f32) aligned whenever the row is.
// `strip`: first sample of the strip in this row; `navail`: pixels of the row from there on (<= 0: outside);
// `fast`: the strip is complete and aligned for vector loads (hoisted by the caller).
// VEC selects the kernel variant: true = rows 16-byte aligned (vector loads), false = any alignment (word loads +
// funnel shift).  Two kernels rather than two paths in one: the aligned kernel is sensitive to its instruction
// footprint (+180 instructions of word path cost it 3 %).
template <int FMT, bool VEC>
__device__ __forceinline__ void load_strip8_row(const uint8_t* __restrict__ strip, int navail, bool row_valid,
                                                bool fast, bool words_ok, uint32_t (&w)[Px<FMT>::kWords / 2]) {
    constexpr int NW = Px<FMT>::kWords / 2;
    if (!VEC && row_valid && words_ok) {
        // complete strip at any byte alignment (row pitch not a multiple of 16 bytes: 500, 1366, 1000 ... pixel wide
        // images): aligned 32-bit loads and one funnel shift per word.  The caller guarantees that the up to three
        // bytes before and after the strip belong to the image buffer (`words_ok`).
        const uintptr_t addr = reinterpret_cast<uintptr_t>(strip);
        const uint32_t* p = reinterpret_cast<const uint32_t*>(addr & ~static_cast<uintptr_t>(3));
        const uint32_t sh = ((uint32_t)addr & 3u) * 8u;
        uint32_t d[NW + 1];
#pragma unroll
        for (int i = 0; i < NW; i++) d[i] = __ldg(p + i);
        d[NW] = sh ? __ldg(p + NW) : 0u;
#pragma unroll
        for (int i = 0; i < NW; i++) w[i] = __funnelshift_r(d[i], d[i + 1], sh);
        return;
    }
    if (VEC && row_valid && fast) {
        if constexpr (FMT == DMMT_RGB_U8) {
            const uint2* p = reinterpret_cast<const uint2*>(strip);
#pragma unroll
            for (int i = 0; i < 3; i++) {
                const uint2 v = __ldg(p + i);
                w[2 * i] = v.x, w[2 * i + 1] = v.y;
            }
        } else {
            const uint4* p = reinterpret_cast<const uint4*>(strip);
#pragma unroll
            for (int i = 0; i < NW / 4; i++) {
                const uint4 v = __ldg(p + i);
                w[4 * i] = v.x, w[4 * i + 1] = v.y, w[4 * i + 2] = v.z, w[4 * i + 3] = v.w;
            }
        }
        return;
    }
#pragma unroll
    for (int i = 0; i < NW; i++) w[i] = 0u;
    if (!row_valid || navail <= 0) return;
    // ragged edge / unaligned pitch: rare, so a rolled loop over a local array keeps the cold code
    // small (the hot path stays contiguous in the instruction cache)
    const int nval = min(8, navail) * 3;  // samples available
    uint32_t tmp[NW];
#pragma unroll 1
    for (int i = 0; i < NW; i++) tmp[i] = 0u;
    if constexpr (FMT == DMMT_RGB_U8) {
        const uint8_t* q = strip;
#pragma unroll 1
        for (int i = 0; i < nval; i++) tmp[i >> 2] |= (uint32_t)q[i] << (8 * (i & 3));
    } else if constexpr (FMT == DMMT_RGB_U16) {
        const uint16_t* q = reinterpret_cast<const uint16_t*>(strip);
#pragma unroll 1
        for (int i = 0; i < nval; i++) tmp[i >> 1] |= (uint32_t)q[i] << (16 * (i & 1));
    } else {
        const uint32_t* q = reinterpret_cast<const uint32_t*>(strip);
#pragma unroll 1
        for (int i = 0; i < nval; i++) tmp[i] = q[i];
    }
#pragma unroll
    for (int i = 0; i < NW; i++) w[i] = tmp[i];
}

// sample i (0..47) of the strip row as the reference's normalised f32 (color.rs:45-53).
// v / max == fma(v, r_hi, v * r_lo) for every v <= max <= 65535 (exhaustively verified on the host
// at plan creation, dmmt_api.cu; tools/exhaustive_div.cu checks all max); `exact` selects IEEE division.
template <int FMT, bool EXACT>
__device__ __forceinline__ float sample_norm(const uint32_t (&w)[Px<FMT>::kWords], int i, float maxf,
                                             float r_hi, float r_lo) {
    if constexpr (FMT == DMMT_RGB_F32_NORM) {
        return __uint_as_float(w[i]);
    } else {
        float v;
        if constexpr (FMT == DMMT_RGB_U8) v = (float)((w[i >> 2] >> (8 * (i & 3))) & 0xFFu);
        else v = (float)((w[i >> 1] >> (16 * (i & 1))) & 0xFFFFu);
        if constexpr (EXACT) return __fdiv_rn(v, maxf);
        else return __fmaf_rn(v, r_hi, __fmul_rn(v, r_lo));
    }
}

struct K1Args {
    const uint8_t* pixels;       // n images, tightly packed
    size_t img_stride_bytes;     // bytes between images
    int W, H;                    // original size (pad region is synthesised)
    int mcus_x;                  // MCUs per row
    float maxf;                  // max_value as f32
    float r_hi, r_lo;            // 1/max split in two f32 (see sample_norm)
    int vec_ok;                  // row pitch and base are 16-byte aligned
    int16_t* coef;               // [n][n_blocks][64]
    size_t coef_img_stride;      // elements between images
    float* dbg;                  // optional pre-quant coefficients [n_blocks][64] natural order (image 0 of launch)
    int check_max;               // samples may exceed max_value: flag DMMT_E_INVALID (color.rs:62-65)
    uint32_t max_rep;            // max_value replicated into every u8 / u16 lane of a word
    ImgMeta* meta;               // [n] error flags
    QuantF qf;                   // divisors
    QuantF rq_hi, rq_lo;         // 1/q split in two f32 (see quantize)
    float neg_zero;              // -0.0f, opaque to ptxas: addend of the never-contracted packed multiply (mulx)
    int force_scalar;            // DMMT_K1_SCALAR=1: use the scalar kernel for P420 too (A/B measurements)
    TileTok fo;                  // fused path: token stream per tile (k1_transform_p420<FMT, true>)
    unsigned int* hist;          // fused path: [n][4][256]
};

// EXACT: IEEE divisions everywhere (fallback when the host-side proof of the fast normalisation
// fails, and for tiles of f32 input that contain non-finite / out-of-range samples).
template <int HR, int VR, int FMT, bool DBG, bool EXACT>
__device__ __forceinline__ void k1_tile(const K1Args& a, float4 (*sY)[4][16], float4 (*sCb)[(HR == 2) ? 2 : 4][16],
                                        float4 (*sCr)[(HR == 2) ? 2 : 4][16], int* s_flag) {
    constexpr int ROWS = 8 * VR;
    constexpr int CCH = (HR == 2) ? 2 : 4;     // 16-byte chunks per strip in a chroma plane row
    constexpr int YPM = HR * VR, BPM = YPM + 2;
    constexpr int MPT = TILE_W / (8 * HR);     // MCUs per tile
    constexpr int NYU = 32 * VR;               // Y blocks per tile
    constexpr int NCU = (HR == 2) ? 16 : 32;   // blocks per chroma plane per tile
    constexpr int NUNITS = NYU + 2 * NCU;

    const int tile_x = blockIdx.x, mrow = blockIdx.y, img = blockIdx.z;
    const uint8_t* __restrict__ pix = a.pixels + (size_t)img * a.img_stride_bytes;
    const size_t pitch = (size_t)a.W * Px<FMT>::kBytes;

    // ---------------- phase A: strip = 16 px x VR rows ----------------
    {
        const int sx = threadIdx.x & 15, sy = threadIdx.x >> 4;  // sy in [0,8)
        const int x0 = tile_x * TILE_W + sx * 16;
        bool bad = false;
        float cbs[16], crs[16];  // chroma of row 0, then the window sums (x outer, y inner: subsampling.rs:108-122)
#pragma unroll 1
        for (int r = 0; r < VR; r++) {
            const int yl = sy * VR + r;
            const int y = mrow * ROWS + yl;
            uint32_t w[Px<FMT>::kWords];
            load_strip_row<FMT>(pix + (size_t)y * pitch, x0, a.W, y < a.H, a.vec_ok != 0, w);
            if constexpr (FMT != DMMT_RGB_F32_NORM) {
                // color.rs:62-65: a component above max panics in the reference (SIMD-in-word compare)
                if (a.check_max) {
#pragma unroll
                    for (int i = 0; i < Px<FMT>::kWords; i++)
                        bad |= (FMT == DMMT_RGB_U8 ? __vcmpgtu4(w[i], a.max_rep) : __vcmpgtu2(w[i], a.max_rep)) != 0u;
                }
            }
            float yv[16], cb[16], cr[16];
#pragma unroll
            for (int p = 0; p < 16; p++) {
                const float nr = sample_norm<FMT, EXACT>(w, 3 * p, a.maxf, a.r_hi, a.r_lo);
                const float ng = sample_norm<FMT, EXACT>(w, 3 * p + 1, a.maxf, a.r_hi, a.r_lo);
                const float nb = sample_norm<FMT, EXACT>(w, 3 * p + 2, a.maxf, a.r_hi, a.r_lo);
                if constexpr (FMT == DMMT_RGB_F32_NORM) {
                    // the fast quantiser is proven for finite coefficients only: fence off wild input
                    bad |= !(fabsf(nr) <= 1024.0f) | !(fabsf(ng) <= 1024.0f) | !(fabsf(nb) <= 1024.0f);
                }
                rgb_to_ycbcr(nr, ng, nb, yv[p], cb[p], cr[p]);
            }
#pragma unroll
            for (int c = 0; c < 4; c++)
                sY[yl][c][sx] = make_float4(yv[4 * c], yv[4 * c + 1], yv[4 * c + 2], yv[4 * c + 3]);
            if (r == 0) {
#pragma unroll
                for (int p = 0; p < 16; p++) cbs[p] = cb[p], crs[p] = cr[p];
            } else {
                // second row of the window: only reachable for VR == 2.  Window order is
                // (x,y),(x,y+1),(x+1,y),(x+1,y+1): keep per-column values apart until the
                // horizontal combine so the additions happen in exactly that order.
#pragma unroll
                for (int p = 0; p < 16; p += 2) {
                    float sb = __fadd_rn(cbs[p], cb[p]);       // ((c00 + c10) + c01) + c11
                    sb = __fadd_rn(sb, cbs[p + 1]);
                    sb = __fadd_rn(sb, cb[p + 1]);
                    cbs[p >> 1] = sb;
                    float sr = __fadd_rn(crs[p], cr[p]);
                    sr = __fadd_rn(sr, crs[p + 1]);
                    sr = __fadd_rn(sr, cr[p + 1]);
                    crs[p >> 1] = sr;
                }
            }
        }
        if constexpr (HR == 2 && VR == 2) {
            // average(): sum / 4.0 (subsampling.rs:231-236) -- exact scaling, same as * 0.25
#pragma unroll
            for (int c = 0; c < 2; c++) {
                sCb[sy][c][sx] = make_float4(__fmul_rn(cbs[4 * c], 0.25f), __fmul_rn(cbs[4 * c + 1], 0.25f),
                                             __fmul_rn(cbs[4 * c + 2], 0.25f), __fmul_rn(cbs[4 * c + 3], 0.25f));
                sCr[sy][c][sx] = make_float4(__fmul_rn(crs[4 * c], 0.25f), __fmul_rn(crs[4 * c + 1], 0.25f),
                                             __fmul_rn(crs[4 * c + 2], 0.25f), __fmul_rn(crs[4 * c + 3], 0.25f));
            }
        } else if constexpr (HR == 2 && VR == 1) {
            // (c[x] + c[x+1]) / 2.0
            float hb[8], hr_[8];
#pragma unroll
            for (int p = 0; p < 8; p++) {
                hb[p] = __fmul_rn(__fadd_rn(cbs[2 * p], cbs[2 * p + 1]), 0.5f);
                hr_[p] = __fmul_rn(__fadd_rn(crs[2 * p], crs[2 * p + 1]), 0.5f);
            }
#pragma unroll
            for (int c = 0; c < 2; c++) {
                sCb[sy][c][sx] = make_float4(hb[4 * c], hb[4 * c + 1], hb[4 * c + 2], hb[4 * c + 3]);
                sCr[sy][c][sx] = make_float4(hr_[4 * c], hr_[4 * c + 1], hr_[4 * c + 2], hr_[4 * c + 3]);
            }
        } else {
            // P444: SubsamplingMethod::Skip (subsampling.rs:50-54)
#pragma unroll
            for (int c = 0; c < 4; c++) {
                sCb[sy][c][sx] = make_float4(cbs[4 * c], cbs[4 * c + 1], cbs[4 * c + 2], cbs[4 * c + 3]);
                sCr[sy][c][sx] = make_float4(crs[4 * c], crs[4 * c + 1], crs[4 * c + 2], crs[4 * c + 3]);
            }
        }
        if (bad) {
            if constexpr (FMT == DMMT_RGB_F32_NORM) *s_flag = 1;
            else if (a.check_max) atomicCAS(&a.meta[img].error, 0, DMMT_E_INVALID);
        }
    }
    __syncthreads();

    // ---------------- phase B: unit = one 8x8 block ----------------
    const int u = threadIdx.x;
    if (u >= NUNITS) return;
    int m, k, comp;
    float d[64];
    if (u < NYU) {
        const int q = u >> 4, sx = u & 15;
        const int byl = q >> 1, p = q & 1;
        comp = 0;
        if constexpr (HR == 2) {
            m = sx;
            k = byl * 2 + p;
        } else {
            m = 2 * sx + p;
            k = 0;
        }
#pragma unroll
        for (int r = 0; r < 8; r++) {
            const float4 lo = sY[byl * 8 + r][2 * p][sx], hi = sY[byl * 8 + r][2 * p + 1][sx];
            d[8 * r] = lo.x, d[8 * r + 1] = lo.y, d[8 * r + 2] = lo.z, d[8 * r + 3] = lo.w;
            d[8 * r + 4] = hi.x, d[8 * r + 5] = hi.y, d[8 * r + 6] = hi.z, d[8 * r + 7] = hi.w;
        }
    } else {
        const int v = u - NYU;
        const int ch = v / NCU, w = v % NCU;
        comp = 1;
        k = YPM + ch;
        int sx, c0;
        if constexpr (HR == 2) {
            sx = w;
            c0 = 0;
            m = sx;
        } else {
            sx = w & 15;
            c0 = 2 * (w >> 4);
            m = 2 * sx + (w >> 4);
        }
        const float4(*pl)[CCH][16] = ch ? sCr : sCb;
#pragma unroll
        for (int r = 0; r < 8; r++) {
            const float4 lo = pl[r][c0][sx], hi = pl[r][c0 + 1][sx];
            d[8 * r] = lo.x, d[8 * r + 1] = lo.y, d[8 * r + 2] = lo.z, d[8 * r + 3] = lo.w;
            d[8 * r + 4] = hi.x, d[8 * r + 5] = hi.y, d[8 * r + 6] = hi.z, d[8 * r + 7] = hi.w;
        }
    }
    const int gmx = tile_x * MPT + m;
    if (gmx >= a.mcus_x) return;  // tile overhangs the padded image

    // arai.rs:95-104: 8 row passes, then 8 column passes
#pragma unroll
    for (int r = 0; r < 8; r++)
        fast_arai(d[8 * r], d[8 * r + 1], d[8 * r + 2], d[8 * r + 3], d[8 * r + 4], d[8 * r + 5],
                  d[8 * r + 6], d[8 * r + 7]);
#pragma unroll
    for (int c = 0; c < 8; c++)
        fast_arai(d[c], d[8 + c], d[16 + c], d[24 + c], d[32 + c], d[40 + c], d[48 + c], d[56 + c]);

    const size_t sblk = ((size_t)mrow * a.mcus_x + gmx) * BPM + k;  // stream block index
    if constexpr (DBG) {
        if (img == 0) {
            float4* o = reinterpret_cast<float4*>(a.dbg + sblk * 64);
#pragma unroll
            for (int i = 0; i < 16; i++) o[i] = make_float4(d[4 * i], d[4 * i + 1], d[4 * i + 2], d[4 * i + 3]);
        }
    }

    // quantise (natural-order table) and emit in zig-zag order
    bool exact = EXACT;
    if constexpr (FMT == DMMT_RGB_F32_NORM && !EXACT) exact = *s_flag != 0;
    uint32_t w[32];
    if (exact) {
#pragma unroll 4
        for (int i = 0; i < 32; i++) {
            const int n0 = zz_at(2 * i), n1 = zz_at(2 * i + 1);
            w[i] = quantize<true>(d[n0], a.qf.q[comp][n0], 0.f, 0.f) |
                   (quantize<true>(d[n1], a.qf.q[comp][n1], 0.f, 0.f) << 16);
        }
    } else {
#pragma unroll
        for (int i = 0; i < 32; i++) {
            const int n0 = zz_at(2 * i), n1 = zz_at(2 * i + 1);
            const uint32_t q0 = quantize<false>(d[n0], 0.f, a.rq_hi.q[comp][n0], a.rq_lo.q[comp][n0]);
            const uint32_t q1 = quantize<false>(d[n1], 0.f, a.rq_hi.q[comp][n1], a.rq_lo.q[comp][n1]);
            w[i] = __byte_perm(q0, q1, 0x5410);
        }
    }
    uint4* out = reinterpret_cast<uint4*>(a.coef + (size_t)img * a.coef_img_stride + sblk * 64);
#pragma unroll
    for (int i = 0; i < 8; i++) out[i] = make_uint4(w[4 * i], w[4 * i + 1], w[4 * i + 2], w[4 * i + 3]);
}

template <int HR, int VR, int FMT, bool DBG, bool EXACT>
__global__ void __launch_bounds__(K1_THREADS, 4) k1_transform(const __grid_constant__ K1Args a) {
    constexpr int ROWS = 8 * VR;
    constexpr int CCH = (HR == 2) ? 2 : 4;
    __shared__ float4 sY[ROWS][4][16];
    __shared__ float4 sCb[8][CCH][16];
    __shared__ float4 sCr[8][CCH][16];
    __shared__ int s_flag;
    if constexpr (FMT == DMMT_RGB_F32_NORM) {
        if (threadIdx.x == 0) s_flag = 0;
        __syncthreads();
    }
    k1_tile<HR, VR, FMT, DBG, EXACT>(a, sY, sCb, sCr, &s_flag);
}

// ==========================================================================================
// P420 fast path: the same arithmetic with sm_100a's packed FP32 instructions (FADD2/FMUL2/FFMA2,
// PTX add/mul/fma.rn.f32x2).  A packed instruction performs two independent IEEE round-to-nearest
// operations, so results are bit-identical to the scalar path while the kernel needs about a third
// fewer issue slots, which moves it from issue-bound to FP32-lane-bound (tools/ubench_pipes.cu).
//
// CAUTION (verified on ptxas 12.9): ptxas contracts mul.rn.f32x2 feeding add/sub.rn.f32x2 into
// FFMA2 even under --fmad=false (and folds fma(a, b, -0.0) with a literal -0.0 back into that
// multiply first).  Therefore a multiplication whose product feeds an addition is issued as `mulx`:
// fma.rn.f32x2(a, c, nz) with nz = -0.0 read from the kernel arguments, i.e. opaque to ptxas.
// a * c + (-0.0) rounds once and equals fl(a * c) for every input including both zeros, and an
// FFMA2 cannot be contracted any further (SASS: one FFMA2 with the constant as a broadcast
// immediate, followed by separate FADD2s).  Plain packed multiplies (mul2) are only used where the
// product feeds a multiply, an fma multiplicand/addend, a conversion or a store.
// tests/test_cuda_parity.py compares every coefficient bit for bit, which would expose a contraction.
typedef unsigned long long f2;  // two f32 in an aligned register pair: {lo, hi}

__device__ __forceinline__ f2 pk(float lo, float hi) {
    f2 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ float lo_of(f2 v) {
    float a, b;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v));
    return a;
}
__device__ __forceinline__ float hi_of(f2 v) {
    float a, b;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v));
    return b;
}
__device__ __forceinline__ f2 bc(float c) { return pk(c, c); }
__device__ __forceinline__ f2 add2(f2 a, f2 b) {
    f2 r;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ f2 sub2(f2 a, f2 b) {
    f2 r;
    asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ f2 mul2(f2 a, f2 b) {  // ONLY where the product does not feed an add/sub
    f2 r;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ f2 fma2(f2 a, f2 b, f2 c) {
    f2 r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
    return r;
}
// product that feeds an addition: fl(a * c) as one FFMA2 that ptxas cannot contract (see CAUTION above)
__device__ __forceinline__ f2 mulx(f2 a, float c, f2 nz) { return fma2(a, bc(c), nz); }

// arai.rs:29-92 on two independent 8-vectors at once.  -v24 never materialises: with u24 = v14 + v15,
// v24 = (-v14) - v15 = -u24 exactly, (-v24) * A2 = u24 * A2 and v24 + v26 = v26 - u24 (same roundings).
// The eight scaled outputs are produced by `out(k, v, S_k)`.
template <class Out>
__device__ __forceinline__ void fast_arai2(f2 x0, f2 x1, f2 x2, f2 x3, f2 x4, f2 x5, f2 x6, f2 x7, f2 nz, Out&& out) {
    const f2 v10 = add2(x0, x7), v11 = add2(x1, x6), v12 = add2(x2, x5), v13 = add2(x3, x4);
    const f2 v14 = sub2(x3, x4), v15 = sub2(x2, x5), v16 = sub2(x1, x6), v17 = sub2(x0, x7);
    const f2 v20 = add2(v10, v13), v21 = add2(v11, v12), v22 = sub2(v11, v12), v23 = sub2(v10, v13);
    const f2 u24 = add2(v14, v15), v25 = add2(v15, v16), v26 = add2(v16, v17);
    const f2 v30 = add2(v20, v21), v31 = sub2(v20, v21), v32 = add2(v22, v23);
    const f2 v42 = mulx(v32, kA1, nz);
    const f2 t5 = mulx(sub2(v26, u24), kA5, nz);
    const f2 v44 = sub2(mulx(u24, kA2, nz), t5);
    const f2 v45 = mulx(v25, kA3, nz);
    const f2 v46 = sub2(mulx(v26, kA4, nz), t5);
    const f2 v52 = add2(v42, v23), v53 = sub2(v23, v42), v55 = add2(v45, v17), v57 = sub2(v17, v45);
    const f2 v64 = add2(v44, v57), v65 = add2(v55, v46), v66 = sub2(v55, v46), v67 = sub2(v57, v44);
    out(0, v30, kS0);
    out(4, v31, kS4);
    out(2, v52, kS2);
    out(6, v53, kS6);
    out(5, v64, kS5);
    out(1, v65, kS1);
    out(7, v66, kS7);
    out(3, v67, kS3);
}

#ifndef K1_I2FP
#define K1_I2FP 1   // bit ch set: channel ch of u8 input is converted by PRMT + I2FP.F32.U32 (integer pipe) instead of I2F.U8
                    // (one channel off the conversion pipe pays; which one is a scheduling matter: R 5.26, G 5.28, B 5.36 ms)
#endif
// raw sample i (0..47) of a strip row as f32 (before normalisation)
template <int FMT, int NW>
__device__ __forceinline__ float sample_raw(const uint32_t (&w)[NW], int i) {
    if constexpr (FMT == DMMT_RGB_F32_NORM) return __uint_as_float(w[i]);
    else if constexpr (FMT == DMMT_RGB_U8) {
        if ((K1_I2FP >> (i % 3)) & 1) {
            return __uint2float_rn(__byte_perm(w[i >> 2], 0u, 0x4440u | (uint32_t)(i & 3)));
        }
        return (float)((w[i >> 2] >> (8 * (i & 3))) & 0xFFu);
    } else return (float)((w[i >> 1] >> (16 * (i & 1))) & 0xFFFFu);
}

template <int COMP>
__device__ __forceinline__ void quantize_block_packed(const K1Args& a, const f2 (&D)[8][4], unsigned short (&qv)[64]) {
#pragma unroll
    for (int r = 0; r < 8; r++)
#pragma unroll
        for (int kp = 0; kp < 4; kp++) {
            const int n0 = 8 * r + 2 * kp;
            const f2 x = fma2(D[r][kp], pk(a.rq_hi.q[COMP][n0], a.rq_hi.q[COMP][n0 + 1]),
                              mul2(D[r][kp], pk(a.rq_lo.q[COMP][n0], a.rq_lo.q[COMP][n0 + 1])));
            const float xl = lo_of(x), xh = hi_of(x);
            const f2 h = pk(__int_as_float((__float_as_int(xl) & 0x80000000) | 0x3EFFFFFF),
                            __int_as_float((__float_as_int(xh) & 0x80000000) | 0x3EFFFFFF));
            const f2 t = add2(x, h);
            asm("cvt.rzi.s16.f32 %0, %1;" : "=h"(qv[n0]) : "f"(lo_of(t)));
            asm("cvt.rzi.s16.f32 %0, %1;" : "=h"(qv[n0 + 1]) : "f"(hi_of(t)));
        }
}

// Phase B of the P420 kernels: block `u` of the tile (stream slot m * 6 + k) -> 64 quantised
// coefficients in NATURAL order.  Returns false for blocks of MCUs beyond the padded image.
template <int FMT>
__device__ __forceinline__ bool p420_block_coefs(const K1Args& a, int u, int mcus_here, const float4 (*sY)[4][32],
                                                 const float4 (*sCb)[4][16], const float4 (*sCr)[4][16],
                                                 const int* s_flag_p, unsigned short (&qv)[64], int& m, int& k,
                                                 int& comp) {
    constexpr int NYU = 64, NCU = 16;
    f2 P[4][8];  // P[j][c] = {d[2j][c], d[2j+1][c]}
    const f2 nz = bc(a.neg_zero);
    if (u < NYU) {
        // warp 0: upper block row of the MCU row, warp 1: lower; lane = block column = 8-pixel strip
        const int byl = u >> 5, bx = u & 31;
        comp = 0, m = bx >> 1, k = byl * 2 + (bx & 1);
#pragma unroll
        for (int j = 0; j < 4; j++)
#pragma unroll
            for (int cc = 0; cc < 4; cc++) {
                const float4 v = sY[4 * byl + j][cc][bx];
                P[j][2 * cc] = pk(v.x, v.y), P[j][2 * cc + 1] = pk(v.z, v.w);
            }
    } else {
        const int v = u - NYU;
        const int ch = v / NCU, sx = v % NCU;
        comp = 1, m = sx, k = 4 + ch;
        const float4(*pl)[4][16] = ch ? sCr : sCb;
#pragma unroll
        for (int j = 0; j < 4; j++)
#pragma unroll
            for (int cc = 0; cc < 4; cc++) {
                const float4 t = pl[j][cc][sx];
                P[j][2 * cc] = pk(t.x, t.y), P[j][2 * cc + 1] = pk(t.z, t.w);
            }
    }
    if (m >= mcus_here) return false;  // tile overhangs the padded image

    // row passes on row pairs; the scaled outputs are written as COLUMN pairs C[r][kp] = {d[r][2kp], d[r][2kp+1]}
    float d[64];
#pragma unroll
    for (int j = 0; j < 4; j++)
        fast_arai2(P[j][0], P[j][1], P[j][2], P[j][3], P[j][4], P[j][5], P[j][6], P[j][7], nz, [&](int kk, f2 v, float S) {
            d[16 * j + kk] = __fmul_rn(lo_of(v), S);
            d[16 * j + 8 + kk] = __fmul_rn(hi_of(v), S);
        });
    // column passes on column pairs; outputs feed the quantiser's multiplies, so the scale is packed
    f2 D[8][4];  // D[r][kp] = {coef[8r + 2kp], coef[8r + 2kp + 1]}
#pragma unroll
    for (int kp = 0; kp < 4; kp++)
        fast_arai2(pk(d[2 * kp], d[2 * kp + 1]), pk(d[8 + 2 * kp], d[8 + 2 * kp + 1]), pk(d[16 + 2 * kp], d[16 + 2 * kp + 1]),
                   pk(d[24 + 2 * kp], d[24 + 2 * kp + 1]), pk(d[32 + 2 * kp], d[32 + 2 * kp + 1]),
                   pk(d[40 + 2 * kp], d[40 + 2 * kp + 1]), pk(d[48 + 2 * kp], d[48 + 2 * kp + 1]),
                   pk(d[56 + 2 * kp], d[56 + 2 * kp + 1]), nz, [&](int r, f2 v, float S) { D[r][kp] = mul2(v, bc(S)); });

    // quantise: x = fma(d, rq_hi, d * rq_lo), round half away from zero, saturate (see quantize<>).
    // `comp` is warp-uniform (warps 0 and 1 luma, warp 2 chroma), so the table is selected by a uniform
    // branch and its entries become uniform-register operands.
    bool exact = false;
    if constexpr (FMT == DMMT_RGB_F32_NORM) exact = *s_flag_p != 0;
    if (!exact) {
        if (comp == 0) quantize_block_packed<0>(a, D, qv);
        else quantize_block_packed<1>(a, D, qv);
    } else {
#pragma unroll
        for (int i = 0; i < 64; i++) {
            const float dv = (i & 1) ? hi_of(D[i >> 3][(i & 7) >> 1]) : lo_of(D[i >> 3][(i & 7) >> 1]);
            qv[i] = (unsigned short)quantize<true>(dv, a.qf.q[comp][i], 0.f, 0.f);
        }
    }
    return true;
}

// categorize.rs:22-41 for a non-zero value: category and the cat low bits of the pattern
__device__ __forceinline__ void k1_cat_bits(int v, int& cat, uint32_t& bits) {
    const uint32_t av = (uint32_t)abs(v);
    uint32_t msb;
    asm("bfind.u32 %0, %1;" : "=r"(msb) : "r"(av));  // position of the leading one (v != 0)
    cat = (int)msb + 1;
    // v > 0: v; v < 0: v - 1 == ~|v|  -> one xor with the sign mask, then the low `cat` bits
    bits = (av ^ (uint32_t)(v >> 31)) & ~(0xFFFFFFFFu << cat);
}
__device__ __forceinline__ uint32_t k1_token(int table, int sym, int nzrl, uint32_t extra) {
    return (uint32_t)sym | ((uint32_t)table << 8) | ((uint32_t)nzrl << 10) | (extra << 16);
}

// Tokens of one block (fused path): DC difference against the previous block of the same component
// inside the tile (categorize.rs:157-161), then the walk over the non-zero AC positions of the staged
// block (categorize.rs:132-151: run of zeros, one 0xF0 per 16 zeros, EOB 0x00 when the tail is zero).
__device__ __forceinline__ bool p420_tokenize_block(uint32_t* __restrict__ dst, bool store, uint32_t off, unsigned short dcq,
                                                    int slot, int m, int k, int comp, uint32_t mlo, uint32_t mhi,
                                                    const short* s_dc, const uint4* s_stage, unsigned int* s_hist,
                                                    uint32_t* dcpos) {
    const int tdc = comp ? T_CDC : T_YDC, tac = tdc + 1;
    unsigned int* const h_dc = s_hist + (comp ? 16 : 0);     // tile histogram layout: SH_YDC / SH_CDC / SH_YAC / SH_CAC
    // AC symbols in the bank-swizzled form of tok_swz (dmmt_common.cuh): symp = sym ^ (run & 15) for both tables;
    // the chroma table's extra "^ 8" lives in tbase, so the walk needs no operation more than the plain symbol did.
    // The tile histogram is indexed by symp (un-swizzled at the flush).
    unsigned int* const h_ac = s_hist + (comp ? 288 : 32);
    const uint32_t tbase = ((uint32_t)tac << 8) | (comp ? 8u : 0u);
    bool ok = true;
    {
        const int ps = k == 0 ? slot - 3 : (k < 4 ? slot - 1 : slot - 6);
        if ((k != 0 && k < 4) || m > 0) {
            const int diff = (int)(short)((short)dcq - s_dc[ps]);
            int cat = 0;
            uint32_t bits = 0;
            if (diff != 0) k1_cat_bits(diff, cat, bits);
            ok &= cat <= 15;
            atomicAdd(&h_dc[cat & 15], 1u);
            if (store) dst[off] = k1_token(tdc, cat & 15, 0, bits);
        } else {
            // predictor is in the previous tile (or is the seed): k2_fix_dc finishes this token
            if (k >= 4) dcpos[k - 4] = off;
            if (store) dst[off] = (uint32_t)dcq << 16;
        }
        ++off;
    }
    const int16_t* sb = reinterpret_cast<const int16_t*>(s_stage + slot * 8);
    const int sw8 = (slot & 7) << 3;
    int prev = 0;
    uint32_t nzrl_total = 0;
#pragma unroll
    for (int half = 0; half < 2; half++) {
        uint32_t mk = __brev(half ? mhi : (mlo & ~1u));  // bit-reversed once: the walk needs one FLO per step
        if (mk) {
            // software-pipelined: the coefficient of the NEXT non-zero position is requested before the current
            // one is categorised (an exhausted mask gives position 32 -- a harmless read inside the plane storage)
            int lz = __clz((int)mk);
            int v = sb[(32 * half + lz) ^ sw8];  // chunk (pos >> 3) lives at slot (pos >> 3) ^ (slot & 7)
            while (true) {
                const int pos = 32 * half + lz;
                mk &= ~(0x80000000u >> lz);
                const int nlz = __clz((int)mk);
                const int nv = sb[(32 * half + nlz) ^ sw8];
                const int run = pos - prev - 1;
                prev = pos;
                int cat;
                uint32_t bits;
                k1_cat_bits(v, cat, bits);
                ok &= cat <= 15;
                const int symp = ((run & 15) << 4) | ((cat ^ run) & 15);
                nzrl_total += (uint32_t)(run >> 4);
                atomicAdd(&h_ac[symp], 1u);
                if (store) dst[off] = ((uint32_t)symp ^ (tbase + ((uint32_t)(run >> 4) << 10))) | (bits << 16);
                ++off;
                if (!mk) break;
                lz = nlz, v = nv;
            }
        }
    }
    if (nzrl_total) atomicAdd(&h_ac[0xFF], nzrl_total);  // ZRL 0xF0
    if (prev != 63) {
        atomicAdd(&h_ac[0], 1u);
        if (store) dst[off] = tbase;  // EOB 0x00
    }
    return ok;
}

#ifndef K1_BULK
#define K1_BULK 1   // 1: the tile's tokens (cp.async.bulk) and symbol counts (cp.reduce.async.bulk .add) leave through the
                    // bulk-copy engine; 0: copy loop + per-bin global reductions
#endif
#ifndef K1_BRED
#define K1_BRED 1   // (with K1_BULK) histogram flush by cp.reduce.async.bulk; 0: per-bin REDG loop
#endif
#ifndef K1_WALK2
#define K1_WALK2 1   // 1: reversed masks from VIMNMX + IMAD, run LUT, 28-instruction walk step; 0: round-1 walk
#endif
// Per-run part of an AC token, one 64-entry table per component class (luma | chroma), in shared memory:
//   F[run] = table << 8 | (run >> 4) << 10 | (run & 15) << 4 | ((run & 15) ^ (chroma ? 8 : 0))
// so that the token of a coefficient of category cat is (bits << 16) | (F[run] ^ cat) -- tok_swz(table, run << 4 | cat)
// plus the ZRL count -- and F[0] itself is the EOB token.
__device__ __forceinline__ uint32_t k1_run_lut_entry(int chroma, int run) {
    const uint32_t tac = chroma ? T_CAC : T_YAC, r = (uint32_t)run & 15u;
    // low half: the token's symbol part; high half: byte offset of the PLAIN symbol row (run nibble << 4) in the tile
    // histogram, which is kept in plain symbol order so that it can be added to the image histogram as it is
    return (tac << 8) | (((uint32_t)run >> 4) << 10) | (r << 4) | (r ^ (chroma ? 8u : 0u)) | ((r << 6) << 16);
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// HOT: the token destination is the tile's shared-memory buffer (32-bit shared addresses, always stored).  The cold
// variant (a tile too dense for that buffer: tokens straight to global memory, or nowhere when the tile does not fit its
// region either) is the same code behind a generic pointer, kept out of line so that it costs no instruction-cache space.
template <bool CHECK, bool HOT>
__device__ __forceinline__ bool p420_walk(uint32_t* __restrict__ dst, bool store, uint32_t off, unsigned short dcq, int slot,
                                          int m, int k, int comp, uint32_t mrl, uint32_t mrh, const short* s_dc,
                                          const uint4* s_stage, unsigned int* s_hist, const uint32_t* s_flut,
                                          uint32_t* dcpos, uint32_t nz_bits) {
    const int tdc = comp ? T_CDC : T_YDC;
    unsigned int* const h_dc = s_hist + (comp ? 16 : 0);
    const uint32_t hac_a = smem_u32(s_hist + (comp ? 288 : 32));   // AC bins of this component class, plain symbol order
    const uint32_t lut_a = smem_u32(s_flut + (comp ? 64 : 0));
    bool ok = true;
    {
        const int ps = k == 0 ? slot - 3 : (k < 4 ? slot - 1 : slot - 6);
        if ((k != 0 && k < 4) || m > 0) {
            const int diff = (int)(short)((short)dcq - s_dc[ps]);
            int cat = 0;
            uint32_t bits = 0;
            if (diff != 0) k1_cat_bits(diff, cat, bits);
            if (CHECK) ok &= cat <= 15;
            atomicAdd(&h_dc[cat & 15], 1u);
            if (store) dst[off] = k1_token(tdc, cat & 15, 0, bits);
        } else {
            // predictor is in the previous tile (or is the seed): k2_fix_dc finishes this token
            if (k >= 4) dcpos[k - 4] = off;
            if (store) dst[off] = (uint32_t)dcq << 16;
        }
        ++off;
    }
    const uint32_t sbase = smem_u32(s_stage + slot * 8);
    const uint32_t swz = (((uint32_t)slot & 7u) << 3) | 1u;
    uint32_t dpa = HOT ? smem_u32(dst + off) : 0u;   // HOT: running shared address of the next token
    int hp = 31;   // 31 - (position of the previous non-zero coefficient), in the coordinates of the current half
    uint32_t nzrl_total = 0;
    // 1 and ~1 derived from the opaque -0.0 argument: as literals ptxas re-materialises them in every step of the walk
    const uint32_t c_one = nz_bits >> 31, c_m2 = ~c_one;
#pragma unroll
    for (int half = 0; half < 2; half++) {
        uint32_t mk = half ? mrh : (mrl & 0x7FFFFFFFu);
        uint32_t cs = (half ? 63u : 31u) ^ swz;   // position p = (31 | 63) - h sits at int16 index (p ^ swz) == h ^ cs
        asm volatile("" : "+r"(cs));              // keep it ONE loop-invariant register
        bool more = mk != 0u;
#pragma unroll 1
        while (more) {
            uint32_t h, f, F;
            int v;
            asm("bfind.u32 %0, %1;" : "=r"(h) : "r"(mk));
            asm volatile("ld.shared.s16 %0, [%1];" : "=r"(v) : "r"(sbase + ((h ^ cs) << 1)));
            const int run = hp - (int)h - 1;
            asm volatile("ld.shared.u32 %0, [%1];" : "=r"(F) : "r"(lut_a + ((uint32_t)run << 2)));
            const uint32_t bit = c_one << h;
            more = bit != mk;
            mk ^= bit;
            hp = (int)h;
            nzrl_total += (uint32_t)run >> 4;
            asm("bfind.u32 %0, %1;" : "=r"(f) : "r"((uint32_t)abs(v)));   // category - 1
            if (CHECK) ok &= f < 15u;
            const uint32_t sym = (F ^ (f + 1u)) & 0xFFFFu;
            asm volatile("red.shared.add.u32 [%0], 1;" ::"r"(hac_a + (F >> 16) + ((f + 1u) << 2)) : "memory");
            // v > 0: v; v < 0: v - 1; the low `cat` bits of it (categorize.rs:22-41)
            const uint32_t bits = (uint32_t)(v + (v >> 31)) & ~(c_m2 << f);
            const uint32_t tok = bits * 65536u + sym;
            if (HOT) {
                asm volatile("st.shared.u32 [%0], %1;" ::"r"(dpa), "r"(tok) : "memory");
                dpa += 4;
            } else {
                if (store) dst[off] = tok;
                ++off;
            }
        }
        hp += 32;
    }
    if (nzrl_total) atomicAdd(&s_hist[(comp ? 288 : 32) + 0xF0], nzrl_total);  // ZRL
    if (hp != 32) {                                                                             // coefficient 63 is zero: EOB 0x00
        atomicAdd(&s_hist[(comp ? 288 : 32)], 1u);
        const uint32_t eob = k1_run_lut_entry(comp, 0) & 0xFFFFu;
        if (HOT) asm volatile("st.shared.u32 [%0], %1;" ::"r"(dpa), "r"(eob) : "memory");
        else if (store) dst[off] = eob;
    }
    return ok;
}
template <bool CHECK>
__device__ __noinline__ bool p420_walk_cold(uint32_t* __restrict__ dst, bool store, uint32_t off, unsigned short dcq, int slot,
                                            int m, int k, int comp, uint32_t mrl, uint32_t mrh, const short* s_dc,
                                            const uint4* s_stage, unsigned int* s_hist, const uint32_t* s_flut,
                                            uint32_t* dcpos, uint32_t nz_bits) {
    return p420_walk<CHECK, false>(dst, store, off, dcq, slot, m, k, comp, mrl, mrh, s_dc, s_stage, s_hist, s_flut, dcpos, nz_bits);
}

constexpr int P420_THREADS = 96;   // one thread per 8x8 block of the tile (16 MCUs x 6 blocks)
// shared-memory histogram of a tile: only the bins that exist, [Y-DC 16 | C-DC 16 | Y-AC 256 | C-AC 256]
constexpr int SH_YDC = 0, SH_CDC = 16, SH_YAC = 32, SH_CAC = 288, SH_BINS = 544;

#ifndef K1_MINB
#define K1_MINB 8
#endif
#ifndef K1_PREFETCH
#define K1_PREFETCH 1
#endif
#ifndef K1_MAGIC
#define K1_MAGIC 0   // bit ch set: channel ch of u8 input is converted with the 2^23 trick instead of I2F
#endif
#ifndef K1_HUNROLL
#define K1_HUNROLL 1   // the two 4-pixel halves of a strip unrolled: -1.1 % (round 2; round 1 measured the opposite on a larger kernel)
#endif
#ifndef K1_INNER
#define K1_INNER 1
#endif
// CTA = 96 threads = 3 warps, every one of them busy in every phase: 8 CTAs (24 warps) per SM at 80 registers.
template <int FMT, bool FUSED, bool VEC>
__global__ void __launch_bounds__(P420_THREADS, FUSED ? K1_MINB : 6) k1_transform_p420(const __grid_constant__ K1Args a) {
    // planes as ROW-PAIR interleaved float4 = {v(x,2j), v(x,2j+1), v(x+1,2j), v(x+1,2j+1)}:
    //   sY[row pair][16-byte chunk of the strip: 2 columns][8-pixel strip], sC*[row pair][chunk][MCU]
    __shared__ __align__(16) float4 s_planes[8 * 4 * 32 + 2 * 4 * 4 * 16];  // 24 KB: Y | Cb | Cr (reused for tokens)
    float4(*sY)[4][32] = reinterpret_cast<float4(*)[4][32]>(s_planes);
    float4(*sCb)[4][16] = reinterpret_cast<float4(*)[4][16]>(s_planes + 1024);
    float4(*sCr)[4][16] = reinterpret_cast<float4(*)[4][16]>(s_planes + 1280);
    // quantised blocks of the tile in stream order (zig-zag, swizzled): own array on the coefficient path,
    // the first half of the (dead) plane storage on the fused path
    __shared__ uint4 s_stage_own[FUSED ? 1 : 96 * 8];
    uint4* s_stage = FUSED ? reinterpret_cast<uint4*>(s_planes) : s_stage_own;
    __shared__ __align__(16) unsigned int s_hist[FUSED ? SH_BINS : 4];   // fused path: symbol counts of the tile
    __shared__ uint32_t s_cnt[FUSED ? 100 : 1];         // tokens per block, then exclusive offsets (+ total)
    __shared__ short s_dc[FUSED ? 96 : 1];
    __shared__ uint32_t s_flut[FUSED && K1_WALK2 ? 128 : 1];   // k1_run_lut_entry: luma | chroma
    __shared__ int s_flag;
    constexpr int BPM = 6, MPT = 16, NUNITS = 96;
    pdl_trigger();   // (compiled out, see dmmt_common.cuh: K2b's CTAs becoming resident while the last tiles drain did not pay)
    if constexpr (FMT == DMMT_RGB_F32_NORM) {
        if (threadIdx.x == 0) s_flag = 0;
        __syncthreads();
    }
    const int tile_x = blockIdx.x, mrow = blockIdx.y, img = blockIdx.z;
    const uint8_t* __restrict__ pix = a.pixels + (size_t)img * a.img_stride_bytes;
    const size_t pitch = (size_t)a.W * Px<FMT>::kBytes;

    // ---------------- phase A: strip = 8 px x 2 rows, both rows packed in one register pair ----------------
    // 256 strips per tile = 8 row pairs x 32 strips; a warp takes one whole row pair per round (rounds 0-1: all
    // three warps, round 2: warps 0 and 1), so the plane stores of a warp are contiguous.
    {
        constexpr int NW = Px<FMT>::kWords / 2;
        const f2 rhi = bc(a.r_hi), rlo = bc(a.r_lo), nz = bc(a.neg_zero);
        bool bad = false;
        // Rolled on purpose (strips, then the two 4-pixel halves of a strip): the loop body is what stays in
        // the instruction cache; fully unrolled, the kernel's hot code exceeds the 32 KB L1.5 I-cache and a
        // quarter of phase B's cycles were instruction-fetch stalls (profiles/).
        // A thread keeps its strip column (sx) and walks down the row pairs sy, sy + 3, sy + 6: everything that
        // depends on the column only is hoisted, the row pointer advances by six rows per strip.
        const int sx = threadIdx.x & 31;
        const int navail = a.W - (tile_x * TILE_W + sx * 8);
        const bool fast = a.vec_ok != 0 && navail >= 8;
        int y = mrow * 16 + 2 * (int)(threadIdx.x >> 5);
        const uint8_t* rowp = pix + (size_t)y * pitch + (size_t)(tile_x * TILE_W + sx * 8) * Px<FMT>::kBytes;
        // unaligned rows: word loads may touch up to 3 bytes before / after the strip -- fine when a pixel of the
        // same image precedes (x0 > 0 or an earlier row; the very first strip only if the image base is aligned)
        // and follows (another pixel of the row or another row)
        const bool head0 = sx > 0 || tile_x > 0 || (reinterpret_cast<uintptr_t>(pix) & 3) == 0;
        // `inner`: the whole tile lies inside the image and the strip is complete and aligned -- no per-row tests
        const bool inner = VEC && K1_INNER && fast && mrow * 16 + 16 <= a.H;
        auto load_rows = [&](const uint8_t* p, int yy, uint32_t (&r0)[NW], uint32_t (&r1)[NW]) {
            if (inner) {
                load_strip8_row<FMT, true>(p, 8, true, true, false, r0);
                load_strip8_row<FMT, true>(p + pitch, 8, true, true, false, r1);
                return;
            }
            const bool w0ok = !VEC && navail >= 8 && (head0 || yy > 0) && (navail >= 9 || yy + 1 < a.H);
            const bool w1ok = !VEC && navail >= 8 && (navail >= 9 || yy + 2 < a.H);
            load_strip8_row<FMT, VEC>(p, navail, yy < a.H, fast, w0ok, r0);
            load_strip8_row<FMT, VEC>(p + pitch, navail, yy + 1 < a.H, fast, w1ok, r1);
        };
        // the rows of the next strip are requested before the current strip is converted (integer formats: the
        // second buffer costs 2 * NW registers, too many for the 96-byte f32 rows)
        constexpr bool PREFETCH = K1_PREFETCH && FMT != DMMT_RGB_F32_NORM;
        uint32_t n0[PREFETCH ? NW : 1], n1[PREFETCH ? NW : 1];
        if constexpr (PREFETCH) load_rows(rowp, y, n0, n1);
#pragma unroll 1
        for (int sy = threadIdx.x >> 5; sy < 8; sy += 3, y += 6, rowp += 6 * pitch) {  // sy = row pair of the MCU row (warp-uniform)
            uint32_t w0[NW], w1[NW];
            if constexpr (PREFETCH) {
#pragma unroll
                for (int i = 0; i < NW; i++) w0[i] = n0[i], w1[i] = n1[i];
                if (sy + 3 < 8) load_rows(rowp + 6 * pitch, y + 6, n0, n1);
            } else {
                load_rows(rowp, y, w0, w1);
            }
            if constexpr (FMT != DMMT_RGB_F32_NORM) {
                if (a.check_max) {  // color.rs:62-65 (SIMD-in-word compare)
#pragma unroll
                    for (int i = 0; i < NW; i++) {
                        bad |= (FMT == DMMT_RGB_U8 ? __vcmpgtu4(w0[i], a.max_rep) : __vcmpgtu2(w0[i], a.max_rep)) != 0u;
                        bad |= (FMT == DMMT_RGB_U8 ? __vcmpgtu4(w1[i], a.max_rep) : __vcmpgtu2(w1[i], a.max_rep)) != 0u;
                    }
                }
            }
            // chroma sample c of the strip is column 4 * (sx & 1) + c of chroma block sx >> 1: chunk
            // 2 * (sx & 1) + (c >> 1), column c & 1 of it; this thread owns half `sy & 1` of each row pair
            float* cbp = reinterpret_cast<float*>(&sCb[sy >> 1][2 * (sx & 1)][sx >> 1]) + (sy & 1);
            float* crp = reinterpret_cast<float*>(&sCr[sy >> 1][2 * (sx & 1)][sx >> 1]) + (sy & 1);
            float4* yp = &sY[sy][0][sx];
#if K1_HUNROLL
#pragma unroll
#else
#pragma unroll 1
#endif
            for (int h = 0; h < 2; h++) {  // pixels 4h .. 4h+3 of the strip = the first NW / 2 words
#pragma unroll
                for (int cl = 0; cl < 2; cl++) {  // chunk 2h + cl = 2 pixels x 2 rows
                    f2 yy[2], cb[2], cr[2];
#pragma unroll
                    for (int q = 0; q < 2; q++) {
                        const int p = 2 * cl + q;
                        f2 n[3];
#pragma unroll
                        for (int ch = 0; ch < 3; ch++) {
                            // (u8 -> f32 stays on I2F: the PRMT 0x4B000000 + FADD2 trick was measured 7 % slower)
                            f2 v;
                            if (FMT == DMMT_RGB_U8 && ((K1_MAGIC >> ch) & 1)) {
                                // 0x4B0000vv is 2^23 + v: one PRMT per sample, one packed subtraction per pair -- takes
                                // these conversions off the quarter-rate I2F pipe
                                const int i = 3 * p + ch;
                                v = add2(pk(__uint_as_float(__byte_perm(w0[i >> 2], 0x4B000000u, 0x7440 | (i & 3))),
                                            __uint_as_float(__byte_perm(w1[i >> 2], 0x4B000000u, 0x7440 | (i & 3)))),
                                         bc(-8388608.0f));
                            } else {
                                v = pk(sample_raw<FMT>(w0, 3 * p + ch), sample_raw<FMT>(w1, 3 * p + ch));
                            }
                            if constexpr (FMT == DMMT_RGB_F32_NORM) {
                                n[ch] = v;
                                bad |= !(fabsf(lo_of(v)) <= 1024.0f) | !(fabsf(hi_of(v)) <= 1024.0f);
                            } else {
                                n[ch] = fma2(v, rhi, mul2(v, rlo));  // v / max, proven per plan (make_k1_consts)
                            }
                        }
                        // color.rs:75-100, products that feed additions as mulx, sums and final scale packed
                        constexpr float kShift = 128.0f / 255.0f;
                        yy[q] = mul2(add2(add2(add2(mulx(n[0], 0.299f, nz), mulx(n[1], 0.587f, nz)), mulx(n[2], 0.114f, nz)),
                                          bc(-kShift)),
                                     bc(255.0f));
                        if constexpr (FMT == DMMT_RGB_F32_NORM) {
                            cb[q] = mul2(add2(add2(mulx(n[0], -0.1687f, nz), mulx(n[1], -0.3312f, nz)), mulx(n[2], 0.5f, nz)),
                                         bc(255.0f));
                            cr[q] = mul2(add2(add2(mulx(n[0], 0.5f, nz), mulx(n[1], -0.4186f, nz)), mulx(n[2], -0.0813f, nz)),
                                         bc(255.0f));
                        } else {
                            // x * 0.5 is exact for normalised integer samples, so s + x * 0.5 == fma(x, 0.5, s) bit for bit:
                            // the two halvings ride on the additions they feed
                            cb[q] = mul2(fma2(n[2], bc(0.5f), add2(mulx(n[0], -0.1687f, nz), mulx(n[1], -0.3312f, nz))), bc(255.0f));
                            cr[q] = mul2(add2(fma2(n[0], bc(0.5f), mulx(n[1], -0.4186f, nz)), mulx(n[2], -0.0813f, nz)), bc(255.0f));
                        }
                    }
                    yp[cl * 32] = make_float4(lo_of(yy[0]), hi_of(yy[0]), lo_of(yy[1]), hi_of(yy[1]));
                    // window (x,y),(x,y+1),(x+1,y),(x+1,y+1), f32 sum from 0, / 4 (subsampling.rs:108-122,231-236)
                    const float sb = __fadd_rn(__fadd_rn(__fadd_rn(lo_of(cb[0]), hi_of(cb[0])), lo_of(cb[1])), hi_of(cb[1]));
                    const float sr = __fadd_rn(__fadd_rn(__fadd_rn(lo_of(cr[0]), hi_of(cr[0])), lo_of(cr[1])), hi_of(cr[1]));
                    cbp[cl * 2] = __fmul_rn(sb, 0.25f);
                    crp[cl * 2] = __fmul_rn(sr, 0.25f);
                }
                yp += 2 * 32, cbp += 64, crp += 64;
#pragma unroll
                for (int i = 0; i < NW / 2; i++) w0[i] = w0[i + NW / 2], w1[i] = w1[i + NW / 2];
            }
        }
        if (bad) {
            if constexpr (FMT == DMMT_RGB_F32_NORM) s_flag = 1;
            else if (a.check_max) atomicCAS(&a.meta[img].error, 0, DMMT_E_INVALID);
        }
    }
    __syncthreads();

    // ---------------- phase B: unit = one 8x8 block ----------------
    const int u = threadIdx.x;
    const int mcus_here = min(MPT, a.mcus_x - tile_x * MPT);  // MCUs of this tile inside the padded image
    unsigned short qv[64];
    int m = 0, k = 0, comp = 0;
    bool active = false;
    if (u < NUNITS) active = p420_block_coefs<FMT>(a, u, mcus_here, sY, sCb, sCr, &s_flag, qv, m, k, comp);
    const int slot = m * BPM + k;

    // zig-zag by register renaming, two i16 per word.  Coefficient path: straight into the tile's staging
    // area (slot = stream order).  Fused path: the words stay in registers until the plane storage, which
    // the staging area and the token buffer reuse, is dead (after the barrier on the block counts).
    uint32_t mlo = 0, mhi = 0;  // occupancy mask of the block (fused path; K1_WALK2: bit-reversed)
    uint32_t wq[FUSED ? 32 : 1];
    if (active) {
        uint32_t mm[2] = {0u, 0u};
        uint32_t G[4] = {0u, 0u, 0u, 0u};
#pragma unroll
        for (int i = 0; i < 8; i++) {
            uint32_t w[4];
            if constexpr (FUSED && K1_WALK2) {
                // halves swapped (even position in the high half): min(w, 0x00010001) has the flag of the odd
                // position in bit 0 and of the even one in bit 16, and ONE multiply-add per word drops both at their
                // place of a bit-reversed mask: word wi of a group of 8 times 2^(15 - 2 wi) + 2^(30 - 2 wi) puts
                // the even flag at bit 31 - 2 wi and the odd one at 30 - 2 wi (the even flag's second image falls off
                // the top, the odd flag's second image lands in the low half, which is discarded).
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    asm("mov.b32 %0, {%1, %2};" : "=r"(w[j]) : "h"(qv[zz_at(8 * i + 2 * j + 1)]), "h"(qv[zz_at(8 * i + 2 * j)]));
                    wq[4 * i + j] = w[j];
                    const int wi = 4 * (i & 1) + j;
                    G[i >> 1] += __vminu2(w[j], 0x00010001u) * ((1u << (15 - 2 * wi)) + (1u << (30 - 2 * wi)));
                }
                continue;
            }
#pragma unroll
            for (int j = 0; j < 4; j++)
                asm("mov.b32 %0, {%1, %2};" : "=r"(w[j]) : "h"(qv[zz_at(8 * i + 2 * j)]), "h"(qv[zz_at(8 * i + 2 * j + 1)]));
            if constexpr (FUSED) {
#pragma unroll
                for (int j = 0; j < 4; j++) wq[4 * i + j] = w[j];
                // 0xFFFF per non-zero half -> one flag byte per coefficient -> 8 mask bits (as in K2)
                const uint32_t f01 = __byte_perm(__vcmpne2(w[0], 0u), __vcmpne2(w[1], 0u), 0x6420);
                const uint32_t f23 = __byte_perm(__vcmpne2(w[2], 0u), __vcmpne2(w[3], 0u), 0x6420);
                const uint32_t b01 = ((f01 & 0x08040201u) * 0x01010101u) >> 24;
                const uint32_t b23 = ((f23 & 0x08040201u) * 0x01010101u) >> 24;
                mm[i >> 2] |= (b01 | (b23 << 4)) << (8 * (i & 3));
            } else {
                s_stage[slot * 8 + (i ^ (slot & 7))] = make_uint4(w[0], w[1], w[2], w[3]);
            }
        }
        mlo = mm[0], mhi = mm[1];
        if constexpr (FUSED && K1_WALK2) {
            mlo = __byte_perm(G[1], G[0], 0x7632);   // positions 0..31 in bits 31..0
            mhi = __byte_perm(G[3], G[2], 0x7632);   // positions 32..63
        }
    }
    if constexpr (!FUSED) {
        // the tile leaves as ONE contiguous, coalesced 12 KB run (16 MCUs x 6 blocks x 128 B in stream order)
        __syncthreads();
        const size_t sblk0 = ((size_t)mrow * a.mcus_x + (size_t)tile_x * MPT) * BPM;
        uint4* out = reinterpret_cast<uint4*>(a.coef + (size_t)img * a.coef_img_stride + sblk0 * 64);
        const int n16 = mcus_here * BPM * 8;
        for (int i = threadIdx.x; i < n16; i += P420_THREADS) {
            const int sl = i >> 3;
            out[i] = s_stage[sl * 8 + ((i & 7) ^ (sl & 7))];
        }
    } else {
        // ---------------- fused tokeniser (replaces K2 on this path) ----------------
        // The quantised block never leaves the SM: its occupancy mask comes from the registers, the
        // walk over the non-zero positions (categorize.rs:132-169) reads the block's own staged words,
        // and the tokens (k2_entropy.cu format) are written compactly in stream order.  The three DC
        // tokens whose predictor lives in the previous tile are left as place-holders carrying the raw
        // DC for k2_fix_dc.
        uint32_t* s_ctok = reinterpret_cast<uint32_t*>(s_planes) + 96 * 8 * 4;  // second half of the plane storage
        constexpr uint32_t S_CTOK_CAP = sizeof(s_planes) / 4 - 96 * 8 * 4;
        for (int i = threadIdx.x; i < SH_BINS / 4; i += P420_THREADS) reinterpret_cast<uint4*>(s_hist)[i] = make_uint4(0, 0, 0, 0);
        if constexpr (K1_WALK2)
            for (int i = threadIdx.x; i < 128; i += P420_THREADS) s_flut[i] = k1_run_lut_entry(i >> 6, i & 63);
        uint32_t cnt = 0;
        if (active) {
            // DC + one token per non-zero AC (ZRLs ride on it) + EOB unless coefficient 63 is non-zero
            if constexpr (K1_WALK2) cnt = 1u + __popc(mlo & 0x7FFFFFFFu) + __popc(mhi) + ((mhi & 1u) ? 0u : 1u);
            else cnt = 1u + __popc(mlo & ~1u) + __popc(mhi) + ((mhi >> 31) ? 0u : 1u);
            s_dc[slot] = (short)qv[0];
        }
        if (u < NUNITS) s_cnt[slot] = cnt;  // every slot is written: blocks beyond the padded image count 0
        __syncthreads();  // everybody has read the planes; s_hist is zero; the 96 counts are complete
        if (active) {  // now the plane storage is free: stage the block for the walk (read back by this thread only)
#pragma unroll
            for (int i = 0; i < 8; i++)
                s_stage[slot * 8 + (i ^ (slot & 7))] = make_uint4(wq[4 * i], wq[4 * i + 1], wq[4 * i + 2], wq[4 * i + 3]);
        }
        // exclusive scan of the 96 per-block counts in stream order, done redundantly by every warp
        // (3 counts per lane + shuffles) so that no second barrier is needed
        uint32_t my_off = 0, total;
        {
            const int l = threadIdx.x & 31;
            const uint32_t c0 = s_cnt[3 * l], c1 = s_cnt[3 * l + 1], c2 = s_cnt[3 * l + 2];
            uint32_t inc = c0 + c1 + c2;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const uint32_t t = __shfl_up_sync(0xffffffffu, inc, d);
                if (l >= d) inc += t;
            }
            total = __shfl_sync(0xffffffffu, inc, 31);
            const uint32_t ex = inc - (c0 + c1 + c2);
            // offset of `slot`: lane slot / 3 holds the prefix of its three slots
            const uint32_t e0 = __shfl_sync(0xffffffffu, ex, slot / 3);
            const uint32_t a0 = __shfl_sync(0xffffffffu, c0, slot / 3), a1 = __shfl_sync(0xffffffffu, c1, slot / 3);
            my_off = e0 + (slot % 3 > 0 ? a0 : 0u) + (slot % 3 > 1 ? a1 : 0u);
        }
        const uint32_t tile = (uint32_t)mrow * a.fo.tiles_x + (uint32_t)tile_x;
        uint32_t* g_tok = a.fo.tok + (size_t)img * a.fo.img_stride_words + (size_t)tile * a.fo.tile_cap;
        const bool fits = total <= a.fo.tile_cap;          // else DMMT_E_OVERFLOW: host retries with the worst-case capacity
        const bool in_smem = total <= S_CTOK_CAP;          // else (very dense tile) tokens go straight to global memory
        if (active) {
            const uint32_t dcp = ((uint32_t)img * a.fo.tiles + tile) * 2;
            bool ok;
            // ONE copy of the walk: the token destination (shared-memory tile buffer, or global memory for a very
            // dense tile) is a CTA-uniform generic pointer -- two specialised copies cost 1.6 % (instruction footprint)
            constexpr bool CHK = FMT == DMMT_RGB_F32_NORM;
            if constexpr (K1_WALK2) {
                if (in_smem) ok = p420_walk<CHK, true>(s_ctok, true, my_off, qv[0], slot, m, k, comp, mlo, mhi, s_dc, s_stage, s_hist, s_flut, a.fo.dcpos + dcp, __float_as_uint(a.neg_zero));
                else ok = p420_walk_cold<CHK>(g_tok, fits, my_off, qv[0], slot, m, k, comp, mlo, mhi, s_dc, s_stage, s_hist, s_flut, a.fo.dcpos + dcp, __float_as_uint(a.neg_zero));
            } else {
                ok = p420_tokenize_block(in_smem ? s_ctok : g_tok, in_smem || fits, my_off, qv[0], slot, m, k, comp, mlo, mhi, s_dc, s_stage, s_hist, a.fo.dcpos + dcp);
            }
            if (!ok) atomicCAS(&a.meta[img].error, 0, DMMT_E_RANGE);
        }
#if K1_BULK
        // the token stores of this thread (generic proxy) must be visible to the bulk-copy engine (async proxy)
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
#endif
        __syncthreads();
        if (in_smem && fits) {  // compact tile -> one coalesced run (tile_cap is a multiple of 8 words)
#if K1_BULK
            // ONE bulk copy (TMA engine, UBLKCP) moves the tile's tokens from shared to global memory: no LDS / STG loop,
            // the three warps go on with the histogram flush; the issuing thread only waits until the engine has READ
            // the shared buffer (the CTA may then retire), not for the global write
            if (threadIdx.x == 32) {
                const uint32_t bytes = ((total + 3u) / 4u) * 16u;
                if (bytes) {
                    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(g_tok), "r"(smem_u32(s_ctok)),
                                 "r"(bytes)
                                 : "memory");
                    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
                }
            }
#else
            uint4* dst = reinterpret_cast<uint4*>(g_tok);
            const uint4* src = reinterpret_cast<const uint4*>(s_ctok);
            for (uint32_t i = threadIdx.x; i < (total + 3) / 4; i += P420_THREADS) dst[i] = src[i];
#endif
        }
        if (threadIdx.x == 0) {
            a.fo.ntok[(size_t)img * a.fo.tiles + tile] = fits ? total : 0u;
            if (!fits) atomicCAS(&a.meta[img].error, 0, DMMT_E_OVERFLOW);
            short* ld = a.fo.last_dc + ((size_t)img * a.fo.tiles + tile) * 4;
            const int lm = (mcus_here - 1) * BPM;
            ld[0] = s_dc[lm + 3], ld[1] = s_dc[lm + 4], ld[2] = s_dc[lm + 5], ld[3] = 0;
        }
        unsigned int* gh = a.hist + (size_t)img * 1024;
#if K1_BULK && K1_WALK2 && K1_BRED
        // the tile's symbol counts are in plain symbol order: four bulk reductions (add.u32, performed at L2) fold them
        // into the image's [4][256] histogram -- Y-DC, Y-AC, C-DC, C-AC -- without a single per-bin instruction
        if (threadIdx.x == 64) {
            const uint32_t hs = smem_u32(s_hist);
            asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.u32 [%0], [%1], 64;" ::"l"(gh + T_YDC * 256), "r"(hs + 4 * SH_YDC) : "memory");
            asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.u32 [%0], [%1], 64;" ::"l"(gh + T_CDC * 256), "r"(hs + 4 * SH_CDC) : "memory");
            asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.u32 [%0], [%1], 1024;" ::"l"(gh + T_YAC * 256), "r"(hs + 4 * SH_YAC) : "memory");
            asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.u32 [%0], [%1], 1024;" ::"l"(gh + T_CAC * 256), "r"(hs + 4 * SH_CAC) : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        }
#else
        for (int q4 = threadIdx.x; q4 < SH_BINS / 4; q4 += P420_THREADS) {  // 4 bins per 128-bit load; most bins of a tile are empty
            const uint4 v = reinterpret_cast<const uint4*>(s_hist)[q4];
            if (v.x | v.y | v.z | v.w) {
                // tile layout -> [4][256] of the image: Y-DC 0.., C-DC 512.., Y-AC 256.., C-AC 768..
                const int b = 4 * q4;
                int g0 = b;
                if (b >= SH_CAC) g0 = T_CAC * 256 + (b - SH_CAC);
                else if (b >= SH_YAC) g0 = T_YAC * 256 + (b - SH_YAC);
                else if (b >= SH_CDC) g0 = T_CDC * 256 + (b - SH_CDC);
                unsigned int* g = gh + g0;
                if (v.x) atomicAdd(g, v.x);
                if (v.y) atomicAdd(g + 1, v.y);
                if (v.z) atomicAdd(g + 2, v.z);
                if (v.w) atomicAdd(g + 3, v.w);
            }
        }
#endif
    }
}

// K2 on the fused path: finishes the DC tokens whose predictor is the last DC of the previous tile
// (tile 0: the seeds, 0 for a whole image -- categorize.rs:157) and counts their symbols.
__global__ void k2_fix_dc(TileTok fo, unsigned int* hist, ImgMeta* meta, const int16_t* seed_dc) {
    const int img = blockIdx.y;
    const uint32_t tile = blockIdx.x * blockDim.x + threadIdx.x;
    if (tile >= fo.tiles) return;
    const size_t ti = (size_t)img * fo.tiles + tile;
    if (fo.ntok[ti] == 0) return;  // overflowed tile (error already flagged)
    uint32_t* tok = fo.tok + (size_t)img * fo.img_stride_words + (size_t)tile * fo.tile_cap;
#pragma unroll
    for (int c = 0; c < 3; c++) {
        const int pred = tile ? (int)fo.last_dc[(ti - 1) * 4 + c] : (seed_dc ? (int)seed_dc[c] : 0);
        const uint32_t pos = c == 0 ? 0u : fo.dcpos[ti * 2 + (c - 1)];
        const int dc = (int)(short)(tok[pos] >> 16);
        const int diff = (int)(short)(dc - pred);
        int cat = 0;
        uint32_t bits = 0;
        if (diff != 0) k1_cat_bits(diff, cat, bits);
        if (cat > 15) atomicCAS(&meta[img].error, 0, DMMT_E_RANGE);
        const int tdc = c ? T_CDC : T_YDC;
        tok[pos] = k1_token(tdc, cat & 15, 0, bits);
        atomicAdd(&hist[(size_t)img * 1024 + tdc * 256 + (cat & 15)], 1u);
    }
}

template <int HR, int VR, int FMT>
cudaError_t launch_fmt(const K1Args& a, dim3 grid, bool dbg, bool exact, cudaStream_t st) {
    if constexpr (HR == 2 && VR == 2) {
        if (!dbg && !exact && !a.force_scalar) {
            // shared memory is what bounds residency, L1 is barely used (streaming loads): take the largest carve-out
            // (function attributes are per device: a process that drives several devices sets them on each)
            static bool carved[64] = {};
            int dev = 0;
            if (cudaGetDevice(&dev) == cudaSuccess && dev >= 0 && dev < 64 && !carved[dev]) {
                (void)cudaFuncSetAttribute(k1_transform_p420<FMT, true, false>, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
                (void)cudaFuncSetAttribute(k1_transform_p420<FMT, true, true>, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
                carved[dev] = true;
            }
            if (a.fo.tok) {
                if (a.vec_ok) k1_transform_p420<FMT, true, true><<<grid, P420_THREADS, 0, st>>>(a);
                else k1_transform_p420<FMT, true, false><<<grid, P420_THREADS, 0, st>>>(a);
            } else {
                if (a.vec_ok) k1_transform_p420<FMT, false, true><<<grid, P420_THREADS, 0, st>>>(a);
                else k1_transform_p420<FMT, false, false><<<grid, P420_THREADS, 0, st>>>(a);
            }
            return cudaGetLastError();
        }
    }
    if (dbg) {
        if (exact) k1_transform<HR, VR, FMT, true, true><<<grid, K1_THREADS, 0, st>>>(a);
        else k1_transform<HR, VR, FMT, true, false><<<grid, K1_THREADS, 0, st>>>(a);
    } else {
        if (exact) k1_transform<HR, VR, FMT, false, true><<<grid, K1_THREADS, 0, st>>>(a);
        else k1_transform<HR, VR, FMT, false, false><<<grid, K1_THREADS, 0, st>>>(a);
    }
    return cudaGetLastError();
}

template <int HR, int VR>
cudaError_t launch_sub(const K1Args& a, dim3 grid, int fmt, bool dbg, bool exact, cudaStream_t st) {
    switch (fmt) {
        case DMMT_RGB_U8: return launch_fmt<HR, VR, DMMT_RGB_U8>(a, grid, dbg, exact, st);
        case DMMT_RGB_U16: return launch_fmt<HR, VR, DMMT_RGB_U16>(a, grid, dbg, exact, st);
        default: return launch_fmt<HR, VR, DMMT_RGB_F32_NORM>(a, grid, dbg, exact, st);
    }
}

}  // namespace

// The fused transform+tokenise kernel exists for 4:2:0 with the proven fast divisions.
bool k1_fused_supported(const Geom& g, const K1Consts& c) {
    static const bool scalar = [] { const char* e = getenv("DMMT_K1_SCALAR"); return e && e[0] == '1'; }();
    return g.hr == 2 && g.vr == 2 && !c.exact && !scalar;
}
uint32_t k1_tiles_x(const Geom& g) { return (uint32_t)((g.mcus_x * 8 * g.hr + TILE_W - 1) / TILE_W); }

cudaError_t launch_k2_fix_dc(const TileTok& fo, int n, unsigned int* hist, ImgMeta* meta, const int16_t* seed_dc,
                             cudaStream_t st) {
    k2_fix_dc<<<dim3((fo.tiles + 127) / 128, n), 128, 0, st>>>(fo, hist, meta, seed_dc);
    return cudaGetLastError();
}

// Host launcher.  n_images equally sized images; dbg != nullptr selects the variant that also
// writes the pre-quantisation coefficients of image 0.
cudaError_t launch_k1(const Geom& g, int fmt, const K1Consts& c, int check_max, const void* d_pixels,
                      size_t img_stride_bytes, int n_images, int16_t* d_coef, size_t coef_img_stride,
                      float* d_dbg, ImgMeta* meta, const TileTok* fused, unsigned int* hist, cudaStream_t st) {
    K1Args a;
    a.fo = TileTok{};
    a.hist = hist;
    if (fused && k1_fused_supported(g, c) && !d_dbg) a.fo = *fused;
    a.pixels = static_cast<const uint8_t*>(d_pixels);
    a.img_stride_bytes = img_stride_bytes;
    a.W = g.W;
    a.H = g.H;
    a.mcus_x = g.mcus_x;
    a.maxf = c.maxf;
    a.r_hi = c.r_hi;
    a.r_lo = c.r_lo;
    a.neg_zero = -0.0f;
    const size_t pb = (fmt == DMMT_RGB_U8) ? 3 : (fmt == DMMT_RGB_U16 ? 6 : 12);
    a.vec_ok = ((reinterpret_cast<uintptr_t>(d_pixels) & 15) == 0) && (((size_t)g.W * pb) % 16 == 0) &&
               (img_stride_bytes % 16 == 0);
    a.coef = d_coef;
    a.coef_img_stride = coef_img_stride;
    a.dbg = d_dbg;
    a.check_max = check_max;
    const uint32_t mv = (uint32_t)c.maxf;
    a.max_rep = fmt == DMMT_RGB_U8 ? (mv & 0xFFu) * 0x01010101u : (mv & 0xFFFFu) * 0x00010001u;
    a.meta = meta;
    a.qf = c.qf;
    a.rq_hi = c.rq_hi;
    a.rq_lo = c.rq_lo;
    {
        static const bool scalar = [] { const char* e = getenv("DMMT_K1_SCALAR"); return e && e[0] == '1'; }();
        a.force_scalar = scalar ? 1 : 0;
    }
    const int pw = g.mcus_x * 8 * g.hr;
    dim3 grid((pw + TILE_W - 1) / TILE_W, g.mcus_y, n_images);
    const bool dbg = d_dbg != nullptr;
    const bool exact = c.exact != 0;
    if (g.hr == 2 && g.vr == 2) return launch_sub<2, 2>(a, grid, fmt, dbg, exact, st);
    if (g.hr == 2 && g.vr == 1) return launch_sub<2, 1>(a, grid, fmt, dbg, exact, st);
    return launch_sub<1, 1>(a, grid, fmt, dbg, exact, st);
}

// Builds the per-plan constants and PROVES the fast normalisation for this max value: for every
// v in [0, max] fma(v, r_hi, v * r_lo) must equal the IEEE quotient v / max (color.rs:45-53);
// otherwise the plan uses the exact-division kernels.
void make_k1_consts(int fmt, int max_value, const uint8_t* q_luma, const uint8_t* q_chroma, K1Consts* c) {
    c->maxf = (float)max_value;
    const double r = 1.0 / (double)max_value;
    c->r_hi = (float)r;
    c->r_lo = (float)(r - (double)c->r_hi);
    c->exact = 0;
    if (fmt != DMMT_RGB_F32_NORM) {
        const int top = fmt == DMMT_RGB_U8 ? 255 : 65535;
        for (int v = 0; v <= top && v <= max_value; v++) {
            const volatile float ref = (float)v / c->maxf;
            const volatile float fast = fmaf((float)v, c->r_hi, (float)v * c->r_lo);
            if (ref != fast) {
                c->exact = 1;
                break;
            }
        }
    }
    for (int t = 0; t < 2; t++)
        for (int i = 0; i < 64; i++) {
            const int q = t ? q_chroma[i] : q_luma[i];
            c->qf.q[t][i] = (float)q;  // `q as f32` (quantizer.rs:60)
            const double rq = 1.0 / (double)(q ? q : 1);
            c->rq_hi.q[t][i] = (float)rq;
            c->rq_lo.q[t][i] = (float)(rq - (double)c->rq_hi.q[t][i]);
            if (q == 0) c->exact = 1;  // no preset has a zero divisor; IEEE division handles it like the reference
        }
}

}  // namespace dmmt
