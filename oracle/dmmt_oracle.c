/*
 * dmmt_oracle.c -- CPU restatement of the dmmt-jpeg-encoder encode hot path.
 * TEST INFRASTRUCTURE ONLY (see dmmt_oracle.h).  Build: -O2 -ffp-contract=off -fno-fast-math
 * (Rust never contracts a*b+c to FMA and never reassociates float ops).
 *
 * Citations are file:line relative to /root/reference.
 */
#include "dmmt_oracle.h"

#include <math.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

static double now_s(void) {
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

/* ---------------------------------------------------------------- colour (src/color.rs) */

/* color.rs:45-53  `value.red as f32 / value.max as f32` */
float orc_normalize(uint16_t v, uint16_t max) { return (float)v / (float)max; }

/* color.rs:75-100.  Strict left-to-right f32; the level shift 128/255 is folded in f32. */
void orc_rgb_to_ycbcr(const float rgb[3], float out[3]) {
    const float red = rgb[0], green = rgb[1], blue = rgb[2];
    float wr = red * 0.299f;
    float wg = green * 0.587f;
    float wb = blue * 0.114f;
    out[0] = (wr + wg + wb - 128.0f / 255.0f) * 255.0f;
    wr = red * -0.1687f;
    wg = green * -0.3312f;
    wb = blue * 0.5f;
    out[1] = (wr + wg + wb) * 255.0f;
    wr = red * 0.5f;
    wg = green * -0.4186f;
    wb = blue * -0.0813f;
    out[2] = (wr + wg + wb) * 255.0f;
}

/* ------------------------------------------------- DCT (src/cosine_transform/arai.rs) */

/* arai.rs:7-26.  Decimal literals exactly as written there (S0 != S4 on purpose). */
#define A1 0.70710678118654752440f /* std::f32::consts::FRAC_1_SQRT_2 = 0x3F3504F3 */
#define A2 0.5411961f
#define A3 A1
#define A4 1.3065629f
#define A5 0.3826834f
#define S0 0.3535533f
#define S1 0.2548978f
#define S2 0.27059805f
#define S3 0.30067244f
#define S4 0.35355338f
#define S5 0.4499881f
#define S6 0.6532815f
#define S7 1.2814577f

/* arai.rs:29-92 */
void orc_fast_arai(float *p, size_t s) {
    const float v00 = p[0], v01 = p[s], v02 = p[2 * s], v03 = p[3 * s];
    const float v04 = p[4 * s], v05 = p[5 * s], v06 = p[6 * s], v07 = p[7 * s];

    const float v10 = v00 + v07, v11 = v01 + v06, v12 = v02 + v05, v13 = v03 + v04;
    const float v14 = v03 - v04, v15 = v02 - v05, v16 = v01 - v06, v17 = v00 - v07;

    const float v20 = v10 + v13, v21 = v11 + v12, v22 = v11 - v12, v23 = v10 - v13;
    const float v24 = -v14 - v15, v25 = v15 + v16, v26 = v16 + v17;

    const float v30 = v20 + v21, v31 = v20 - v21, v32 = v22 + v23;

    const float v42 = v32 * A1;
    const float v44 = -v24 * A2 - (v24 + v26) * A5;
    const float v45 = v25 * A3;
    const float v46 = v26 * A4 - (v26 + v24) * A5;

    const float v52 = v42 + v23, v53 = v23 - v42, v55 = v45 + v17, v57 = v17 - v45;

    const float v64 = v44 + v57, v65 = v55 + v46, v66 = v55 - v46, v67 = v57 - v44;

    p[0] = v30 * S0;
    p[4 * s] = v31 * S4;
    p[2 * s] = v52 * S2;
    p[6 * s] = v53 * S6;
    p[5 * s] = v64 * S5;
    p[s] = v65 * S1;
    p[7 * s] = v66 * S7;
    p[3 * s] = v67 * S3;
}

/* arai.rs:95-104 rows (stride 1) then columns (stride 8) */
void orc_dct8x8(float *b) {
    for (int i = 0; i < 8; i++) orc_fast_arai(b + 8 * i, 1);
    for (int i = 0; i < 8; i++) orc_fast_arai(b + i, 8);
}

/* ------------------------------------------------------------------ quantisation tables */
/* Values of quantization_tables.rs:8-230 in the enum order of :233-243, each pair packed
 * as [luma 64][chroma 64], natural (row-major) order. */
static const uint8_t QT[7][2][64] = {
    /* 0 Specification (JPEG Annex K) */
    {{16, 11, 10, 16, 24, 40, 51, 61, 12, 12, 14, 19, 26, 58, 60, 55, 14, 13, 16, 24, 40, 57,
      69, 56, 14, 17, 22, 29, 51, 87, 80, 62, 18, 22, 37, 56, 68, 109, 103, 77, 24, 35, 55, 64,
      81, 104, 113, 92, 49, 64, 78, 87, 103, 121, 120, 101, 72, 92, 95, 98, 112, 100, 103, 99},
     {17, 18, 24, 47, 99, 99, 99, 99, 18, 21, 26, 66, 99, 99, 99, 99, 24, 26, 56, 99, 99, 99,
      99, 99, 47, 66, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99,
      99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99}},
    /* 1 Flat */
    {{16, 16, 16, 16, 16, 16, 16, 16, 16, 16, 16, 16, 16, 16, 16, 16, 16, 16, 16, 16, 16, 16,
      16, 16, 16, 16, 16, 16, 16, 16, 16, 16, 16, 16, 16, 16, 16, 16, 16, 16, 16, 16, 16, 16,
      16, 16, 16, 16, 16, 16, 16, 16, 16, 16, 16, 16, 16, 16, 16, 16, 16, 16, 16, 16},
     {16, 16, 16, 16, 16, 16, 16, 16, 16, 16, 16, 16, 16, 16, 16, 16, 16, 16, 16, 16, 16, 16,
      16, 16, 16, 16, 16, 16, 16, 16, 16, 16, 16, 16, 16, 16, 16, 16, 16, 16, 16, 16, 16, 16,
      16, 16, 16, 16, 16, 16, 16, 16, 16, 16, 16, 16, 16, 16, 16, 16, 16, 16, 16, 16}},
    /* 2 MSSIM-Kodak-Tuned */
    {{12, 17, 20, 21, 30, 34, 56, 63, 18, 20, 20, 26, 28, 51, 61, 55, 19, 20, 21, 26, 33, 58,
      69, 55, 26, 26, 26, 30, 46, 87, 86, 66, 31, 33, 36, 40, 46, 96, 100, 73, 40, 35, 46, 62,
      81, 100, 111, 91, 46, 66, 76, 86, 102, 121, 120, 101, 68, 90, 90, 96, 113, 102, 105, 103},
     {8,  12, 15, 15, 86, 96, 96, 98, 13, 13, 15, 26, 90, 96, 99, 98, 12, 15, 18, 96, 99, 99,
      99, 99, 17, 16, 90, 96, 99, 99, 99, 99, 96, 96, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99,
      99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99}},
    /* 3 PSNR-HVS-N-Kodak-Tuned */
    {{9,  10, 12, 14, 27, 32,  51,  62,  11, 12, 14, 19, 27, 44,  59,  73,
      12, 14, 18, 25, 42, 59,  79,  78,  17, 18, 25, 42, 61, 92,  87,  92,
      23, 28, 42, 75, 79, 112, 112, 99,  40, 42, 59, 84, 88, 124, 132, 111,
      42, 64, 78, 95, 105, 126, 125, 99, 70, 75, 100, 102, 116, 100, 107, 98},
     {9,  10,  17, 19, 62, 89, 91, 97, 12, 13, 18, 29, 84, 91, 88, 98,
      14, 19,  29, 93, 95, 95, 98, 97, 20, 26, 84, 88, 95, 95, 98, 94,
      26, 86,  91, 93, 97, 99, 98, 99, 99, 100, 98, 99, 99, 99, 99, 99,
      99, 99,  99, 99, 99, 99, 99, 99, 97, 97, 99, 99, 99, 99, 97, 99}},
    /* 4 DCTune-Perceptual-Optimization */
    {{7,  8,  10, 14, 23,  44,  95,  241, 8,  8,  11, 15,  25,  47,  102, 255,
      10, 11, 13, 19, 31,  58,  127, 255, 14, 15, 19, 27,  44,  83,  181, 255,
      23, 25, 31, 44, 72,  136, 255, 255, 44, 47, 58, 83,  136, 255, 255, 255,
      95, 102, 127, 181, 255, 255, 255, 255, 241, 255, 255, 255, 255, 255, 255, 255},
     {7,  8,  10, 14, 23,  44,  95,  241, 8,  8,  11, 15,  25,  47,  102, 255,
      10, 11, 13, 19, 31,  58,  127, 255, 14, 15, 19, 27,  44,  83,  181, 255,
      23, 25, 31, 44, 72,  136, 255, 255, 44, 47, 58, 83,  136, 255, 255, 255,
      95, 102, 127, 181, 255, 255, 255, 255, 241, 255, 255, 255, 255, 255, 255, 255}},
    /* 5 A-visual-detection-model */
    {{15, 11, 11, 12, 15, 19, 25, 32, 11, 13, 10, 10, 12, 15, 19, 24, 11, 10, 14, 14, 16, 18,
      22, 27, 12, 10, 14, 18, 21, 24, 28, 33, 15, 12, 16, 21, 26, 31, 36, 42, 19, 15, 18, 24,
      31, 38, 45, 53, 25, 19, 22, 28, 36, 45, 55, 65, 32, 24, 27, 33, 42, 53, 65, 77},
     {15, 11, 11, 12, 15, 19, 25, 32, 11, 13, 10, 10, 12, 15, 19, 24, 11, 10, 14, 14, 16, 18,
      22, 27, 12, 10, 14, 18, 21, 24, 28, 33, 15, 12, 16, 21, 26, 31, 36, 42, 19, 15, 18, 24,
      31, 38, 45, 53, 25, 19, 22, 28, 36, 45, 55, 65, 32, 24, 27, 33, 42, 53, 65, 77}},
    /* 6 An-improved-detection-model */
    {{14, 10, 11, 14, 19, 25, 34, 45, 10, 11, 11, 12, 15, 20, 26, 33, 11, 11, 15, 18, 21, 25,
      31, 38, 14, 12, 18, 24, 28, 33, 39, 47, 19, 15, 21, 28, 36, 43, 51, 59, 25, 20, 25, 33,
      43, 54, 64, 74, 34, 26, 31, 39, 51, 64, 77, 91, 45, 33, 38, 47, 59, 74, 91, 108},
     {14, 10, 11, 14, 19, 25, 34, 45, 10, 11, 11, 12, 15, 20, 26, 33, 11, 11, 15, 18, 21, 25,
      31, 38, 14, 12, 18, 24, 28, 33, 39, 47, 19, 15, 21, 28, 36, 43, 51, 59, 25, 20, 25, 33,
      43, 54, 64, 74, 34, 26, 31, 39, 51, 64, 77, 91, 45, 33, 38, 47, 59, 74, 91, 108}},
};

const uint8_t *orc_qtable(int preset, int chroma) {
    if (preset < 0 || preset > 6) return NULL;
    return QT[preset][chroma ? 1 : 0];
}

/* frequency_block.rs:1-5 */
static const uint8_t ZZ[64] = {0,  1,  8,  16, 9,  2,  3,  10, 17, 24, 32, 25, 18, 11, 4,  5,
                               12, 19, 26, 33, 40, 48, 41, 34, 27, 20, 13, 6,  7,  14, 21, 28,
                               35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23, 30, 37, 44, 51,
                               58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63};
const uint8_t *orc_zigzag(void) { return ZZ; }

/* quantizer.rs:60  `(d / q as f32).round() as i16` : IEEE divide, round half away from
 * zero, saturating cast (NaN -> 0). */
int16_t orc_quantize(float d, uint8_t q) {
    float r = roundf(d / (float)q);
    if (r != r) return 0;
    if (r >= 32767.0f) return 32767;
    if (r <= -32768.0f) return -32768;
    return (int16_t)r;
}

/* ------------------------------------------------------------- categorize.rs:21-63 */
int orc_categorize(int16_t v, uint16_t *pattern) {
    if (v == 0) { /* categorize.rs:48-53,66-73 */
        if (pattern) *pattern = 0;
        return 0;
    }
    unsigned a = (unsigned)(v < 0 ? -(int)v : (int)v); /* unsigned_abs */
    int cat = 0;
    while ((a >> cat) != 0) cat++; /* 16 - leading_zeros(u16) */
    if (cat > 15) return -1;       /* categorize.rs:25-30 panics (only v = -32768) */
    unsigned pat = (v > 0) ? (unsigned)v : ((1u << cat) - 1u - a); /* :34-41 */
    if (pattern) *pattern = (uint16_t)(pat << (16 - cat));         /* :43-46 left align */
    return cat;
}

/* categorize.rs:132-151 */
int orc_rle_tokens(const int16_t *seq, int n, uint8_t *out_zeros, int16_t *out_value) {
    int count = 0, zeros = 0;
    for (int i = 0; i < n; i++) {
        if (seq[i] == 0) {
            zeros++;
        } else {
            while (zeros > 15) {
                out_zeros[count] = 15;
                out_value[count++] = 0;
                zeros -= 16;
            }
            out_zeros[count] = (uint8_t)zeros;
            out_value[count++] = seq[i];
            zeros = 0;
        }
    }
    if (zeros != 0) {
        out_zeros[count] = 0;
        out_value[count++] = 0;
    }
    return count;
}

/* subsampling.rs:32-46 rates ; padder.rs:13-14 ; transformer.rs:48-49 */
static void preset_rates(int preset, int *hr, int *vr) {
    *hr = (preset == ORC_P444) ? 1 : 2;
    *vr = (preset == ORC_P420) ? 2 : 1;
}
void orc_padded_dims(int w, int h, int preset, int *pw, int *ph) {
    int hr, vr;
    preset_rates(preset, &hr, &vr);
    int mw = 8 * hr, mh = 8 * vr;
    *pw = (w + mw - 1) / mw * mw;
    *ph = (h + mh - 1) / mh * mh;
}

/* block_entangler.rs:69-77 */
size_t orc_quadfold_index(size_t i, size_t line_length) {
    size_t on_quad = i / 4;
    size_t line = (i % 4) / 2;
    return i - (on_quad + line) * 2 + line_length * line;
}

/* ------------------------------------------------ Huffman: length_limited.rs:37-134 */
typedef struct {
    uint64_t f;
    uint8_t pkg; /* NodeKind: Leaf(0) < Package(1), length_limited.rs:22-26 */
} pm_node;

int orc_package_merge(const uint64_t *freqs, int n, int limit, int *lengths) {
    if (n <= 0) return -1; /* `code_length - 1` underflows (length_limited.rs:80) */
    if (limit < 63 && (uint64_t)n > (1ull << limit)) return -2; /* :43-49 */
    for (int i = 1; i < n; i++)
        if (freqs[i] < freqs[i - 1]) return -3; /* :39-42 */
    /* calculate_packages :63-73 : list[0] = leaves; list[k] = sorted(pairs(list[k-1]) U leaves) */
    size_t cap = (size_t)2 * n + 2;
    pm_node *lists = (pm_node *)malloc(sizeof(pm_node) * cap * (size_t)limit);
    int *llen = (int *)malloc(sizeof(int) * (size_t)limit);
    for (int i = 0; i < n; i++) {
        lists[i].f = freqs[i];
        lists[i].pkg = 0;
    }
    llen[0] = n;
    for (int k = 1; k < limit; k++) {
        const pm_node *prev = lists + (size_t)(k - 1) * cap;
        pm_node *cur = lists + (size_t)k * cap;
        int np = llen[k - 1] / 2; /* chunks_exact(2) drops an odd tail (:104-109) */
        /* BinaryHeap::into_sorted_vec over (frequency, kind): merge, leaf first on ties */
        int a = 0, b = 0, o = 0;
        while (a < n || b < np) {
            uint64_t pf = (b < np) ? prev[2 * b].f + prev[2 * b + 1].f : 0;
            if (b >= np || (a < n && freqs[a] <= pf)) {
                cur[o].f = freqs[a++];
                cur[o++].pkg = 0;
            } else {
                cur[o].f = pf;
                cur[o++].pkg = 1;
                b++;
            }
        }
        llen[k] = o;
    }
    /* calculate_solution :75-89,117-133 ; sum_up_codeword_lengths :91-102 */
    for (int i = 0; i < n; i++) lengths[i] = 0;
    size_t npk = (size_t)n - 1;
    int rc = 0;
    for (int k = limit - 1; k >= 0; k--) {
        size_t count = npk * 2;
        if (count > (size_t)llen[k]) { /* slice index panic in the reference */
            rc = -4;
            break;
        }
        const pm_node *cur = lists + (size_t)k * cap;
        size_t leaves = 0, pk = 0;
        for (size_t i = 0; i < count; i++) {
            if (cur[i].pkg) pk++;
            else leaves++;
        }
        for (size_t i = 0; i < leaves; i++) lengths[i] += 1;
        npk = pk;
    }
    free(lists);
    free(llen);
    return rc;
}

/* symbol_counting.rs:25-32 (ascending symbol, drop zeros), :92-94 (stable sort by freq),
 * :85-90 (limit, then symlens[0].length += 1) */
int orc_build_table(const uint64_t *hist, int nsym, int limit, int plus_one, uint8_t *symbols,
                    int *lengths) {
    uint64_t f[256];
    int n = 0;
    for (int s = 0; s < nsym; s++)
        if (hist[s] > 0) {
            symbols[n] = (uint8_t)s;
            f[n++] = hist[s];
        }
    /* stable insertion sort by frequency */
    for (int i = 1; i < n; i++) {
        uint64_t kf = f[i];
        uint8_t ks = symbols[i];
        int j = i - 1;
        while (j >= 0 && f[j] > kf) {
            f[j + 1] = f[j];
            symbols[j + 1] = symbols[j];
            j--;
        }
        f[j + 1] = kf;
        symbols[j + 1] = ks;
    }
    if (n == 0) return 0;
    int rc = orc_package_merge(f, n, limit, lengths);
    if (rc < 0) return rc;
    if (plus_one) lengths[0] += 1;
    return n;
}

/* huffman/encoder.rs:37-157 */
int orc_canonical_codes(const uint8_t *symbols, const int *lengths, int n, uint16_t *code_lut,
                        uint8_t *len_lut) {
    memset(code_lut, 0, 256 * sizeof(uint16_t));
    memset(len_lut, 0, 256);
    if (n == 0) return -1;   /* :74-76 */
    if (n > 255) return -2;  /* :78-80 */
    for (int i = 1; i < n; i++)
        if (lengths[i] > lengths[i - 1]) return -3; /* :82-84 descending length */
    if (lengths[0] > 16) return -4;                 /* :86-92 */
    uint32_t code = 0;
    for (int i = n - 1; i >= 0; i--) {
        uint8_t s = symbols[i];
        if (s == 255) return -6; /* LUT has Symbol::MAX = 255 entries (:33,152): index OOB */
        if (i != n - 1) {
            if (len_lut[s]) return -5; /* :129-136 duplicate symbol */
            code += 1u << (16 - lengths[i + 1]); /* :116-119 (prev = element i+1) */
            if (code > 0xFFFF) return -7;        /* u16 add overflow (debug panic) */
        }
        code_lut[s] = (uint16_t)code;
        len_lut[s] = (uint8_t)lengths[i];
    }
    return 0;
}

/* ------------------------------------------------------ binary_stream.rs / injector */
static void bw_emit(orc_bitwriter *bw, uint8_t b) {
    if (bw->len + 2 > bw->cap) {
        bw->cap = bw->cap ? bw->cap * 2 : 4096;
        bw->data = (uint8_t *)realloc(bw->data, bw->cap);
    }
    bw->data[bw->len++] = b;
    if (bw->stuff && b == 0xFF) bw->data[bw->len++] = 0x00; /* segment_marker_injector.rs:22-28 */
}
void orc_bw_init(orc_bitwriter *bw, int ones, int stuff) {
    memset(bw, 0, sizeof *bw);
    bw->init_val = ones ? 0xFF : 0x00; /* binary_stream.rs:20 */
    bw->buffer = bw->init_val;
    bw->stuff = stuff;
}
/* binary_stream.rs:38-66 (aligned whole-byte fast path, then bit-by-bit MSB first) */
void orc_bw_write_bits(orc_bitwriter *bw, const uint8_t *buf, size_t count) {
    size_t off = 0;
    if (bw->used == 0) {
        size_t q = count / 8;
        for (size_t i = 0; i < q; i++) bw_emit(bw, buf[i]);
        off = q * 8;
    }
    for (size_t bi = off; bi < count; bi++) {
        int bit = (buf[bi / 8] >> (7 - (bi % 8))) & 1;
        uint8_t m = (uint8_t)(0x80u >> bw->used);
        if (bit) bw->buffer |= m;
        else bw->buffer &= (uint8_t)~m;
        if (++bw->used == 8) {
            bw_emit(bw, bw->buffer);
            bw->used = 0;
            bw->buffer = bw->init_val;
        }
    }
}
/* binary_stream.rs:89-96 */
void orc_bw_flush(orc_bitwriter *bw) {
    if (bw->used != 0) {
        bw_emit(bw, bw->buffer);
        bw->buffer = bw->init_val;
        bw->used = 0;
    }
}
void orc_bw_free(orc_bitwriter *bw) {
    free(bw->data);
    memset(bw, 0, sizeof *bw);
}
size_t orc_stuff_bytes(const uint8_t *in, size_t n, uint8_t *out) {
    size_t o = 0;
    for (size_t i = 0; i < n; i++) {
        out[o++] = in[i];
        if (in[i] == 0xFF) out[o++] = 0;
    }
    return o;
}

/* ------------------------------------------------------------ image/reader/ppm.rs */
/* error codes: -1 missing token, -2 token parse failed, -3 incomplete pixel, -4 size mismatch,
 * -5 sample > max (color.rs:62-65 panics), -6 not P3 */
typedef struct {
    const uint8_t *p, *end;
} tokz;
/* ppm.rs:41-78.  '#' starts a comment anywhere (even inside a token, which then continues
 * after the newline); whitespace = u8::is_ascii_whitespace (space \t \n \x0C \r). */
static int next_token(tokz *t, char *buf, size_t cap) {
    size_t n = 0;
    int in_comment = 0;
    while (t->p < t->end) {
        uint8_t c = *t->p++;
        if (in_comment) {
            if (c == '\n') in_comment = 0;
            continue;
        }
        if (c == '#') {
            in_comment = 1;
            continue;
        }
        if (c == ' ' || c == '\t' || c == '\n' || c == '\x0C' || c == '\r') {
            if (n) break;
        } else if (n + 1 < cap) {
            buf[n++] = (char)c;
        } else {
            n = cap; /* over-long token: will fail to parse */
        }
    }
    if (n >= cap) {
        buf[0] = 'x';
        buf[1] = 0;
        return 1;
    }
    buf[n] = 0;
    return n > 0;
}
/* Rust `str::parse::<u16>()`: optional '+', decimal digits, no overflow */
static int parse_u16(const char *s, int *out) {
    if (*s == '+') s++;
    if (!*s) return 0;
    long v = 0;
    for (; *s; s++) {
        if (*s < '0' || *s > '9') return 0;
        v = v * 10 + (*s - '0');
        if (v > 65535) return 0;
    }
    *out = (int)v;
    return 1;
}
int orc_parse_ppm(const uint8_t *text, size_t n, int *w, int *h, int *max, uint16_t **samples) {
    tokz t = {text, text + n};
    char tok[64];
    *samples = NULL;
    if (!next_token(&t, tok, sizeof tok)) return -1;
    if (strcmp(tok, "P3") != 0) return -6; /* ppm.rs:177-184 */
    if (!next_token(&t, tok, sizeof tok)) return -1;
    if (!parse_u16(tok, w)) return -2;
    if (!next_token(&t, tok, sizeof tok)) return -1;
    if (!parse_u16(tok, h)) return -2;
    if (!next_token(&t, tok, sizeof tok)) return -1;
    if (!parse_u16(tok, max)) return -2;
    size_t cap = (size_t)(*w) * (size_t)(*h) * 3 + 3, cnt = 0;
    uint16_t *buf = (uint16_t *)malloc(sizeof(uint16_t) * (cap ? cap : 3));
    while (next_token(&t, tok, sizeof tok)) { /* ppm.rs:224-237 */
        int v;
        if (!parse_u16(tok, &v)) {
            free(buf);
            return -2;
        }
        if (cnt == cap) {
            cap = cap * 2 + 3;
            buf = (uint16_t *)realloc(buf, sizeof(uint16_t) * cap);
        }
        buf[cnt++] = (uint16_t)v;
    }
    if (cnt % 3 != 0) { /* ppm.rs:239-244 */
        free(buf);
        return -3;
    }
    if (cnt / 3 != (size_t)(*w) * (size_t)(*h)) { /* ppm.rs:165-175 */
        free(buf);
        return -4;
    }
    for (size_t i = 0; i < cnt; i++)
        if (buf[i] > *max) { /* color.rs:62-65 */
            free(buf);
            return -5;
        }
    *samples = buf;
    return 0;
}

/* ------------------------------------------------------------------- whole path */

typedef struct {
    float *base;
    size_t first_block, n_blocks;
} dct_job;
static void *dct_worker(void *arg) {
    dct_job *j = (dct_job *)arg;
    for (size_t b = 0; b < j->n_blocks; b++) orc_dct8x8(j->base + 64 * (j->first_block + b));
    return NULL;
}
/* cosine_transform.rs:55-73 + transformer.rs:126-148: 700-block jobs on a pool.  The pool
 * only schedules; results do not depend on it.  Here: nthreads workers, jobs dealt
 * round-robin (chunk 700) so the fan-out has the reference's granularity. */
typedef struct {
    float *base;
    size_t n_blocks;
    int tid, nthreads;
} dct_pool_arg;
static void *dct_pool_worker(void *arg) {
    dct_pool_arg *a = (dct_pool_arg *)arg;
    const size_t chunk = 700;
    size_t njobs = (a->n_blocks + chunk - 1) / chunk;
    for (size_t j = (size_t)a->tid; j < njobs; j += (size_t)a->nthreads) {
        dct_job job = {a->base, j * chunk,
                       (j * chunk + chunk <= a->n_blocks) ? chunk : a->n_blocks - j * chunk};
        dct_worker(&job);
    }
    return NULL;
}
static void dct_channel(float *base, size_t n_blocks, int nthreads) {
    if (nthreads <= 1) {
        dct_job job = {base, 0, n_blocks};
        dct_worker(&job);
        return;
    }
    pthread_t th[256];
    dct_pool_arg args[256];
    if (nthreads > 256) nthreads = 256;
    for (int t = 0; t < nthreads; t++) {
        args[t] = (dct_pool_arg){base, n_blocks, t, nthreads};
        pthread_create(&th[t], NULL, dct_pool_worker, &args[t]);
    }
    for (int t = 0; t < nthreads; t++) pthread_join(th[t], NULL);
}

static inline void load_norm_rgb(const void *pixels, int fmt, size_t idx, float maxf, float rgb[3]) {
    if (fmt == ORC_FMT_U8) {
        const uint8_t *p = (const uint8_t *)pixels + 3 * idx;
        rgb[0] = (float)p[0] / maxf; /* color.rs:48-50 */
        rgb[1] = (float)p[1] / maxf;
        rgb[2] = (float)p[2] / maxf;
    } else if (fmt == ORC_FMT_U16) {
        const uint16_t *p = (const uint16_t *)pixels + 3 * idx;
        rgb[0] = (float)p[0] / maxf;
        rgb[1] = (float)p[1] / maxf;
        rgb[2] = (float)p[2] / maxf;
    } else {
        const float *p = (const float *)pixels + 3 * idx;
        rgb[0] = p[0];
        rgb[1] = p[1];
        rgb[2] = p[2];
    }
}

static void put(uint8_t **p, const void *src, size_t n) {
    memcpy(*p, src, n);
    *p += n;
}
/* jpeg/encoder.rs:137-153 : marker, u16 BE length = 2 + content, content */
static void put_segment(uint8_t **p, uint8_t marker, const uint8_t *content, size_t n) {
    uint8_t hd[4] = {0xFF, marker, (uint8_t)((n + 2) >> 8), (uint8_t)((n + 2) & 0xFF)};
    put(p, hd, 4);
    put(p, content, n);
}

/* padder.rs:18-38 : the dot at (x, y) of the padded image -- the source dot inside the original
 * width x height, RGBColorFormat::default() = black (0,0,0) to the right of it and below it. */
static inline void padded_dot(const void *pixels, int fmt, size_t x, size_t y, size_t width, size_t height,
                              float maxf, float rgb[3]) {
    rgb[0] = rgb[1] = rgb[2] = 0.0f;
    if (x < width && y < height) load_norm_rgb(pixels, fmt, y * width + x, maxf, rgb);
}

/* padder.rs:12-42 PaddedImage::new(image, pad_nearest_width, pad_nearest_height) on normalised f32 dots.
 * out (may be NULL to query the size) receives padded_w * padded_h RGB triples; returns their number. */
size_t orc_pad_image(const float *rgb, int width, int height, int nearest_w, int nearest_h, float *out,
                     int *padded_w, int *padded_h) {
    const size_t pw = ((size_t)width + (size_t)nearest_w - 1) / (size_t)nearest_w * (size_t)nearest_w; /* :13 */
    const size_t ph = ((size_t)height + (size_t)nearest_h - 1) / (size_t)nearest_h * (size_t)nearest_h; /* :14 */
    if (padded_w) *padded_w = (int)pw;
    if (padded_h) *padded_h = (int)ph;
    if (out)
        for (size_t y = 0; y < ph; y++)
            for (size_t x = 0; x < pw; x++)
                padded_dot(rgb, ORC_FMT_F32_NORM, x, y, (size_t)width, (size_t)height, 1.0f, out + 3 * (y * pw + x));
    return pw * ph;
}

/* subsampling.rs:206-236 ChannelColumnView::nth on row `sy`, column `sx` of the subsampled view
 * (row_index = sy * vr, column_index = sx * hr): Skip = the dot itself; Average = rect() -- x outer,
 * y inner, clamped to the last column / row (:108-122) -- summed left to right from 0 and divided by
 * the element count (:231-236).  Returns -1 for the reference's `None` (past the edge). */
static inline int subsample_value(const float *plane, size_t w, size_t h, int hr, int vr, int average, size_t sx,
                                  size_t sy, float *out) {
    const size_t col = sx * (size_t)hr, row = sy * (size_t)vr;
    if (row >= h || col >= w) return -1; /* :179-181, :213-215 */
    if (!average) {
        *out = plane[row * w + col];
        return 0;
    }
    float sum = 0.0f;
    for (int dx = 0; dx < hr; dx++) {
        size_t cx = col + (size_t)dx;
        if (cx > w - 1) cx = w - 1;
        for (int dy = 0; dy < vr; dy++) {
            size_t cy = row + (size_t)dy;
            if (cy > h - 1) cy = h - 1;
            sum = sum + plane[cy * w + cx];
        }
    }
    *out = sum / (float)(hr * vr);
    return 0;
}
int orc_subsample_value(const float *plane, int w, int h, int hr, int vr, int average, int sx, int sy,
                        float *out) {
    return subsample_value(plane, (size_t)w, (size_t)h, hr, vr, average, (size_t)sx, (size_t)sy, out);
}

/* subsampling.rs:136-140,150-166,286-309 subsample_to_square_structure(square): every row / column
 * of the subsampled view goes to result[square_row * (row_length * square) + square_col * square^2 +
 * y * square + x].  out holds (w / hr) * (h / vr) items; returns that number. */
size_t orc_subsample_retile(const float *plane, int w, int h, int hr, int vr, int average, int square,
                            float *out) {
    const size_t sw = (size_t)(w / hr), sh = (size_t)(h / vr), sq = (size_t)square;
    /* same mapping, walked square row by square row so that no index needs a division */
    for (size_t y0 = 0, qrow = 0; y0 < sh; y0 += sq, qrow++)
        for (size_t y = 0; y < sq && y0 + y < sh; y++)
            for (size_t x0 = 0, qcol = 0; x0 < sw; x0 += sq, qcol++) {
                float *dst = out + qrow * (sw * sq) + qcol * (sq * sq) + y * sq;
                for (size_t x = 0; x < sq && x0 + x < sw; x++)
                    subsample_value(plane, (size_t)w, (size_t)h, hr, vr, average, x0 + x, y0 + y, dst + x);
            }
    return sw * sh;
}

/* symbol_counting.rs:55-64 : one CategorizedBlock into the DC / AC counters -- the DC symbol is the
 * category of the (already differenced) DC value, every token counts (zeros << 4) | category.
 * Returns <0 where categorize.rs panics (-32768). */
int orc_count_block(int16_t dc_value, const uint8_t *tok_zeros, const int16_t *tok_value, int ntok,
                    uint64_t *dc_hist, uint64_t *ac_hist) {
    int cat = orc_categorize(dc_value, NULL);
    if (cat < 0) return -1;
    dc_hist[cat]++;
    for (int t = 0; t < ntok; t++) {
        int c = orc_categorize(tok_value[t], NULL);
        if (c < 0) return -1;
        ac_hist[(tok_zeros[t] << 4) | c]++;
    }
    return 0;
}

int orc_encode(const void *pixels, int fmt, int width, int height, int max_value, int preset,
               int bits_per_channel, int qpreset, int nthreads, uint8_t **jpeg, size_t *jpeg_len,
               orc_result *res, int keep_planes) {
    if (width <= 0 || height <= 0 || width > 65535 || height > 65535) return -10;
    if (preset < 0 || preset > 2 || qpreset < 0 || qpreset > 6) return -11;
    const uint8_t *QL = QT[qpreset][0], *QC = QT[qpreset][1];
    int hr, vr, pw, ph;
    preset_rates(preset, &hr, &vr);
    orc_padded_dims(width, height, preset, &pw, &ph);
    if (pw > 65535 || ph > 65535) return -12; /* u16 overflow in padder.rs:13-14 */
    const size_t PW = (size_t)pw, PH = (size_t)ph;
    const size_t cw = PW / (size_t)hr, chh = PH / (size_t)vr;
    const size_t ybl = (PW / 8) * (PH / 8), cbl = (cw / 8) * (chh / 8);
    const float maxf = (float)(uint16_t)max_value;
    double t0 = now_s();

    /* padder.rs:12-42 (black RGB pad) + transformer.rs:61-85 (YCbCr, planar split) */
    float *py = (float *)malloc(sizeof(float) * PW * PH);
    float *pcb = (float *)malloc(sizeof(float) * PW * PH);
    float *pcr = (float *)malloc(sizeof(float) * PW * PH);
    for (size_t y = 0; y < PH; y++)
        for (size_t x = 0; x < PW; x++) {
            float rgb[3], o[3];
            padded_dot(pixels, fmt, x, y, (size_t)width, (size_t)height, maxf, rgb);
            orc_rgb_to_ycbcr(rgb, o);
            py[y * PW + x] = o[0];
            pcb[y * PW + x] = o[1];
            pcr[y * PW + x] = o[2];
        }

    /* transformer.rs:87-124 + subsampling.rs:102-122,206-236,286-309 */
    float *ty = (float *)malloc(sizeof(float) * ybl * 64);
    float *tcb = (float *)malloc(sizeof(float) * cbl * 64);
    float *tcr = (float *)malloc(sizeof(float) * cbl * 64);
    /* luma: SubsamplingConfig 1x1 / Skip (transformer.rs:87-100); chroma: the preset's rates,
     * Average unless P444 (subsampling.rs:48-54) */
    orc_subsample_retile(py, pw, ph, 1, 1, 0, 8, ty);
    orc_subsample_retile(pcb, pw, ph, hr, vr, preset != ORC_P444, 8, tcb);
    orc_subsample_retile(pcr, pw, ph, hr, vr, preset != ORC_P444, 8, tcr);
    free(py);
    free(pcb);
    free(pcr);

    /* transformer.rs:126-148 */
    dct_channel(ty, ybl, nthreads);
    dct_channel(tcr, cbl, nthreads);
    dct_channel(tcb, cbl, nthreads);

    /* quantizer.rs:53-62 (natural-order table cycled over block-contiguous data), then
     * block_entangler.rs (P420 luma quad folding), then the MCU interleave of
     * block_fold_iterator.rs:75-148 -> one stream-ordered array, zig-zag inside a block. */
    const size_t bw8 = PW / 8, cbw8 = cw / 8;
    const size_t n_mcus = cbl, ypm = (size_t)(hr * vr), bpm = ypm + 2;
    const size_t nsb = n_mcus * bpm;
    int16_t *stream = (int16_t *)malloc(sizeof(int16_t) * nsb * 64);
    for (size_t m = 0; m < n_mcus; m++) {
        size_t my = m / cbw8, mx = m % cbw8;
        for (size_t k = 0; k < bpm; k++) {
            const float *src;
            const uint8_t *Q;
            if (k < ypm) {
                size_t by, bx;
                if (preset == ORC_P420) { /* TL,TR,BL,BR (block_entangler.rs:69-77) */
                    by = 2 * my + k / 2;
                    bx = 2 * mx + k % 2;
                } else if (preset == ORC_P422) { /* raster pairs */
                    by = my;
                    bx = 2 * mx + k;
                } else {
                    by = my;
                    bx = mx;
                }
                src = ty + 64 * (by * bw8 + bx);
                Q = QL;
            } else {
                src = (k == ypm ? tcb : tcr) + 64 * m;
                Q = QC;
            }
            int16_t *dst = stream + 64 * (m * bpm + k);
            for (int i = 0; i < 64; i++) dst[i] = orc_quantize(src[ZZ[i]], Q[ZZ[i]]);
        }
    }
    double t1 = now_s();

    /* categorize.rs:153-169 (per-component DC chain in stream order) + symbol_counting.rs:55-74 */
    uint64_t hist[4][256];
    memset(hist, 0, sizeof hist);
    {
        int last_dc[3] = {0, 0, 0};
        uint8_t tz[64];
        int16_t tv[64];
        for (size_t s = 0; s < nsb; s++) {
            size_t k = s % bpm;
            int comp = k < ypm ? 0 : (k == ypm ? 1 : 2);
            int tb = comp ? 2 : 0;
            const int16_t *blk = stream + 64 * s;
            int16_t diff = (int16_t)(blk[0] - last_dc[comp]);
            last_dc[comp] = blk[0];
            int nt = orc_rle_tokens(blk + 1, 63, tz, tv);
            if (orc_count_block(diff, tz, tv, nt, hist[tb], hist[tb + 1]) < 0) return -13;
        }
    }

    /* symbol_counting.rs:85-90 ; transformer.rs:214-217 */
    uint8_t tsym[4][256];
    int tlen[4][256], tn[4];
    uint16_t code[4][256];
    uint8_t clen[4][256];
    for (int t = 0; t < 4; t++) {
        tn[t] = orc_build_table(hist[t], (t & 1) ? 256 : 16, 15, 1, tsym[t], tlen[t]);
        if (tn[t] <= 0) return -14;
        if (orc_canonical_codes(tsym[t], tlen[t], tn[t], code[t], clen[t]) < 0) return -15;
    }

    /* jpeg/encoder.rs:125-135 headers */
    size_t hdr_cap = 1024 + 4 * 300;
    uint8_t *hdr = (uint8_t *)malloc(hdr_cap), *hp = hdr;
    {
        const uint8_t soi[2] = {0xFF, 0xD8};
        put(&hp, soi, 2);
        const uint8_t app0[14] = {'J', 'F', 'I', 'F', 0, 1, 2, 0, 0, 0x48, 0, 0x48, 0, 0}; /* :211-225 */
        put_segment(&hp, 0xE0, app0, 14);
        for (int t = 0; t < 2; t++) { /* :190-209 */
            uint8_t dqt[65];
            dqt[0] = (uint8_t)t;
            for (int i = 0; i < 64; i++) dqt[1 + i] = (t ? QC : QL)[ZZ[i]];
            put_segment(&hp, 0xDB, dqt, 65);
        }
        const uint8_t sof[15] = {(uint8_t)bits_per_channel, /* :227-245 */
                                 (uint8_t)(height >> 8), (uint8_t)height, (uint8_t)(width >> 8),
                                 (uint8_t)width, 3, 1, (uint8_t)((hr << 4) | vr), 0, 2, 0x11, 1,
                                 3, 0x11, 1};
        put_segment(&hp, 0xC0, sof, 15);
        /* :183-188 order LumaAC(0x11) LumaDC(0x00) ChromaAC(0x13) ChromaDC(0x02) ; :78-84 ids */
        const int order[4] = {1, 0, 3, 2};
        const uint8_t ids[4] = {0x00, 0x11, 0x02, 0x13};
        for (int o = 0; o < 4; o++) {
            int t = order[o];
            uint8_t dht[1 + 16 + 256];
            memset(dht, 0, sizeof dht);
            dht[0] = ids[t];
            for (int i = 0; i < tn[t]; i++) dht[1 + tlen[t][i] - 1]++;                   /* :92-98 */
            for (int i = 0; i < tn[t]; i++) dht[17 + i] = tsym[t][tn[t] - 1 - i];        /* :177 rev */
            put_segment(&hp, 0xC4, dht, (size_t)17 + (size_t)tn[t]);
        }
        const uint8_t sos[10] = {3, 1, 0x01, 2, 0x23, 3, 0x23, 0, 0x3F, 0}; /* :247-262 */
        put_segment(&hp, 0xDA, sos, 10);
    }
    size_t header_bytes = (size_t)(hp - hdr);

    /* jpeg/encoder.rs:264-404 scan */
    orc_bitwriter bw;
    orc_bw_init(&bw, 1, 1);
    uint64_t scan_bits = 0;
    size_t raw_bytes = 0;
    {
        int last_dc[3] = {0, 0, 0};
        uint8_t tz[64];
        int16_t tv[64];
        for (size_t s = 0; s < nsb; s++) {
            size_t k = s % bpm;
            int comp = k < ypm ? 0 : (k == ypm ? 1 : 2);
            int tb = comp ? 2 : 0;
            const int16_t *blk = stream + 64 * s;
            int16_t diff = (int16_t)(blk[0] - last_dc[comp]);
            last_dc[comp] = blk[0];
            uint16_t pat;
            int cat = orc_categorize(diff, &pat);
            uint8_t be[2];
            /* write_symbol_and_category :376-384 : code word then category bits, each a
             * left-aligned u16 in big-endian bytes (to_bytes) with bit_len bits */
            be[0] = (uint8_t)(code[tb][cat] >> 8);
            be[1] = (uint8_t)code[tb][cat];
            orc_bw_write_bits(&bw, be, clen[tb][cat]);
            be[0] = (uint8_t)(pat >> 8);
            be[1] = (uint8_t)pat;
            orc_bw_write_bits(&bw, be, (size_t)cat);
            scan_bits += (uint64_t)clen[tb][cat] + (uint64_t)cat;
            int nt = orc_rle_tokens(blk + 1, 63, tz, tv);
            for (int t = 0; t < nt; t++) {
                int c = orc_categorize(tv[t], &pat);
                int sym = (tz[t] << 4) | c;
                be[0] = (uint8_t)(code[tb + 1][sym] >> 8);
                be[1] = (uint8_t)code[tb + 1][sym];
                orc_bw_write_bits(&bw, be, clen[tb + 1][sym]);
                be[0] = (uint8_t)(pat >> 8);
                be[1] = (uint8_t)pat;
                orc_bw_write_bits(&bw, be, (size_t)c);
                scan_bits += (uint64_t)clen[tb + 1][sym] + (uint64_t)c;
            }
        }
        orc_bw_flush(&bw);
        raw_bytes = (size_t)((scan_bits + 7) / 8);
    }

    size_t total = header_bytes + bw.len + 2;
    uint8_t *out = (uint8_t *)malloc(total), *op = out;
    put(&op, hdr, header_bytes);
    put(&op, bw.data, bw.len);
    const uint8_t eoi[2] = {0xFF, 0xD9};
    put(&op, eoi, 2);
    double t2 = now_s();

    if (res) {
        memset(res, 0, sizeof *res);
        res->width = width;
        res->height = height;
        res->padded_width = pw;
        res->padded_height = ph;
        res->hr = hr;
        res->vr = vr;
        res->y_blocks = ybl;
        res->c_blocks = cbl;
        res->n_mcus = n_mcus;
        res->n_stream_blocks = nsb;
        memcpy(res->hist, hist, sizeof hist);
        memcpy(res->table_n, tn, sizeof tn);
        memcpy(res->table_sym, tsym, sizeof tsym);
        memcpy(res->table_len, tlen, sizeof tlen);
        res->scan_bits = scan_bits;
        res->scan_bytes_unstuffed = raw_bytes;
        res->scan_bytes_stuffed = bw.len;
        res->header_bytes = header_bytes;
        res->t_transform_s = t1 - t0;
        res->t_encode_s = t2 - t1;
        if (keep_planes) {
            res->dct_y = ty;
            res->dct_cb = tcb;
            res->dct_cr = tcr;
            res->stream = stream;
            ty = tcb = tcr = NULL;
            stream = NULL;
        }
    }
    free(ty);
    free(tcb);
    free(tcr);
    free(stream);
    free(hdr);
    orc_bw_free(&bw);
    *jpeg = out;
    *jpeg_len = total;
    return 0;
}

void orc_result_free(orc_result *res) {
    if (!res) return;
    free(res->dct_y);
    free(res->dct_cb);
    free(res->dct_cr);
    free(res->stream);
    res->dct_y = res->dct_cb = res->dct_cr = NULL;
    res->stream = NULL;
}
void orc_free(void *p) { free(p); }
