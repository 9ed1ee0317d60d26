/*
 * dmmt_oracle.h -- CPU restatement of the dmmt-jpeg-encoder encode hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  This is the parity checker for the CUDA path; only
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may link or call it.  The product (dmmt_jpeg_encoder_b200/csrc) never does.
 *
 * PARITY PINNING: the reference (Rust) cannot be built in this environment (no
 * rustc/cargo) and it ships NO whole-file golden vectors, so
 * whole-file parity is "unpinned" by the reference itself.  Every stage below is
 * pinned against the reference's own unit-test known-answer vectors
 * (tests/test_oracle_kats.py) and the whole-file output is cross-checked against an
 * independent restatement's SHA-256 table (tests/golden/).
 *
 * Each function cites the reference file:line (relative to /root/reference) it follows.
 */
#ifndef DMMT_ORACLE_H
#define DMMT_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum { ORC_FMT_F32_NORM = 0, ORC_FMT_U8 = 1, ORC_FMT_U16 = 2 };
enum { ORC_P444 = 0, ORC_P422 = 1, ORC_P420 = 2 };

/* --- scalar stage primitives (KAT-tested one by one) ------------------------------- */
float orc_normalize(uint16_t v, uint16_t max);                    /* color.rs:45-53   */
void orc_rgb_to_ycbcr(const float rgb[3], float ycbcr[3]);        /* color.rs:75-100  */
void orc_fast_arai(float *p, size_t stride);                      /* arai.rs:29-92    */
void orc_dct8x8(float *block);                                    /* arai.rs:95-104   */
int16_t orc_quantize(float d, uint8_t q);                         /* quantizer.rs:60  */
const uint8_t *orc_qtable(int preset, int chroma);                /* quantization_tables.rs:286-327 */
const uint8_t *orc_zigzag(void);                                  /* frequency_block.rs:1-5 */
/* categorize.rs:21-63 : returns category; *pattern = left-aligned u16 pattern */
int orc_categorize(int16_t v, uint16_t *pattern);
/* categorize.rs:132-151 : AC run-length tokens of a 63-long zig-zag tail.
 * out_zeros/out_value sized >= 64; returns token count (EOB = (0,0), ZRL = (15,0)). */
int orc_rle_tokens(const int16_t *seq, int n, uint8_t *out_zeros, int16_t *out_value);
/* padder.rs:12-16 + transformer.rs:48-51 */
void orc_padded_dims(int w, int h, int preset, int *pw, int *ph);
/* block_entangler.rs:69-77 : position i of the 2-line buffer -> source index */
size_t orc_quadfold_index(size_t i, size_t line_length);

/* padder.rs:12-42 : pad normalised RGB dots to multiples of (nearest_w, nearest_h) with black;
 * out == NULL only reports the size.  Returns the number of dots. */
size_t orc_pad_image(const float *rgb, int width, int height, int nearest_w, int nearest_h, float *out,
                     int *padded_w, int *padded_h);
/* subsampling.rs:206-236 : value (sx, sy) of the subsampled view (average != 0: window mean with
 * border clamp); -1 = the reference's None */
int orc_subsample_value(const float *plane, int w, int h, int hr, int vr, int average, int sx, int sy,
                        float *out);
/* subsampling.rs:136-140,286-309 : subsample_to_square_structure(square) */
size_t orc_subsample_retile(const float *plane, int w, int h, int hr, int vr, int average, int square,
                            float *out);
/* symbol_counting.rs:55-64 : one categorised block (DC value, AC tokens) into the counters */
int orc_count_block(int16_t dc_value, const uint8_t *tok_zeros, const int16_t *tok_value, int ntok,
                    uint64_t *dc_hist, uint64_t *ac_hist);

/* --- Huffman (length_limited.rs, symbol_counting.rs, huffman/encoder.rs) ---------------- */
/* length_limited.rs:37-134 : package-merge; freqs ascending; returns 0 ok, <0 on the
 * reference's panics (n==0, n > 2^limit). */
int orc_package_merge(const uint64_t *sorted_freqs, int n, int limit, int *lengths);
/* symbol_counting.rs:25-32,67-70,85-94 : hist[nsym] -> (symbols,lengths) sorted by
 * ascending frequency (stable => ties by ascending symbol), package-merge(limit),
 * then lengths[0] += plus_one.  Returns n (number of symbols with freq>0). */
int orc_build_table(const uint64_t *hist, int nsym, int limit, int plus_one,
                    uint8_t *symbols, int *lengths);
/* huffman/encoder.rs:37-157 : canonical codes walking the table from its END.
 * code_lut[sym] = left-aligned u16, len_lut[sym] = length (0 = absent).
 * returns 0, or <0 where the reference panics (unsorted, len>16, n==0, n>255, dup). */
int orc_canonical_codes(const uint8_t *symbols, const int *lengths, int n,
                        uint16_t *code_lut, uint8_t *len_lut);

/* --- bit writer + stuffing (binary_stream.rs, segment_marker_injector.rs) ------------- */
typedef struct {
    uint8_t *data;
    size_t len, cap;
    uint8_t buffer, used, init_val;
    int stuff; /* 1: run every emitted byte through the 0xFF->0xFF00 injector */
} orc_bitwriter;
void orc_bw_init(orc_bitwriter *bw, int flush_with_ones, int stuff);
void orc_bw_write_bits(orc_bitwriter *bw, const uint8_t *buf, size_t count); /* binary_stream.rs:38-66 */
void orc_bw_flush(orc_bitwriter *bw);                                        /* binary_stream.rs:89-96 */
void orc_bw_free(orc_bitwriter *bw);
size_t orc_stuff_bytes(const uint8_t *in, size_t n, uint8_t *out);           /* segment_marker_injector.rs:13-30 */

/* --- P3 PPM loader (image/reader/ppm.rs) ------------------------------------------------ */
/* returns 0 ok; <0 error code mirroring error.rs variants. samples: malloc'd u16 RGB. */
int orc_parse_ppm(const uint8_t *text, size_t n, int *w, int *h, int *max, uint16_t **samples);

/* --- whole path ------------------------------------------------------------------------- */
typedef struct {
    int width, height, padded_width, padded_height, hr, vr;
    size_t y_blocks, c_blocks, n_mcus, n_stream_blocks;
    /* pre-quantisation DCT coefficients, block-contiguous, raster block order, natural
     * index 8*row+col inside a block (= reference memory layout after transformer.rs:192) */
    float *dct_y, *dct_cb, *dct_cr;
    /* quantised coefficients, zig-zag order inside a block, STREAM order of blocks
     * (MCU-interleaved: Y*hr*vr, Cb, Cr per MCU) */
    int16_t *stream;
    uint64_t hist[4][256];      /* 0:Y-DC 1:Y-AC 2:C-DC 3:C-AC (DC uses first 16) */
    int table_n[4];             /* number of symbols per table */
    uint8_t table_sym[4][256];  /* ascending-frequency order (reference's Vec<SymbolCodeLength>) */
    int table_len[4][256];
    uint64_t scan_bits;         /* unpadded entropy-coded bits */
    size_t scan_bytes_unstuffed, scan_bytes_stuffed, header_bytes;
    double t_transform_s, t_encode_s; /* wall-clock split for the CPU baseline */
} orc_result;

/* Encode one image. pixels: interleaved RGB in `fmt`. nthreads>1 fans the DCT out over
 * 700-block jobs like transformer.rs:126-148 (everything else single-threaded, as in the
 * reference).  *jpeg is malloc'd (free with orc_free).  res may be NULL; if keep_planes
 * is non-zero the intermediate arrays in res are kept (free with orc_result_free). */
int orc_encode(const void *pixels, int fmt, int width, int height, int max_value,
               int preset, int bits_per_channel, int qpreset, int nthreads,
               uint8_t **jpeg, size_t *jpeg_len, orc_result *res, int keep_planes);
void orc_result_free(orc_result *res);
void orc_free(void *p);

#ifdef __cplusplus
}
#endif
#endif
