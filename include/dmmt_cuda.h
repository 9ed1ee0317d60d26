/*
 * dmmt_cuda.h -- C ABI of the B200-native (sm_100a) encode hot path of dmmt-jpeg-encoder.
 *
 * This is the drop-in boundary: the body of the reference's
 *     impl ImageWriter for JpegImageWriter { fn write_image(&mut self) }   (src/image/writer/jpeg.rs:64-75)
 * i.e. Transformer::transform (src/image/writer/jpeg/transformer.rs:188-221) followed by
 * Encoder::encode (src/image/writer/jpeg/encoder.rs:125-135) becomes ONE call, dmmt_encode();
 * the returned bytes are what the reference `write_all`s into its `T: Write`.
 * INTEGRATION.md shows the Rust `extern "C"` block / build.rs that binds these symbols.
 *
 * Conventions: plain pointers and sizes only; 0 = OK, negative = error (dmmt_strerror);
 * no exceptions cross the boundary; there is NO CPU fallback: without a CUDA device every
 * entry point that computes returns DMMT_E_NODEVICE.  A context is bound to one device and
 * one stream and is not thread-safe (one host thread per context).
 */
#ifndef DMMT_CUDA_H
#define DMMT_CUDA_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- error codes (mapping onto the reference's crate::Error, src/error.rs:4-23) ---------- */
#define DMMT_OK 0
#define DMMT_E_INVALID (-1)   /* bad argument (null, zero-sized image, unknown preset, sample > max: color.rs:62-65 panics) */
#define DMMT_E_NODEVICE (-2)  /* no usable CUDA device; never falls back to the CPU */
#define DMMT_E_CUDA (-3)      /* CUDA runtime failure (see dmmt_last_cuda_error) */
#define DMMT_E_NCCL (-4)      /* collective failure in the sharded path */
#define DMMT_E_NOMEM (-5)     /* host or device allocation failed */
#define DMMT_E_OVERFLOW (-6)  /* entropy-coded scan exceeded the plan's scan capacity, or its stuffed form the plan's output slot
                               * (sized from that capacity): retry with dmmt_plan_set_scan_capacity(worst case), which also
                               * sizes the slot for a scan of nothing but 0xFF bytes.  dmmt_encode, dmmt_plan_encode_host*,
                               * dmmt_encode_sharded and the Python drivers do this themselves. */
#define DMMT_E_SYMBOL (-7)    /* Error::HuffmanSymbolNotPresentInTranslator (error.rs:21) */
#define DMMT_E_RANGE (-8)     /* coefficient not categorisable (categorize.rs:25-30 panics on -32768) */
#define DMMT_E_WRITE (-9)     /* Error::FailedToWriteImageData / FailedToWriteBlock family: output arena too small */
#define DMMT_E_SIZE (-10)     /* geometry beyond the reference's u16 image model (image.rs:8-9, padder.rs:13-14) */

/* ---- reference option enums -------------------------------------------------------------- */
/* ChromaSubsamplingPreset, src/image/subsampling.rs:11-16 */
#define DMMT_P444 0
#define DMMT_P422 1
#define DMMT_P420 2
/* QuantizationTablePreset in enum order, src/image/writer/jpeg/quantization_tables.rs:232-243 */
#define DMMT_Q_SPECIFICATION 0
#define DMMT_Q_FLAT 1
#define DMMT_Q_MSSIM_KODAK_TUNED 2
#define DMMT_Q_PSNR_HVS_N_KODAK_TUNED 3
#define DMMT_Q_DCTUNE_PERCEPTUAL_OPTIMIZATION 4
#define DMMT_Q_A_VISUAL_DETECTION_MODEL 5
#define DMMT_Q_AN_IMPROVED_DETECTION_MODEL 6

/* Pixel formats.  F32_NORM is the exact stand-in for the reference's Image<f32>
 * (src/image.rs:7-11: Vec<RGBColorFormat<f32>>, already v/max normalised by the reader,
 * src/image/reader/ppm.rs:153-157 + src/color.rs:45-53).  U8/U16 are the throughput formats:
 * the device performs the same IEEE `v as f32 / max as f32`. */
typedef enum { DMMT_RGB_F32_NORM = 0, DMMT_RGB_U8 = 1, DMMT_RGB_U16 = 2 } dmmt_fmt;

typedef struct dmmt_ctx dmmt_ctx;   /* one device + one stream + cached plans */
typedef struct dmmt_plan dmmt_plan; /* fixed geometry/options, owns all device scratch for n images */

typedef struct {
    uint16_t width, height;   /* Image<T>::{width,height} are u16 (image.rs:8-9) */
    uint16_t max_value;       /* PPM max value (ignored for F32_NORM) */
    dmmt_fmt fmt;
    const void *pixels;       /* interleaved R,G,B, row-major, tightly packed */
    int pixels_on_device;     /* 0: host memory, 1: device memory of the context's device */
} dmmt_image;

/* JpegTransformationOptions, src/image/writer/jpeg.rs:25-29 */
typedef struct {
    uint8_t subsampling;      /* DMMT_P444 / P422 / P420 */
    uint8_t bits_per_channel; /* copied into SOF0 only (encoder.rs:235) */
    uint8_t qtable_preset;    /* DMMT_Q_* */
} dmmt_options;

/* ---- contexts ---------------------------------------------------------------------------- */
int dmmt_device_count(void);                                /* 0 when no GPU / no driver */
int dmmt_ctx_create(int device, dmmt_ctx **out);            /* own non-blocking stream */
/* Same, but all work is issued on the caller's cudaStream_t (e.g. torch's current stream). */
int dmmt_ctx_create_on_stream(int device, void *cuda_stream, dmmt_ctx **out);
void dmmt_ctx_destroy(dmmt_ctx *);
int dmmt_ctx_synchronize(dmmt_ctx *);
void *dmmt_ctx_stream(dmmt_ctx *);                          /* the cudaStream_t in use */

/* ---- the drop-in call: replaces JpegImageWriter::write_image (jpeg.rs:64-75) -------------- */
/* Synchronous at return.  *jpeg is malloc'd by the callee; release with dmmt_free. */
int dmmt_encode(dmmt_ctx *, const dmmt_image *, const dmmt_options *, uint8_t **jpeg, size_t *len);
/* n independent images; image i is encoded on ctxs[i % nctx] (round-robin).
 * Images sharing geometry+format on one context are encoded by ONE batched launch chain. */
int dmmt_encode_batch(dmmt_ctx *const *ctxs, int nctx, const dmmt_image *imgs, int n,
                      const dmmt_options *, uint8_t **jpegs, size_t *lens);
/* One large image split by MCU rows over nctx devices of this process (single-process
 * multi-device; the one-process-per-GPU variant is the dmmt_shard_* API below).  With peer access
 * between the devices (NVLink / NVSwitch) every exchanged value stays in device memory: the shards
 * read each other's last DCs, histograms, bit counts, tails and byte counts over peer loads, K4 of
 * every shard stores straight into the file on ctxs[0]'s device, and there is ONE host
 * synchronisation; without peer access the values pass through the host.  Retries once with the
 * worst-case scan capacity on DMMT_E_OVERFLOW, like dmmt_encode. */
int dmmt_encode_sharded(dmmt_ctx *const *ctxs, int nctx, const dmmt_image *, const dmmt_options *,
                        uint8_t **jpeg, size_t *len);
void dmmt_free(void *);
const char *dmmt_strerror(int code);
const char *dmmt_last_cuda_error(void); /* thread-local text of the last CUDA failure */

/* ---- plans: fixed geometry, device-resident launch chain over up to n_images images ------- */
int dmmt_plan_create(dmmt_ctx *, uint16_t width, uint16_t height, dmmt_fmt fmt, uint16_t max_value,
                     const dmmt_options *, int n_images, dmmt_plan **out);
void dmmt_plan_destroy(dmmt_plan *);
size_t dmmt_plan_pixel_bytes(const dmmt_plan *);  /* bytes of ONE input image */
size_t dmmt_plan_out_stride(const dmmt_plan *);   /* bytes reserved per image in the output arena */
/* Scan capacity per image in bytes (default: 128 B per 8x8 block); worst case is
 * dmmt_plan_worst_case_scan_bytes().  Reallocates scratch. */
int dmmt_plan_set_scan_capacity(dmmt_plan *, size_t bytes_per_image);
size_t dmmt_plan_worst_case_scan_bytes(const dmmt_plan *);
/* Asynchronous on the plan's stream.  d_pixels: n_images (<= plan size) images back to back in
 * device memory; d_out: n_images * out_stride bytes; d_lens: n_images u64 (file length per
 * image; 0 + error flag on failure).  Use dmmt_plan_status after synchronising to read the
 * device-side error flags. */
int dmmt_plan_encode_device(dmmt_plan *, const void *d_pixels, int n_images, uint8_t *d_out,
                            uint64_t *d_lens);
int dmmt_plan_status(dmmt_plan *);                /* synchronises; first device-side error or 0 */
/* A call that repeats the device pointers of the previous one is replayed from a CUDA graph of the launch chain
 * (memset, K1, DC fix-up, K2b, scan zeroing, K3, K4: one graph launch instead of eight launches -- what the latency
 * of a single frame, BASELINE config 3, is made of).  On by default; 0 switches back to plain launches. */
int dmmt_plan_set_graph(dmmt_plan *, int enabled);
/* Host-buffer end-to-end: H2D of the pixels, encode, D2H of lengths + bytes.  jpegs[i] malloc'd
 * (release with dmmt_free).  Grows the scan capacity and retries once on DMMT_E_OVERFLOW. */
int dmmt_plan_encode_host(dmmt_plan *, const void *h_pixels, int n_images, uint8_t **jpegs,
                          size_t *lens);
/* Same, but the files are packed back to back (16-byte aligned starts) into the caller's host
 * arena (ideally pinned, see dmmt_host_alloc): file i = h_out[h_offsets[i] .. + h_lens[i]). */
int dmmt_plan_encode_host_into(dmmt_plan *, const void *h_pixels, int n_images, uint8_t *h_out,
                               uint64_t out_cap, uint64_t *h_offsets, uint64_t *h_lens);

/* ---- batches: many equally sized images, pipelined in sub-batches over `depth` streams ------ */
/* This is the throughput path (BASELINE config 4): sub-batch k runs on slot k % depth, so the
 * H2D copy, the kernel chain and the D2H copy of neighbouring sub-batches overlap. */
typedef struct dmmt_batch dmmt_batch;
int dmmt_batch_create(dmmt_ctx *, uint16_t width, uint16_t height, dmmt_fmt fmt, uint16_t max_value,
                      const dmmt_options *, int sub_batch, int depth, dmmt_batch **out);
void dmmt_batch_destroy(dmmt_batch *);
/* Inputs resident in device memory (n images back to back), outputs stay in device memory:
 * files packed back to back into d_dense (file i at d_offsets[i], length d_lens[i];
 * d_offsets has n + 1 entries, the last one is the total).  Asynchronous; follow with
 * dmmt_batch_status. */
int dmmt_batch_encode_device(dmmt_batch *, const void *d_pixels, int n, uint8_t *d_dense,
                             uint64_t dense_cap, uint64_t *d_offsets, uint64_t *d_lens);
/* Host pixels (ideally pinned) -> host files, packed into h_out like
 * dmmt_plan_encode_host_into.  Synchronous at return. */
int dmmt_batch_encode_host(dmmt_batch *, const void *h_pixels, int n, uint8_t *h_out, uint64_t out_cap,
                           uint64_t *h_offsets, uint64_t *h_lens);
int dmmt_batch_status(dmmt_batch *);              /* synchronises all slots; first error or 0 */
/* scan capacity of every slot (see dmmt_plan_set_scan_capacity); use after DMMT_E_OVERFLOW */
int dmmt_batch_set_scan_capacity(dmmt_batch *, size_t bytes_per_image);
size_t dmmt_batch_worst_case_scan_bytes(const dmmt_batch *);
int dmmt_batch_last_launch_count(const dmmt_batch *);
int dmmt_batch_uses_fused_path(const dmmt_batch *);
/* per-kernel timings (DMMT_T_*) summed over the sub-batches of the last dmmt_batch_encode_device
 * call; profiling serialises the slots. */
int dmmt_batch_set_profiling(dmmt_batch *, int enabled);
int dmmt_batch_last_timings(dmmt_batch *, float *ms, int n);
/* page-locked host memory for the host-buffer paths */
int dmmt_host_alloc(size_t bytes, void **out);
void dmmt_host_free(void *);

/* per-kernel CUDA-event timings of the last dmmt_plan_encode_* call (profiling must be enabled
 * first; it adds event records between the kernels).  ms[] order: DMMT_T_* */
#define DMMT_T_K1_TRANSFORM 0 /* fused normalise/pad/YCbCr/subsample/DCT/quantise/zig-zag */
#define DMMT_T_K2_HISTOGRAM 1 /* tokenise + histogram (generic path) / tile-boundary DC fix-up (fused path) */
#define DMMT_T_K2B_TABLES 2
#define DMMT_T_K3_PACK 3      /* token packing + decoupled look-back scan (incl. scan zeroing) */
#define DMMT_T_K4_STUFF 4
#define DMMT_T_K5_COMPACT 5 /* packing of the output arena (0 when not run) */
#define DMMT_T_TOTAL 6
#define DMMT_T_COUNT 7
int dmmt_plan_set_profiling(dmmt_plan *, int enabled);
/* 1: run the generic kernels (K1 writes the coefficient stream, K2 tokenises it) even where the fused
 * 4:2:0 fast path (K1 tokenises in registers, no coefficient stream) applies; needed before
 * dmmt_plan_fetch(DMMT_FETCH_COEF).  Default 0. */
int dmmt_plan_set_generic_path(dmmt_plan *, int generic);
int dmmt_plan_uses_fused_path(const dmmt_plan *);  /* 1 when the fused 4:2:0 kernels run */
int dmmt_plan_last_timings(dmmt_plan *, float *ms, int n);
/* number of kernels launched by the last encode call on this plan */
int dmmt_plan_last_launch_count(const dmmt_plan *);

/* ---- test / measurement hooks (not part of the drop-in surface) --------------------------- */
#define DMMT_FETCH_COEF 0      /* i16 [n_stream_blocks][64], zig-zag, MCU-interleaved stream order */
#define DMMT_FETCH_HIST 1      /* u32 [4][256]: Y-DC, Y-AC, C-DC, C-AC */
#define DMMT_FETCH_TABLES 2    /* u8  [2][4][256]: symbols[4][256] then lengths[4][256], each table in the reference's Vec<SymbolCodeLength> order; counts via DMMT_FETCH_META */
#define DMMT_FETCH_SCAN 3      /* unstuffed, 1-padded scan bytes */
#define DMMT_FETCH_META 4      /* dmmt_image_meta */
#define DMMT_FETCH_TOKEN_COUNT 5 /* u64: tokens (4 bytes each) the transform / tokenise stage wrote for this image */
typedef struct {
    uint64_t scan_bits;        /* entropy-coded bits before padding */
    uint64_t out_len;          /* whole file length */
    uint32_t header_len;
    uint32_t n_symbols[4];
    int32_t error;
    uint32_t n_stream_blocks;
    uint32_t reserved;
} dmmt_image_meta;
/* copies an intermediate of image `index` of the last encode to host memory */
int dmmt_plan_fetch(dmmt_plan *, int what, int index, void *dst, size_t cap_bytes, size_t *got);
/* pre-quantisation DCT coefficients (f32, natural order inside a block, stream order of blocks)
 * of image `index`: runs the debug variant of K1 on d_pixels (device) and copies to host. */
int dmmt_plan_debug_dct(dmmt_plan *, const void *d_pixels, int index, float *dst, size_t cap_floats);
size_t dmmt_plan_stream_blocks(const dmmt_plan *);
/* test hook, K4 alone: stuffs the `n` bytes of an unstuffed scan (host memory) the way the last stage of dmmt_encode
 * does (segment_marker_injector.rs:13-30: 0x00 after every 0xFF; EOI appended, encoder.rs:164-167) into a device buffer
 * that starts `misalign` (0..15) bytes after a 16-byte boundary, and copies the result to `out`. */
int dmmt_debug_stuff(dmmt_ctx *, const uint8_t *scan, size_t n, int misalign, uint8_t *out, size_t cap, size_t *got);

/* ---- one-process-per-GPU sharding of ONE image by MCU rows -------------------- */
/* A shard plan covers MCU rows [mcu_row_begin, mcu_row_end) of a full_width x full_height image.
 * The caller exchanges the small values between the phases with its own collective
 * (torch.distributed / NCCL allgather + allreduce); every phase is synchronous at return. */
typedef struct dmmt_shard dmmt_shard;
int dmmt_shard_create(dmmt_ctx *, uint16_t full_width, uint16_t full_height, dmmt_fmt fmt,
                      uint16_t max_value, const dmmt_options *, int mcu_row_begin, int mcu_row_end,
                      dmmt_shard **out);
void dmmt_shard_destroy(dmmt_shard *);
int dmmt_shard_mcu_rows_total(uint16_t full_height, const dmmt_options *);
size_t dmmt_shard_pixel_bytes(const dmmt_shard *);       /* bytes of this shard's pixel rows */
size_t dmmt_shard_pixel_offset(const dmmt_shard *);      /* byte offset of those rows in the full image */
/* phase 1: K1 on the shard's rows (device pointer). last_dc = quantised DC of the shard's last Y, Cb, Cr block. */
int dmmt_shard_transform(dmmt_shard *, const void *d_pixels, int16_t last_dc[3]);
/* phase 2: K2 with the previous shard's last DCs as predictors; hist = 4*256 u64 local counts. */
int dmmt_shard_histogram(dmmt_shard *, const int16_t seed_dc[3], uint64_t hist[1024]);
/* phase 3: K2b from the GLOBAL (all-reduced) histogram; returns this shard's entropy-coded bits. */
int dmmt_shard_tables(dmmt_shard *, const uint64_t global_hist[1024], uint64_t *local_bits);
/* phase 4: K3 at the shard's global bit offset. is_last: append the 1-padding.
 * tail_bits/tail_nbits: the trailing partial byte of this shard (bits beyond the last whole
 * byte boundary of the global stream), which the next shard must OR into its first byte. */
int dmmt_shard_pack(dmmt_shard *, uint64_t global_bit_offset, int is_last, uint8_t *tail_byte,
                    int *tail_nbits);
/* phase 5: K4 on the bytes this shard owns (global bytes whose LAST bit lies in the shard),
 * with the previous shard's tail byte OR-ed into the first one.  is_first: prepend the header;
 * is_last: append EOI.  Result stays on the device: *d_bytes / *n_bytes. */
int dmmt_shard_stuff(dmmt_shard *, uint8_t prev_tail_byte, int prev_tail_nbits, int is_first,
                     int is_last, const uint8_t **d_bytes, uint64_t *n_bytes);

/* Scan capacity of a shard (see dmmt_plan_set_scan_capacity).  A phase that reports DMMT_E_OVERFLOW means: give EVERY
 * shard of the image dmmt_shard_worst_case_scan_bytes() and run the phases again from the transform.  The decision
 * must be taken by all ranks together (all-reduce the status), dmmt_encode_sharded and sharded.py do so. */
int dmmt_shard_set_scan_capacity(dmmt_shard *, size_t bytes);
size_t dmmt_shard_worst_case_scan_bytes(const dmmt_shard *);
int dmmt_shard_launch_count(const dmmt_shard *);  /* kernels launched by the phases so far */
/* Device-resident exchange: the same five phases, ASYNCHRONOUS on the context's stream, with every exchanged
 * value in device memory, so the caller's collectives (NCCL all-gather / all-reduce on the same stream) need
 * no host round trip between the phases.  Layouts: last_dc / seed_dc int32[4] (Y, Cb, Cr, pad), hist
 * int64[1024], local bits / global bit offset int64, tail int32[2] = {byte, valid leading bits};
 * dmmt_shard_launch_stuff takes the all-gathered tails [world][2], exclusive bit offsets [world] and bit
 * counts [world] and writes this shard's stuffed byte count; *d_bytes is where the bytes will be. */
int dmmt_shard_launch_transform(dmmt_shard *, const void *d_pixels, int32_t *d_last_dc4);
int dmmt_shard_launch_histogram(dmmt_shard *, const int32_t *d_seed_dc4 /* NULL on the first shard */, int64_t *d_hist1024);
int dmmt_shard_launch_tables(dmmt_shard *, const int64_t *d_global_hist1024, int64_t *d_local_bits);
int dmmt_shard_launch_pack(dmmt_shard *, const int64_t *d_global_bit_offset, int is_last, int32_t *d_tail2);
int dmmt_shard_launch_stuff(dmmt_shard *, const int32_t *d_all_tail2, const int64_t *d_all_bit_offsets,
                            const int64_t *d_all_bits, int rank, int world, const uint8_t **d_bytes, int64_t *d_n_bytes);
int dmmt_shard_status(dmmt_shard *);              /* synchronises; device-side error of the phases so far or 0 */
/* the same flag as an int64 in device memory, asynchronous: all-gather it with the byte counts so that every rank
 * learns of a failed shard (a failed shard reports 0 bytes, which is also a legitimate count) */
int dmmt_shard_launch_error(dmmt_shard *, int64_t *d_err);

/* Peer-memory gather (one process per GPU on one NVLink / NVSwitch node): instead of dmmt_shard_launch_stuff +
 * a send / recv of the shard outputs, the destination rank allocates the whole file once
 * (dmmt_device_alloc, plain device memory) and exports it (dmmt_peer_export -> 64 opaque bytes the caller
 * hands to the other processes); they map it (dmmt_peer_open) and K4 of every shard stores its bytes at their
 * final place in that file over NVLink.  Phase 5 splits in two around one more tiny exchange:
 *   5a dmmt_shard_launch_count_bytes: stuffed size of this shard (header on rank 0, EOI on the last) ->
 *      all-gather + exclusive sum = byte offset of every shard in the file;
 *   5b dmmt_shard_launch_stuff_into:  K4 into d_file + *d_byte_offset; d_result2 = {end offset, error}, whose
 *      all-gather is both the completion barrier for the destination rank and the status of the encode. */
#define DMMT_PEER_HANDLE_BYTES 64
size_t dmmt_shard_out_stride(const dmmt_shard *); /* most bytes this shard can write: sum over shards = a safe file size */
int dmmt_device_alloc(dmmt_ctx *, size_t bytes, void **d_ptr);
int dmmt_device_free(dmmt_ctx *, void *d_ptr);
int dmmt_peer_export(dmmt_ctx *, void *d_ptr, uint8_t handle[DMMT_PEER_HANDLE_BYTES]);
int dmmt_peer_open(dmmt_ctx *, const uint8_t handle[DMMT_PEER_HANDLE_BYTES], void **d_ptr);
int dmmt_peer_close(dmmt_ctx *, void *d_ptr);
int dmmt_shard_launch_count_bytes(dmmt_shard *, const int32_t *d_all_tail2, const int64_t *d_all_bit_offsets,
                                  const int64_t *d_all_bits, int rank, int world, int64_t *d_n_bytes);
int dmmt_shard_launch_stuff_into(dmmt_shard *, const int64_t *d_all_bit_offsets, int rank, int world, uint8_t *d_file,
                                 size_t file_capacity, const int64_t *d_byte_offset, int64_t *d_result2);

/* measurement hook: with DMMT_SHARDED_TIMING=1 in the environment, the wall-clock milliseconds the phase section of this
 * thread's last dmmt_encode_sharded took on the peer-memory path (every H2D copy finished before it, its one host
 * synchronisation at the end; the retry after DMMT_E_OVERFLOW overwrites it); -1 if never measured */
double dmmt_encode_sharded_last_ms(void);

/* Mailbox exchange: the small values the shards exchange between the phases, moved by the library's own kernels over
 * peer memory instead of a collective library (one process per GPU on one NVLink / NVSwitch node).  Every rank allocates
 * dmmt_mailbox_bytes(world) of ZEROED device memory (dmmt_device_alloc + cudaMemset, once), exports it and maps the
 * others' (dmmt_peer_export / dmmt_peer_open).  dmmt_shard_launch_post stores n_words64 (<= 1024) 64-bit words into
 * row `rank` of slot `slot` (0 .. DMMT_MAILBOX_SLOTS-1) of EVERY mailbox and releases the row with `seq`;
 * dmmt_shard_launch_collect, launched after it on the same stream, waits until all rows of the slot carry a sequence
 * number >= seq and writes d_out: mode 0 the element-wise sum over the ranks (n words), 1 the rows one after the other
 * (world x n words), 2 (n = 1) the rows followed by their exclusive prefix sums (2 x world words).  `seq` must be the
 * same on every rank and grow from one encode to the next (start at 1); a slot is used once per encode.  A peer that
 * never posts makes the collect give up after 10 s with DMMT_E_NCCL in the shard's error flag (dmmt_shard_launch_error)
 * rather than hang the device.  Asynchronous like the other dmmt_shard_launch_* calls. */
#define DMMT_MAILBOX_SLOTS 8
size_t dmmt_mailbox_bytes(int world);
int dmmt_shard_launch_post(dmmt_shard *, void *const *d_mailboxes /* [world], own one included */, int rank, int world, int slot,
                           unsigned long long seq, const void *d_src, int n_words64);
int dmmt_shard_launch_collect(dmmt_shard *, void *d_own_mailbox, int world, int slot, unsigned long long seq, int mode,
                              int n_words64, int64_t *d_out);

/* ---- host ingest: ASCII P3 reader (replaces PPMImageReader::read_image, src/image/reader/ppm.rs:9-251) ----
 * Pure host code (no device needed): tokenises `len` bytes of a P3 file with the reference's rules (`#` comments
 * anywhere, Rust's ASCII whitespace set, every number a u16) and returns the raw samples (interleaved R,G,B u16,
 * malloc'd, release with dmmt_free) -- feed them to dmmt_encode as DMMT_RGB_U16 with *max_value; the device performs
 * the reader's `v as f32 / max as f32` (ppm.rs:153-157, color.rs:45-53).  `threads` > 1 parses a comment-free sample
 * section in parallel (the CLI's -t/--threads, src/cli.rs:104-109).  Returns DMMT_OK or one of DMMT_PPM_* (> 0) with
 * *detail = the header token index 0..3 (P3, width, height, max value; 4 = "Color Component Value") for
 * MISSING_TOKEN / BAD_TOKEN and `n % 3` for INCOMPLETE_PIXEL; the texts of src/error.rs are dmmt_ppm_strerror's. */
#define DMMT_PPM_MISSING_TOKEN 1    /* Error::PPMFileDoesNotContainRequiredToken (error.rs:6) */
#define DMMT_PPM_BAD_TOKEN 2        /* Error::ParsingOfTokenFailed (error.rs:8) */
#define DMMT_PPM_INCOMPLETE_PIXEL 3 /* Error::IncompletePixelParsed (error.rs:10) */
#define DMMT_PPM_SIZE_MISMATCH 4    /* Error::MismatchOfSizeBetweenHeaderAndValues (error.rs:12) */
#define DMMT_PPM_SAMPLE_ABOVE_MAX 5 /* the reference panics (color.rs:62-65) */
int dmmt_ppm_parse(const char *text, size_t len, int threads, uint16_t *width, uint16_t *height, uint16_t *max_value,
                   uint16_t **samples, size_t *n_samples, int *detail);
/* Display text of the reader error (status, detail) into buf (NUL-terminated, truncated to cap); returns buf. */
const char *dmmt_ppm_strerror(int status, int detail, char *buf, size_t cap);

#ifdef __cplusplus
}
#endif
#endif /* DMMT_CUDA_H */
