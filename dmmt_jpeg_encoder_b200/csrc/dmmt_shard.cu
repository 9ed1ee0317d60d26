// dmmt_shard.cu -- ONE image split by MCU rows over several devices / ranks (BASELINE config 5,
// SURVEY 8e).  A shard owns MCU rows [begin, end): a contiguous range of the stream order of all
// three components (block_fold_iterator.rs:125-148, block_entangler.rs:69-77), so K1 is local and
// the entropy stage needs four tiny exchanges, which the caller performs with its own collective
// (torch.distributed / NCCL all-gather + all-reduce in the one-process-per-GPU driver, plain host
// code in dmmt_encode_sharded):
//   1. last quantised DC of (Y, Cb, Cr) of every shard  -> predictor seeds (categorize.rs:157-161)
//   2. sum of the 4 symbol histograms                   -> image-global tables (transformer.rs:201-217)
//   3. entropy-coded bits of every shard                -> global bit offsets (binary_stream.rs:38-66)
//   4. stuffed byte counts                              -> final byte offsets of the shard outputs
#include <algorithm>
#include <chrono>
#include <cstdlib>
#include <cstring>
#include <new>

#include "dmmt_internal.h"
#include "dmmt_qtables.h"

using namespace dmmt;

struct dmmt_shard {
    dmmt_plan* plan = nullptr;
    int full_W = 0, full_H = 0;
    int row_begin = 0, row_end = 0;  // MCU rows
    int px_row_begin = 0, px_rows = 0;
    unsigned long long local_bits = 0;
    unsigned long long seed_bits = 0;
    bool have_seed = false;
    unsigned int* h_hist = nullptr;  // pinned [1024]
    int* d_prev_tail = nullptr;      // device-resident exchange: tail bits to OR into this shard's first byte
    unsigned long long* d_count_ctr = nullptr;  // [2] self-resetting counters of k_shard_count_bytes
};

static int mcu_rows_of(int H, int subsampling) {
    const int vr = subsampling == DMMT_P420 ? 2 : 1;
    return (H + 8 * vr - 1) / (8 * vr);
}

extern "C" int dmmt_shard_mcu_rows_total(uint16_t full_height, const dmmt_options* opt) {
    if (!opt || opt->subsampling > DMMT_P420 || full_height == 0) return DMMT_E_INVALID;
    return mcu_rows_of(full_height, opt->subsampling);
}

extern "C" void dmmt_shard_destroy(dmmt_shard* s) {
    if (!s) return;
    if (s->plan) {
        (void)cudaSetDevice(s->plan->ctx->device);
        dmmt_plan_destroy(s->plan);
    }
    if (s->h_hist) (void)cudaFreeHost(s->h_hist);
    (void)cudaFree(s->d_prev_tail);
    (void)cudaFree(s->d_count_ctr);
    delete s;
}

extern "C" int dmmt_shard_create(dmmt_ctx* ctx, uint16_t full_width, uint16_t full_height, dmmt_fmt fmt,
                                 uint16_t max_value, const dmmt_options* opt, int mcu_row_begin, int mcu_row_end,
                                 dmmt_shard** out) {
    if (!ctx || !opt || !out || opt->subsampling > DMMT_P420 || full_width == 0 || full_height == 0)
        return DMMT_E_INVALID;
    *out = nullptr;
    const int total = mcu_rows_of(full_height, opt->subsampling);
    if (mcu_row_begin < 0 || mcu_row_end > total || mcu_row_begin >= mcu_row_end) return DMMT_E_INVALID;
    const int vr = opt->subsampling == DMMT_P420 ? 2 : 1;
    dmmt_shard* s = new (std::nothrow) dmmt_shard();
    if (!s) return DMMT_E_NOMEM;
    s->full_W = full_width, s->full_H = full_height;
    s->row_begin = mcu_row_begin, s->row_end = mcu_row_end;
    s->px_row_begin = mcu_row_begin * 8 * vr;
    s->px_rows = std::min<int>(full_height, mcu_row_end * 8 * vr) - s->px_row_begin;  // >= 1
    // the whole padded image must fit the reference's u16 model (padder.rs:6-7)
    if (total * 8 * vr > 65535) {
        delete s;
        return DMMT_E_SIZE;
    }
    int rc = dmmt_plan_create_impl(ctx, full_width, s->px_rows, mcu_row_end - mcu_row_begin, full_width, full_height,
                                   (int)fmt, fmt == DMMT_RGB_F32_NORM ? 1 : max_value, opt, 1, ctx->stream, false,
                                   &s->plan);
    if (rc == DMMT_OK && cudaHostAlloc(&s->h_hist, 1024 * sizeof(unsigned int), cudaHostAllocDefault) != cudaSuccess)
        rc = DMMT_E_NOMEM;
    if (rc == DMMT_OK && cudaMalloc(&s->d_prev_tail, sizeof(int)) != cudaSuccess) rc = DMMT_E_NOMEM;
    if (rc == DMMT_OK && (cudaMalloc(&s->d_count_ctr, 2 * sizeof(unsigned long long)) != cudaSuccess ||
                          cudaMemset(s->d_count_ctr, 0, 2 * sizeof(unsigned long long)) != cudaSuccess))
        rc = DMMT_E_NOMEM;
    if (rc != DMMT_OK) {
        dmmt_shard_destroy(s);
        return rc;
    }
    *out = s;
    return DMMT_OK;
}

extern "C" size_t dmmt_shard_pixel_bytes(const dmmt_shard* s) { return s ? s->plan->pixel_bytes : 0; }
extern "C" size_t dmmt_shard_out_stride(const dmmt_shard* s) { return s ? s->plan->out_stride : 0; }
extern "C" size_t dmmt_shard_pixel_offset(const dmmt_shard* s) {
    if (!s) return 0;
    const size_t pb = s->plan->fmt == DMMT_RGB_U8 ? 3 : (s->plan->fmt == DMMT_RGB_U16 ? 6 : 12);
    return (size_t)s->px_row_begin * s->full_W * pb;
}

// ---- phase 1 ---------------------------------------------------------------------------------
static int shard_transform_launch(dmmt_shard* s, const void* d_pixels) {
    dmmt_plan* p = s->plan;
    DMMT_CUDA(cudaSetDevice(p->ctx->device));
    DMMT_CUDA(cudaMemsetAsync(p->zero_region, 0, p->zero_bytes, p->stream));
    const int check_max = (p->fmt == DMMT_RGB_U8 && p->max_value < 255) || (p->fmt == DMMT_RGB_U16 && p->max_value < 65535);
    DMMT_CUDA(launch_k1(p->g, p->fmt, p->k1c, check_max, d_pixels, p->pixel_bytes, 1, p->coef, p->coef_stride, nullptr,
                        p->meta, p->fused ? &p->fo : nullptr, p->hist, p->stream));
    if (p->fused) {
        // quantised DC of the shard's last Y, Cb, Cr block = last_dc of its last tile
        DMMT_CUDA(cudaMemcpyAsync(p->d_last_dc, p->fo.last_dc + (size_t)(p->fo.tiles - 1) * 4, 3 * sizeof(int16_t),
                                  cudaMemcpyDeviceToDevice, p->stream));
    } else {
        DMMT_CUDA(launch_last_dc(p->g, p->coef, p->d_last_dc, p->stream));
    }
    p->last_launches = 2;
    p->last_n = 1;
    return DMMT_OK;
}
static int shard_transform_collect(dmmt_shard* s, int16_t last_dc[3]) {
    dmmt_plan* p = s->plan;
    DMMT_CUDA(cudaSetDevice(p->ctx->device));
    DMMT_CUDA(cudaMemcpyAsync(last_dc, p->d_last_dc, 3 * sizeof(int16_t), cudaMemcpyDeviceToHost, p->stream));
    DMMT_CUDA(cudaStreamSynchronize(p->stream));
    return DMMT_OK;
}
extern "C" int dmmt_shard_transform(dmmt_shard* s, const void* d_pixels, int16_t last_dc[3]) {
    if (!s || !d_pixels || !last_dc) return DMMT_E_INVALID;
    DMMT_TRY(shard_transform_launch(s, d_pixels));
    return shard_transform_collect(s, last_dc);
}

// ---- phase 2 ---------------------------------------------------------------------------------
static int shard_histogram_launch(dmmt_shard* s, const int16_t seed_dc[3]) {
    dmmt_plan* p = s->plan;
    DMMT_CUDA(cudaSetDevice(p->ctx->device));
    DMMT_CUDA(cudaMemcpyAsync(p->d_seed_dc, seed_dc, 3 * sizeof(int16_t), cudaMemcpyHostToDevice, p->stream));
    s->have_seed = true;
    if (p->fused) DMMT_CUDA(launch_k2_fix_dc(p->fo, 1, p->hist, p->meta, p->d_seed_dc, p->stream));
    else DMMT_CUDA(launch_k2(p->g, p->coef, p->coef_stride, 1, p->hist, p->meta, p->d_seed_dc, p->tb, p->stream));
    p->last_launches += 1;
    DMMT_CUDA(cudaMemcpyAsync(s->h_hist, p->hist, 1024 * sizeof(unsigned int), cudaMemcpyDeviceToHost, p->stream));
    return DMMT_OK;
}
static int shard_histogram_collect(dmmt_shard* s, uint64_t hist[1024]) {
    DMMT_CUDA(cudaSetDevice(s->plan->ctx->device));
    DMMT_CUDA(cudaStreamSynchronize(s->plan->stream));
    for (int i = 0; i < 1024; i++) hist[i] = s->h_hist[i];
    return DMMT_OK;
}
extern "C" int dmmt_shard_histogram(dmmt_shard* s, const int16_t seed_dc[3], uint64_t hist[1024]) {
    if (!s || !seed_dc || !hist) return DMMT_E_INVALID;
    DMMT_TRY(shard_histogram_launch(s, seed_dc));
    return shard_histogram_collect(s, hist);
}

// ---- phase 3 ---------------------------------------------------------------------------------
static int shard_tables_launch(dmmt_shard* s, const uint64_t global_hist[1024]) {
    dmmt_plan* p = s->plan;
    DMMT_CUDA(cudaSetDevice(p->ctx->device));
    if (!p->d_out_own) DMMT_CUDA(cudaMalloc(&p->d_out_own, p->out_stride));
    DMMT_CUDA(cudaMemcpyAsync(p->d_ghist, global_hist, 1024 * 8, cudaMemcpyHostToDevice, p->stream));
    K2bHostArgs b{};
    b.hist = p->hist, b.ghist = p->d_ghist, b.enc = p->enc, b.lens = p->lens, b.meta = p->meta;
    b.out = p->d_out_own, b.out_stride = p->out_stride;
    b.scan_cap_bits = (unsigned long long)p->scan_cap_bytes * 8 - 8;  // room for the seed bits
    b.W = p->sof_W, b.H = p->sof_H, b.bits_per_channel = p->opt.bits_per_channel;
    b.qtab_luma = kQuantPresets[p->opt.qtable_preset][0];
    b.qtab_chroma = kQuantPresets[p->opt.qtable_preset][1];
    b.write_header = 1;  // every shard writes the (identical) header into its arena; only the first one ships it
    b.lcount = p->lcount, b.fix = nullptr;
    DMMT_CUDA(launch_k2b(p->g, b, 1, p->stream));
    p->last_launches += 1;
    return DMMT_OK;
}
static int shard_tables_collect(dmmt_shard* s, uint64_t* local_bits) {
    dmmt_plan* p = s->plan;
    DMMT_CUDA(cudaSetDevice(p->ctx->device));
    ImgMeta m{};
    DMMT_CUDA(cudaMemcpyAsync(&m, p->meta, sizeof m, cudaMemcpyDeviceToHost, p->stream));
    DMMT_CUDA(cudaStreamSynchronize(p->stream));
    if (m.error) return m.error;
    s->local_bits = m.scan_bits;
    *local_bits = m.scan_bits;
    return DMMT_OK;
}
extern "C" int dmmt_shard_tables(dmmt_shard* s, const uint64_t global_hist[1024], uint64_t* local_bits) {
    if (!s || !global_hist || !local_bits) return DMMT_E_INVALID;
    DMMT_TRY(shard_tables_launch(s, global_hist));
    return shard_tables_collect(s, local_bits);
}

// ---- phase 4 ---------------------------------------------------------------------------------
static int shard_pack_launch(dmmt_shard* s, uint64_t global_bit_offset, int is_last) {
    dmmt_plan* p = s->plan;
    DMMT_CUDA(cudaSetDevice(p->ctx->device));
    s->seed_bits = global_bit_offset & 7;  // the shard's buffer starts at the byte holding its first bit
    const int zero_blocks = (int)std::min<size_t>(std::max<size_t>(p->scan_cap_bytes / 65536, 1), 1024);
    DMMT_CUDA(launch_zero_scan(p->scan, p->scan_stride_words, p->meta, 1, s->seed_bits, zero_blocks, p->stream));
    {
        TokBuf tb = p->tb;
        if (p->fused) tb.chunk_cap = p->fo.tile_cap;
        DMMT_CUDA(launch_k3(p->fused ? p->n_chunks3f : p->n_chunks3, p->fused ? p->fo.tiles : 0u, 1, tb, p->enc, p->meta,
                            p->lb3, p->tk3, p->scan, p->scan_stride_words, s->seed_bits, is_last ? 1 : 0, p->stream));
    }
    p->last_launches += 2;
    return DMMT_OK;
}
static int shard_pack_collect(dmmt_shard* s, int is_last, uint8_t* tail_byte, int* tail_nbits) {
    dmmt_plan* p = s->plan;
    DMMT_CUDA(cudaSetDevice(p->ctx->device));
    const unsigned long long end = s->seed_bits + s->local_bits;
    *tail_byte = 0;
    *tail_nbits = is_last ? 0 : (int)(end & 7);
    if (*tail_nbits)
        DMMT_CUDA(cudaMemcpyAsync(tail_byte, reinterpret_cast<const uint8_t*>(p->scan) + end / 8, 1,
                                  cudaMemcpyDeviceToHost, p->stream));
    DMMT_CUDA(cudaStreamSynchronize(p->stream));
    return DMMT_OK;
}
extern "C" int dmmt_shard_pack(dmmt_shard* s, uint64_t global_bit_offset, int is_last, uint8_t* tail_byte,
                               int* tail_nbits) {
    if (!s || !tail_byte || !tail_nbits) return DMMT_E_INVALID;
    DMMT_TRY(shard_pack_launch(s, global_bit_offset, is_last));
    return shard_pack_collect(s, is_last, tail_byte, tail_nbits);
}

// ---- phase 5 ---------------------------------------------------------------------------------
// Owned bytes: local bytes [0, floor(end / 8)) -- plus the padded last byte on the last shard.
// Local byte 0 also carries the previous shard's tail bits (OR-ed in here).
static int shard_stuff_launch(dmmt_shard* s, uint8_t prev_tail_byte, int is_first, int is_last) {
    dmmt_plan* p = s->plan;
    DMMT_CUDA(cudaSetDevice(p->ctx->device));
    const unsigned long long end = s->seed_bits + s->local_bits;
    const unsigned long long owned = is_last ? (end + 7) / 8 : end / 8;
    K4HostArgs k{};
    k.scan = reinterpret_cast<const uint8_t*>(p->scan), k.scan_stride_bytes = p->scan_stride_words * 4;
    k.meta = p->meta, k.lb_state = p->lb4, k.ticket = p->tk4, k.max_chunks = p->max_chunks4;
    k.out = p->d_out_own, k.out_stride = p->out_stride, k.out_lens = p->d_lens;
    k.first_byte = 0, k.n_bytes_override = (long long)owned, k.seed_bits = s->seed_bits;
    k.prepend_header = is_first, k.append_eoi = is_last;
    k.or_first_byte = prev_tail_byte;
    const uint32_t grid = std::max<uint32_t>(1, k4_max_chunks((size_t)owned));
    DMMT_CUDA(launch_k4(k, 1, grid, p->stream));
    p->last_launches += 1;
    DMMT_CUDA(cudaMemcpyAsync(p->h_lens, p->d_lens, 8, cudaMemcpyDeviceToHost, p->stream));
    return DMMT_OK;
}
static int shard_stuff_collect(dmmt_shard* s, const uint8_t** d_bytes, uint64_t* n_bytes) {
    dmmt_plan* p = s->plan;
    DMMT_CUDA(cudaSetDevice(p->ctx->device));
    ImgMeta m{};
    DMMT_CUDA(cudaMemcpyAsync(&m, p->meta, sizeof m, cudaMemcpyDeviceToHost, p->stream));
    DMMT_CUDA(cudaStreamSynchronize(p->stream));
    if (m.error) return m.error;
    *d_bytes = p->d_out_own;
    *n_bytes = p->h_lens[0];
    return DMMT_OK;
}
extern "C" int dmmt_shard_stuff(dmmt_shard* s, uint8_t prev_tail_byte, int prev_tail_nbits, int is_first,
                                int is_last, const uint8_t** d_bytes, uint64_t* n_bytes) {
    if (!s || !d_bytes || !n_bytes) return DMMT_E_INVALID;
    if ((unsigned long long)(prev_tail_nbits & 7) != s->seed_bits) return DMMT_E_INVALID;  // exchange out of step
    DMMT_TRY(shard_stuff_launch(s, prev_tail_byte, is_first, is_last));
    return shard_stuff_collect(s, d_bytes, n_bytes);
}

// ---- device-resident exchange: the same five phases, asynchronous, every exchanged value in device
// memory, so the caller can run its collectives (NCCL through torch.distributed) on the same stream
// without a single host round trip between the phases.  Layouts: last_dc / seed_dc int32[4] (Y, Cb, Cr,
// pad), hist int64[1024], bits / bit_offset int64, tail int32[2] = {byte, valid leading bits}.
extern "C" int dmmt_shard_launch_transform(dmmt_shard* s, const void* d_pixels, int32_t* d_last_dc4) {
    if (!s || !d_pixels || !d_last_dc4) return DMMT_E_INVALID;
    DMMT_TRY(shard_transform_launch(s, d_pixels));
    dmmt_plan* p = s->plan;
    DMMT_CUDA(launch_shard_widen(p->d_last_dc, d_last_dc4, nullptr, nullptr, nullptr, nullptr, p->stream));
    return DMMT_OK;
}

extern "C" int dmmt_shard_launch_histogram(dmmt_shard* s, const int32_t* d_seed_dc4, int64_t* d_hist1024) {
    if (!s || !d_hist1024) return DMMT_E_INVALID;
    dmmt_plan* p = s->plan;
    DMMT_CUDA(cudaSetDevice(p->ctx->device));
    DMMT_CUDA(launch_shard_narrow_seed(d_seed_dc4, p->d_seed_dc, p->stream));  // null: first shard, seeds 0
    s->have_seed = true;
    if (p->fused) DMMT_CUDA(launch_k2_fix_dc(p->fo, 1, p->hist, p->meta, p->d_seed_dc, p->stream));
    else DMMT_CUDA(launch_k2(p->g, p->coef, p->coef_stride, 1, p->hist, p->meta, p->d_seed_dc, p->tb, p->stream));
    DMMT_CUDA(launch_shard_widen(nullptr, nullptr, p->hist, reinterpret_cast<long long*>(d_hist1024), nullptr, nullptr,
                                 p->stream));
    p->last_launches += 3;
    return DMMT_OK;
}

extern "C" int dmmt_shard_launch_tables(dmmt_shard* s, const int64_t* d_global_hist, int64_t* d_local_bits) {
    if (!s || !d_global_hist || !d_local_bits) return DMMT_E_INVALID;
    dmmt_plan* p = s->plan;
    DMMT_CUDA(cudaSetDevice(p->ctx->device));
    if (!p->d_out_own) DMMT_CUDA(cudaMalloc(&p->d_out_own, p->out_stride));
    K2bHostArgs b{};
    b.hist = p->hist, b.ghist = reinterpret_cast<const unsigned long long*>(d_global_hist), b.enc = p->enc, b.lens = p->lens;
    b.meta = p->meta, b.out = p->d_out_own, b.out_stride = p->out_stride;
    b.scan_cap_bits = (unsigned long long)p->scan_cap_bytes * 8 - 8;
    b.W = p->sof_W, b.H = p->sof_H, b.bits_per_channel = p->opt.bits_per_channel;
    b.qtab_luma = kQuantPresets[p->opt.qtable_preset][0];
    b.qtab_chroma = kQuantPresets[p->opt.qtable_preset][1];
    b.write_header = 1;
    b.lcount = p->lcount, b.fix = nullptr;
    DMMT_CUDA(launch_k2b(p->g, b, 1, p->stream));
    DMMT_CUDA(launch_shard_widen(nullptr, nullptr, nullptr, nullptr, p->meta, reinterpret_cast<long long*>(d_local_bits),
                                 p->stream));
    p->last_launches += 2;
    return DMMT_OK;
}

extern "C" int dmmt_shard_launch_pack(dmmt_shard* s, const int64_t* d_global_bit_offset, int is_last, int32_t* d_tail2) {
    if (!s || !d_global_bit_offset || !d_tail2) return DMMT_E_INVALID;
    dmmt_plan* p = s->plan;
    DMMT_CUDA(cudaSetDevice(p->ctx->device));
    const auto* off = reinterpret_cast<const unsigned long long*>(d_global_bit_offset);
    const int zero_blocks = (int)std::min<size_t>(std::max<size_t>(p->scan_cap_bytes / 65536, 1), 1024);
    DMMT_CUDA(launch_zero_scan(p->scan, p->scan_stride_words, p->meta, 1, 0ull, zero_blocks, p->stream, off));
    TokBuf tb = p->tb;
    if (p->fused) tb.chunk_cap = p->fo.tile_cap;
    DMMT_CUDA(launch_k3(p->fused ? p->n_chunks3f : p->n_chunks3, p->fused ? p->fo.tiles : 0u, 1, tb, p->enc, p->meta, p->lb3,
                        p->tk3, p->scan, p->scan_stride_words, 0ull, is_last ? 1 : 0, p->stream, off));
    DMMT_CUDA(launch_shard_tail(reinterpret_cast<const uint8_t*>(p->scan), p->meta, off, is_last, d_tail2, p->stream));
    p->last_launches += 3;
    return DMMT_OK;
}

extern "C" int dmmt_shard_launch_stuff(dmmt_shard* s, const int32_t* d_all_tail2, const int64_t* d_all_bit_offsets,
                                       const int64_t* d_all_bits, int rank, int world, const uint8_t** d_bytes,
                                       int64_t* d_n_bytes) {
    if (!s || !d_all_tail2 || !d_all_bit_offsets || !d_all_bits || !d_bytes || !d_n_bytes || rank < 0 || rank >= world)
        return DMMT_E_INVALID;
    dmmt_plan* p = s->plan;
    DMMT_CUDA(cudaSetDevice(p->ctx->device));
    DMMT_CUDA(launch_shard_prev_tail(d_all_tail2, reinterpret_cast<const long long*>(d_all_bit_offsets),
                                     reinterpret_cast<const long long*>(d_all_bits), rank, s->d_prev_tail, p->stream));
    K4HostArgs k{};
    k.scan = reinterpret_cast<const uint8_t*>(p->scan), k.scan_stride_bytes = p->scan_stride_words * 4;
    k.meta = p->meta, k.lb_state = p->lb4, k.ticket = p->tk4, k.max_chunks = p->max_chunks4;
    k.out = p->d_out_own, k.out_stride = p->out_stride, k.out_lens = reinterpret_cast<unsigned long long*>(d_n_bytes);
    k.first_byte = 0, k.n_bytes_override = -1, k.seed_bits = 0;
    k.prepend_header = rank == 0, k.append_eoi = rank == world - 1, k.or_first_byte = 0;
    k.seed_src = reinterpret_cast<const unsigned long long*>(d_all_bit_offsets) + rank;
    k.owned_mode = rank == world - 1 ? 2 : 1;
    k.or_first_src = s->d_prev_tail;
    DMMT_CUDA(launch_k4(k, 1, p->max_chunks4, p->stream));
    p->last_launches += 2;
    *d_bytes = p->d_out_own;
    return DMMT_OK;
}

// ---- peer-memory gather: K4 of every shard writes straight into the destination rank's file --------------
// The file buffer is plain cudaMalloc memory on the destination device; the other ranks (one process per GPU)
// map it through CUDA IPC and their K4 stores travel over NVLink, so there is no separate gather step.
extern "C" int dmmt_device_alloc(dmmt_ctx* ctx, size_t bytes, void** d_ptr) {
    if (!ctx || !d_ptr || bytes == 0) return DMMT_E_INVALID;
    *d_ptr = nullptr;
    DMMT_CUDA(cudaSetDevice(ctx->device));
    if (cudaMalloc(d_ptr, bytes) != cudaSuccess) {
        (void)cudaGetLastError();
        return DMMT_E_NOMEM;
    }
    return DMMT_OK;
}
extern "C" int dmmt_device_free(dmmt_ctx* ctx, void* d_ptr) {
    if (!ctx) return DMMT_E_INVALID;
    DMMT_CUDA(cudaSetDevice(ctx->device));
    DMMT_CUDA(cudaFree(d_ptr));
    return DMMT_OK;
}
static_assert(sizeof(cudaIpcMemHandle_t) == DMMT_PEER_HANDLE_BYTES, "peer handle size");
extern "C" int dmmt_peer_export(dmmt_ctx* ctx, void* d_ptr, uint8_t handle[DMMT_PEER_HANDLE_BYTES]) {
    if (!ctx || !d_ptr || !handle) return DMMT_E_INVALID;
    DMMT_CUDA(cudaSetDevice(ctx->device));
    cudaIpcMemHandle_t h;
    DMMT_CUDA(cudaIpcGetMemHandle(&h, d_ptr));
    std::memcpy(handle, &h, sizeof h);
    return DMMT_OK;
}
extern "C" int dmmt_peer_open(dmmt_ctx* ctx, const uint8_t handle[DMMT_PEER_HANDLE_BYTES], void** d_ptr) {
    if (!ctx || !handle || !d_ptr) return DMMT_E_INVALID;
    *d_ptr = nullptr;
    DMMT_CUDA(cudaSetDevice(ctx->device));
    cudaIpcMemHandle_t h;
    std::memcpy(&h, handle, sizeof h);
    DMMT_CUDA(cudaIpcOpenMemHandle(d_ptr, h, cudaIpcMemLazyEnablePeerAccess));
    return DMMT_OK;
}
extern "C" int dmmt_peer_close(dmmt_ctx* ctx, void* d_ptr) {
    if (!ctx || !d_ptr) return DMMT_E_INVALID;
    DMMT_CUDA(cudaSetDevice(ctx->device));
    DMMT_CUDA(cudaIpcCloseMemHandle(d_ptr));
    return DMMT_OK;
}

// phase 5a: the stuffed size of this shard's bytes, known BEFORE K4 runs, so that the byte offsets of all
// shards in the file can be exchanged first and K4 can write to its final place
extern "C" int dmmt_shard_launch_count_bytes(dmmt_shard* s, const int32_t* d_all_tail2, const int64_t* d_all_bit_offsets,
                                             const int64_t* d_all_bits, int rank, int world, int64_t* d_n_bytes) {
    if (!s || !d_all_tail2 || !d_all_bit_offsets || !d_all_bits || !d_n_bytes || rank < 0 || rank >= world)
        return DMMT_E_INVALID;
    dmmt_plan* p = s->plan;
    DMMT_CUDA(cudaSetDevice(p->ctx->device));
    DMMT_CUDA(launch_shard_prev_tail(d_all_tail2, reinterpret_cast<const long long*>(d_all_bit_offsets),
                                     reinterpret_cast<const long long*>(d_all_bits), rank, s->d_prev_tail, p->stream));
    DMMT_CUDA(launch_shard_count_bytes(reinterpret_cast<const uint8_t*>(p->scan), p->meta,
                                       reinterpret_cast<const unsigned long long*>(d_all_bit_offsets) + rank,
                                       rank == world - 1 ? 2 : 1, s->d_prev_tail, rank == 0, rank == world - 1,
                                       s->d_count_ctr, reinterpret_cast<long long*>(d_n_bytes), p->stream));
    p->last_launches += 2;
    return DMMT_OK;
}

// phase 5b: K4 straight into the file at *d_byte_offset (exclusive sum of the all-gathered counts of 5a);
// d_file may be another device's memory (dmmt_peer_open).  d_result2 = {end offset in the file, error}.
extern "C" int dmmt_shard_launch_stuff_into(dmmt_shard* s, const int64_t* d_all_bit_offsets, int rank, int world,
                                            uint8_t* d_file, size_t file_capacity, const int64_t* d_byte_offset,
                                            int64_t* d_result2) {
    if (!s || !d_all_bit_offsets || !d_file || !d_byte_offset || !d_result2 || rank < 0 || rank >= world)
        return DMMT_E_INVALID;
    dmmt_plan* p = s->plan;
    DMMT_CUDA(cudaSetDevice(p->ctx->device));
    if (rank == 0) DMMT_CUDA(launch_shard_copy_header(p->d_out_own, p->meta, d_file, file_capacity, p->stream));
    K4HostArgs k{};
    k.scan = reinterpret_cast<const uint8_t*>(p->scan), k.scan_stride_bytes = p->scan_stride_words * 4;
    k.meta = p->meta, k.lb_state = p->lb4, k.ticket = p->tk4, k.max_chunks = p->max_chunks4;
    k.out = d_file, k.out_stride = file_capacity, k.out_lens = nullptr;
    k.first_byte = 0, k.n_bytes_override = -1, k.seed_bits = 0;
    k.prepend_header = rank == 0, k.append_eoi = rank == world - 1, k.or_first_byte = 0;
    k.seed_src = reinterpret_cast<const unsigned long long*>(d_all_bit_offsets) + rank;
    k.owned_mode = rank == world - 1 ? 2 : 1;
    k.or_first_src = s->d_prev_tail;
    k.base_src = reinterpret_cast<const unsigned long long*>(d_byte_offset);
    DMMT_CUDA(launch_k4(k, 1, p->max_chunks4, p->stream));
    DMMT_CUDA(launch_shard_result(p->meta, reinterpret_cast<long long*>(d_result2), p->stream));
    p->last_launches += rank == 0 ? 3 : 2;
    return DMMT_OK;
}

// the device-side error flag of the phases so far as an int64 in device memory (asynchronous): travels with the
// byte counts in the caller's all-gather, so that EVERY rank learns of a failed shard
extern "C" int dmmt_shard_launch_error(dmmt_shard* s, int64_t* d_err) {
    if (!s || !d_err) return DMMT_E_INVALID;
    dmmt_plan* p = s->plan;
    DMMT_CUDA(cudaSetDevice(p->ctx->device));
    DMMT_CUDA(launch_shard_widen(nullptr, nullptr, nullptr, nullptr, p->meta, nullptr, p->stream, reinterpret_cast<long long*>(d_err)));
    p->last_launches += 1;
    return DMMT_OK;
}

// device-side error flag of the shard's phases so far (synchronises)
extern "C" int dmmt_shard_status(dmmt_shard* s) {
    if (!s) return DMMT_E_INVALID;
    dmmt_plan* p = s->plan;
    DMMT_CUDA(cudaSetDevice(p->ctx->device));
    ImgMeta m{};
    DMMT_CUDA(cudaMemcpyAsync(&m, p->meta, sizeof m, cudaMemcpyDeviceToHost, p->stream));
    DMMT_CUDA(cudaStreamSynchronize(p->stream));
    return m.error;
}

extern "C" int dmmt_shard_launch_count(const dmmt_shard* s) { return s ? s->plan->last_launches : 0; }

// scan capacity of the shard's plan (see dmmt_plan_set_scan_capacity): DMMT_E_OVERFLOW from any phase means "grow and
// run the phases again from dmmt_shard_transform / dmmt_shard_launch_transform"
// ---- mailbox exchange (one process per GPU, no collective library on the data path) ----
extern "C" size_t dmmt_mailbox_bytes(int world) {
    return world > 0 && world <= DMMT_MAX_PEER_SHARDS ? mailbox_bytes(world, DMMT_MAILBOX_SLOTS) : 0;
}
extern "C" int dmmt_shard_launch_post(dmmt_shard* s, void* const* d_mailboxes, int rank, int world, int slot,
                                      unsigned long long seq, const void* d_src, int n_words64) {
    if (!s || !d_mailboxes || !d_src || world < 1 || world > DMMT_MAX_PEER_SHARDS || rank < 0 || rank >= world || slot < 0 ||
        slot >= DMMT_MAILBOX_SLOTS || n_words64 < 1 || n_words64 > 1024 || seq == 0)
        return DMMT_E_INVALID;
    dmmt_plan* p = s->plan;
    DMMT_CUDA(cudaSetDevice(p->ctx->device));
    PeerPtrs pp{};
    for (int r = 0; r < world; r++) {
        if (!d_mailboxes[r]) return DMMT_E_INVALID;
        pp.p[r] = d_mailboxes[r];
    }
    DMMT_CUDA(launch_mailbox_post(pp, world, rank, slot, seq, d_src, n_words64, p->stream));
    p->last_launches += 1;
    return DMMT_OK;
}
extern "C" int dmmt_shard_launch_collect(dmmt_shard* s, void* d_own_mailbox, int world, int slot, unsigned long long seq,
                                         int mode, int n_words64, int64_t* d_out) {
    if (!s || !d_own_mailbox || !d_out || world < 1 || world > DMMT_MAX_PEER_SHARDS || slot < 0 || slot >= DMMT_MAILBOX_SLOTS ||
        mode < 0 || mode > 2 || n_words64 < 1 || n_words64 > 1024 || (mode == 2 && n_words64 != 1) || seq == 0)
        return DMMT_E_INVALID;
    dmmt_plan* p = s->plan;
    DMMT_CUDA(cudaSetDevice(p->ctx->device));
    DMMT_CUDA(launch_mailbox_collect(d_own_mailbox, world, slot, seq, mode, n_words64, reinterpret_cast<long long*>(d_out),
                                     p->meta, p->stream));
    p->last_launches += 1;
    return DMMT_OK;
}

extern "C" size_t dmmt_shard_worst_case_scan_bytes(const dmmt_shard* s) {
    return s ? dmmt_plan_worst_case_scan_bytes(s->plan) : 0;
}
extern "C" int dmmt_shard_set_scan_capacity(dmmt_shard* s, size_t bytes) {
    if (!s) return DMMT_E_INVALID;
    return dmmt_plan_set_scan_capacity(s->plan, bytes);
}

// ---- single-process driver: shards on the given contexts (devices may repeat) ------------------
// Exchange buffers of one shard, in ITS device's memory; the other shards read them over peer access.
struct ShardXchg {
    int32_t last4[4];
    long long hist[1024];
    long long ghist[1024];
    long long bits;
    long long all_bits_offs[2 * DMMT_MAX_PEER_SHARDS];   // [0, n): bit counts, [n, 2n): exclusive bit offsets
    int32_t tail2[2];
    long long all_tail2[DMMT_MAX_PEER_SHARDS];           // {byte, bits} pairs as one 64-bit word each
    long long n_bytes;
    long long all_n_offs[2 * DMMT_MAX_PEER_SHARDS];      // [0, n): stuffed byte counts, [n, 2n): byte offsets in the file
    long long res2[2];
};

namespace {
struct ShardedJob {
    dmmt_ctx* const* ctxs = nullptr;
    int ns = 0;
    std::vector<dmmt_shard*> sh;
    std::vector<uint8_t*> d_px;
    std::vector<ShardXchg*> x;
    std::vector<cudaEvent_t> ev;   // one per shard, re-recorded phase by phase
    cudaEvent_t ev_hub = nullptr;  // on ctxs[0]: "every shard has reached this exchange"
    uint8_t* d_file = nullptr;
    size_t file_cap = 0;
    ~ShardedJob() {
        for (int i = 0; i < ns; i++) {
            (void)cudaSetDevice(ctxs[i]->device);
            (void)cudaStreamSynchronize(ctxs[i]->stream);
        }
        for (int i = 0; i < ns; i++) {
            (void)cudaSetDevice(ctxs[i]->device);
            if (i < (int)d_px.size() && d_px[i]) (void)cudaFree(d_px[i]);
            if (i < (int)x.size() && x[i]) (void)cudaFree(x[i]);
            if (i < (int)ev.size() && ev[i]) (void)cudaEventDestroy(ev[i]);
            if (i < (int)sh.size()) dmmt_shard_destroy(sh[i]);
        }
        if (d_file) {
            (void)cudaSetDevice(ctxs[0]->device);
            (void)cudaFree(d_file);
        }
        if (ev_hub) {
            (void)cudaSetDevice(ctxs[0]->device);
            (void)cudaEventDestroy(ev_hub);
        }
    }
};

// every pair of distinct devices of the job can read / write each other's memory
int enable_peer_access(dmmt_ctx* const* ctxs, int ns, bool* ok) {
    *ok = true;
    for (int i = 0; i < ns && *ok; i++)
        for (int j = 0; j < ns && *ok; j++) {
            const int di = ctxs[i]->device, dj = ctxs[j]->device;
            if (di == dj) continue;
            int can = 0;
            DMMT_CUDA(cudaDeviceCanAccessPeer(&can, di, dj));
            if (!can) {
                *ok = false;
                break;
            }
            DMMT_CUDA(cudaSetDevice(di));
            const cudaError_t e = cudaDeviceEnablePeerAccess(dj, 0);
            if (e == cudaErrorPeerAccessAlreadyEnabled) (void)cudaGetLastError();
            else if (e != cudaSuccess) DMMT_CUDA(e);
        }
    return DMMT_OK;
}

// stream of shard r waits for the event of shard `of` (recorded on another device's stream of this process)
int wait_for(ShardedJob& J, int r, int of) {
    DMMT_CUDA(cudaSetDevice(J.ctxs[r]->device));
    DMMT_CUDA(cudaStreamWaitEvent(J.ctxs[r]->stream, J.ev[of], 0));
    return DMMT_OK;
}
int record(ShardedJob& J, int r) {
    DMMT_CUDA(cudaSetDevice(J.ctxs[r]->device));
    DMMT_CUDA(cudaEventRecord(J.ev[r], J.ctxs[r]->stream));
    return DMMT_OK;
}
// one exchange step: every shard has recorded its event; every stream waits for all of them -- through a hub: the first
// shard's stream waits for the others and records one event they all wait for (2 n cross-device waits instead of n^2:
// at 8 devices the 320 waits of an encode were most of its 11 ms)
int all_wait_all(ShardedJob& J) {
    if (J.ns <= 2 || !J.ev_hub) {
        for (int r = 0; r < J.ns; r++)
            for (int o = 0; o < J.ns; o++)
                if (o != r) DMMT_TRY(wait_for(J, r, o));
        return DMMT_OK;
    }
    for (int o = 1; o < J.ns; o++) DMMT_TRY(wait_for(J, 0, o));
    DMMT_CUDA(cudaEventRecord(J.ev_hub, J.ctxs[0]->stream));   // wait_for left device 0 current
    for (int r = 1; r < J.ns; r++) {
        DMMT_CUDA(cudaSetDevice(J.ctxs[r]->device));
        DMMT_CUDA(cudaStreamWaitEvent(J.ctxs[r]->stream, J.ev_hub, 0));
    }
    return DMMT_OK;
}

// The five phases with every exchanged value in device memory: asynchronous launches on the contexts' streams,
// ordered across devices by events, the shards' K4 writing straight into the file on ctxs[0]'s device.  One host
// synchronisation at the end.  *err receives the first device-side error of any shard.
int sharded_run_peer(ShardedJob& J, uint64_t* file_len, int* err) {
    const int n = J.ns;
    PeerPtrs pp{};
    auto ptrs = [&](auto member) {
        for (int j = 0; j < n; j++) pp.p[j] = member(J.x[j]);
        return pp;
    };
    // phase 1 + exchange 1 (last DCs -> predictor seeds of the next shard, categorize.rs:157-161)
    for (int r = 0; r < n; r++) {
        DMMT_TRY(dmmt_shard_launch_transform(J.sh[r], J.d_px[r], J.x[r]->last4));
        DMMT_TRY(record(J, r));
    }
    // phase 2 + exchange 2 (sum of the histograms -> image-global tables, transformer.rs:201-217)
    for (int r = 0; r < n; r++) {
        if (r) DMMT_TRY(wait_for(J, r, r - 1));
        DMMT_TRY(dmmt_shard_launch_histogram(J.sh[r], r ? J.x[r - 1]->last4 : nullptr, reinterpret_cast<int64_t*>(J.x[r]->hist)));
    }
    for (int r = 0; r < n; r++) DMMT_TRY(record(J, r));
    DMMT_TRY(all_wait_all(J));
    ptrs([](ShardXchg* x) { return static_cast<const void*>(x->hist); });
    for (int r = 0; r < n; r++) {
        DMMT_CUDA(cudaSetDevice(J.ctxs[r]->device));
        DMMT_CUDA(launch_peer_exchange(pp, n, 1024, 0, J.x[r]->ghist, J.ctxs[r]->stream));
        // phase 3 + exchange 3 (bit counts -> global bit offsets)
        DMMT_TRY(dmmt_shard_launch_tables(J.sh[r], reinterpret_cast<const int64_t*>(J.x[r]->ghist), reinterpret_cast<int64_t*>(&J.x[r]->bits)));
        DMMT_TRY(record(J, r));
    }
    DMMT_TRY(all_wait_all(J));
    ptrs([](ShardXchg* x) { return static_cast<const void*>(&x->bits); });
    for (int r = 0; r < n; r++) {
        DMMT_CUDA(cudaSetDevice(J.ctxs[r]->device));
        DMMT_CUDA(launch_peer_exchange(pp, n, 1, 2, J.x[r]->all_bits_offs, J.ctxs[r]->stream));
        // phase 4 + exchange of the trailing partial bytes
        DMMT_TRY(dmmt_shard_launch_pack(J.sh[r], reinterpret_cast<const int64_t*>(J.x[r]->all_bits_offs + n + r), r == n - 1, J.x[r]->tail2));
        DMMT_TRY(record(J, r));
    }
    DMMT_TRY(all_wait_all(J));
    ptrs([](ShardXchg* x) { return static_cast<const void*>(x->tail2); });
    for (int r = 0; r < n; r++) {
        DMMT_CUDA(cudaSetDevice(J.ctxs[r]->device));
        DMMT_CUDA(launch_peer_exchange(pp, n, 1, 1, J.x[r]->all_tail2, J.ctxs[r]->stream));
        // phase 5a + exchange 4 (stuffed byte counts -> byte offsets in the file), BEFORE K4 runs
        DMMT_TRY(dmmt_shard_launch_count_bytes(J.sh[r], reinterpret_cast<const int32_t*>(J.x[r]->all_tail2),
                                               reinterpret_cast<const int64_t*>(J.x[r]->all_bits_offs + n),
                                               reinterpret_cast<const int64_t*>(J.x[r]->all_bits_offs), r, n,
                                               reinterpret_cast<int64_t*>(&J.x[r]->n_bytes)));
        DMMT_TRY(record(J, r));
    }
    DMMT_TRY(all_wait_all(J));
    ptrs([](ShardXchg* x) { return static_cast<const void*>(&x->n_bytes); });
    std::vector<long long> res((size_t)2 * n, 0);
    for (int r = 0; r < n; r++) {
        DMMT_CUDA(cudaSetDevice(J.ctxs[r]->device));
        DMMT_CUDA(launch_peer_exchange(pp, n, 1, 2, J.x[r]->all_n_offs, J.ctxs[r]->stream));
        // phase 5b: K4 into the file, at the shard's final byte offset (peer stores over NVLink for r > 0)
        DMMT_TRY(dmmt_shard_launch_stuff_into(J.sh[r], reinterpret_cast<const int64_t*>(J.x[r]->all_bits_offs + n), r, n, J.d_file,
                                              J.file_cap, reinterpret_cast<const int64_t*>(J.x[r]->all_n_offs + n + r),
                                              reinterpret_cast<int64_t*>(J.x[r]->res2)));
        DMMT_CUDA(cudaMemcpyAsync(&res[2 * r], J.x[r]->res2, 16, cudaMemcpyDeviceToHost, J.ctxs[r]->stream));
    }
    for (int r = 0; r < n; r++) {  // the only host synchronisation: completion + status of every shard
        DMMT_CUDA(cudaSetDevice(J.ctxs[r]->device));
        DMMT_CUDA(cudaStreamSynchronize(J.ctxs[r]->stream));
    }
    *err = DMMT_OK;
    for (int r = 0; r < n; r++)
        if (res[2 * r + 1]) {
            *err = (int)res[2 * r + 1];
            break;
        }
    *file_len = (uint64_t)res[2 * (n - 1)];
    return DMMT_OK;
}

// Fallback when some pair of devices has no peer access: exchanged values pass through the host (4 synchronisations),
// the shard outputs are gathered by D2H copies.  *err as above.
int sharded_run_host(ShardedJob& J, uint8_t** out, uint64_t* out_len, int* err) {
    const int ns = J.ns;
    *err = DMMT_OK;
#define SH_PHASE(expr)                              \
    do {                                            \
        const int rc__ = (expr);                    \
        if (rc__ > DMMT_OK || rc__ == DMMT_E_CUDA || rc__ == DMMT_E_NOMEM || rc__ == DMMT_E_INVALID) return rc__; \
        if (rc__ != DMMT_OK) {                      \
            *err = rc__;                            \
            return DMMT_OK;                         \
        }                                           \
    } while (0)
    std::vector<int16_t> last_dc((size_t)ns * 3), seed((size_t)ns * 3, 0);
    for (int r = 0; r < ns; r++) SH_PHASE(shard_transform_launch(J.sh[r], J.d_px[r]));
    for (int r = 0; r < ns; r++) SH_PHASE(shard_transform_collect(J.sh[r], &last_dc[3 * r]));
    for (int r = 1; r < ns; r++)
        for (int c = 0; c < 3; c++) seed[3 * r + c] = last_dc[3 * (r - 1) + c];
    std::vector<uint64_t> h((size_t)1024), g((size_t)1024, 0);
    for (int r = 0; r < ns; r++) SH_PHASE(shard_histogram_launch(J.sh[r], &seed[3 * r]));
    for (int r = 0; r < ns; r++) {
        SH_PHASE(shard_histogram_collect(J.sh[r], h.data()));
        for (int i = 0; i < 1024; i++) g[i] += h[i];
    }
    std::vector<uint64_t> bits((size_t)ns), bit_off((size_t)ns + 1, 0);
    for (int r = 0; r < ns; r++) SH_PHASE(shard_tables_launch(J.sh[r], g.data()));
    int first_err = DMMT_OK;
    for (int r = 0; r < ns; r++) {  // collect from EVERY shard: one of them overflowing must not strand the others
        const int rc = shard_tables_collect(J.sh[r], &bits[r]);
        if (rc != DMMT_OK && first_err == DMMT_OK) first_err = rc;
    }
    SH_PHASE(first_err);
    for (int r = 0; r < ns; r++) bit_off[r + 1] = bit_off[r] + bits[r];
    std::vector<uint8_t> tail((size_t)ns, 0);
    std::vector<int> tail_n((size_t)ns, 0);
    for (int r = 0; r < ns; r++) SH_PHASE(shard_pack_launch(J.sh[r], bit_off[r], r == ns - 1));
    for (int r = 0; r < ns; r++) SH_PHASE(shard_pack_collect(J.sh[r], r == ns - 1, &tail[r], &tail_n[r]));
    // a shard that does not complete a byte hands its predecessor's bits on
    for (int r = 1; r < ns; r++)
        if (r < ns - 1 && (bit_off[r] & 7) + bits[r] < 8) tail[r] |= tail[r - 1];
    std::vector<const uint8_t*> d_bytes((size_t)ns, nullptr);
    std::vector<uint64_t> n_bytes((size_t)ns, 0);
    for (int r = 0; r < ns; r++) SH_PHASE(shard_stuff_launch(J.sh[r], r ? tail[r - 1] : 0, r == 0, r == ns - 1));
    first_err = DMMT_OK;
    for (int r = 0; r < ns; r++) {
        const int rc = shard_stuff_collect(J.sh[r], &d_bytes[r], &n_bytes[r]);
        if (rc != DMMT_OK && first_err == DMMT_OK) first_err = rc;
    }
    SH_PHASE(first_err);
#undef SH_PHASE
    uint64_t total = 0;
    for (int r = 0; r < ns; r++) total += n_bytes[r];
    uint8_t* buf = static_cast<uint8_t*>(malloc(total ? total : 1));
    if (!buf) return DMMT_E_NOMEM;
    uint64_t at = 0;
    for (int r = 0; r < ns; r++) {
        cudaError_t e = cudaSetDevice(J.ctxs[r]->device);
        if (e == cudaSuccess && n_bytes[r])
            e = cudaMemcpyAsync(buf + at, d_bytes[r], n_bytes[r], cudaMemcpyDeviceToHost, J.ctxs[r]->stream);
        if (e != cudaSuccess) {
            dmmt_set_cuda_error(e, "gather of the shard outputs", __FILE__, __LINE__);
            free(buf);
            return DMMT_E_CUDA;
        }
        at += n_bytes[r];
    }
    for (int r = 0; r < ns; r++) {
        (void)cudaSetDevice(J.ctxs[r]->device);
        (void)cudaStreamSynchronize(J.ctxs[r]->stream);
    }
    *out = buf, *out_len = total;
    return DMMT_OK;
}
}  // namespace

static thread_local double g_sharded_last_ms = -1.0;
extern "C" double dmmt_encode_sharded_last_ms(void) { return g_sharded_last_ms; }

extern "C" int dmmt_encode_sharded(dmmt_ctx* const* ctxs, int nctx, const dmmt_image* im, const dmmt_options* o,
                                   uint8_t** jpeg, size_t* len) {
    if (!ctxs || nctx <= 0 || !im || !o || !jpeg || !len || !im->pixels || o->subsampling > DMMT_P420)
        return DMMT_E_INVALID;
    *jpeg = nullptr, *len = 0;
    if (im->width == 0 || im->height == 0) return DMMT_E_INVALID;
    if (im->pixels_on_device) return DMMT_E_INVALID;  // the shards live on different devices: host pixels only
    for (int i = 0; i < nctx; i++)
        if (!ctxs[i]) return DMMT_E_INVALID;
    const int rows = mcu_rows_of(im->height, o->subsampling);
    const int ns = std::min(std::min(nctx, rows), DMMT_MAX_PEER_SHARDS);
    ShardedJob J;
    J.ctxs = ctxs, J.ns = ns;
    J.sh.assign((size_t)ns, nullptr), J.d_px.assign((size_t)ns, nullptr), J.x.assign((size_t)ns, nullptr);
    J.ev.assign((size_t)ns, nullptr);
    bool peer = false;
    DMMT_TRY(enable_peer_access(ctxs, ns, &peer));
    DMMT_CUDA(cudaSetDevice(ctxs[0]->device));
    DMMT_CUDA(cudaEventCreateWithFlags(&J.ev_hub, cudaEventDisableTiming));
    if (const char* e = getenv("DMMT_SHARDED_HOST_EXCHANGE")) peer = peer && e[0] != '1';  // tests: force the fallback
    // MCU rows [r*rows/ns, (r+1)*rows/ns)
    for (int r = 0; r < ns; r++) {
        const int b = (int)((long long)r * rows / ns), e = (int)((long long)(r + 1) * rows / ns);
        DMMT_TRY(dmmt_shard_create(ctxs[r], im->width, im->height, im->fmt, im->max_value, o, b, e, &J.sh[r]));
        DMMT_CUDA(cudaSetDevice(ctxs[r]->device));
        DMMT_CUDA(cudaMalloc(&J.d_px[r], dmmt_shard_pixel_bytes(J.sh[r])));
        DMMT_CUDA(cudaMalloc(&J.x[r], sizeof(ShardXchg)));
        // the shard's own output slot (header, non-peer stuffing) now, not lazily inside the phases: cudaMalloc
        // synchronises its device, which serialised the shards (8 devices: 11 ms instead of 1)
        if (!J.sh[r]->plan->d_out_own) DMMT_CUDA(cudaMalloc(&J.sh[r]->plan->d_out_own, J.sh[r]->plan->out_stride));
        DMMT_CUDA(cudaEventCreateWithFlags(&J.ev[r], cudaEventDisableTiming));
        DMMT_CUDA(cudaMemcpyAsync(J.d_px[r], static_cast<const uint8_t*>(im->pixels) + dmmt_shard_pixel_offset(J.sh[r]),
                                  dmmt_shard_pixel_bytes(J.sh[r]), cudaMemcpyHostToDevice, ctxs[r]->stream));
    }
    for (int attempt = 0; attempt < 2; attempt++) {
        int err = DMMT_OK;
        uint8_t* buf = nullptr;
        uint64_t total = 0;
        if (peer) {
            if (!J.d_file) {
                for (int r = 0; r < ns; r++) J.file_cap += dmmt_shard_out_stride(J.sh[r]);
                DMMT_CUDA(cudaSetDevice(ctxs[0]->device));
                DMMT_CUDA(cudaMalloc(&J.d_file, J.file_cap));
            }
            // measurement hook (DMMT_SHARDED_TIMING=1): wall clock of the phase section alone -- every H2D copy done
            // before it starts, the one host synchronisation of sharded_run_peer at its end
            static const bool timing = [] {
                const char* e = getenv("DMMT_SHARDED_TIMING");
                return e && e[0] == '1';
            }();
            std::chrono::steady_clock::time_point t0;
            if (timing) {
                for (int r = 0; r < ns; r++) {
                    DMMT_CUDA(cudaSetDevice(ctxs[r]->device));
                    DMMT_CUDA(cudaStreamSynchronize(ctxs[r]->stream));
                }
                t0 = std::chrono::steady_clock::now();
            }
            DMMT_TRY(sharded_run_peer(J, &total, &err));
            if (timing) g_sharded_last_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
            if (err == DMMT_OK) {
                buf = static_cast<uint8_t*>(malloc(total ? total : 1));
                if (!buf) return DMMT_E_NOMEM;
                DMMT_CUDA(cudaSetDevice(ctxs[0]->device));
                const cudaError_t e = cudaMemcpy(buf, J.d_file, total, cudaMemcpyDeviceToHost);
                if (e != cudaSuccess) {
                    free(buf);
                    DMMT_CUDA(e);
                }
            }
        } else {
            DMMT_TRY(sharded_run_host(J, &buf, &total, &err));
        }
        if (err == DMMT_OK) {
            *jpeg = buf, *len = (size_t)total;
            return DMMT_OK;
        }
        if (err != DMMT_E_OVERFLOW || attempt == 1) return err;
        // denser than the default 128 B per block somewhere (the reference encodes any input): every shard gets the
        // worst-case capacity and the phases run again from the transform
        for (int r = 0; r < ns; r++) {
            DMMT_CUDA(cudaSetDevice(ctxs[r]->device));
            DMMT_CUDA(cudaStreamSynchronize(ctxs[r]->stream));
        }
        for (int r = 0; r < ns; r++) DMMT_TRY(dmmt_shard_set_scan_capacity(J.sh[r], dmmt_shard_worst_case_scan_bytes(J.sh[r])));
        if (J.d_file) {
            DMMT_CUDA(cudaSetDevice(ctxs[0]->device));
            DMMT_CUDA(cudaFree(J.d_file));
            J.d_file = nullptr, J.file_cap = 0;
        }
    }
    return DMMT_E_OVERFLOW;
}
