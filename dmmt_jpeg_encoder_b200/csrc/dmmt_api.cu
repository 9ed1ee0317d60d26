// dmmt_api.cu -- the C ABI of include/dmmt_cuda.h: contexts, plans (one launch chain over n
// equally sized images), single-image encode, and the debug/measurement hooks.
//
// Replaces the body of JpegImageWriter::write_image (src/image/writer/jpeg.rs:64-75 of the
// reference): Transformer::transform (transformer.rs:188-221) = K1 + K2 + K2b, Encoder::encode
// (encoder.rs:125-135) = K2b's header part + K3 + K4.  There is no CPU path in this file: every
// computing entry point needs a CUDA device and fails with DMMT_E_NODEVICE / DMMT_E_CUDA otherwise.
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>

#include "dmmt_internal.h"
#include "dmmt_qtables.h"
#include "ppm_parse.hpp"

using namespace dmmt;

// ------------------------------------------------------------------------------------------
static thread_local char g_cuda_err[512] = "";

void dmmt_set_cuda_error(cudaError_t e, const char* what, const char* file, int line) {
    const char* base = strrchr(file, '/');
    snprintf(g_cuda_err, sizeof g_cuda_err, "%s (%s) at %s:%d: %s", cudaGetErrorName(e),
             cudaGetErrorString(e), base ? base + 1 : file, line, what);
    (void)cudaGetLastError();  // clear the sticky-less error state
}

extern "C" const char* dmmt_last_cuda_error(void) { return g_cuda_err; }

extern "C" const char* dmmt_strerror(int code) {
    switch (code) {
        case DMMT_OK: return "ok";
        case DMMT_E_INVALID: return "invalid argument (null / zero size / unknown preset / sample above max value)";
        case DMMT_E_NODEVICE: return "no usable CUDA device (this library has no CPU fallback)";
        case DMMT_E_CUDA: return "CUDA runtime failure";
        case DMMT_E_NCCL: return "collective failure";
        case DMMT_E_NOMEM: return "out of memory";
        case DMMT_E_OVERFLOW: return "entropy-coded scan exceeds the plan's scan capacity";
        case DMMT_E_SYMBOL: return "Huffman symbol not present in translator";
        case DMMT_E_RANGE: return "coefficient cannot be categorized";
        case DMMT_E_WRITE: return "failed to write image data (output arena too small)";
        case DMMT_E_SIZE: return "image geometry exceeds the u16 image model";
        default: return "unknown error";
    }
}

extern "C" void dmmt_free(void* p) { free(p); }

// ---- host ingest (ppm_parse.hpp): no device involved
extern "C" int dmmt_ppm_parse(const char* text, size_t len, int threads, uint16_t* width, uint16_t* height,
                              uint16_t* max_value, uint16_t** samples, size_t* n_samples, int* detail) {
    if (!text || !width || !height || !max_value || !samples || !n_samples) return DMMT_E_INVALID;
    *samples = nullptr, *n_samples = 0;
    dmmt_ppm::Result r;
    try {
        r = dmmt_ppm::parse(text, len, threads > 1 ? (unsigned)threads : 1u);
    } catch (const std::bad_alloc&) {
        return DMMT_E_NOMEM;
    } catch (...) {
        return DMMT_E_INVALID;
    }
    if (detail) *detail = r.detail;
    *width = r.width, *height = r.height, *max_value = r.max_value;
    if (r.status != dmmt_ppm::OK) return (int)r.status;
    *n_samples = r.samples.size();
    *samples = r.samples.release();  // malloc'd by the tokenizer: dmmt_free releases it
    return DMMT_OK;
}

extern "C" const char* dmmt_ppm_strerror(int status, int detail, char* buf, size_t cap) {
    if (!buf || !cap) return buf;
    static const char* const kTok[5] = {"P3 Header", "Width Header", "Height Header", "Max Value Header", "Color Component Value"};
    const char* tok = kTok[detail >= 0 && detail < 5 ? detail : 4];
    switch (status) {
        case DMMT_OK: snprintf(buf, cap, "ok"); break;
        case DMMT_PPM_MISSING_TOKEN: snprintf(buf, cap, "Expected token '%s' not found in PPM file", tok); break;
        case DMMT_PPM_BAD_TOKEN: snprintf(buf, cap, "Parsing of token '%s' failed", tok); break;
        case DMMT_PPM_INCOMPLETE_PIXEL: snprintf(buf, cap, "Incomplete pixel parsed. Expected 3 components, but got %d.", detail); break;
        case DMMT_PPM_SIZE_MISMATCH: snprintf(buf, cap, "Nubmer of pixels do not match the size, provided in header"); break;
        case DMMT_PPM_SAMPLE_ABOVE_MAX: snprintf(buf, cap, "color component exceeds the max value (color.rs:62-65)"); break;
        default: snprintf(buf, cap, "%s", dmmt_strerror(status)); break;
    }
    return buf;
}

extern "C" int dmmt_host_alloc(size_t bytes, void** out) {
    if (!out) return DMMT_E_INVALID;
    *out = nullptr;
    DMMT_CUDA(cudaHostAlloc(out, bytes ? bytes : 1, cudaHostAllocDefault));
    return DMMT_OK;
}
extern "C" void dmmt_host_free(void* p) {
    if (p) (void)cudaFreeHost(p);
}

// ------------------------------------------------------------------------------------------
// contexts
extern "C" int dmmt_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        (void)cudaGetLastError();
        return 0;
    }
    return n;
}

static int ctx_create_impl(int device, cudaStream_t st, bool own, dmmt_ctx** out) {
    if (!out) return DMMT_E_INVALID;
    *out = nullptr;
    const int n = dmmt_device_count();
    if (n <= 0) return DMMT_E_NODEVICE;
    if (device < 0 || device >= n) return DMMT_E_INVALID;
    DMMT_CUDA(cudaSetDevice(device));
    dmmt_ctx* c = new (std::nothrow) dmmt_ctx();
    if (!c) return DMMT_E_NOMEM;
    c->device = device;
    if (own) {
        cudaError_t e = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking);
        if (e != cudaSuccess) {
            dmmt_set_cuda_error(e, "cudaStreamCreateWithFlags", __FILE__, __LINE__);
            delete c;
            return DMMT_E_CUDA;
        }
        c->own_stream = true;
    } else {
        c->stream = st;
    }
    *out = c;
    return DMMT_OK;
}

extern "C" int dmmt_ctx_create(int device, dmmt_ctx** out) { return ctx_create_impl(device, nullptr, true, out); }
extern "C" int dmmt_ctx_create_on_stream(int device, void* cuda_stream, dmmt_ctx** out) {
    return ctx_create_impl(device, static_cast<cudaStream_t>(cuda_stream), false, out);
}
extern "C" void dmmt_ctx_destroy(dmmt_ctx* c) {
    if (!c) return;
    (void)cudaSetDevice(c->device);
    for (dmmt_plan* p : c->cache) dmmt_plan_destroy(p);
    if (c->own_stream && c->stream) (void)cudaStreamDestroy(c->stream);
    delete c;
}
extern "C" int dmmt_ctx_synchronize(dmmt_ctx* c) {
    if (!c) return DMMT_E_INVALID;
    DMMT_CUDA(cudaSetDevice(c->device));
    DMMT_CUDA(cudaStreamSynchronize(c->stream));
    return DMMT_OK;
}
extern "C" void* dmmt_ctx_stream(dmmt_ctx* c) { return c ? c->stream : nullptr; }

// ------------------------------------------------------------------------------------------
// plans
static size_t fmt_bytes(int fmt) { return fmt == DMMT_RGB_U8 ? 3 : (fmt == DMMT_RGB_U16 ? 6 : 12); }
static size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

static void plan_free_scratch(dmmt_plan* p) {
    (void)cudaFree(p->coef), p->coef = nullptr;
    (void)cudaFree(p->zero_region), p->zero_region = nullptr;
    (void)cudaFree(p->enc), p->enc = nullptr;
    (void)cudaFree(p->lens), p->lens = nullptr;
    (void)cudaFree(p->lcount), p->lcount = nullptr;
    (void)cudaFree(p->scan), p->scan = nullptr;
    (void)cudaFree(p->tb.tok), p->tb.tok = nullptr;
    (void)cudaFree(p->tb.ntok), p->tb.ntok = nullptr;
    (void)cudaFree(p->fo.last_dc), p->fo.last_dc = nullptr;
    (void)cudaFree(p->fo.dcpos), p->fo.dcpos = nullptr;
    (void)cudaFree(p->d_lens), p->d_lens = nullptr;
    (void)cudaFree(p->d_offsets), p->d_offsets = nullptr;
    (void)cudaFree(p->d_seed_dc), p->d_seed_dc = nullptr;
    (void)cudaFree(p->d_ghist), p->d_ghist = nullptr;
    (void)cudaFree(p->d_pixels_own), p->d_pixels_own = nullptr;
    (void)cudaFree(p->d_out_own), p->d_out_own = nullptr;
    (void)cudaFree(p->d_dense), p->d_dense = nullptr;
    p->dense_cap = 0;
    if (p->h_lens) (void)cudaFreeHost(p->h_lens), p->h_lens = nullptr;
    if (p->h_offsets) (void)cudaFreeHost(p->h_offsets), p->h_offsets = nullptr;
}

// (re)allocates everything whose size depends on the scan capacity.  All-or-nothing: the new buffers are allocated
// first and swapped in, sizes and pointers together, only when every allocation has succeeded; on failure the plan
// keeps its old buffers and capacity (a plan with new sizes and null buffers would fault on the next encode).
static int plan_alloc_scan(dmmt_plan* p, size_t scan_cap_bytes) {
    dmmt_plan_drop_graph(p);  // a captured chain points at the old buffers (before the copy: `n` must not carry the handle)
    dmmt_plan n = *p;         // the sizes of the new layout are computed on a copy
    n.scan_cap_bytes = align_up(std::max<size_t>(scan_cap_bytes, 64), 64);
    n.scan_stride_words = n.scan_cap_bytes / 4 + 32;  // + slack: look-ahead loads of K4, seed byte
    // token stream: 32 tokens per block by default (the bytes of the coefficient stream); the worst
    // case is 65 per block (68 allocated) once the scan capacity has been grown to the worst case.
    // The same buffer serves the generic path (regions = K2's 256-block chunks) and the fused 4:2:0
    // path (regions = K1's 96-block tiles).
    const bool worst = n.scan_cap_bytes >= (size_t)n.g.n_blocks * 209;
    // output slot: header + stuffed scan + EOI.  By default the stuffing is given 1/8 of the scan; once the scan
    // capacity is the true maximum the slot must hold a scan of nothing but 0xFF bytes too (every byte doubled)
    n.out_stride = align_up(1024 + n.scan_cap_bytes + (worst ? n.scan_cap_bytes : n.scan_cap_bytes / 8) + 64, 16);
    n.max_chunks4 = k4_max_chunks(n.scan_cap_bytes + 8);
    const uint32_t per_block = worst ? 68u : 32u;
    n.tb.chunk_cap = tok_blocks_per_chunk() * per_block;
    n.fused = k1_fused_supported(n.g, n.k1c) && !n.force_generic;
    n.fo.tiles_x = k1_tiles_x(n.g);
    n.fo.tiles = n.fo.tiles_x * (uint32_t)n.g.mcus_y;
    n.fo.tile_cap = 96u * per_block;
    n.n_chunks3f = (n.fo.tiles + 7) / 8;
    const size_t words_generic = (size_t)n.n_chunks3 * n.tb.chunk_cap;
    const size_t words_fused = (size_t)n.fo.tiles * n.fo.tile_cap;
    const size_t img_words = std::max(words_generic, words_fused);
    n.tb.img_stride_words = img_words;
    n.fo.img_stride_words = img_words;
    const size_t n_regions = std::max<size_t>(n.n_chunks3, n.fo.tiles);
    // one zero-initialised region per run: hist | meta | lb3 | lb4 | tk3 | tk4
    const size_t o_hist = 0;
    const size_t o_meta = o_hist + align_up((size_t)n.n * 1024 * sizeof(unsigned int), 16);
    const size_t o_lb3 = o_meta + align_up((size_t)n.n * sizeof(ImgMeta), 16);
    const size_t o_lb4 = o_lb3 + (size_t)n.n * std::max(n.n_chunks3, n.fo.tiles) * 8;  // K3: per chunk, or per tile (fused)
    const size_t o_tk3 = o_lb4 + (size_t)n.n * n.max_chunks4 * 8;
    const size_t o_tk4 = o_tk3 + align_up((size_t)n.n * 4, 16);
    n.zero_bytes = o_tk4 + align_up((size_t)n.n * 4, 16);

    uint32_t *scan = nullptr, *tok = nullptr, *ntok = nullptr, *dcpos = nullptr;
    int16_t* last_dc = nullptr;
    uint8_t* zero_region = nullptr;
    cudaError_t e = cudaMalloc(&scan, (size_t)n.n * n.scan_stride_words * 4);
    if (e == cudaSuccess) e = cudaMalloc(&tok, (size_t)n.n * img_words * 4);
    if (e == cudaSuccess) e = cudaMalloc(&ntok, (size_t)n.n * n_regions * 4);
    if (e == cudaSuccess) e = cudaMalloc(&last_dc, (size_t)n.n * n.fo.tiles * 4 * sizeof(int16_t));
    if (e == cudaSuccess) e = cudaMalloc(&dcpos, (size_t)n.n * n.fo.tiles * 2 * sizeof(uint32_t));
    if (e == cudaSuccess) e = cudaMalloc(&zero_region, n.zero_bytes);
    if (e != cudaSuccess) {
        dmmt_set_cuda_error(e, "cudaMalloc of the plan's scan / token scratch", __FILE__, __LINE__);
        (void)cudaFree(scan), (void)cudaFree(tok), (void)cudaFree(ntok), (void)cudaFree(last_dc), (void)cudaFree(dcpos);
        (void)cudaFree(zero_region);
        return e == cudaErrorMemoryAllocation ? DMMT_E_NOMEM : DMMT_E_CUDA;
    }
    // success: release the old buffers (and the arenas sized by the old out_stride) and install the new layout
    (void)cudaFree(p->scan), (void)cudaFree(p->tb.tok), (void)cudaFree(p->tb.ntok), (void)cudaFree(p->zero_region);
    (void)cudaFree(p->fo.last_dc), (void)cudaFree(p->fo.dcpos), (void)cudaFree(p->d_out_own), (void)cudaFree(p->d_dense);
    *p = n;
    p->d_out_own = nullptr, p->d_dense = nullptr, p->dense_cap = 0;
    p->scan = scan;
    p->tb.tok = p->fo.tok = tok;
    p->tb.ntok = p->fo.ntok = ntok;
    p->fo.last_dc = last_dc, p->fo.dcpos = dcpos;
    p->zero_region = zero_region;
    p->hist = reinterpret_cast<unsigned int*>(p->zero_region + o_hist);
    p->meta = reinterpret_cast<ImgMeta*>(p->zero_region + o_meta);
    p->lb3 = reinterpret_cast<unsigned long long*>(p->zero_region + o_lb3);
    p->lb4 = reinterpret_cast<unsigned long long*>(p->zero_region + o_lb4);
    p->tk3 = reinterpret_cast<unsigned int*>(p->zero_region + o_tk3);
    p->tk4 = reinterpret_cast<unsigned int*>(p->zero_region + o_tk4);
    return DMMT_OK;
}

static int make_geom(int W, int H, int subsampling, int mcus_y_override, Geom* g) {
    if (W <= 0 || H <= 0) return DMMT_E_INVALID;
    switch (subsampling) {  // subsampling.rs:32-46
        case DMMT_P444: g->hr = 1, g->vr = 1; break;
        case DMMT_P422: g->hr = 2, g->vr = 1; break;
        case DMMT_P420: g->hr = 2, g->vr = 2; break;
        default: return DMMT_E_INVALID;
    }
    g->W = W, g->H = H;
    g->mcus_x = (W + 8 * g->hr - 1) / (8 * g->hr);  // padder.rs:13-14, transformer.rs:48-51
    g->mcus_y = mcus_y_override > 0 ? mcus_y_override : (H + 8 * g->vr - 1) / (8 * g->vr);
    g->ypm = g->hr * g->vr;
    g->bpm = g->ypm + 2;
    const unsigned long long n_mcus = (unsigned long long)g->mcus_x * g->mcus_y;
    if (n_mcus * g->bpm > 0xFFFFFFFFull) return DMMT_E_SIZE;
    g->n_mcus = (uint32_t)n_mcus;
    g->n_blocks = (uint32_t)(n_mcus * g->bpm);
    return DMMT_OK;
}

int dmmt_plan_create_impl(dmmt_ctx* ctx, int W, int H_rows, int mcus_y_override, int sof_W, int sof_H,
                          int fmt, int max_value, const dmmt_options* opt, int n_images, cudaStream_t st,
                          bool own_stream, dmmt_plan** out) {
    if (!ctx || !opt || !out || n_images <= 0) return DMMT_E_INVALID;
    *out = nullptr;
    if (fmt != DMMT_RGB_F32_NORM && fmt != DMMT_RGB_U8 && fmt != DMMT_RGB_U16) return DMMT_E_INVALID;
    if (opt->qtable_preset > DMMT_Q_AN_IMPROVED_DETECTION_MODEL) return DMMT_E_INVALID;
    if (fmt != DMMT_RGB_F32_NORM && max_value <= 0) return DMMT_E_INVALID;
    if (fmt == DMMT_RGB_U8 && max_value > 255) return DMMT_E_INVALID;
    Geom g{};
    DMMT_TRY(make_geom(W, H_rows, opt->subsampling, mcus_y_override, &g));
    // PaddedImage keeps u16 padded sizes (padder.rs:6-7): larger geometries wrap in the reference
    if (g.mcus_x * 8 * g.hr > 65535 || sof_W > 65535 || sof_H > 65535 ||
        (mcus_y_override <= 0 && g.mcus_y * 8 * g.vr > 65535))
        return DMMT_E_SIZE;
    DMMT_CUDA(cudaSetDevice(ctx->device));
    dmmt_plan* p = new (std::nothrow) dmmt_plan();
    if (!p) return DMMT_E_NOMEM;
    p->ctx = ctx;
    p->stream = st;
    p->own_stream = own_stream;
    p->g = g;
    p->W = W, p->H = H_rows, p->sof_W = sof_W, p->sof_H = sof_H;
    p->fmt = fmt, p->max_value = max_value, p->opt = *opt, p->n = n_images;
    p->pixel_bytes = (size_t)W * H_rows * fmt_bytes(fmt);
    p->coef_stride = (size_t)g.n_blocks * 64;
    p->n_chunks3 = k3_chunks(g);
    make_k1_consts(fmt, max_value, kQuantPresets[opt->qtable_preset][0], kQuantPresets[opt->qtable_preset][1], &p->k1c);
    int rc = DMMT_OK;
    auto fail = [&](int code) {
        plan_free_scratch(p);
        if (p->own_stream && p->stream) (void)cudaStreamDestroy(p->stream);
        delete p;
        return code;
    };
#define PLAN_CUDA(expr)                                          \
    do {                                                         \
        cudaError_t e__ = (expr);                                \
        if (e__ != cudaSuccess) {                                \
            dmmt_set_cuda_error(e__, #expr, __FILE__, __LINE__); \
            return fail(e__ == cudaErrorMemoryAllocation ? DMMT_E_NOMEM : DMMT_E_CUDA); \
        }                                                        \
    } while (0)
    // the coefficient stream (6 B per padded pixel) only exists on the generic path: a fused 4:2:0 plan allocates it
    // when somebody asks for that path (dmmt_plan_set_generic_path, dmmt_plan_debug_dct)
    if (!k1_fused_supported(g, p->k1c)) PLAN_CUDA(cudaMalloc(&p->coef, (size_t)n_images * p->coef_stride * sizeof(int16_t)));
    PLAN_CUDA(cudaMalloc(&p->enc, (size_t)n_images * sizeof(EncTables)));
    PLAN_CUDA(cudaMalloc(&p->lens, (size_t)n_images * sizeof(LenTables)));
    PLAN_CUDA(cudaMalloc(&p->lcount, (size_t)n_images * 64));
    PLAN_CUDA(cudaMalloc(&p->d_lens, (size_t)n_images * 8));
    PLAN_CUDA(cudaMalloc(&p->d_offsets, (size_t)(n_images + 1) * 8));
    PLAN_CUDA(cudaMalloc(&p->d_seed_dc, 8 * sizeof(int16_t)));
    p->d_last_dc = p->d_seed_dc + 4;
    PLAN_CUDA(cudaMalloc(&p->d_ghist, 1024 * 8));
    PLAN_CUDA(cudaHostAlloc(&p->h_lens, (size_t)n_images * 8, cudaHostAllocDefault));
    PLAN_CUDA(cudaHostAlloc(&p->h_offsets, (size_t)(n_images + 1) * 8, cudaHostAllocDefault));
#undef PLAN_CUDA
    rc = plan_alloc_scan(p, (size_t)g.n_blocks * 128);
    if (rc != DMMT_OK) return fail(rc);
    *out = p;
    return DMMT_OK;
}

extern "C" int dmmt_plan_create(dmmt_ctx* ctx, uint16_t width, uint16_t height, dmmt_fmt fmt,
                                uint16_t max_value, const dmmt_options* opt, int n_images, dmmt_plan** out) {
    if (!ctx) return DMMT_E_INVALID;
    return dmmt_plan_create_impl(ctx, width, height, 0, width, height, (int)fmt, max_value, opt, n_images,
                                 ctx->stream, false, out);
}

extern "C" void dmmt_plan_destroy(dmmt_plan* p) {
    if (!p) return;
    (void)cudaSetDevice(p->ctx->device);
    (void)cudaStreamSynchronize(p->stream);
    dmmt_plan_drop_graph(p);
    plan_free_scratch(p);
    if (p->ev_valid)
        for (auto& e : p->ev) (void)cudaEventDestroy(e);
    if (p->own_stream && p->stream) (void)cudaStreamDestroy(p->stream);
    delete p;
}

extern "C" size_t dmmt_plan_pixel_bytes(const dmmt_plan* p) { return p ? p->pixel_bytes : 0; }
extern "C" size_t dmmt_plan_out_stride(const dmmt_plan* p) { return p ? p->out_stride : 0; }
extern "C" size_t dmmt_plan_stream_blocks(const dmmt_plan* p) { return p ? p->g.n_blocks : 0; }
extern "C" size_t dmmt_plan_worst_case_scan_bytes(const dmmt_plan* p) {
    // per block: DC code <= 16 + 11 bits, 63 x (AC code <= 16 + 10 bits) = 1665 bits -> 209 B
    return p ? (size_t)p->g.n_blocks * 209 + 64 : 0;
}
extern "C" int dmmt_plan_set_scan_capacity(dmmt_plan* p, size_t bytes_per_image) {
    if (!p || bytes_per_image == 0) return DMMT_E_INVALID;
    DMMT_CUDA(cudaSetDevice(p->ctx->device));
    DMMT_CUDA(cudaStreamSynchronize(p->stream));
    return plan_alloc_scan(p, bytes_per_image);
}

extern "C" int dmmt_plan_uses_fused_path(const dmmt_plan* p) { return p && p->fused ? 1 : 0; }

static int plan_ensure_coef(dmmt_plan* p) {
    if (!p->coef) DMMT_CUDA(cudaMalloc(&p->coef, (size_t)p->n * p->coef_stride * sizeof(int16_t)));
    return DMMT_OK;
}

extern "C" int dmmt_plan_set_generic_path(dmmt_plan* p, int generic) {
    if (!p) return DMMT_E_INVALID;
    DMMT_CUDA(cudaSetDevice(p->ctx->device));
    DMMT_CUDA(cudaStreamSynchronize(p->stream));
    if (generic) DMMT_TRY(plan_ensure_coef(p));
    dmmt_plan_drop_graph(p);
    p->force_generic = generic ? 1 : 0;
    p->fused = k1_fused_supported(p->g, p->k1c) && !p->force_generic;
    return DMMT_OK;
}

extern "C" int dmmt_plan_set_profiling(dmmt_plan* p, int enabled) {
    if (!p) return DMMT_E_INVALID;
    DMMT_CUDA(cudaSetDevice(p->ctx->device));
    if (enabled && !p->ev_valid) {
        for (auto& e : p->ev) DMMT_CUDA(cudaEventCreate(&e));
        p->ev_valid = true;
    }
    p->profiling = enabled != 0;
    return DMMT_OK;
}

// The launch chain.  seed_dc/seed_bits/ghist/flags are only used by the shard phases.
int dmmt_plan_chain(dmmt_plan* p, const void* d_pixels, int n, uint8_t* d_out, unsigned long long* d_lens) {
    cudaStream_t st = p->stream;
    const bool prof = p->profiling && p->ev_valid;
    int launches = 0;
    auto mark = [&](int i) -> cudaError_t { return prof ? cudaEventRecord(p->ev[i], st) : cudaSuccess; };
    DMMT_CUDA(cudaMemsetAsync(p->zero_region, 0, p->zero_bytes, st));
    DMMT_CUDA(mark(0));
    const int check_max = (p->fmt == DMMT_RGB_U8 && p->max_value < 255) ||
                          (p->fmt == DMMT_RGB_U16 && p->max_value < 65535);
    DMMT_CUDA(launch_k1(p->g, p->fmt, p->k1c, check_max, d_pixels, p->pixel_bytes, n, p->coef, p->coef_stride, nullptr,
                        p->meta, p->fused ? &p->fo : nullptr, p->hist, st));
    launches += 1;
    DMMT_CUDA(mark(1));
    // fused path: the tile-boundary DC tokens are finished by the DC-table CTAs of K2b (no kernel of its own)
    if (!p->fused) {
        DMMT_CUDA(launch_k2(p->g, p->coef, p->coef_stride, n, p->hist, p->meta, nullptr, p->tb, st));
        launches += 1;
    }
    DMMT_CUDA(mark(2));
    K2bHostArgs b{};
    b.hist = p->hist, b.ghist = nullptr, b.enc = p->enc, b.lens = p->lens, b.meta = p->meta;
    b.out = d_out, b.out_stride = p->out_stride;
    b.scan_cap_bits = (unsigned long long)p->scan_cap_bytes * 8;
    b.W = p->sof_W, b.H = p->sof_H, b.bits_per_channel = p->opt.bits_per_channel;
    b.qtab_luma = kQuantPresets[p->opt.qtable_preset][0];
    b.qtab_chroma = kQuantPresets[p->opt.qtable_preset][1];
    b.write_header = 1;
    b.lcount = p->lcount;
    b.fix = p->fused ? &p->fo : nullptr;
    DMMT_CUDA(launch_k2b(p->g, b, n, st));
    launches += 1;
    DMMT_CUDA(mark(3));
    const int zero_blocks = (int)std::min<size_t>(std::max<size_t>(p->scan_cap_bytes / 65536, 1), 128);
    DMMT_CUDA(launch_zero_scan(p->scan, p->scan_stride_words, p->meta, n, 0ull, zero_blocks, st));
    {
        TokBuf tb = p->tb;
        if (p->fused) tb.chunk_cap = p->fo.tile_cap;
        DMMT_CUDA(launch_k3(p->fused ? p->n_chunks3f : p->n_chunks3, p->fused ? p->fo.tiles : 0u, n, tb, p->enc, p->meta,
                            p->lb3, p->tk3, p->scan, p->scan_stride_words, 0ull, 1, st));
    }
    launches += 2;
    DMMT_CUDA(mark(4));
    K4HostArgs k{};
    k.scan = reinterpret_cast<const uint8_t*>(p->scan), k.scan_stride_bytes = p->scan_stride_words * 4;
    k.meta = p->meta, k.lb_state = p->lb4, k.ticket = p->tk4, k.max_chunks = p->max_chunks4;
    k.out = d_out, k.out_stride = p->out_stride, k.out_lens = d_lens;
    k.first_byte = 0, k.n_bytes_override = -1, k.seed_bits = 0, k.prepend_header = 1, k.append_eoi = 1;
    k.or_first_byte = 0;
    DMMT_CUDA(launch_k4(k, n, p->max_chunks4, st));
    launches += 1;
    DMMT_CUDA(mark(5));
    p->last_launches = launches;
    p->last_n = n;
    return DMMT_OK;
}

void dmmt_plan_drop_graph(dmmt_plan* p) {
    if (p->gexec) (void)cudaGraphExecDestroy(p->gexec);
    p->gexec = nullptr;
    p->g_pixels = nullptr, p->g_out = nullptr, p->g_lens = nullptr, p->g_n = 0;
}

int dmmt_plan_chain_replay(dmmt_plan* p, const void* d_pixels, int n, uint8_t* d_out, unsigned long long* d_lens) {
    const bool same = p->g_pixels == d_pixels && p->g_n == n && p->g_out == d_out && p->g_lens == d_lens;
    if (!p->graphs || (p->profiling && p->ev_valid)) return dmmt_plan_chain(p, d_pixels, n, d_out, d_lens);
    if (p->gexec && same) {
        DMMT_CUDA(cudaGraphLaunch(p->gexec, p->stream));
        p->last_launches = p->g_launches;
        p->last_n = n;
        return DMMT_OK;
    }
    if (p->gexec) dmmt_plan_drop_graph(p);
    if (!same) {  // first call with these arguments: plain launches (this also runs every one-time initialisation)
        p->g_pixels = d_pixels, p->g_n = n, p->g_out = d_out, p->g_lens = d_lens;
        return dmmt_plan_chain(p, d_pixels, n, d_out, d_lens);
    }
    // second call with the same arguments: capture the chain, instantiate, launch
    cudaGraph_t graph = nullptr;
    if (cudaStreamBeginCapture(p->stream, cudaStreamCaptureModeThreadLocal) != cudaSuccess) {
        (void)cudaGetLastError();  // a stream that cannot be captured (legacy default stream): keep launching plainly
        p->graphs = 0;
        return dmmt_plan_chain(p, d_pixels, n, d_out, d_lens);
    }
    const int rc = dmmt_plan_chain(p, d_pixels, n, d_out, d_lens);
    const cudaError_t e = cudaStreamEndCapture(p->stream, &graph);
    if (rc != DMMT_OK || e != cudaSuccess || !graph) {
        if (graph) (void)cudaGraphDestroy(graph);
        (void)cudaGetLastError();
        p->graphs = 0;
        return rc != DMMT_OK ? rc : dmmt_plan_chain(p, d_pixels, n, d_out, d_lens);
    }
    const cudaError_t ei = cudaGraphInstantiate(&p->gexec, graph, 0);
    (void)cudaGraphDestroy(graph);
    if (ei != cudaSuccess) {
        (void)cudaGetLastError();
        p->gexec = nullptr, p->graphs = 0;
        return dmmt_plan_chain(p, d_pixels, n, d_out, d_lens);
    }
    p->g_launches = p->last_launches;
    DMMT_CUDA(cudaGraphLaunch(p->gexec, p->stream));
    return DMMT_OK;
}

extern "C" int dmmt_plan_set_graph(dmmt_plan* p, int enabled) {
    if (!p) return DMMT_E_INVALID;
    DMMT_CUDA(cudaSetDevice(p->ctx->device));
    DMMT_CUDA(cudaStreamSynchronize(p->stream));
    dmmt_plan_drop_graph(p);
    p->graphs = enabled ? 1 : 0;
    return DMMT_OK;
}

extern "C" int dmmt_plan_encode_device(dmmt_plan* p, const void* d_pixels, int n_images, uint8_t* d_out,
                                       uint64_t* d_lens) {
    if (!p || !d_pixels || !d_out || n_images <= 0 || n_images > p->n) return DMMT_E_INVALID;
    DMMT_CUDA(cudaSetDevice(p->ctx->device));
    DMMT_TRY(dmmt_plan_chain_replay(p, d_pixels, n_images, d_out,
                                    d_lens ? reinterpret_cast<unsigned long long*>(d_lens) : p->d_lens));
    if (p->profiling && p->ev_valid) {
        DMMT_CUDA(cudaEventRecord(p->ev[6], p->stream));  // no K5 here: ev[5] == ev[6]
    }
    return DMMT_OK;
}

// internal: chain + K5 compaction into `dense` (used by the host paths and by batches)
int dmmt_plan_encode_compact(dmmt_plan* p, const void* d_pixels, int n, uint8_t* d_dense,
                             unsigned long long dense_cap, unsigned long long* d_offsets,
                             unsigned long long* d_lens, int chained, int* sticky_err) {
    if (!p->d_out_own) DMMT_CUDA(cudaMalloc(&p->d_out_own, (size_t)p->n * p->out_stride));
    DMMT_TRY(dmmt_plan_chain(p, d_pixels, n, p->d_out_own, d_lens));
    DMMT_CUDA(launch_k5_compact(p->d_out_own, p->out_stride, d_lens, n, d_dense, dense_cap, d_offsets,
                                p->meta, chained, sticky_err, p->stream));
    p->last_launches += 2;
    if (p->profiling && p->ev_valid) DMMT_CUDA(cudaEventRecord(p->ev[6], p->stream));
    return DMMT_OK;
}

static int first_error(dmmt_plan* p, int n) {
    std::vector<ImgMeta> m((size_t)n);
    DMMT_CUDA(cudaMemcpyAsync(m.data(), p->meta, (size_t)n * sizeof(ImgMeta), cudaMemcpyDeviceToHost, p->stream));
    DMMT_CUDA(cudaStreamSynchronize(p->stream));
    for (int i = 0; i < n; i++)
        if (m[i].error) return m[i].error;
    return DMMT_OK;
}

extern "C" int dmmt_plan_status(dmmt_plan* p) {
    if (!p) return DMMT_E_INVALID;
    DMMT_CUDA(cudaSetDevice(p->ctx->device));
    DMMT_CUDA(cudaStreamSynchronize(p->stream));
    if (p->last_n <= 0) return DMMT_OK;
    return first_error(p, p->last_n);
}

extern "C" int dmmt_plan_last_launch_count(const dmmt_plan* p) { return p ? p->last_launches : 0; }

extern "C" int dmmt_plan_last_timings(dmmt_plan* p, float* ms, int n) {
    if (!p || !ms || n < DMMT_T_COUNT) return DMMT_E_INVALID;
    if (!p->profiling || !p->ev_valid) return DMMT_E_INVALID;
    DMMT_CUDA(cudaSetDevice(p->ctx->device));
    DMMT_CUDA(cudaEventSynchronize(p->ev[6]));
    for (int i = 0; i < 5; i++) DMMT_CUDA(cudaEventElapsedTime(&ms[i], p->ev[i], p->ev[i + 1]));
    DMMT_CUDA(cudaEventElapsedTime(&ms[DMMT_T_K5_COMPACT], p->ev[5], p->ev[6]));
    DMMT_CUDA(cudaEventElapsedTime(&ms[DMMT_T_TOTAL], p->ev[0], p->ev[6]));
    return DMMT_OK;
}

// ---- host-buffer paths -------------------------------------------------------------------
static int plan_ensure_host_arenas(dmmt_plan* p) {
    if (!p->d_pixels_own) DMMT_CUDA(cudaMalloc(&p->d_pixels_own, (size_t)p->n * p->pixel_bytes));
    if (!p->d_out_own) DMMT_CUDA(cudaMalloc(&p->d_out_own, (size_t)p->n * p->out_stride));
    if (p->n > 1 && !p->d_dense) {
        p->dense_cap = (size_t)p->n * p->out_stride;
        DMMT_CUDA(cudaMalloc(&p->d_dense, p->dense_cap));
    }
    return DMMT_OK;
}

// one attempt; *total receives the packed size.  On success the packed files are in h_out.
static int plan_host_attempt(dmmt_plan* p, const void* h_pixels, int n, uint8_t* h_out, uint64_t out_cap,
                             uint64_t* h_offsets, uint64_t* h_lens) {
    cudaStream_t st = p->stream;
    DMMT_TRY(plan_ensure_host_arenas(p));
    DMMT_CUDA(cudaMemcpyAsync(p->d_pixels_own, h_pixels, (size_t)n * p->pixel_bytes, cudaMemcpyHostToDevice, st));
    if (p->n == 1) {
        // single image: no packing pass, copy straight out of the arena
        DMMT_TRY(dmmt_plan_chain_replay(p, p->d_pixels_own, 1, p->d_out_own, p->d_lens));
        if (p->profiling && p->ev_valid) DMMT_CUDA(cudaEventRecord(p->ev[6], st));
        DMMT_CUDA(cudaMemcpyAsync(p->h_lens, p->d_lens, 8, cudaMemcpyDeviceToHost, st));
        const int rc = first_error(p, 1);  // synchronises
        if (rc != DMMT_OK) return rc;
        const uint64_t len = p->h_lens[0];
        if (len > out_cap) return DMMT_E_WRITE;
        DMMT_CUDA(cudaMemcpyAsync(h_out, p->d_out_own, len, cudaMemcpyDeviceToHost, st));
        DMMT_CUDA(cudaStreamSynchronize(st));
        h_offsets[0] = 0, h_lens[0] = len;
        return DMMT_OK;
    }
    DMMT_TRY(dmmt_plan_encode_compact(p, p->d_pixels_own, n, p->d_dense, p->dense_cap, p->d_offsets, p->d_lens, 0, nullptr));
    DMMT_CUDA(cudaMemcpyAsync(p->h_lens, p->d_lens, (size_t)n * 8, cudaMemcpyDeviceToHost, st));
    DMMT_CUDA(cudaMemcpyAsync(p->h_offsets, p->d_offsets, (size_t)(n + 1) * 8, cudaMemcpyDeviceToHost, st));
    const int rc = first_error(p, n);  // synchronises
    if (rc != DMMT_OK) return rc;
    const uint64_t total = p->h_offsets[n];
    if (total > out_cap) return DMMT_E_WRITE;
    DMMT_CUDA(cudaMemcpyAsync(h_out, p->d_dense, total, cudaMemcpyDeviceToHost, st));
    DMMT_CUDA(cudaStreamSynchronize(st));
    memcpy(h_offsets, p->h_offsets, (size_t)n * 8);
    memcpy(h_lens, p->h_lens, (size_t)n * 8);
    return DMMT_OK;
}

extern "C" int dmmt_plan_encode_host_into(dmmt_plan* p, const void* h_pixels, int n_images, uint8_t* h_out,
                                          uint64_t out_cap, uint64_t* h_offsets, uint64_t* h_lens) {
    if (!p || !h_pixels || !h_out || !h_offsets || !h_lens || n_images <= 0 || n_images > p->n)
        return DMMT_E_INVALID;
    DMMT_CUDA(cudaSetDevice(p->ctx->device));
    int rc = plan_host_attempt(p, h_pixels, n_images, h_out, out_cap, h_offsets, h_lens);
    if (rc == DMMT_E_OVERFLOW) {  // denser than the default 128 B per block: size for the worst case
        DMMT_TRY(dmmt_plan_set_scan_capacity(p, dmmt_plan_worst_case_scan_bytes(p)));
        rc = plan_host_attempt(p, h_pixels, n_images, h_out, out_cap, h_offsets, h_lens);
    }
    return rc;
}

extern "C" int dmmt_plan_encode_host(dmmt_plan* p, const void* h_pixels, int n_images, uint8_t** jpegs,
                                     size_t* lens) {
    if (!p || !h_pixels || !jpegs || !lens || n_images <= 0 || n_images > p->n) return DMMT_E_INVALID;
    for (int i = 0; i < n_images; i++) jpegs[i] = nullptr, lens[i] = 0;
    DMMT_CUDA(cudaSetDevice(p->ctx->device));
    for (int attempt = 0; attempt < 2; attempt++) {
        const size_t cap = (size_t)n_images * p->out_stride;
        uint8_t* tmp = static_cast<uint8_t*>(malloc(cap));
        if (!tmp) return DMMT_E_NOMEM;
        std::vector<uint64_t> off((size_t)n_images), ln((size_t)n_images);
        int rc = plan_host_attempt(p, h_pixels, n_images, tmp, cap, off.data(), ln.data());
        if (rc == DMMT_E_OVERFLOW && attempt == 0) {
            free(tmp);
            DMMT_TRY(dmmt_plan_set_scan_capacity(p, dmmt_plan_worst_case_scan_bytes(p)));
            continue;
        }
        if (rc == DMMT_OK) {
            for (int i = 0; i < n_images; i++) {
                jpegs[i] = static_cast<uint8_t*>(malloc(ln[i] ? ln[i] : 1));
                if (!jpegs[i]) {
                    for (int j = 0; j < i; j++) free(jpegs[j]), jpegs[j] = nullptr;
                    free(tmp);
                    return DMMT_E_NOMEM;
                }
                memcpy(jpegs[i], tmp + off[i], ln[i]);
                lens[i] = ln[i];
            }
        }
        free(tmp);
        return rc;
    }
    return DMMT_E_OVERFLOW;
}

// ---- the drop-in call -----------------------------------------------------------------------
static dmmt_plan* ctx_find_plan(dmmt_ctx* c, const dmmt_image* im, const dmmt_options* o) {
    for (size_t i = 0; i < c->cache.size(); i++) {
        dmmt_plan* p = c->cache[i];
        if (p->W == im->width && p->H == im->height && p->fmt == (int)im->fmt &&
            (im->fmt == DMMT_RGB_F32_NORM || p->max_value == im->max_value) &&
            p->opt.subsampling == o->subsampling && p->opt.bits_per_channel == o->bits_per_channel &&
            p->opt.qtable_preset == o->qtable_preset) {
            std::rotate(c->cache.begin(), c->cache.begin() + i, c->cache.begin() + i + 1);
            return p;
        }
    }
    return nullptr;
}

extern "C" int dmmt_encode(dmmt_ctx* c, const dmmt_image* im, const dmmt_options* o, uint8_t** jpeg, size_t* len) {
    if (!c || !im || !o || !jpeg || !len || !im->pixels) return DMMT_E_INVALID;
    *jpeg = nullptr, *len = 0;
    if (im->width == 0 || im->height == 0) return DMMT_E_INVALID;
    DMMT_CUDA(cudaSetDevice(c->device));
    dmmt_plan* p = ctx_find_plan(c, im, o);
    if (!p) {
        DMMT_TRY(dmmt_plan_create(c, im->width, im->height, im->fmt,
                                  im->fmt == DMMT_RGB_F32_NORM ? (uint16_t)1 : im->max_value, o, 1, &p));
        c->cache.insert(c->cache.begin(), p);
        if (c->cache.size() > 4) {
            dmmt_plan_destroy(c->cache.back());
            c->cache.pop_back();
        }
    }
    if (!im->pixels_on_device) return dmmt_plan_encode_host(p, im->pixels, 1, jpeg, len);
    // pixels already on this context's device
    for (int attempt = 0; attempt < 2; attempt++) {
        if (!p->d_out_own) DMMT_CUDA(cudaMalloc(&p->d_out_own, p->out_stride));
        DMMT_TRY(dmmt_plan_chain_replay(p, im->pixels, 1, p->d_out_own, p->d_lens));
        if (p->profiling && p->ev_valid) DMMT_CUDA(cudaEventRecord(p->ev[6], p->stream));
        DMMT_CUDA(cudaMemcpyAsync(p->h_lens, p->d_lens, 8, cudaMemcpyDeviceToHost, p->stream));
        int rc = first_error(p, 1);
        if (rc == DMMT_E_OVERFLOW && attempt == 0) {
            DMMT_TRY(dmmt_plan_set_scan_capacity(p, dmmt_plan_worst_case_scan_bytes(p)));
            continue;
        }
        if (rc != DMMT_OK) return rc;
        const size_t n = (size_t)p->h_lens[0];
        uint8_t* buf = static_cast<uint8_t*>(malloc(n ? n : 1));
        if (!buf) return DMMT_E_NOMEM;
        cudaError_t e = cudaMemcpyAsync(buf, p->d_out_own, n, cudaMemcpyDeviceToHost, p->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(p->stream);
        if (e != cudaSuccess) {
            dmmt_set_cuda_error(e, "D2H of the encoded file", __FILE__, __LINE__);
            free(buf);
            return DMMT_E_CUDA;
        }
        *jpeg = buf, *len = n;
        return DMMT_OK;
    }
    return DMMT_E_OVERFLOW;
}

extern "C" int dmmt_encode_batch(dmmt_ctx* const* ctxs, int nctx, const dmmt_image* imgs, int n,
                                 const dmmt_options* o, uint8_t** jpegs, size_t* lens) {
    if (!ctxs || nctx <= 0 || !imgs || n < 0 || !o || !jpegs || !lens) return DMMT_E_INVALID;
    for (int i = 0; i < n; i++) jpegs[i] = nullptr, lens[i] = 0;
    // Generic entry point: arbitrary mixed geometries, image i on ctxs[i % nctx].  The
    // pipelined equal-geometry path is dmmt_batch_* (dmmt_batch.cu).
    for (int i = 0; i < n; i++) {
        const int rc = dmmt_encode(ctxs[i % nctx], &imgs[i], o, &jpegs[i], &lens[i]);
        if (rc != DMMT_OK) {
            for (int j = 0; j < i; j++) free(jpegs[j]), jpegs[j] = nullptr, lens[j] = 0;
            return rc;
        }
    }
    return DMMT_OK;
}

// ---- test / measurement hooks ---------------------------------------------------------------
extern "C" int dmmt_plan_fetch(dmmt_plan* p, int what, int index, void* dst, size_t cap_bytes, size_t* got) {
    if (!p || !dst || index < 0 || index >= p->n) return DMMT_E_INVALID;
    DMMT_CUDA(cudaSetDevice(p->ctx->device));
    DMMT_CUDA(cudaStreamSynchronize(p->stream));
    const void* src = nullptr;
    size_t bytes = 0;
    ImgMeta m{};
    DMMT_CUDA(cudaMemcpy(&m, p->meta + index, sizeof m, cudaMemcpyDeviceToHost));
    switch (what) {
        case DMMT_FETCH_COEF:
            if (!p->coef) return DMMT_E_INVALID;  // the fused path never wrote one: dmmt_plan_set_generic_path first
            src = p->coef + (size_t)index * p->coef_stride, bytes = p->coef_stride * sizeof(int16_t);
            break;
        case DMMT_FETCH_HIST: src = p->hist + (size_t)index * 1024, bytes = 1024 * sizeof(unsigned int); break;
        case DMMT_FETCH_TABLES: src = p->lens + index, bytes = sizeof(LenTables); break;
        case DMMT_FETCH_SCAN:
            src = p->scan + (size_t)index * p->scan_stride_words, bytes = (size_t)((m.scan_bits + 7) / 8);
            break;
        case DMMT_FETCH_META: src = p->meta + index, bytes = sizeof(ImgMeta); break;
        case DMMT_FETCH_TOKEN_COUNT: {
            // tokens K1 / K2 wrote for this image (sum over its token regions): the live measure of the token traffic
            const size_t regions = p->fused ? p->fo.tiles : p->n_chunks3;
            std::vector<uint32_t> nt(regions);
            DMMT_CUDA(cudaMemcpy(nt.data(), p->tb.ntok + (size_t)index * regions, regions * 4, cudaMemcpyDeviceToHost));
            unsigned long long total = 0;
            for (uint32_t v : nt) total += v;
            if (got) *got = sizeof total;
            if (cap_bytes < sizeof total) return DMMT_E_WRITE;
            memcpy(dst, &total, sizeof total);
            return DMMT_OK;
        }
        default: return DMMT_E_INVALID;
    }
    if (got) *got = bytes;
    if (bytes > cap_bytes) return DMMT_E_WRITE;
    DMMT_CUDA(cudaMemcpy(dst, src, bytes, cudaMemcpyDeviceToHost));
    return DMMT_OK;
}

extern "C" int dmmt_plan_debug_dct(dmmt_plan* p, const void* d_pixels, int index, float* dst, size_t cap_floats) {
    if (!p || !d_pixels || !dst || index < 0 || index >= p->n) return DMMT_E_INVALID;
    if (cap_floats < p->coef_stride) return DMMT_E_WRITE;
    DMMT_CUDA(cudaSetDevice(p->ctx->device));
    DMMT_TRY(plan_ensure_coef(p));
    float* d_dbg = nullptr;
    DMMT_CUDA(cudaMalloc(&d_dbg, p->coef_stride * sizeof(float)));
    const uint8_t* px = static_cast<const uint8_t*>(d_pixels) + (size_t)index * p->pixel_bytes;
    cudaError_t e = launch_k1(p->g, p->fmt, p->k1c, 0, px, p->pixel_bytes, 1,
                              p->coef + (size_t)index * p->coef_stride, p->coef_stride, d_dbg, p->meta + index,
                              nullptr, nullptr, p->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(dst, d_dbg, p->coef_stride * sizeof(float), cudaMemcpyDeviceToHost, p->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(p->stream);
    (void)cudaFree(d_dbg);
    if (e != cudaSuccess) {
        dmmt_set_cuda_error(e, "debug K1", __FILE__, __LINE__);
        return DMMT_E_CUDA;
    }
    return DMMT_OK;
}

extern "C" int dmmt_debug_stuff(dmmt_ctx* ctx, const uint8_t* scan, size_t n, int misalign, uint8_t* out, size_t cap,
                                size_t* got) {
    if (!ctx || (!scan && n) || !out || !got || misalign < 0 || misalign > 15) return DMMT_E_INVALID;
    DMMT_CUDA(cudaSetDevice(ctx->device));
    const size_t out_cap = 2 * n + 64;
    const uint32_t chunks = k4_max_chunks(n + 1);
    uint8_t *d_scan = nullptr, *d_out = nullptr;
    ImgMeta* d_meta = nullptr;
    unsigned long long* d_lb = nullptr;
    unsigned int* d_tk = nullptr;
    auto release = [&]() {
        (void)cudaFree(d_scan), (void)cudaFree(d_out), (void)cudaFree(d_meta), (void)cudaFree(d_lb), (void)cudaFree(d_tk);
    };
    ImgMeta m{};
    m.scan_bits = 8ull * n;
    cudaError_t e = cudaMalloc(&d_scan, n + 64);
    if (e == cudaSuccess) e = cudaMalloc(&d_out, out_cap + 16);
    if (e == cudaSuccess) e = cudaMalloc(&d_meta, sizeof(ImgMeta));
    if (e == cudaSuccess) e = cudaMalloc(&d_lb, (size_t)chunks * 8);
    if (e == cudaSuccess) e = cudaMalloc(&d_tk, 4);
    if (e == cudaSuccess) e = cudaMemsetAsync(d_scan, 0xA5, n + 64, ctx->stream);  // bytes past the end are not zero
    if (e == cudaSuccess && n) e = cudaMemcpyAsync(d_scan, scan, n, cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess) e = cudaMemsetAsync(d_out, 0xEE, out_cap + 16, ctx->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(d_meta, &m, sizeof(m), cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess) e = cudaMemsetAsync(d_lb, 0, (size_t)chunks * 8, ctx->stream);
    if (e == cudaSuccess) e = cudaMemsetAsync(d_tk, 0, 4, ctx->stream);
    if (e == cudaSuccess) {
        K4HostArgs k{};
        k.scan = d_scan, k.scan_stride_bytes = n + 64, k.meta = d_meta, k.lb_state = d_lb, k.ticket = d_tk;
        k.max_chunks = chunks, k.out = d_out + misalign, k.out_stride = out_cap, k.out_lens = nullptr;
        k.first_byte = 0, k.n_bytes_override = (long long)n, k.seed_bits = 0, k.prepend_header = 0, k.append_eoi = 1;
        e = launch_k4(k, 1, chunks, ctx->stream);
    }
    if (e == cudaSuccess) e = cudaMemcpyAsync(&m, d_meta, sizeof(m), cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    int rc = DMMT_OK;
    if (e == cudaSuccess) {
        if (m.error) rc = m.error;
        else if (m.out_len > cap) rc = DMMT_E_WRITE;
        else {
            *got = (size_t)m.out_len;
            // one byte before and after the file as well: the kernel must not touch them
            e = cudaMemcpy(out, d_out + misalign, (size_t)m.out_len, cudaMemcpyDeviceToHost);
            uint8_t guard[2] = {0, 0};
            if (e == cudaSuccess && misalign) e = cudaMemcpy(&guard[0], d_out + misalign - 1, 1, cudaMemcpyDeviceToHost);
            if (e == cudaSuccess) e = cudaMemcpy(&guard[1], d_out + misalign + m.out_len, 1, cudaMemcpyDeviceToHost);
            if (e == cudaSuccess && ((misalign && guard[0] != 0xEE) || guard[1] != 0xEE)) rc = DMMT_E_WRITE;  // a byte outside the file was written
        }
    }
    release();
    if (e != cudaSuccess) {
        dmmt_set_cuda_error(e, "debug K4", __FILE__, __LINE__);
        return DMMT_E_CUDA;
    }
    return rc;
}
