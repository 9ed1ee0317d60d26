// dmmt_internal.h -- host-side objects behind the opaque handles of include/dmmt_cuda.h.
#pragma once

#include <string>
#include <vector>

#include "dmmt_kernels.h"

// thread-local text of the last CUDA failure (dmmt_last_cuda_error)
void dmmt_set_cuda_error(cudaError_t e, const char* what, const char* file, int line);

#define DMMT_CUDA(expr)                                           \
    do {                                                          \
        cudaError_t e__ = (expr);                                 \
        if (e__ != cudaSuccess) {                                 \
            dmmt_set_cuda_error(e__, #expr, __FILE__, __LINE__);  \
            return e__ == cudaErrorMemoryAllocation ? DMMT_E_NOMEM : DMMT_E_CUDA; \
        }                                                         \
    } while (0)

#define DMMT_TRY(expr)           \
    do {                         \
        int rc__ = (expr);       \
        if (rc__ != DMMT_OK) return rc__; \
    } while (0)

struct dmmt_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    // small cache of single-image plans used by dmmt_encode (most recent first)
    std::vector<dmmt_plan*> cache;
};

struct dmmt_plan {
    dmmt_ctx* ctx = nullptr;
    cudaStream_t stream = nullptr;  // defaults to the context's stream; batch slots own theirs
    bool own_stream = false;
    dmmt::Geom g{};
    int W = 0, H = 0;               // ORIGINAL size of this plan's pixel rows (shard: rows of the shard)
    int sof_W = 0, sof_H = 0;       // size written to SOF0 (shard: the full image)
    int fmt = 0, max_value = 255;
    dmmt_options opt{};
    int n = 0;                      // images per launch chain
    size_t pixel_bytes = 0;         // one image
    size_t scan_cap_bytes = 0;      // per image
    size_t scan_stride_words = 0;
    size_t out_stride = 0;
    size_t coef_stride = 0;         // int16 elements per image
    uint32_t n_chunks3 = 0, max_chunks4 = 0;
    dmmt::K1Consts k1c{};

    // device scratch
    int16_t* coef = nullptr;
    uint8_t* zero_region = nullptr;
    size_t zero_bytes = 0;
    unsigned int* hist = nullptr;
    dmmt::ImgMeta* meta = nullptr;
    unsigned long long* lb3 = nullptr;
    unsigned long long* lb4 = nullptr;
    unsigned int* tk3 = nullptr;
    unsigned int* tk4 = nullptr;
    dmmt::EncTables* enc = nullptr;
    dmmt::LenTables* lens = nullptr;
    uint8_t* lcount = nullptr;                // [n][4][16] K2b scratch (codes per length)
    uint32_t* scan = nullptr;
    dmmt::TokBuf tb{};                        // K2 -> K3 token stream (generic path)
    dmmt::TileTok fo{};                       // fused K1 -> K3 token stream (4:2:0 fast path), shares tb.tok
    bool fused = false;                       // this plan runs the fused path
    int force_generic = 0;                    // dmmt_plan_set_generic_path
    uint32_t n_chunks3f = 0;                  // look-back chunks of K3 in tile mode
    unsigned long long* d_lens = nullptr;     // [n]
    unsigned long long* d_offsets = nullptr;  // [n + 1]
    int16_t* d_seed_dc = nullptr;             // [3] shard predictors
    int16_t* d_last_dc = nullptr;             // [3]
    unsigned long long* d_ghist = nullptr;    // [4][256] shard global histogram

    // lazily created arenas of the host-buffer paths
    uint8_t* d_pixels_own = nullptr;
    uint8_t* d_out_own = nullptr;
    uint8_t* d_dense = nullptr;
    size_t dense_cap = 0;
    unsigned long long* h_lens = nullptr;     // pinned [n]
    unsigned long long* h_offsets = nullptr;  // pinned [n + 1]

    // CUDA graph of the launch chain for repeated calls with the same arguments (dmmt_plan_chain_replay)
    cudaGraphExec_t gexec = nullptr;
    const void* g_pixels = nullptr;      // arguments of the captured (or, before capture, of the previous) call
    int g_n = 0;
    uint8_t* g_out = nullptr;
    unsigned long long* g_lens = nullptr;
    int g_launches = 0;
    int graphs = 1;                      // dmmt_plan_set_graph

    // profiling
    bool profiling = false;
    cudaEvent_t ev[DMMT_T_COUNT + 1] = {};
    bool ev_valid = false;
    int last_launches = 0;
    int last_n = 0;
};

// ---- internals shared between dmmt_api.cu, dmmt_batch.cu and dmmt_shard.cu -------------------
// plan over `H_rows` pixel rows (a whole image, or the rows of one MCU-row shard when
// mcus_y_override > 0); sof_W/sof_H go into SOF0.
int dmmt_plan_create_impl(dmmt_ctx* ctx, int W, int H_rows, int mcus_y_override, int sof_W, int sof_H,
                          int fmt, int max_value, const dmmt_options* opt, int n_images, cudaStream_t st,
                          bool own_stream, dmmt_plan** out);
// memset + K1 + K2 + K2b + K3 + K4 on the plan's stream (asynchronous)
int dmmt_plan_chain(dmmt_plan* p, const void* d_pixels, int n, uint8_t* d_out, unsigned long long* d_lens);
// the same chain, replayed from a CUDA graph once a call repeats the arguments of the previous one (one launch instead
// of eight: what a latency-bound single frame is made of); plain launches while profiling or when graphs are off
int dmmt_plan_chain_replay(dmmt_plan* p, const void* d_pixels, int n, uint8_t* d_out, unsigned long long* d_lens);
void dmmt_plan_drop_graph(dmmt_plan* p);
// chain + K5 packing into `dense`
int dmmt_plan_encode_compact(dmmt_plan* p, const void* d_pixels, int n, uint8_t* d_dense,
                             unsigned long long dense_cap, unsigned long long* d_offsets,
                             unsigned long long* d_lens, int chained, int* sticky_err);
