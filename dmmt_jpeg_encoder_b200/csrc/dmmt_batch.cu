// dmmt_batch.cu -- pipelined encode of many equally sized images (BASELINE config 4: a batch of
// 1920x1080 frames).  The batch is cut into sub-batches; sub-batch k runs on slot k % depth, a
// slot being a plan with its own stream and scratch, so the H2D copy, the K1..K5 chain and the
// D2H copy of neighbouring sub-batches overlap and the scratch footprint stays bounded no matter
// how many images are encoded.  All slots fork from / join into the context's stream, so CUDA
// events recorded on that stream bracket the whole batch.
//
// The reference encodes one image per process invocation (src/lib.rs:59-77); this is the
// data-parallel widening named by BASELINE.json (image i of the batch -> device i mod N is done
// by the caller: one dmmt_batch per device/rank).
#include <algorithm>
#include <cstring>
#include <new>

#include "dmmt_internal.h"

using namespace dmmt;

struct dmmt_batch {
    dmmt_ctx* ctx = nullptr;
    int sub = 0, depth = 0;
    std::vector<dmmt_plan*> slots;
    std::vector<cudaEvent_t> ev_k5;    // K5 of the slot's current sub-batch has been issued up to here
    std::vector<cudaEvent_t> ev_meta;  // lens/offsets of the slot's sub-batch are on the host
    cudaEvent_t ev_fork = nullptr;
    int* d_err = nullptr;              // sticky first device-side error of the last call
    int* h_err = nullptr;              // pinned
    int last_launches = 0;
    bool profiling = false;
    float acc_ms[DMMT_T_COUNT] = {};
    // host path bookkeeping
    struct Pending {
        bool active = false;
        int first = 0, m = 0;
    };
    std::vector<Pending> pending;
    // CUDA graph of one whole dmmt_batch_encode_device call (all sub-batches, all slot streams): captured when a call
    // repeats the arguments of the previous one, replayed while they stay the same (DMMT_BATCH_GRAPH=0: never)
    struct {
        const void* px = nullptr;
        const void *dense = nullptr, *offs = nullptr, *lens = nullptr;
        uint64_t cap = 0;
        int n = 0;
    } last;                          // arguments of the previous device-resident call
    cudaGraphExec_t gexec = nullptr; // built for `last`
    int g_launches = 0;
};

static void batch_drop_graph(dmmt_batch* b) {
    if (b->gexec) (void)cudaGraphExecDestroy(b->gexec);
    b->gexec = nullptr;
    b->last = {};
}

extern "C" void dmmt_batch_destroy(dmmt_batch* b) {
    if (!b) return;
    (void)cudaSetDevice(b->ctx->device);
    batch_drop_graph(b);
    for (dmmt_plan* p : b->slots) dmmt_plan_destroy(p);
    for (auto e : b->ev_k5) (void)cudaEventDestroy(e);
    for (auto e : b->ev_meta) (void)cudaEventDestroy(e);
    if (b->ev_fork) (void)cudaEventDestroy(b->ev_fork);
    (void)cudaFree(b->d_err);
    if (b->h_err) (void)cudaFreeHost(b->h_err);
    delete b;
}

extern "C" int dmmt_batch_create(dmmt_ctx* ctx, uint16_t width, uint16_t height, dmmt_fmt fmt,
                                 uint16_t max_value, const dmmt_options* opt, int sub_batch, int depth,
                                 dmmt_batch** out) {
    if (!ctx || !opt || !out || sub_batch <= 0 || depth <= 0 || depth > 16) return DMMT_E_INVALID;
    *out = nullptr;
    DMMT_CUDA(cudaSetDevice(ctx->device));
    dmmt_batch* b = new (std::nothrow) dmmt_batch();
    if (!b) return DMMT_E_NOMEM;
    b->ctx = ctx, b->sub = sub_batch, b->depth = depth;
    b->pending.resize(depth);
    int rc = DMMT_OK;
    for (int s = 0; s < depth && rc == DMMT_OK; s++) {
        cudaStream_t st = nullptr;
        cudaError_t e = cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking);
        if (e != cudaSuccess) {
            dmmt_set_cuda_error(e, "cudaStreamCreateWithFlags", __FILE__, __LINE__);
            rc = DMMT_E_CUDA;
            break;
        }
        dmmt_plan* p = nullptr;
        rc = dmmt_plan_create_impl(ctx, width, height, 0, width, height, (int)fmt, max_value, opt, sub_batch, st,
                                   true, &p);
        if (rc != DMMT_OK) {
            (void)cudaStreamDestroy(st);
            break;
        }
        b->slots.push_back(p);
        cudaEvent_t e1 = nullptr, e2 = nullptr;
        if (cudaEventCreateWithFlags(&e1, cudaEventDisableTiming) != cudaSuccess ||
            cudaEventCreateWithFlags(&e2, cudaEventDisableTiming) != cudaSuccess) {
            rc = DMMT_E_CUDA;
            break;
        }
        b->ev_k5.push_back(e1), b->ev_meta.push_back(e2);
    }
    if (rc == DMMT_OK && cudaEventCreateWithFlags(&b->ev_fork, cudaEventDisableTiming) != cudaSuccess) rc = DMMT_E_CUDA;
    if (rc == DMMT_OK && cudaMalloc(&b->d_err, sizeof(int)) != cudaSuccess) rc = DMMT_E_NOMEM;
    if (rc == DMMT_OK && cudaHostAlloc(&b->h_err, sizeof(int), cudaHostAllocDefault) != cudaSuccess) rc = DMMT_E_NOMEM;
    if (rc != DMMT_OK) {
        dmmt_batch_destroy(b);
        return rc;
    }
    *out = b;
    return DMMT_OK;
}

extern "C" int dmmt_batch_set_scan_capacity(dmmt_batch* b, size_t bytes_per_image) {
    if (!b) return DMMT_E_INVALID;
    batch_drop_graph(b);  // the scratch buffers the graph points at are replaced
    for (dmmt_plan* p : b->slots) DMMT_TRY(dmmt_plan_set_scan_capacity(p, bytes_per_image));
    return DMMT_OK;
}

extern "C" size_t dmmt_batch_worst_case_scan_bytes(const dmmt_batch* b) {
    return b ? dmmt_plan_worst_case_scan_bytes(b->slots[0]) : 0;
}

extern "C" int dmmt_batch_set_profiling(dmmt_batch* b, int enabled) {
    if (!b) return DMMT_E_INVALID;
    batch_drop_graph(b);
    for (dmmt_plan* p : b->slots) DMMT_TRY(dmmt_plan_set_profiling(p, enabled));
    b->profiling = enabled != 0;
    return DMMT_OK;
}

extern "C" int dmmt_batch_last_launch_count(const dmmt_batch* b) { return b ? b->last_launches : 0; }
extern "C" int dmmt_batch_uses_fused_path(const dmmt_batch* b) { return b && !b->slots.empty() && b->slots[0]->fused ? 1 : 0; }

extern "C" int dmmt_batch_last_timings(dmmt_batch* b, float* ms, int n) {
    if (!b || !ms || n < DMMT_T_COUNT || !b->profiling) return DMMT_E_INVALID;
    memcpy(ms, b->acc_ms, sizeof b->acc_ms);
    return DMMT_OK;
}

// all slot streams wait for what has been issued on the context's stream so far
static int fork_slots(dmmt_batch* b) {
    DMMT_CUDA(cudaEventRecord(b->ev_fork, b->ctx->stream));
    for (dmmt_plan* p : b->slots) DMMT_CUDA(cudaStreamWaitEvent(p->stream, b->ev_fork, 0));
    return DMMT_OK;
}
// the context's stream waits for everything issued on the slots
static int join_slots(dmmt_batch* b) {
    for (int s = 0; s < b->depth; s++) {
        DMMT_CUDA(cudaEventRecord(b->ev_k5[s], b->slots[s]->stream));
        DMMT_CUDA(cudaStreamWaitEvent(b->ctx->stream, b->ev_k5[s], 0));
    }
    return DMMT_OK;
}

// issues one whole device-resident batch on the context's stream and the slot streams (forked from / joined into it)
static int batch_issue_device(dmmt_batch* b, const void* d_pixels, int n, uint8_t* d_dense, uint64_t dense_cap,
                              uint64_t* d_offsets, uint64_t* d_lens) {
    auto* offs = reinterpret_cast<unsigned long long*>(d_offsets);
    auto* lens = reinterpret_cast<unsigned long long*>(d_lens);
    DMMT_CUDA(cudaMemsetAsync(b->d_err, 0, sizeof(int), b->ctx->stream));
    DMMT_CUDA(cudaMemsetAsync(offs, 0, 8, b->ctx->stream));  // offsets[0] = 0: start of the chain
    DMMT_TRY(fork_slots(b));
    const size_t img_bytes = b->slots[0]->pixel_bytes;
    const int n_sub = (n + b->sub - 1) / b->sub;
    int launches = 0;
    if (b->profiling) memset(b->acc_ms, 0, sizeof b->acc_ms);
    for (int k = 0; k < n_sub; k++) {
        const int slot = k % b->depth;
        dmmt_plan* p = b->slots[slot];
        const int first = k * b->sub, m = std::min(b->sub, n - first);
        if (b->profiling && k >= b->depth) {  // the slot's events are about to be re-recorded
            float ms[DMMT_T_COUNT];
            DMMT_TRY(dmmt_plan_last_timings(p, ms, DMMT_T_COUNT));
            for (int i = 0; i < DMMT_T_COUNT; i++) b->acc_ms[i] += ms[i];
        }
        // With profiling the slots are serialised so that the per-kernel event times do not overlap.
        if (b->profiling && k > 0) DMMT_CUDA(cudaStreamWaitEvent(p->stream, b->ev_k5[(k - 1) % b->depth], 0));
        if (!p->d_out_own) DMMT_CUDA(cudaMalloc(&p->d_out_own, (size_t)p->n * p->out_stride));
        DMMT_TRY(dmmt_plan_chain(p, static_cast<const uint8_t*>(d_pixels) + (size_t)first * img_bytes, m,
                                 p->d_out_own, lens + first));
        // K5 of sub-batch k continues the offset chain of sub-batch k-1 (other slot, other stream):
        // only K5 is ordered behind it, the K1..K4 chains of the slots run concurrently.
        if (k > 0) DMMT_CUDA(cudaStreamWaitEvent(p->stream, b->ev_k5[(k - 1) % b->depth], 0));
        DMMT_CUDA(launch_k5_compact(p->d_out_own, p->out_stride, lens + first, m, d_dense, dense_cap,
                                    offs + first, p->meta, 1, b->d_err, p->stream));
        p->last_launches += 2;
        if (p->profiling && p->ev_valid) DMMT_CUDA(cudaEventRecord(p->ev[6], p->stream));
        DMMT_CUDA(cudaEventRecord(b->ev_k5[slot], p->stream));
        launches += p->last_launches;
    }
    if (b->profiling) {
        for (int s = 0; s < std::min(b->depth, n_sub); s++) {
            float ms[DMMT_T_COUNT];
            DMMT_TRY(dmmt_plan_last_timings(b->slots[s], ms, DMMT_T_COUNT));
            for (int i = 0; i < DMMT_T_COUNT; i++) b->acc_ms[i] += ms[i];
        }
    }
    DMMT_TRY(join_slots(b));
    b->last_launches = launches;
    return DMMT_OK;
}

static bool batch_graphs_enabled() {
    static const bool on = [] {
        const char* e = getenv("DMMT_BATCH_GRAPH");
        return !(e && e[0] == '0');
    }();
    return on;
}

extern "C" int dmmt_batch_encode_device(dmmt_batch* b, const void* d_pixels, int n, uint8_t* d_dense,
                                        uint64_t dense_cap, uint64_t* d_offsets, uint64_t* d_lens) {
    if (!b || !d_pixels || n <= 0 || !d_dense || !d_offsets || !d_lens) return DMMT_E_INVALID;
    DMMT_CUDA(cudaSetDevice(b->ctx->device));
    const bool same = b->last.px == d_pixels && b->last.n == n && b->last.dense == d_dense && b->last.cap == dense_cap &&
                      b->last.offs == d_offsets && b->last.lens == d_lens;
    if (b->profiling || !batch_graphs_enabled()) return batch_issue_device(b, d_pixels, n, d_dense, dense_cap, d_offsets, d_lens);
    if (b->gexec && same) {
        DMMT_CUDA(cudaGraphLaunch(b->gexec, b->ctx->stream));
        b->last_launches = b->g_launches;
        return DMMT_OK;
    }
    if (b->gexec) batch_drop_graph(b);
    if (!same) {  // first call with these arguments: plain launches (lazy allocations and one-time set-up happen here)
        const int rc = batch_issue_device(b, d_pixels, n, d_dense, dense_cap, d_offsets, d_lens);
        b->last.px = d_pixels, b->last.n = n, b->last.dense = d_dense, b->last.cap = dense_cap;
        b->last.offs = d_offsets, b->last.lens = d_lens;
        if (rc != DMMT_OK) b->last = {};
        return rc;
    }
    // second identical call: capture it (the slot streams join the capture through the fork / join events), replay it
    cudaGraph_t graph = nullptr;
    if (cudaStreamBeginCapture(b->ctx->stream, cudaStreamCaptureModeThreadLocal) != cudaSuccess) {
        // a stream that cannot be captured (the legacy default stream, or one the caller is capturing already)
        (void)cudaGetLastError();
        b->last = {};
        return batch_issue_device(b, d_pixels, n, d_dense, dense_cap, d_offsets, d_lens);
    }
    const int rc = batch_issue_device(b, d_pixels, n, d_dense, dense_cap, d_offsets, d_lens);
    const cudaError_t ce = cudaStreamEndCapture(b->ctx->stream, &graph);
    if (rc != DMMT_OK || ce != cudaSuccess || !graph) {
        if (graph) (void)cudaGraphDestroy(graph);
        (void)cudaGetLastError();
        b->last = {};
        if (rc != DMMT_OK) return rc;
        return batch_issue_device(b, d_pixels, n, d_dense, dense_cap, d_offsets, d_lens);  // capture refused: plain launches
    }
    const cudaError_t ie = cudaGraphInstantiate(&b->gexec, graph, 0);
    (void)cudaGraphDestroy(graph);
    if (ie != cudaSuccess) {
        (void)cudaGetLastError();
        b->gexec = nullptr, b->last = {};
        return batch_issue_device(b, d_pixels, n, d_dense, dense_cap, d_offsets, d_lens);
    }
    b->g_launches = b->last_launches;
    DMMT_CUDA(cudaGraphLaunch(b->gexec, b->ctx->stream));
    return DMMT_OK;
}

extern "C" int dmmt_batch_status(dmmt_batch* b) {
    if (!b) return DMMT_E_INVALID;
    DMMT_CUDA(cudaSetDevice(b->ctx->device));
    for (dmmt_plan* p : b->slots) DMMT_CUDA(cudaStreamSynchronize(p->stream));
    DMMT_CUDA(cudaMemcpyAsync(b->h_err, b->d_err, sizeof(int), cudaMemcpyDeviceToHost, b->ctx->stream));
    DMMT_CUDA(cudaStreamSynchronize(b->ctx->stream));
    return *b->h_err;
}

// ---- host path --------------------------------------------------------------------------------
// waits for the slot's lens/offsets, then queues the D2H copy of its packed files
static int drain_slot(dmmt_batch* b, int slot, uint8_t* h_out, uint64_t out_cap, uint64_t* h_offsets,
                      uint64_t* h_lens, uint64_t* run) {
    dmmt_batch::Pending& pd = b->pending[slot];
    if (!pd.active) return DMMT_OK;
    pd.active = false;
    dmmt_plan* p = b->slots[slot];
    DMMT_CUDA(cudaEventSynchronize(b->ev_meta[slot]));
    const uint64_t total = p->h_offsets[pd.m];
    if (*run + total > out_cap) return DMMT_E_WRITE;
    if (total) DMMT_CUDA(cudaMemcpyAsync(h_out + *run, p->d_dense, total, cudaMemcpyDeviceToHost, p->stream));
    for (int i = 0; i < pd.m; i++) {
        h_offsets[pd.first + i] = *run + p->h_offsets[i];
        h_lens[pd.first + i] = p->h_lens[i];
    }
    *run += total;
    return DMMT_OK;
}

extern "C" int dmmt_batch_encode_host(dmmt_batch* b, const void* h_pixels, int n, uint8_t* h_out,
                                      uint64_t out_cap, uint64_t* h_offsets, uint64_t* h_lens) {
    if (!b || !h_pixels || n <= 0 || !h_out || !h_offsets || !h_lens) return DMMT_E_INVALID;
    DMMT_CUDA(cudaSetDevice(b->ctx->device));
    DMMT_CUDA(cudaMemsetAsync(b->d_err, 0, sizeof(int), b->ctx->stream));
    DMMT_TRY(fork_slots(b));
    const size_t img_bytes = b->slots[0]->pixel_bytes;
    const int n_sub = (n + b->sub - 1) / b->sub;
    uint64_t run = 0;
    int launches = 0;
    int rc = DMMT_OK;
    for (int k = 0; k < n_sub && rc == DMMT_OK; k++) {
        const int slot = k % b->depth;
        dmmt_plan* p = b->slots[slot];
        const int first = k * b->sub, m = std::min(b->sub, n - first);
        rc = drain_slot(b, slot, h_out, out_cap, h_offsets, h_lens, &run);  // sub-batch k - depth
        if (rc != DMMT_OK) break;
        if (!p->d_pixels_own) DMMT_CUDA(cudaMalloc(&p->d_pixels_own, (size_t)p->n * p->pixel_bytes));
        if (!p->d_dense) {
            p->dense_cap = (size_t)p->n * p->out_stride;
            DMMT_CUDA(cudaMalloc(&p->d_dense, p->dense_cap));
        }
        DMMT_CUDA(cudaMemcpyAsync(p->d_pixels_own, static_cast<const uint8_t*>(h_pixels) + (size_t)first * img_bytes,
                                  (size_t)m * img_bytes, cudaMemcpyHostToDevice, p->stream));
        rc = dmmt_plan_encode_compact(p, p->d_pixels_own, m, p->d_dense, p->dense_cap, p->d_offsets, p->d_lens, 0,
                                      b->d_err);
        if (rc != DMMT_OK) break;
        launches += p->last_launches;
        DMMT_CUDA(cudaMemcpyAsync(p->h_lens, p->d_lens, (size_t)m * 8, cudaMemcpyDeviceToHost, p->stream));
        DMMT_CUDA(cudaMemcpyAsync(p->h_offsets, p->d_offsets, (size_t)(m + 1) * 8, cudaMemcpyDeviceToHost, p->stream));
        DMMT_CUDA(cudaEventRecord(b->ev_meta[slot], p->stream));
        b->pending[slot].active = true, b->pending[slot].first = first, b->pending[slot].m = m;
    }
    // flush the remaining sub-batches in issue order
    for (int k = std::max(0, n_sub - b->depth); k < n_sub && rc == DMMT_OK; k++)
        rc = drain_slot(b, k % b->depth, h_out, out_cap, h_offsets, h_lens, &run);
    for (auto& pd : b->pending) pd.active = false;
    b->last_launches = launches;
    const int st = dmmt_batch_status(b);  // synchronises every slot (also on the error path)
    if (rc != DMMT_OK) return rc;
    return st;
}
