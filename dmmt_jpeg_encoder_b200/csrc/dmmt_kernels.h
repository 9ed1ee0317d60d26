// dmmt_kernels.h -- host launchers of the sm_100a kernels (k1_transform.cu, k2_entropy.cu),
// called by the C-ABI layer (dmmt_api.cu).  Everything is asynchronous on the given stream.
#pragma once

#include "dmmt_common.cuh"

namespace dmmt {

// K1 (k1_transform.cu)
struct K1Consts {
    float maxf, r_hi, r_lo;  // max value and 1/max split in two f32
    int exact;               // 1: use the IEEE-division kernels (fast forms not proven for this plan)
    QuantF qf, rq_hi, rq_lo; // divisors and 1/q split in two f32
};
// Fused K1 output (4:2:0 fast path): one token region per 256x16-pixel tile, tiles in stream order.
struct TileTok {
    uint32_t* tok;            // [n][tiles][tile_cap]
    size_t img_stride_words;
    uint32_t tile_cap;        // tokens per tile (multiple of 8)
    uint32_t* ntok;           // [n][tiles]
    int16_t* last_dc;         // [n][tiles][4]: quantised DC of the tile's last Y, Cb, Cr block
    uint32_t* dcpos;          // [n][tiles][2]: token index of the tile's first Cb / Cr DC token
    uint32_t tiles_x, tiles;
};
bool k1_fused_supported(const Geom& g, const K1Consts& c);
uint32_t k1_tiles_x(const Geom& g);
cudaError_t launch_k2_fix_dc(const TileTok& fo, int n, unsigned int* hist, ImgMeta* meta, const int16_t* seed_dc,
                             cudaStream_t st);
void make_k1_consts(int fmt, int max_value, const uint8_t* q_luma, const uint8_t* q_chroma, K1Consts* c);
cudaError_t launch_k1(const Geom& g, int fmt, const K1Consts& c, int check_max, const void* d_pixels,
                      size_t img_stride_bytes, int n_images, int16_t* d_coef, size_t coef_img_stride,
                      float* d_dbg, ImgMeta* meta, const TileTok* fused, unsigned int* hist, cudaStream_t st);

// K2 (k2_entropy.cu): tokens of chunk c (256 stream blocks) of image i live at
// tok + i * img_stride_words + c * chunk_cap; ntok[i * n_chunks + c] of them are valid.
struct TokBuf {
    uint32_t* tok;
    size_t img_stride_words;
    uint32_t chunk_cap;        // tokens per chunk (multiple of 4)
    uint32_t* ntok;
};
uint32_t tok_blocks_per_chunk();
cudaError_t launch_k2(const Geom& g, const int16_t* coef, size_t coef_img_stride, int n,
                      unsigned int* hist, ImgMeta* meta, const int16_t* seed_dc, const TokBuf& tb,
                      cudaStream_t st);

struct K2bHostArgs {
    unsigned int* hist;               // [n][4][256] u32 local counts (the fused DC fix-up adds its symbols)
    const unsigned long long* ghist;  // optional [4][256] u64 global counts (sharded mode)
    EncTables* enc;
    LenTables* lens;
    ImgMeta* meta;
    uint8_t* out;
    size_t out_stride;
    unsigned long long scan_cap_bits;
    int W, H, bits_per_channel;       // ORIGINAL size for SOF0
    const uint8_t* qtab_luma;         // natural order
    const uint8_t* qtab_chroma;
    int write_header;
    uint8_t* lcount;                  // [n][4][16] scratch: codes per length (DHT)
    const TileTok* fix;               // fused 4:2:0 chain: the DC-table CTAs finish the tile-boundary DC tokens first
                                      // (nullptr: done already, by launch_k2_fix_dc or by K2)
};
cudaError_t launch_k2b(const Geom& g, const K2bHostArgs& h, int n, cudaStream_t st);

cudaError_t launch_zero_scan(uint32_t* scan, size_t stride_words, const ImgMeta* meta, int n,
                             unsigned long long seed_bits, int blocks_per_image, cudaStream_t st,
                             const unsigned long long* seed_src = nullptr);

uint32_t k3_chunks(const Geom& g);
uint32_t k4_max_chunks(size_t scan_cap_bytes);

// n_segs == 0: chunk c = token region c (K2's 256-block chunks), sliced over the 8 warps of a CTA;
// n_segs > 0 (fused K1): region = tile, chunk c = tiles [8c, 8c + 8), one warp per tile.
cudaError_t launch_k3(uint32_t n_chunks, uint32_t n_segs, int n, const TokBuf& tb, const EncTables* enc, ImgMeta* meta,
                      unsigned long long* lb_state, unsigned int* ticket, uint32_t* scan, size_t scan_stride_words,
                      unsigned long long seed_bits, int pad_ones, cudaStream_t st,
                      const unsigned long long* seed_src = nullptr);

struct K4HostArgs {
    const uint8_t* scan;
    size_t scan_stride_bytes;
    ImgMeta* meta;
    unsigned long long* lb_state;
    unsigned int* ticket;
    uint32_t max_chunks;
    uint8_t* out;
    size_t out_stride;
    unsigned long long* out_lens;
    unsigned long long first_byte;
    long long n_bytes_override;
    unsigned long long seed_bits;
    int prepend_header, append_eoi;
    uint8_t or_first_byte;
    const unsigned long long* seed_src = nullptr;  // device-resident shard exchange (see K4Args)
    int owned_mode = 0;
    const int* or_first_src = nullptr;
    const unsigned long long* base_src = nullptr;  // peer-memory gather: byte offset of the shard in the file
};
cudaError_t launch_k4(const K4HostArgs& h, int n, uint32_t grid_chunks, cudaStream_t st);

cudaError_t launch_last_dc(const Geom& g, const int16_t* coef, int16_t* d_out3, cudaStream_t st);
// single-process shard exchange over peer memory (see k_peer_exchange)
constexpr int DMMT_MAX_PEER_SHARDS = 64;
struct PeerPtrs {
    const void* p[DMMT_MAX_PEER_SHARDS];
};
cudaError_t launch_peer_exchange(const PeerPtrs& srcs, int n, int elems, int mode, long long* out, cudaStream_t st);
// mailbox exchange between processes (see k_mailbox_post / k_mailbox_collect)
size_t mailbox_bytes(int world, int slots);
cudaError_t launch_mailbox_post(const PeerPtrs& boxes, int world, int rank, int slot, unsigned long long seq, const void* src,
                                int n_words, cudaStream_t st);
cudaError_t launch_mailbox_collect(void* box, int world, int slot, unsigned long long seq, int mode, int n_words, long long* out,
                                   ImgMeta* meta, cudaStream_t st);
// device-resident shard exchange helpers
cudaError_t launch_shard_widen(const int16_t* last_dc3, int* out4, const unsigned int* hist, long long* hist64,
                               const ImgMeta* meta, long long* bits_out, cudaStream_t st, long long* err_out = nullptr);
cudaError_t launch_shard_count_bytes(const uint8_t* scan, const ImgMeta* meta, const unsigned long long* seed_src,
                                     int owned_mode, const int* or_first_src, int is_first, int is_last,
                                     unsigned long long* ctr2, long long* n_bytes, cudaStream_t st);
cudaError_t launch_shard_copy_header(const uint8_t* own_out, const ImgMeta* meta, uint8_t* file, size_t capacity,
                                     cudaStream_t st);
cudaError_t launch_shard_result(const ImgMeta* meta, long long* out2, cudaStream_t st);
cudaError_t launch_shard_narrow_seed(const int* seed4, int16_t* seed3, cudaStream_t st);
cudaError_t launch_shard_tail(const uint8_t* scan, const ImgMeta* meta, const unsigned long long* bit_offset, int is_last,
                              int* tail2, cudaStream_t st);
cudaError_t launch_shard_prev_tail(const int* all_tail2, const long long* all_offs, const long long* all_bits, int rank,
                                   int* out, cudaStream_t st);

// K5: packs the n files of an output arena ([n][out_stride]) back to back (16-byte aligned
// starts) into `dense`; offsets[0..n) and offsets[n] (= end) are written on the device.  chained:
// start at the value already stored in offsets[0] (end of the previous sub-batch).
cudaError_t launch_k5_compact(const uint8_t* out, size_t out_stride, const unsigned long long* lens,
                              int n, uint8_t* dense, unsigned long long dense_cap,
                              unsigned long long* offsets, ImgMeta* meta, int chained,
                              int* sticky_err, cudaStream_t st);

}  // namespace dmmt
