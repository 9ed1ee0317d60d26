// k2_entropy.cu -- the entropy stage of the sm_100a encode path:
//   K2  tokenise (DC diff / RLE / category) + symbol histogram
//                                            (categorize.rs:21-169, symbol_counting.rs:55-74)
//   K2b Huffman table construction + headers (symbol_counting.rs:85-94, huffman/length_limited.rs:37-134,
//                                             huffman/encoder.rs:37-157, jpeg/encoder.rs:137-262)
//   K3  tokens -> per-warp private bit buffers (one pass) -> decoupled look-back exclusive scan over
//       chunks -> funnel-shift copy to the final bit phase
//                                            (jpeg/encoder.rs:264-404, binary_stream.rs:38-96)
//   K4  0xFF byte stuffing as a scan-compaction (segment_marker_injector.rs:13-30) + EOI
// All file:line citations are relative to /root/reference/src.
//
// Generic path: the input is K1's coefficient stream: int16 [n_blocks][64], zig-zag inside a block,
// MCU-interleaved stream order.  (On the fused 4:2:0 path K1 tokenises on-chip and K2 is replaced by
// k2_fix_dc in k1_transform.cu; K2b..K5 are shared.)  K2 stages 256 blocks (32 KB) per CTA into shared memory with a 16-byte-chunk XOR
// swizzle (chunk c of block b at slot c ^ (b & 7)), so the coalesced global loads AND the
// thread-per-block 128-bit shared loads are both conflict-free; each thread then walks only the
// NON-ZERO coefficients of its block (occupancy mask + ffs) ONCE, emitting one 32-bit token per
// coded coefficient (the reference's CategorizedBlock, categorize.rs:101-104, flattened) and counting
// symbols.  K3 never sees coefficients: every lane owns runs of 8 consecutive tokens, so the serial,
// divergent walk happens once per image and the bit packing is balanced.
//
// Token word: bits 0-7 symbol, 8-9 table (T_*), 10-11 number of ZRL (0xF0) codes that precede the
// symbol (categorize.rs:139-142), 16-31 the category's extra bits.  bits 0-9 index the encoder LUT; the symbol
// field of AC tokens is bank-swizzled (tok_swz, dmmt_common.cuh).
#include "dmmt_kernels.h"

namespace dmmt {

namespace {

constexpr int EB = 256;           // blocks per CTA chunk in K2/K3 (= threads)
#ifndef K4_THREADS_N
#define K4_THREADS_N 256
#endif
constexpr int K4_THREADS = K4_THREADS_N;
constexpr int K4_BYTES_PER_THREAD = 32;
constexpr int K4_CHUNK = K4_THREADS * K4_BYTES_PER_THREAD;  // 8 KB of unstuffed scan per CTA

// ---- staging: 256 blocks x 128 B -> swizzled shared memory --------------------------------
__device__ __forceinline__ void stage_chunk(uint4* s_coef, const int16_t* __restrict__ coef,
                                            uint32_t first_block, uint32_t n_blocks) {
    const uint4* g = reinterpret_cast<const uint4*>(coef) + (size_t)first_block * 8;
    const uint32_t avail = (n_blocks - first_block < EB) ? (n_blocks - first_block) : EB;
#pragma unroll
    for (int i = 0; i < 8; i++) {
        const uint32_t gi = i * EB + threadIdx.x;  // 16-byte chunk index within the CTA chunk
        const uint32_t b = gi >> 3, c = gi & 7;
        uint4 v = make_uint4(0, 0, 0, 0);
        if (b < avail) v = __ldg(g + gi);
        s_coef[b * 8 + (c ^ (b & 7))] = v;
    }
}

__device__ __forceinline__ int s_coef_at(const uint4* s_coef, int b, int pos) {
    const int16_t* p = reinterpret_cast<const int16_t*>(s_coef + b * 8 + ((pos >> 3) ^ (b & 7)));
    return p[pos & 7];
}

// previous block of the same component in stream order (categorize.rs:157-161 runs one DC chain
// per component over the entangled/stream order); -1 => first block of the component
__device__ __forceinline__ long long prev_block_same_comp(uint32_t s, int ypm, int bpm) {
    const uint32_t m = s / bpm, k = s - m * bpm;
    if ((int)k < ypm) {
        if (k > 0) return (long long)s - 1;
        return m ? (long long)(m - 1) * bpm + ypm - 1 : -1;
    }
    return m ? (long long)s - bpm : -1;
}

// categorize.rs:22-41 for a non-zero value: category and the cat low bits of the pattern
__device__ __forceinline__ void cat_bits(int v, int& cat, uint32_t& bits) {
    const int a = abs(v);
    cat = 32 - __clz(a);
    bits = (uint32_t)(v > 0 ? v : v - 1) & ((1u << cat) - 1u);
}

// =========================================== K2 ===========================================
struct K2Args {
    const int16_t* coef;
    size_t coef_img_stride;
    uint32_t n_blocks;
    uint32_t n_chunks;
    int ypm, bpm;
    unsigned int* hist;        // [n][4][256]
    ImgMeta* meta;             // [n]
    const int16_t* seed_dc;    // optional [3] predictors for the first block of each component (shards)
    TokBuf tb;
};

// 64-bit occupancy mask of the thread's own block (bit i set <=> zig-zag coefficient i != 0)
__device__ __forceinline__ void block_mask32(const uint4* s_coef, int b, uint32_t& lo, uint32_t& hi) {
    uint32_t m[2] = {0u, 0u};
#pragma unroll
    for (int c = 0; c < 8; c++) {
        const uint4 v = s_coef[b * 8 + (c ^ (b & 7))];
        // per word: 0xFFFF in each non-zero half -> pick one byte of each half of two words -> 4 flag bytes
        const uint32_t f01 = __byte_perm(__vcmpne2(v.x, 0u), __vcmpne2(v.y, 0u), 0x6420);
        const uint32_t f23 = __byte_perm(__vcmpne2(v.z, 0u), __vcmpne2(v.w, 0u), 0x6420);
        // bytes (0xFF / 0x00) -> bits 0..3 via a multiply that sums the masked bytes into the top byte
        const uint32_t b01 = ((f01 & 0x08040201u) * 0x01010101u) >> 24;
        const uint32_t b23 = ((f23 & 0x08040201u) * 0x01010101u) >> 24;
        m[c >> 2] |= (b01 | (b23 << 4)) << (8 * (c & 3));
    }
    lo = m[0], hi = m[1];
}

__device__ __forceinline__ uint32_t make_token(int table, int sym, int nzrl, uint32_t extra) {
    return (uint32_t)sym | ((uint32_t)table << 8) | ((uint32_t)nzrl << 10) | (extra << 16);
}

__global__ void __launch_bounds__(EB) k2_tokenize(const K2Args a) {
    __shared__ uint4 s_coef[EB * 8];
    __shared__ unsigned int s_hist[4 * 256];
    __shared__ uint32_t s_warp[EB / 32 + 1];
    const int img = blockIdx.y;
    const int16_t* __restrict__ coef = a.coef + (size_t)img * a.coef_img_stride;
    const uint32_t chunk = blockIdx.x, first = chunk * EB;
    for (int i = threadIdx.x; i < 1024; i += EB) s_hist[i] = 0;
    stage_chunk(s_coef, coef, first, a.n_blocks);
    __syncthreads();
    const uint32_t s = first + threadIdx.x;
    const bool active = s < a.n_blocks;
    const int b = threadIdx.x;
    uint32_t mlo = 0, mhi = 0, cnt = 0;
    if (active) {
        block_mask32(s_coef, b, mlo, mhi);
        // tokens of the block: DC + one per non-zero AC (ZRLs ride on it) + EOB unless coefficient 63 != 0
        cnt = 1u + __popc(mlo & ~1u) + __popc(mhi) + ((mhi >> 31) ? 0u : 1u);
    }
    uint32_t total;
    uint32_t off = block_exclusive_scan<EB>(cnt, s_warp, &total);
    const bool fits = total <= a.tb.chunk_cap;
    if (threadIdx.x == 0) {
        a.tb.ntok[(size_t)img * a.n_chunks + chunk] = fits ? total : 0u;
        if (!fits) atomicCAS(&a.meta[img].error, 0, DMMT_E_OVERFLOW);  // host retries with the worst-case capacity
    }
    if (active) {
        uint32_t* __restrict__ tok = a.tb.tok + (size_t)img * a.tb.img_stride_words + (size_t)chunk * a.tb.chunk_cap;
        const uint32_t m_idx = s / a.bpm, k = s - m_idx * a.bpm;
        const int comp = (int)k < a.ypm ? 0 : ((int)k == a.ypm ? 1 : 2);
        const int tdc = comp ? T_CDC : T_YDC, tac = tdc + 1;
        const long long pb = prev_block_same_comp(s, a.ypm, a.bpm);
        const int dc_pred = pb >= 0 ? (int)coef[(size_t)pb * 64] : (a.seed_dc ? (int)a.seed_dc[comp] : 0);
        bool ok = true;
        {   // DC: categorize.rs:159-160
            const int dc = s_coef_at(s_coef, b, 0);
            const int diff = (int)(int16_t)(dc - dc_pred);
            int cat = 0;
            uint32_t bits = 0;
            if (diff != 0) cat_bits(diff, cat, bits);
            ok &= cat <= 15;
            atomicAdd(&s_hist[tdc * 256 + (cat & 15)], 1u);
            if (fits) tok[off] = make_token(tdc, cat & 15, 0, bits);
            ++off;
        }
        // AC: categorize.rs:132-151 (run of zeros; 0xF0 per 16 zeros; EOB 0x00 when the tail is zero)
        int prev = 0;
        uint32_t nzrl_total = 0;
#pragma unroll
        for (int half = 0; half < 2; half++) {
            uint32_t m = half ? mhi : (mlo & ~1u);
            while (m) {
                const int pos = 32 * half + __ffs((int)m) - 1;
                m &= m - 1;
                const int run = pos - prev - 1;
                prev = pos;
                const int v = s_coef_at(s_coef, b, pos);
                int cat;
                uint32_t bits;
                cat_bits(v, cat, bits);
                ok &= cat <= 15;
                const int sym = ((run & 15) << 4) | (cat & 15);
                nzrl_total += (uint32_t)(run >> 4);
                atomicAdd(&s_hist[tac * 256 + sym], 1u);
                if (fits) tok[off] = make_token(tac, (int)tok_swz((uint32_t)tac, (uint32_t)sym), run >> 4, bits);
                ++off;
            }
        }
        if (nzrl_total) atomicAdd(&s_hist[tac * 256 + 0xF0], nzrl_total);
        if (prev != 63) {
            atomicAdd(&s_hist[tac * 256], 1u);
            if (fits) tok[off] = make_token(tac, (int)tok_swz((uint32_t)tac, 0x00u), 0, 0u);
        }
        if (!ok) atomicCAS(&a.meta[img].error, 0, DMMT_E_RANGE);
    }
    __syncthreads();
    unsigned int* gh = a.hist + (size_t)img * 1024;
    for (int i = threadIdx.x; i < 1024; i += EB) {
        const unsigned int v = s_hist[i];
        if (v) atomicAdd(&gh[i], v);
    }
}

// =========================================== K2b ==========================================
// One CTA (512 threads) per (table, image).  Package-merge with limit 15 in shared memory:
// every level is a parallel merge-by-rank of the sorted leaves with the pairwise packages of the
// previous level (Leaf < Package on equal frequency, length_limited.rs:7-26,104-115): threads 0..255 place the
// packages, threads 256..511 the leaves, at the same time.  On the fused 4:2:0 chain the two DC-table CTAs first finish
// the tile-boundary DC tokens (what k2_fix_dc does as a kernel of its own on the sharded path) -- the AC tables, which
// take longer anyway, do not depend on them.  The CTA that finishes last writes the header of the image.
struct K2bArgs {
    unsigned int* hist;            // [n][4][256] local counts (u32); the DC fix-up adds its symbols
    const unsigned long long* ghist;  // optional [4][256] u64 GLOBAL counts (sharded mode), else nullptr
    EncTables* enc;                // [n]
    LenTables* lens;               // [n]
    ImgMeta* meta;                 // [n]
    uint8_t* out;                  // [n][out_stride]  (header is written at offset 0)
    size_t out_stride;
    unsigned long long scan_cap_bits;
    int W, H;                      // ORIGINAL width/height for SOF0 (transformer.rs:210-211)
    int hr, vr;
    int bits_per_channel;
    uint8_t qzz[2][64];            // quantisation tables in zig-zag order for DQT
    uint32_t n_stream_blocks;
    int write_header;
    uint8_t* lcount;               // [n][4][16]: codes per length of every table, for the DHT segments
    int fix_dc;                    // 1: finish the tile-boundary DC tokens of the fused K1 (fo) first
    TileTok fo;
};

constexpr int PM_LIMIT = 15;
constexpr int K2B_THREADS = 512;

__global__ void __launch_bounds__(K2B_THREADS) k2b_tables(const __grid_constant__ K2bArgs a) {
    __shared__ unsigned long long s_freq[256];        // sorted leaf frequencies
    __shared__ uint8_t s_sym[256];                    // sorted symbols (ascending frequency, ties by symbol)
    __shared__ unsigned long long s_list[2][512];     // ping-pong merged lists (frequencies)
    __shared__ uint16_t s_pkgpos[PM_LIMIT][256];      // position of package j in level k's merged list
    __shared__ int s_npk[PM_LIMIT];                   // number of packages in level k
    __shared__ int s_nleaves[PM_LIMIT];               // back-trace: leaves taken at level k
    __shared__ int s_len[256];
    __shared__ uint32_t s_scan[256];
    __shared__ unsigned int s_lcount[17];
    __shared__ unsigned int s_fix[16];
    __shared__ unsigned long long s_bits;
    __shared__ uint32_t s_warp[K2B_THREADS / 32 + 1];
    __shared__ int s_err, s_is_last;

    pdl_wait();
    pdl_trigger();
    // grid (image, slot): the two AC tables of every image are dispatched before any DC table -- the AC tables take the
    // whole 30 us, the DC tables (<= 12 symbols) a third of it, and with a launch that does not fit one wave the AC CTAs of
    // the later images must not queue behind DC CTAs of the earlier ones
    const int t = (blockIdx.y < 2 ? 1 : 0) + 2 * (blockIdx.y & 1), img = blockIdx.x, tid = threadIdx.x;
    const int nsym_max = (t & 1) ? 256 : 16;
    unsigned int* lh = a.hist + (size_t)img * 1024;
    ImgMeta* meta = a.meta + img;

    if (tid < 17) s_lcount[tid] = 0;
    if (tid < 16) s_fix[tid] = 0;
    if (tid == 0) s_bits = 0ull, s_err = 0;
    __syncthreads();
    if (a.fix_dc && !(t & 1)) {
        // DC difference of the first Y (table 0) / Cb and Cr (table 2) block of every tile against the last DC of the
        // previous tile (tile 0: the chain starts at 0, categorize.rs:157): token patched, symbol counted
        const TileTok& fo = a.fo;
        for (uint32_t tile = tid; tile < fo.tiles; tile += K2B_THREADS) {
            const size_t ti = (size_t)img * fo.tiles + tile;
            if (fo.ntok[ti] == 0) continue;  // overflowed tile (error already flagged)
            uint32_t* tok = fo.tok + (size_t)img * fo.img_stride_words + (size_t)tile * fo.tile_cap;
            for (int c = (t == T_YDC ? 0 : 1); c < (t == T_YDC ? 1 : 3); c++) {
                const int pred = tile ? (int)fo.last_dc[(ti - 1) * 4 + c] : 0;
                const uint32_t pos = c == 0 ? 0u : fo.dcpos[ti * 2 + (c - 1)];
                const int dc = (int)(short)(tok[pos] >> 16);
                const int diff = (int)(short)(dc - pred);
                int cat = 0;
                uint32_t bits = 0;
                if (diff != 0) cat_bits(diff, cat, bits);
                if (cat > 15) atomicCAS(&meta->error, 0, DMMT_E_RANGE);
                tok[pos] = make_token(t, cat & 15, 0, bits);
                atomicAdd(&s_fix[cat & 15], 1u);
            }
        }
        __syncthreads();
        if (tid < 16 && s_fix[tid]) lh[t * 256 + tid] += s_fix[tid];  // this CTA is the only one that touches these bins
        __syncthreads();
    }
    const unsigned long long myf = (tid < nsym_max)
                                       ? (a.ghist ? a.ghist[t * 256 + tid] : (unsigned long long)lh[t * 256 + tid])
                                       : 0ull;
    if (tid < 256) s_list[0][tid] = myf;  // scratch: raw frequencies by symbol
    const int n = __syncthreads_count(myf != 0ull);  // symbol_counting.rs:25-32: symbols with a non-zero count
    // stable sort by frequency of the ascending-symbol list (symbol_counting.rs:92-94): rank sort
    if (myf != 0ull) {
        int rank = 0;
        for (int j = 0; j < nsym_max; j++) {
            const unsigned long long fj = s_list[0][j];
            rank += (fj != 0ull) && (fj < myf || (fj == myf && j < tid));
        }
        s_freq[rank] = myf;
        s_sym[rank] = (uint8_t)tid;
    }
    __syncthreads();
    bool ok_table = n != 0;
    if (!ok_table && tid == 0) atomicCAS(&meta->error, 0, DMMT_E_INVALID);  // no blocks at all: the reference panics (symbol_counting.rs:88)

    // ---- package-merge levels (length_limited.rs:63-73,104-115) ----
    int cur = 1;
    int len_prev = n;
    if (tid < n) s_list[1][tid] = s_freq[tid];
    if (tid == 0) s_npk[0] = 0;
    __syncthreads();
    for (int k = 1; k < PM_LIMIT && ok_table; k++) {
        const unsigned long long* prev = s_list[cur];
        unsigned long long* nxt = s_list[cur ^ 1];   // last read in level k - 1, before that level's closing barrier
        const int np = len_prev >> 1;  // chunks_exact(2)
        if (tid < np) {
            // package j = prev[2j] + prev[2j + 1]; leaves with freq <= pf come first (leaf wins ties)
            const unsigned long long pf = prev[2 * tid] + prev[2 * tid + 1];
            int lo = 0, hi = n;
            while (lo < hi) {
                const int mid = (lo + hi) >> 1;
                if (s_freq[mid] <= pf) lo = mid + 1;
                else hi = mid;
            }
            const int pos = tid + lo;
            nxt[pos] = pf;
            s_pkgpos[k][tid] = (uint16_t)pos;
        } else if (tid >= 256 && tid - 256 < n) {
            // packages with freq < leaf come first.  Package frequencies are non-decreasing in j
            // (pair sums of a sorted list), so binary search over j on the fly.
            const int i = tid - 256;
            const unsigned long long lf = s_freq[i];
            int lo = 0, hi = np;
            while (lo < hi) {
                const int mid = (lo + hi) >> 1;
                const unsigned long long pm = prev[2 * mid] + prev[2 * mid + 1];
                if (pm < lf) lo = mid + 1;
                else hi = mid;
            }
            nxt[i + lo] = lf;
        }
        if (tid == 0) s_npk[k] = np;
        len_prev = n + np;
        cur ^= 1;
        __syncthreads();
    }
    // ---- back-trace (length_limited.rs:75-89,117-133): one warp, per level a 32-ary search in two rounds over the
    // (increasing) package positions instead of a binary search ----
    if (tid < 32 && ok_table) {
        int p = n - 1;
        for (int k = PM_LIMIT - 1; k >= 0; k--) {
            const int count = 2 * p;
            const int npk = s_npk[k];
            if (count > n + npk) {  // slice panic in the reference; impossible for n <= 2^15
                if (tid == 0) s_err = DMMT_E_INVALID, s_nleaves[k] = 0;
                p = 0;
                continue;
            }
            // lo = number of packages with position < count
            const int i1 = tid * 8 + 7;
            const int c1 = __popc(__ballot_sync(0xffffffffu, i1 < npk && (int)s_pkgpos[k][i1] < count));
            const int i2 = c1 * 8 + tid;
            const int lo = c1 * 8 + __popc(__ballot_sync(0xffffffffu, tid < 8 && i2 < npk && (int)s_pkgpos[k][i2] < count));
            if (tid == 0) s_nleaves[k] = count - lo;
            p = lo;
        }
    }
    __syncthreads();
    // sum_up_codeword_lengths (length_limited.rs:91-102) then the +1 quirk (symbol_counting.rs:88)
    int mylen = 0;
    if (tid < n) {
#pragma unroll
        for (int k = 0; k < PM_LIMIT; k++) mylen += (tid < s_nleaves[k]);
        if (tid == 0) mylen += 1;
        s_len[tid] = mylen;
        atomicAdd(&s_lcount[mylen <= 16 ? mylen : 0], 1u);
        if (mylen < 1 || mylen > 16) s_err = DMMT_E_INVALID;
    }
    // ---- canonical codes (huffman/encoder.rs:45-67,116-119): walk from the END of the table:
    // code(i) = sum_{j > i} 2^(16 - len_j)  -> exclusive scan over the reversed order ----
    const int ri = n - 1 - tid;  // reversed index handled by this thread
    uint32_t inc = 0;
    __syncthreads();
    if (tid < n) inc = 1u << (16 - min(max(s_len[ri], 0), 16));
    uint32_t total;
    const uint32_t excl = block_exclusive_scan<K2B_THREADS>(tid < n ? inc : 0u, s_warp, &total);
    if (tid < n) {
        const int len = s_len[ri];
        if (excl > 0xFFFFu) s_err = DMMT_E_INVALID;
        const uint32_t code = (excl & 0xFFFFu) >> (16 - min(max(len, 0), 16));  // right-aligned
        s_scan[ri] = ((uint32_t)len << 16) | code;
    }
    __syncthreads();
    // LUT by symbol + length tables + this table's share of the scan bits
    EncTables* enc = a.enc + img;
    LenTables* lt = a.lens + img;
    if (tid < 256) {
        enc->e[t][tid] = 0u;
        lt->sym[t][tid] = 0;
        lt->len[t][tid] = 0;
    }
    __syncthreads();
    if (tid < n) {
        const int sym = s_sym[tid];
        enc->e[t][sym] = s_scan[tid];
        lt->sym[t][tid] = (uint8_t)sym;
        lt->len[t][tid] = (uint8_t)s_len[tid];
        // bits contributed by this symbol in THIS image/shard: local count x (code length + category bits)
        const unsigned long long cnt = lh[t * 256 + sym];
        atomicAdd(&s_bits, cnt * (unsigned long long)(s_len[tid] + (sym & 15)));
    }
    if (tid < 16) a.lcount[((size_t)img * 4 + t) * 16 + tid] = (uint8_t)s_lcount[tid + 1];
    __syncthreads();
    if (tid == 0) {
        atomicAdd(&meta->scan_bits, s_bits);
        meta->n_symbols[t] = (uint32_t)n;
        if (s_err) atomicCAS(&meta->error, 0, s_err);
        __threadfence();  // tables, counts and bits of this table are visible before the arrival is
        s_is_last = atomicAdd(&meta->reserved, 1u) == 3u;
    }
    __syncthreads();
    if (!s_is_last) return;
    // ---- the last of the image's four CTAs: capacity check and headers (jpeg/encoder.rs:125-262) ----
    __threadfence();
    if (tid == 0 && *reinterpret_cast<volatile unsigned long long*>(&meta->scan_bits) + 8 > a.scan_cap_bits)
        atomicCAS(&meta->error, 0, DMMT_E_OVERFLOW);
    if (!a.write_header) return;
    // layout: SOI(2) APP0(18) DQT(69) DQT(69) SOF0(19) | DHT YAC, YDC, CAC, CDC (21 + n each) | SOS(14)
    uint8_t* out = a.out + (size_t)img * a.out_stride;
    const volatile uint32_t* ns = meta->n_symbols;
    const uint32_t nY_AC = ns[T_YAC], nY_DC = ns[T_YDC], nC_AC = ns[T_CAC], nC_DC = ns[T_CDC];
    const uint32_t hdr_len = 177 + 84 + nY_AC + nY_DC + nC_AC + nC_DC + 14;
    {
        // DHT: FF C4, len = 2 + 17 + n, Tc/Th id (encoder.rs:78-84), counts[16] (:92-98),
        // symbols in REVERSE table order (:177); a quarter of the CTA per table, in file order
        const int q = tid >> 7, j = tid & 127;
        const int order[4] = {T_YAC, T_YDC, T_CAC, T_CDC};
        const uint8_t ids[4] = {0x00, 0x11, 0x02, 0x13};
        const int tt = order[q];
        const uint32_t nn = ns[tt];
        uint32_t off = 177;
        if (q >= 1) off += 21 + nY_AC;
        if (q >= 2) off += 21 + nY_DC;
        if (q >= 3) off += 21 + nC_AC;
        uint8_t* p = out + off;
        if (j == 0) {
            p[0] = 0xFF, p[1] = 0xC4;
            p[2] = (uint8_t)((19 + nn) >> 8), p[3] = (uint8_t)((19 + nn) & 0xFF);
            p[4] = ids[tt];
        }
        if (j < 16) p[5 + j] = __ldcg(&a.lcount[((size_t)img * 4 + tt) * 16 + j]);
        for (uint32_t i = j; i < nn; i += 128) p[21 + i] = __ldcg(&lt->sym[tt][nn - 1 - i]);
    }
    if (tid == 0) {
        uint8_t* p = out;
        const uint8_t fixed[20] = {0xFF, 0xD8, 0xFF, 0xE0, 0x00, 0x10, 'J', 'F', 'I', 'F', 0,
                                   0x01, 0x02, 0x00, 0x00, 0x48, 0x00, 0x48, 0x00, 0x00};
        for (int i = 0; i < 20; i++) p[i] = fixed[i];
        p = out + 158;  // SOF0 (encoder.rs:227-245)
        const uint8_t sof[19] = {0xFF, 0xC0, 0x00, 0x11, (uint8_t)a.bits_per_channel,
                                 (uint8_t)(a.H >> 8), (uint8_t)a.H, (uint8_t)(a.W >> 8), (uint8_t)a.W,
                                 0x03, 0x01, (uint8_t)((a.hr << 4) | a.vr), 0x00,
                                 0x02, 0x11, 0x01, 0x03, 0x11, 0x01};
        for (int i = 0; i < 19; i++) p[i] = sof[i];
        p = out + hdr_len - 14;  // SOS (encoder.rs:247-262)
        const uint8_t sos[14] = {0xFF, 0xDA, 0x00, 0x0C, 0x03, 0x01, 0x01, 0x02, 0x23, 0x03, 0x23,
                                 0x00, 0x3F, 0x00};
        for (int i = 0; i < 14; i++) p[i] = sos[i];
        meta->header_len = hdr_len;
        meta->n_stream_blocks = a.n_stream_blocks;
    }
    if (tid >= 256 && tid < 256 + 138) {  // two DQT segments (encoder.rs:190-209), 69 bytes each
        const int x = tid - 256, which = x / 69, i = x % 69;
        uint8_t v;
        if (i == 0) v = 0xFF;
        else if (i == 1) v = 0xDB;
        else if (i == 2) v = 0x00;
        else if (i == 3) v = 0x43;
        else if (i == 4) v = (uint8_t)which;
        else v = a.qzz[which][i - 5];
        out[20 + x] = v;
    }
}

// Zeroes the words of the unstuffed scan buffers that K3 will OR into (sizes are only known on
// the device): grid (Z, n), each row strides over that image's words.
__global__ void k_zero_scan(uint32_t* scan, size_t scan_img_stride_words, const ImgMeta* meta,
                            unsigned long long seed_bits, const unsigned long long* seed_src) {
    pdl_wait();
    pdl_trigger();
    if (seed_src) seed_bits = *seed_src & 7ull;  // device-resident shard exchange
    const int img = blockIdx.y;
    if (meta[img].error) return;
    const unsigned long long words = (meta[img].scan_bits + seed_bits + 8 + 31) / 32 + 1;
    uint32_t* p = scan + (size_t)img * scan_img_stride_words;
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < words;
         i += (unsigned long long)gridDim.x * blockDim.x)
        p[i] = 0u;
}

// =========================================== K3 ===========================================
struct K3Args {
    uint32_t n_chunks;
    uint32_t n_segs;               // tile mode: token regions (tiles) per image, 8 per chunk, one per warp; else 0
    TokBuf tb;
    const EncTables* enc;          // [n]
    ImgMeta* meta;                 // [n]
    unsigned long long* lb_state;  // [n][n_chunks]  zero-initialised per run
    unsigned int* ticket;          // [n]            zero-initialised per run
    uint32_t* scan;                // [n][scan_img_stride_words] big-endian bit stream, zero-initialised by k_zero_scan
    size_t scan_img_stride_words;
    unsigned long long seed_bits;  // bit offset of this shard's first bit inside its first byte/word (< 32)
    int pad_ones;                  // append the 1-padding after the last block (binary_stream.rs:89-96)
    const unsigned long long* seed_src;  // optional: global bit offset of the shard in device memory (its low 3 bits are the seed)
    uint32_t n_items;              // n * n_chunks work items (see k3_pack)
};

// code bits of one token: ZRL codes + symbol code + category bits (encoder.rs:356-404)
__device__ __forceinline__ uint32_t token_bits(uint32_t t, uint32_t e, uint32_t zl_y, uint32_t zl_c) {
    const uint32_t nz = (t >> 10) & 3u;
    return (e >> 16) + (t & 15u) + nz * ((t & 0x200u) ? zl_c : zl_y);
}

// ORs `len` (<= 32) bits, right-aligned in `v`, at bit position `pos` of a big-endian word stream.
template <bool Shared>
__device__ __forceinline__ void or_bits(uint32_t* words, unsigned long long pos, uint32_t v, uint32_t len) {
    const uint32_t sh = (uint32_t)(pos & 31);
    const unsigned long long x = (unsigned long long)v << (64u - len - sh);  // len + sh <= 63
    const uint32_t hi = (uint32_t)(x >> 32), lo = (uint32_t)x;
    uint32_t* w = words + (pos >> 5);
    if (Shared) {
        if (hi) atomicOr(w, hi);
        if (lo) atomicOr(w + 1, lo);
    } else {
        if (hi) atomicOr(w, bswap32(hi));
        if (lo) atomicOr(w + 1, bswap32(lo));
    }
}

#ifndef K3_LUT_COPIES
#define K3_LUT_COPIES 1   // copies of the fast path's code LUT in shared memory (2: odd lanes read the second one)
#endif
constexpr int K3_RUN = 8;                    // consecutive tokens per lane and step
constexpr int K3_STEP = 32 * K3_RUN;         // tokens per warp step
constexpr int K3_HALF_STEP = K3_STEP / 2;    // the fast path's tail step: 4 tokens per lane

// tokens [base, base + 8) into registers; tokens at or beyond `end` read as 0
__device__ __forceinline__ void load_run(const uint32_t* __restrict__ tok, uint32_t base, uint32_t end,
                                         uint32_t (&t)[K3_RUN]) {
    if (base + K3_RUN <= end) {
        const uint4* p = reinterpret_cast<const uint4*>(tok + base);  // chunk bases and `base` are multiples of 8
        const uint4 v0 = __ldg(p), v1 = __ldg(p + 1);
        t[0] = v0.x, t[1] = v0.y, t[2] = v0.z, t[3] = v0.w, t[4] = v1.x, t[5] = v1.y, t[6] = v1.z, t[7] = v1.w;
    } else {
#pragma unroll
        for (int i = 0; i < K3_RUN; i++) t[i] = (base + i < end) ? __ldg(tok + base + i) : 0u;
    }
}

// A warp walks its token range 256 tokens per step, every lane owning a RUN of 8 consecutive tokens
// (2 x 128-bit loads): the bit lengths are scanned inside the warp and each lane then appends its
// run into a 64-bit accumulator, ORing finished 32-bit words into the bit buffer -- only the first
// and last word of a run are shared with the neighbouring lanes, so the atomics rarely collide.
//   Shared: `words` is shared memory (host byte order), else the global big-endian stream.
//   cap_bits: emission stops (counting continues) once the range would run past it -> *overflow.
// Returns the bits of the whole range.
// appends `len` (<= 31) bits to a lane's 64-bit accumulator; finished 32-bit words are OR-ed out
template <bool Shared>
struct LaneSink {
    uint32_t* w;
    unsigned long long acc;  // left-aligned pending bits
    int fill;                // valid bits in acc (< 32 between appends)
    __device__ __forceinline__ void init(uint32_t* words, unsigned long long pos) {
        w = words + (pos >> 5);
        acc = 0ull;
        fill = (int)(pos & 31);
    }
    __device__ __forceinline__ void put(uint32_t code, int len) {
        acc |= (unsigned long long)code << (64 - fill - len);
        fill += len;
        if (fill >= 32) {
            const uint32_t v = (uint32_t)(acc >> 32);
            if (v) atomicOr(w, Shared ? v : bswap32(v));
            ++w;
            acc <<= 32;
            fill -= 32;
        }
    }
    __device__ __forceinline__ void flush() {
        if (fill > 0) {
            const uint32_t v = (uint32_t)(acc >> 32);
            if (v) atomicOr(w, Shared ? v : bswap32(v));
        }
    }
};

// LUT entry per (table, symbol): {code << cat, len + cat}; a symbol without a code has bit 31 of .x set (and .y = 0),
// except the pad token's entry, which is {0, 0}
constexpr uint32_t K3_PAD_TOKEN = T_YDC * 256u + 255u;  // DC symbols are categories (<= 15): never a real token

// one warp step of the general path: t[] = the lane's run (n valid tokens); returns the bits of the step
template <bool Shared>
__device__ __forceinline__ uint32_t emit_step(const uint32_t (&t)[K3_RUN], int n, const uint2* s_enc2, uint32_t zl_y,
                                              uint32_t zl_c, uint32_t* words, unsigned long long bitpos,
                                              unsigned long long cap_bits, bool& sym_ok, bool& overflow) {
    const int lane = threadIdx.x & 31;
    uint32_t val[K3_RUN], ln[K3_RUN];
    // code word + category bits of every token (encoder.rs:356-404); ZRLs are added below (rare)
    uint32_t nb = 0, nzf = 0, flags = 0;
#pragma unroll
    for (int i = 0; i < K3_RUN; i++) {
        const uint2 e = s_enc2[t[i] & 0x3FFu];
        ln[i] = i < n ? e.y : 0u;
        val[i] = (e.x & 0x7FFFFFFFu) | (t[i] >> 16);
        if (i < n) flags |= e.x, nzf |= t[i];
        nb += ln[i];
    }
    if (flags >> 31) sym_ok = false;  // Error::HuffmanSymbolNotPresentInTranslator
    nzf &= 0xC00u;
    if (nzf) {
#pragma unroll
        for (int i = 0; i < K3_RUN; i++)
            if (i < n) nb += ((t[i] >> 10) & 3u) * ((t[i] & 0x200u) ? zl_c : zl_y);
    }
    uint32_t inc = nb;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t u = __shfl_up_sync(0xffffffffu, inc, d);
        if (lane >= d) inc += u;
    }
    const uint32_t step_bits = __shfl_sync(0xffffffffu, inc, 31);
    if (bitpos + step_bits > cap_bits) overflow = true;  // warp-uniform
    if (nb && !overflow) {
        LaneSink<Shared> sink;
        sink.init(words, bitpos + (inc - nb));
        if (!nzf) {
#pragma unroll
            for (int i = 0; i < K3_RUN; i++)
                if (ln[i]) sink.put(val[i], (int)ln[i]);
        } else {
#pragma unroll
            for (int i = 0; i < K3_RUN; i++) {
                if (ln[i]) {
                    uint32_t nz = (t[i] >> 10) & 3u;
                    if (nz) {  // ZRL codes first (categorize.rs:139-142); symbol 0xF0 has category 0
                        const uint2 z = s_enc2[(t[i] & 0x300u) | tok_swz(t[i] >> 8 & 3u, 0xF0u)];
                        for (; nz; --nz) sink.put(z.x & 0xFFFFu, (int)z.y);
                    }
                    sink.put(val[i], (int)ln[i]);
                }
            }
        }
        sink.flush();
    }
    return step_bits;
}

template <bool Shared>
__device__ __forceinline__ unsigned long long emit_range(const uint32_t* __restrict__ tok, uint32_t begin,
                                                         uint32_t end, const uint2* s_enc2, uint32_t zl_y,
                                                         uint32_t zl_c, uint32_t* words, unsigned long long bitpos,
                                                         unsigned long long cap_bits, bool& sym_ok, bool& overflow) {
    const int lane = threadIdx.x & 31;
    const unsigned long long start = bitpos;
    for (uint32_t wbase = begin; wbase < end; wbase += K3_STEP) {
        const uint32_t base = wbase + lane * K3_RUN;
        const int n = base < end ? (int)min((uint32_t)K3_RUN, end - base) : 0;  // valid tokens of this lane's run
        uint32_t t[K3_RUN];
        load_run(tok, base, end, t);
        bitpos += emit_step<Shared>(t, n, s_enc2, zl_y, zl_c, words, bitpos, cap_bits, sym_ok, overflow);
    }
    return bitpos - start;
}

// K3's first pass (private shared-memory buffer).  The common step -- no ZRL prefix in the warp's 256 tokens and
// every PAIR of neighbouring tokens at most 32 bits long -- needs no per-token bookkeeping at all: s_enc2 holds
// {code << cat, len + cat} per (table, symbol), so a token is one 64-bit LUT read and one OR; a lane merges its
// 8 tokens into 4 pairs (32-bit) and 2 quads (64-bit) with plain shifts and ORs each quad into the bit buffer
// as up to three words.  Tokens beyond the range read as K3_PAD_TOKEN, whose LUT entry is {0, 0}; a symbol
// without a code has bit 31 set in its entry (the output of such an image is discarded, encoder.rs:381-386).
// Any other step (ZRLs, very long codes) goes through emit_step.
// PTX shifts clamp the amount at 32 (a C shift by 32 is undefined); the pair / quad merges below rely on that
__device__ __forceinline__ uint32_t shl32(uint32_t v, uint32_t n) {
    uint32_t r;
    asm("shl.b32 %0, %1, %2;" : "=r"(r) : "r"(v), "r"(n));
    return r;
}
__device__ __forceinline__ uint32_t shr32(uint32_t v, uint32_t n) {
    uint32_t r;
    asm("shr.u32 %0, %1, %2;" : "=r"(r) : "r"(v), "r"(n));
    return r;
}
// OR into a shared-memory word (unconditionally: ptxas turns a predicated reduction into a branch around it)
__device__ __forceinline__ void red_or_shared(uint32_t saddr, uint32_t v) {
#ifdef K3_COND_RED   // A/B: skip zero values (compare + branch around the reduction): 2.5 % slower
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %1, 0;\n\t@p red.shared.or.b32 [%0], %1;\n\t}" ::"r"(saddr), "r"(v) : "memory");
#else
    asm volatile("red.shared.or.b32 [%0], %1;" ::"r"(saddr), "r"(v) : "memory");
#endif
}
// ORs the `len` (<= 64) bits held right-aligned in (hi:lo) at bit position `pos` of the warp's private bit buffer
// (shared address `wsa`, host byte order): up to three words
__device__ __forceinline__ void or_quad(uint32_t wsa, uint32_t pos, uint32_t hi, uint32_t lo, uint32_t len) {
    const unsigned long long left = (((unsigned long long)hi << 32) | lo) << ((64u - len) & 63u);  // len = 0 comes with 0
    const uint32_t lhi = (uint32_t)(left >> 32), llo = (uint32_t)left;
    const uint32_t sft = pos & 31u;
    const uint32_t wa = wsa + ((pos >> 5) << 2);
#ifdef K3_EXPERIMENT_NO_ATOMICS   // measurement only (wrong output): what the shared-memory reductions cost
    if ((lhi >> sft) + __funnelshift_r(llo, lhi, sft) + __funnelshift_r(0u, llo, sft) == 0xFFFFFFFFu) red_or_shared(wa, 1u);
#else
    red_or_shared(wa, lhi >> sft);
    red_or_shared(wa + 4, __funnelshift_r(llo, lhi, sft));
    red_or_shared(wa + 8, __funnelshift_r(0u, llo, sft));
#endif
}
// `RUN` (8 or 4) tokens of every lane = one warp step of the fast path; tokens beyond the range are pad tokens
template <int RUN>
__device__ __forceinline__ void load_run_fast(const uint32_t* __restrict__ tok, uint32_t base, uint32_t end, uint32_t (&t)[K3_RUN]) {
    if (base + RUN <= end) {
        const uint4* p = reinterpret_cast<const uint4*>(tok + base);  // region bases and `base` are multiples of RUN
        const uint4 v0 = __ldg(p);
        t[0] = v0.x, t[1] = v0.y, t[2] = v0.z, t[3] = v0.w;
        if constexpr (RUN == 8) {
            const uint4 v1 = __ldg(p + 1);
            t[4] = v1.x, t[5] = v1.y, t[6] = v1.z, t[7] = v1.w;
        }
    } else {
#pragma unroll
        for (int i = 0; i < RUN; i++) t[i] = (base + i < end) ? __ldg(tok + base + i) : K3_PAD_TOKEN;
    }
}
template <int RUN>
__device__ __forceinline__ uint32_t emit_step_fast(const uint32_t (&t)[K3_RUN], uint32_t base, uint32_t end, const uint2* s_enc2,
                                                   uint32_t zl_y, uint32_t zl_c, uint32_t* words, uint32_t wsa, uint32_t bitpos,
                                                   uint32_t cap, bool& sym_ok, bool& overflow) {
    const int lane = threadIdx.x & 31;
    uint32_t val[RUN], ln[RUN];
    uint32_t tor = 0u, flags = 0u;
    // (K3_LUT_COPIES = 2: odd lanes read a second copy of the LUT, which halves the lanes that can collide on a bank pair)
    const uint2* lut = s_enc2 + (K3_LUT_COPIES > 1 ? (lane & 1) * 2048 : 0);
    // ONE 64-bit LUT read per token, indexed by table | symbol | "one ZRL in front" (token bit 10): the second half of
    // the LUT holds every AC symbol with its table's ZRL code (categorize.rs:139-142) already in front of it, so a
    // run of 16..31 zeros costs nothing here.  Entries that do not fit 31 bits read as length 64 and fail the pair
    // test below; two or three ZRLs (token bit 11, runs of 32.., rare) take the general path as well.
#pragma unroll
    for (int i = 0; i < RUN; i++) {
        const uint2 e = lut[t[i] & 0x7FFu];
        val[i] = e.x | (t[i] >> 16);
        ln[i] = e.y;
        tor |= t[i];
        // a symbol without a code has bit 31 of its entry set (and length 0); category bits only reach bit 15
        flags |= val[i];
    }
    const bool hard = (tor & 0x800u) != 0u;
    uint32_t lp[RUN / 2], lmax = 0u, nb = 0u;
#pragma unroll
    for (int i = 0; i < RUN / 2; i++) {
        lp[i] = ln[2 * i] + ln[2 * i + 1];
        lmax = max(lmax, lp[i]);
        nb += lp[i];
    }
    const bool simple = !hard && lmax <= 32u;
    if (__all_sync(0xffffffffu, simple)) {
        if (flags >> 31) sym_ok = false;  // Error::HuffmanSymbolNotPresentInTranslator
        uint32_t inc = nb;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t u = __shfl_up_sync(0xffffffffu, inc, d);
            if (lane >= d) inc += u;
        }
        const uint32_t step_bits = __shfl_sync(0xffffffffu, inc, 31);
        if (bitpos + step_bits > cap) overflow = true;  // warp-uniform
        if (!overflow) {
            // (a "no code" marker in bit 31 of a value ends up above the quad's valid bits and is shifted out by
            // or_quad's left alignment; the output of such an image is discarded anyway, encoder.rs:381-386)
            uint32_t at = bitpos + (inc - nb);
#pragma unroll
            for (int qd = 0; qd < RUN / 4; qd++) {
                const uint32_t pa = shl32(val[4 * qd], ln[4 * qd + 1]) | val[4 * qd + 1];
                const uint32_t pb = shl32(val[4 * qd + 2], ln[4 * qd + 3]) | val[4 * qd + 3];
                const uint32_t lb = lp[2 * qd + 1], lq = lp[2 * qd] + lb;
                or_quad(wsa, at, shr32(pa, 32u - lb), shl32(pa, lb) | pb, lq);
                at += lq;
            }
        }
        return step_bits;
    }
    uint32_t t8[K3_RUN];
#pragma unroll
    for (int i = 0; i < K3_RUN; i++) t8[i] = i < RUN ? t[i] : K3_PAD_TOKEN;
    const int n = base < end ? (int)min((uint32_t)RUN, end - base) : 0;
    return emit_step<true>(t8, n, s_enc2, zl_y, zl_c, words, bitpos, cap, sym_ok, overflow);
}
// The last 128 or fewer tokens of a range go as a HALF step (4 tokens per lane, one quad): a tile of the `photo` batch
// holds 343 tokens on average, which two full steps would pad to 512.
// first step of a range: its tokens requested ahead of time (`full`: a full step follows, else the half step)
__device__ __forceinline__ void emit_first_run(const uint32_t* __restrict__ tok, uint32_t begin, uint32_t end, uint32_t (&tn)[K3_RUN],
                                               bool& full) {
    const int lane = threadIdx.x & 31;
#pragma unroll
    for (int i = 0; i < K3_RUN; i++) tn[i] = K3_PAD_TOKEN;
    full = begin < end && end - begin > K3_HALF_STEP;  // warp-uniform
    if (begin < end) {
        if (full) load_run_fast<8>(tok, begin + lane * 8, end, tn);
        else load_run_fast<4>(tok, begin + lane * 4, end, tn);
    }
}
// `tn` / `full` come from emit_first_run(tok, begin, end, ..) -- called by the caller as early as it knows the range
__device__ __forceinline__ unsigned long long emit_range_fast(const uint32_t* __restrict__ tok, uint32_t begin,
                                                              uint32_t end, const uint2* s_enc2, uint32_t zl_y, uint32_t zl_c, uint32_t* words,
                                                              unsigned long long cap_bits, bool& sym_ok, bool& overflow,
                                                              uint32_t (&tn)[K3_RUN], bool full) {
    const int lane = threadIdx.x & 31;
    const uint32_t wsa = (uint32_t)__cvta_generic_to_shared(words);
    const uint32_t cap = (uint32_t)cap_bits;
    uint32_t bitpos = 0u;  // the private buffer holds < 2^32 bits
    // the tokens of step k + 1 are requested before step k is processed (the chain of bit positions makes the
    // steps sequential, so the load latency would otherwise be exposed once per step)
    uint32_t wbase = begin;
    while (wbase < end) {
        uint32_t t[K3_RUN];
#pragma unroll
        for (int i = 0; i < K3_RUN; i++) t[i] = tn[i];
        const bool cur_full = full;
        const uint32_t nxt = wbase + (cur_full ? (uint32_t)K3_STEP : (uint32_t)K3_HALF_STEP);
        if (nxt < end) {
            full = end - nxt > K3_HALF_STEP;
            if (full) load_run_fast<8>(tok, nxt + lane * 8, end, tn);
            else load_run_fast<4>(tok, nxt + lane * 4, end, tn);
        }
        if (cur_full) bitpos += emit_step_fast<8>(t, wbase + lane * 8, end, s_enc2, zl_y, zl_c, words, wsa, bitpos, cap, sym_ok, overflow);
        else bitpos += emit_step_fast<4>(t, wbase + lane * 4, end, s_enc2, zl_y, zl_c, words, wsa, bitpos, cap, sym_ok, overflow);
        wbase = nxt;
    }
    return bitpos;
}

// Installs LUT entries of raw code-table entry `e` (len << 16 | code) of (table, symbol) = i: the token's own
// {code << cat, len + cat}, and at + 0x400 the same token behind ONE ZRL of its table (raw entry `ez`)
__device__ __forceinline__ void k3_install(uint2* s_enc2, uint32_t i, uint32_t e, uint32_t ez) {
    const uint32_t len = e >> 16, code = e & 0xFFFFu, cat = i & 15u, tbl = i >> 8;
    const uint32_t at = (i & 0x300u) | tok_swz(tbl, i & 255u);
    const uint2 own = len ? make_uint2(code << cat, len + cat) : make_uint2(i == K3_PAD_TOKEN ? 0u : 0x80000000u, 0u);
    const uint32_t zlen = ez >> 16, tot = len + cat + zlen;
    const bool fast = (tbl == T_YAC || tbl == T_CAC) && len && zlen && tot <= 31u;
    const uint2 zrl = fast ? make_uint2(((ez & 0xFFFFu) << (len + cat)) | (code << cat), tot) : make_uint2(0u, 64u);
#pragma unroll
    for (int c = 0; c < K3_LUT_COPIES; c++) s_enc2[c * 2048 + at] = own, s_enc2[c * 2048 + (0x400u | at)] = zrl;
}

constexpr size_t K3_LUT_BYTES = K3_LUT_COPIES * 2 * 4 * 256 * sizeof(uint2);  // dynamic shared memory of both K3 kernels
constexpr int K3_WBUF_WORDS = 576;  // per-warp private bit buffer: 18432 bits; two of them per warp (pipelined chunks)

__device__ __forceinline__ uint32_t k3_load_ntok(const K3Args& a, int img, uint32_t chunk, int wid) {
    if (a.n_segs) {
        const uint32_t seg = chunk * (EB / 32) + wid;
        return seg < a.n_segs ? a.tb.ntok[(size_t)img * a.n_segs + seg] : 0u;
    }
    return a.tb.ntok[(size_t)img * a.n_chunks + chunk];
}

// K3.  Work items = the chunks of ALL images, image-major, taken by one ticket in START order by a grid sized
// to what the device holds at once.  A CTA works on TWO chunks at a time: it packs chunk i into one set of
// per-warp buffers and publishes its bit count, and only then resolves the look-back of its previous chunk
// and copies that one out -- by then the predecessors of the previous chunk have long published, so nobody
// waits for a neighbour that is still packing.  The next ticket, the next image's code table and the next
// token counts are requested while the current chunk is being packed.
__global__ void __launch_bounds__(EB, 4) k3_pack(const K3Args a) {
    __shared__ __align__(16) uint32_t s_wbuf[2][EB / 32][K3_WBUF_WORDS];
    extern __shared__ __align__(16) uint2 s_enc2[];   // K3_LUT_BYTES: [one ZRL in front][table][symbol, swizzled]
    __shared__ uint32_t s_wsum[EB / 32];
    __shared__ unsigned long long s_prefix;
    __shared__ unsigned int s_next;
    __shared__ int s_err_next;
    __shared__ unsigned int s_ovf_tag;  // = 1 + item in which some warp's range did not fit its buffer

    pdl_wait();
    pdl_trigger();
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    if (tid == 0) {
        s_next = atomicAdd(a.ticket, 1u);
        s_ovf_tag = 0u;
    }
    __syncthreads();
    uint32_t item = s_next;
    if (item >= a.n_items) return;
    int cur_img = (int)(item / a.n_chunks);
    // images flagged by K1/K2/K2b (range / capacity) are skipped; K3 itself only ever adds DMMT_E_SYMBOL
    if (tid == 0) s_err_next = *reinterpret_cast<volatile int32_t*>(&a.meta[cur_img].error);
    for (int i = tid; i < 1024; i += EB) {
        k3_install(s_enc2, (uint32_t)i, a.enc[cur_img].e[i >> 8][i & 255], a.enc[cur_img].e[i >> 8][0xF0]);
    }
    uint32_t ntok_cur = k3_load_ntok(a, cur_img, item % a.n_chunks, wid);
    __syncthreads();
    int err = s_err_next;
    uint32_t zl_y = s_enc2[T_YAC * 256 + tok_swz(T_YAC, 0xF0u)].y, zl_c = s_enc2[T_CAC * 256 + tok_swz(T_CAC, 0xF0u)].y;
    const unsigned long long seed = a.seed_src ? (*a.seed_src & 7ull) : a.seed_bits;

    bool pend = false;  // the previous chunk is packed in s_wbuf[buf ^ 1] and waits for its position
    int p_img = 0, buf = 0;
    uint32_t p_chunk = 0, p_wbase = 0, p_wbits = 0, p_cbits = 0;

    while (true) {
        const bool have = item < a.n_items;  // false only in the last round, which drains the pending chunk
        const int img = cur_img;
        const uint32_t chunk = have ? item % a.n_chunks : 0u;
        const bool active = have && (err == 0 || err == DMMT_E_SYMBOL);
        unsigned int nxt = 0xFFFFFFFFu;
        if (tid == 0 && have) nxt = atomicAdd(a.ticket, 1u);  // in flight while this chunk is packed

        // token range of this warp: generic mode = a slice (multiple of 256 tokens, so the 128-bit loads
        // stay aligned) of the chunk's region; tile mode (fused K1) = the whole region of tile 8 * chunk + warp
        const uint32_t* __restrict__ tok = a.tb.tok + (size_t)img * a.tb.img_stride_words;
        uint32_t begin = 0, end = 0;
        if (active) {
            if (a.n_segs) {
                tok += (size_t)(chunk * (EB / 32) + wid) * a.tb.chunk_cap;
                end = ntok_cur;
            } else {
                tok += (size_t)chunk * a.tb.chunk_cap;
                const uint32_t per_warp = ((ntok_cur + EB / 32 * K3_STEP - 1) / (EB / 32 * K3_STEP)) * K3_STEP;
                begin = min(ntok_cur, wid * per_warp), end = min(ntok_cur, begin + per_warp);
            }
        }
        // ONE pass over the tokens: the warp packs its range into its private buffer from bit 0
        uint32_t* wbuf = s_wbuf[buf][wid];
        uint32_t wbits = 0;
        if (active) {
            for (int i = lane; i < K3_WBUF_WORDS / 4; i += 32) reinterpret_cast<uint4*>(wbuf)[i] = make_uint4(0, 0, 0, 0);
            __syncwarp();
            bool sym_ok = true, ovf = false;
            uint32_t tn[K3_RUN];
            bool full;
            emit_first_run(tok, begin, end, tn, full);
            wbits = (uint32_t)emit_range_fast(tok, begin, end, s_enc2, zl_y, zl_c, wbuf,
                                              (unsigned long long)K3_WBUF_WORDS * 32 - 64, sym_ok, ovf, tn, full);
            if (!sym_ok) atomicCAS(&a.meta[img].error, 0, DMMT_E_SYMBOL);
            if (ovf && lane == 0) s_ovf_tag = item + 1u;
        }
        if (lane == 0) s_wsum[wid] = wbits;
        if (tid == 0) s_next = nxt;
        __syncthreads();  // (a) bit counts, overflow tag and the next ticket are visible
        uint32_t wbase = 0, cbits = 0;
#pragma unroll
        for (int w = 0; w < EB / 32; w++) {
            const uint32_t v = s_wsum[w];
            if (w < wid) wbase += v;
            cbits += v;
        }
        const bool cur_ovf = active && s_ovf_tag == item + 1u;
        const uint32_t nitem = s_next;
        const bool nhave = nitem < a.n_items;
        const int nimg = nhave ? (int)(nitem / a.n_chunks) : cur_img;
        if (wid == 0) {
            if (active && lane == 0) lookback_publish_aggregate(a.lb_state + (size_t)img * a.n_chunks, (int)chunk, cbits);
            if (pend) {
                const unsigned long long ex =
                    lookback_resolve_warp(a.lb_state + (size_t)p_img * a.n_chunks, (int)p_chunk, p_cbits);
                if (lane == 0) s_prefix = ex;
            }
        }
        // requests for the next chunk: code table and error flag if the image changes, token count
        uint32_t e_next[1024 / EB], z_next[1024 / EB];
        if (nimg != cur_img) {
#pragma unroll
            for (int k = 0; k < 1024 / EB; k++) {
                e_next[k] = a.enc[nimg].e[(tid + k * EB) >> 8][(tid + k * EB) & 255];
                z_next[k] = a.enc[nimg].e[(tid + k * EB) >> 8][0xF0];
            }
            if (tid == 0) s_err_next = *reinterpret_cast<volatile int32_t*>(&a.meta[nimg].error);
        }
        const uint32_t ntok_next = nhave ? k3_load_ntok(a, nimg, nitem % a.n_chunks, wid) : 0u;
        __syncthreads();  // (b) the pending chunk's prefix is known
        if (pend) {
            // shifted copy of the private buffer to its place: destination word k holds relative bits
            // [32k - s, 32k - s + 32); the first and last word are shared with the neighbours
            uint32_t* gscan = a.scan + (size_t)p_img * a.scan_img_stride_words;
            const uint32_t* src = s_wbuf[buf ^ 1][wid];
            const unsigned long long g0 = seed + s_prefix, p0 = g0 + p_wbase;
            const uint32_t sft = (uint32_t)(p0 & 31);
            const uint32_t n_dst = (sft + p_wbits + 31) >> 5;
            uint32_t* dst = gscan + (p0 >> 5);
            for (uint32_t k = lane; k < n_dst; k += 32) {
                const uint32_t cur = k < (uint32_t)K3_WBUF_WORDS ? src[k] : 0u;
                const uint32_t prv = k ? src[k - 1] : 0u;
                const uint32_t v = bswap32(__funnelshift_r(cur, prv, sft));
                if (k == 0 || k == n_dst - 1) {
                    if (v) atomicOr(dst + k, v);
                } else {
                    dst[k] = v;
                }
            }
            if (p_chunk == a.n_chunks - 1 && a.pad_ones && tid == 0) {
                const uint32_t pad = (uint32_t)((8 - ((g0 + p_cbits) & 7)) & 7);  // binary_stream.rs:89-96
                if (pad) or_bits<false>(gscan, g0 + p_cbits, (1u << pad) - 1u, pad);
            }
        }
        if (cur_ovf) {
            // a range too dense for the private buffer (rare): resolve this chunk right away and make a second
            // pass straight into the zeroed stream while its code table is still loaded
            __syncthreads();
            if (wid == 0) {
                const unsigned long long ex = lookback_resolve_warp(a.lb_state + (size_t)img * a.n_chunks, (int)chunk, cbits);
                if (lane == 0) s_prefix = ex;
            }
            __syncthreads();
            uint32_t* gscan = a.scan + (size_t)img * a.scan_img_stride_words;
            const unsigned long long g0 = seed + s_prefix;
            bool d0 = true, d1 = false;
            (void)emit_range<false>(tok, begin, end, s_enc2, zl_y, zl_c, gscan, g0 + wbase, ~0ull, d0, d1);
            if (chunk == a.n_chunks - 1 && a.pad_ones && tid == 0) {
                const uint32_t pad = (uint32_t)((8 - ((g0 + cbits) & 7)) & 7);
                if (pad) or_bits<false>(gscan, g0 + cbits, (1u << pad) - 1u, pad);
            }
            __syncthreads();  // the table may be replaced below
            pend = false;
        } else {
            pend = active;
            p_img = img, p_chunk = chunk, p_wbase = wbase, p_wbits = wbits, p_cbits = cbits;
        }
        if (!nhave && !pend) return;
        if (nimg != cur_img) {  // every warp is past its emission: install the next image's table
#pragma unroll
            for (int k = 0; k < 1024 / EB; k++) {
                k3_install(s_enc2, (uint32_t)tid + k * EB, e_next[k], z_next[k]);
            }
        }
        __syncthreads();  // (c) table, error flag ready; s_wsum / s_next / s_prefix may be rewritten
        if (nimg != cur_img) {
            cur_img = nimg;
            err = s_err_next;
            zl_y = s_enc2[T_YAC * 256 + tok_swz(T_YAC, 0xF0u)].y, zl_c = s_enc2[T_CAC * 256 + tok_swz(T_CAC, 0xF0u)].y;
        }
        item = nitem;
        ntok_cur = ntok_next;
        buf ^= 1;
    }
}

// K3 in TILE mode (fused K1: one token region per 256x16-pixel tile).  The unit of work, of the ticket and of the
// decoupled look-back is ONE TILE taken by ONE WARP: a warp takes the next tile of the CTA's current image, packs it
// into a private shared-memory buffer from bit 0, publishes its bit count and only then resolves the look-back of its
// PREVIOUS tile and copies that one out -- so no warp ever waits for another warp of its CTA (the chunk-of-8-tiles
// version spent 29 % of its stall samples at CTA barriers, waiting for the slowest of eight tiles), and a single
// 4K frame (338 tiles) spreads over 338 warps at once.  A CTA keeps one image's code table in shared memory; CTA c
// starts at image c * n / grid and moves on to the next images when its own has no tiles left (barriers only there).
#ifndef K3_MINB
#define K3_MINB 3
#endif
__global__ void __launch_bounds__(EB, K3_MINB) k3_pack_tiles(const K3Args a) {
    __shared__ __align__(16) uint32_t s_wbuf[2][EB / 32][K3_WBUF_WORDS];
    extern __shared__ __align__(16) uint2 s_enc2[];   // K3_LUT_BYTES: [one ZRL in front][table][symbol, swizzled]
    __shared__ int s_err_img;
    __shared__ unsigned int s_left;

    pdl_wait();
    pdl_trigger();
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int n = (int)(a.n_items / a.n_chunks);  // images of this launch (n_chunks = tiles per image here)
    const uint32_t tiles = a.n_segs;
    const int first_img = (int)((unsigned long long)blockIdx.x * (unsigned)n / gridDim.x);
    const unsigned long long seed = a.seed_src ? (*a.seed_src & 7ull) : a.seed_bits;
    const int visits = n < 3 ? n : 3;  // own image, then help the next two
    const bool ahead = a.n_items >= 2u * gridDim.x * (EB / 32);
    for (int v = 0; v < visits; v++) {
        const int img = (first_img + v) % n;
        if (tid == 0) {
            s_left = *reinterpret_cast<volatile unsigned int*>(&a.ticket[img]) < tiles ? 1u : 0u;
            s_err_img = *reinterpret_cast<volatile int32_t*>(&a.meta[img].error);
        }
        __syncthreads();  // also: every warp is done with the previous image's table
        const int err = s_err_img;
        const bool img_active = s_left != 0u && (err == 0 || err == DMMT_E_SYMBOL);
        if (!img_active) continue;  // nothing left (or an image flagged by K1 / K2b: skipped, K3 only ever adds DMMT_E_SYMBOL)
        for (int i = tid; i < 1024; i += EB) {
            k3_install(s_enc2, (uint32_t)i, a.enc[img].e[i >> 8][i & 255], a.enc[img].e[i >> 8][0xF0]);
        }
        __syncthreads();
        const uint32_t zl_y = s_enc2[T_YAC * 256 + tok_swz(T_YAC, 0xF0u)].y, zl_c = s_enc2[T_CAC * 256 + tok_swz(T_CAC, 0xF0u)].y;
        unsigned long long* lb = a.lb_state + (size_t)img * a.n_chunks;
        uint32_t* gscan = a.scan + (size_t)img * a.scan_img_stride_words;
        const uint32_t* __restrict__ tok_img = a.tb.tok + (size_t)img * a.tb.img_stride_words;

        bool pend = false;  // the previous tile of this warp is packed in s_wbuf[buf ^ 1][wid] and waits for its position
        int buf = 0;
        uint32_t p_tile = 0, p_bits = 0;
        // the ticket and the token count of the NEXT tile are requested while the current one is packed / copied out
        // (a ticket holder only ever waits for lower tickets, all of which are held by running warps) -- unless there
        // are about as many warps as tiles (a single frame): a ticket taken ahead would then keep a tile from an idle warp
        uint32_t tile = 0;
        if (lane == 0) tile = atomicAdd(&a.ticket[img], 1u);
        tile = __shfl_sync(0xffffffffu, tile, 0);
        uint32_t ntok = tile < tiles ? a.tb.ntok[(size_t)img * tiles + tile] : 0u;
        uint32_t tn[K3_RUN];  // first tokens of `tile`: requested before the previous tile is copied out
        bool tn_full = false;
        if (tile < tiles) emit_first_run(tok_img + (size_t)tile * a.tb.chunk_cap, 0u, ntok, tn, tn_full);
        while (true) {
            const bool have = tile < tiles;
            uint32_t next_raw = 0;
            if (have && ahead && lane == 0) next_raw = atomicAdd(&a.ticket[img], 1u);
            uint32_t wbits = 0;
            bool cur_ovf = false;
            const uint32_t* __restrict__ tok = tok_img + (size_t)tile * a.tb.chunk_cap;
            if (have) {
                uint32_t* wbuf = s_wbuf[buf][wid];
                for (int i = lane; i < K3_WBUF_WORDS / 4; i += 32) reinterpret_cast<uint4*>(wbuf)[i] = make_uint4(0, 0, 0, 0);
                __syncwarp();
                bool sym_ok = true, ovf = false;
                wbits = (uint32_t)emit_range_fast(tok, 0u, ntok, s_enc2, zl_y, zl_c, wbuf,
                                                  (unsigned long long)K3_WBUF_WORDS * 32 - 64, sym_ok, ovf, tn, tn_full);
                if (!sym_ok) atomicCAS(&a.meta[img].error, 0, DMMT_E_SYMBOL);
                cur_ovf = ovf;
                __syncwarp();
                if (lane == 0) lookback_publish_aggregate(lb, (int)tile, wbits);
            }
            if (have && !ahead && lane == 0) next_raw = atomicAdd(&a.ticket[img], 1u);
            const uint32_t tile_next = have ? __shfl_sync(0xffffffffu, next_raw, 0) : tile;
            const uint32_t ntok_next = tile_next < tiles ? a.tb.ntok[(size_t)img * tiles + tile_next] : 0u;
            if (have && tile_next < tiles) emit_first_run(tok_img + (size_t)tile_next * a.tb.chunk_cap, 0u, ntok_next, tn, tn_full);
            if (pend) {
                // shifted copy of the private buffer to its place: destination word k holds relative bits
                // [32k - s, 32k - s + 32); the first and last word are shared with the neighbours
                const unsigned long long ex = lookback_resolve_warp(lb, (int)p_tile, p_bits);
                const uint32_t* src = s_wbuf[buf ^ 1][wid];
                const unsigned long long p0 = seed + ex;
                const uint32_t sft = (uint32_t)(p0 & 31);
                const uint32_t n_dst = (sft + p_bits + 31) >> 5;   // <= 575: the buffer is filled to 64 bits short of its 576 words
                uint32_t* dst = gscan + (p0 >> 5);
                for (uint32_t k = lane + 1; k + 1 < n_dst; k += 32)   // interior words: this tile's bits only
                    dst[k] = bswap32(__funnelshift_r(src[k], src[k - 1], sft));
                if (lane == 0 && n_dst) {                             // first word: shared with the previous tile
                    const uint32_t val = bswap32(src[0] >> sft);
                    if (val) atomicOr(dst, val);
                }
                if (lane == 1 && n_dst > 1) {                         // last word: shared with the next tile
                    const uint32_t val = bswap32(__funnelshift_r(src[n_dst - 1], src[n_dst - 2], sft));
                    if (val) atomicOr(dst + (n_dst - 1), val);
                }
                if (p_tile == tiles - 1 && a.pad_ones && lane == 0) {
                    const uint32_t pad = (uint32_t)((8 - ((p0 + p_bits) & 7)) & 7);  // binary_stream.rs:89-96
                    if (pad) or_bits<false>(gscan, p0 + p_bits, (1u << pad) - 1u, pad);
                }
                pend = false;
                __syncwarp();
            }
            if (!have) break;
            if (cur_ovf) {
                // a tile too dense for the private buffer (rare): resolve it right away and make a second pass straight
                // into the zeroed stream
                const unsigned long long ex = lookback_resolve_warp(lb, (int)tile, wbits);
                bool d0 = true, d1 = false;
                (void)emit_range<false>(tok, 0u, ntok, s_enc2, zl_y, zl_c, gscan, seed + ex, ~0ull, d0, d1);
                if (tile == tiles - 1 && a.pad_ones && lane == 0) {
                    const uint32_t pad = (uint32_t)((8 - ((seed + ex + wbits) & 7)) & 7);
                    if (pad) or_bits<false>(gscan, seed + ex + wbits, (1u << pad) - 1u, pad);
                }
            } else {
                pend = true;
                p_tile = tile, p_bits = wbits;
                buf ^= 1;
            }
            tile = tile_next, ntok = ntok_next;
        }
    }
}

// =========================================== K4 ===========================================
// Byte stuffing (segment_marker_injector.rs:13-30) as scan + compaction.  A CTA takes 8 KB chunks of
// the unstuffed scan by ticket (so a waiting chunk's predecessors are always running or done):
//   1. coalesced 128-bit loads into a padded shared tile, thread = 32 consecutive bytes
//   2. per-thread 0xFF count (SIMD-in-word compare), CTA scan, decoupled look-back over chunks
//   3. every thread ORs its bytes (plus the inserted zeros) into a zeroed shared output tile laid out
//      congruent (mod 16) to the destination, word by word with 64-bit shifts
//   4. 128-bit copies of the tile to the file, byte stores only at the ragged head/tail
// The owner of the last byte also writes EOI (encoder.rs:164-167) and the file length.
struct K4Args {
    const uint8_t* scan;           // [n][scan_img_stride_bytes]
    size_t scan_img_stride_bytes;
    ImgMeta* meta;                 // [n]
    unsigned long long* lb_state;  // [n][max_chunks]
    unsigned int* ticket;          // [n]
    uint32_t max_chunks;
    uint8_t* out;                  // [n][out_stride]
    size_t out_stride;
    unsigned long long* out_lens;  // [n] optional
    // shard controls (whole-image encode: first_byte = 0, use meta bits, header from meta, eoi = 1)
    unsigned long long first_byte; // first unstuffed byte this call owns
    long long n_bytes_override;    // < 0: ceil((seed_bits + scan_bits) / 8) - first_byte
    unsigned long long seed_bits;
    int prepend_header;            // output starts after the header written by K2b
    int append_eoi;
    uint8_t or_first_byte;         // previous shard's tail bits, OR-ed into the first owned byte
    // device-resident shard exchange: seed from the shard's global bit offset, owned bytes from the
    // seed + scan bits (1: whole bytes only, 2: including the padded last byte), tail byte from memory
    const unsigned long long* seed_src;
    int owned_mode;
    const int* or_first_src;
    // peer-memory gather: `out` is the whole file (possibly another device's memory) and the shard's
    // bytes start *base_src bytes into it (after the header on the first shard)
    const unsigned long long* base_src;
    // what an output slot that is too small is reported as: DMMT_E_OVERFLOW when the slot is the plan's own (its size
    // follows from the scan capacity: the caller grows the plan and retries), DMMT_E_WRITE for a caller-sized file
    int slot_err;
};

#ifndef K4_GATHER
#define K4_GATHER 1
#endif
#if !K4_GATHER   // round 1's push form, compiled only for A/B (-DK4_GATHER=0)
constexpr int K4_ROW = K4_BYTES_PER_THREAD + 16;              // padded row: conflict-free 128-bit access
constexpr int K4_OUT_WORDS = (2 * K4_CHUNK + 16 + 16) / 4;    // worst case: every byte is 0xFF

#ifndef K4_MINB
#define K4_MINB 8
#endif
__global__ void __launch_bounds__(K4_THREADS, K4_MINB) k4_stuff(const K4Args a) {
    __shared__ __align__(16) uint8_t s_in[K4_THREADS * K4_ROW];
    __shared__ __align__(16) uint32_t s_out[K4_OUT_WORDS];
    __shared__ uint32_t s_warp[K4_THREADS / 32 + 1];
    __shared__ unsigned long long s_prefix;
    __shared__ unsigned int s_chunk;
    pdl_wait();
    const int img = blockIdx.y, tid = threadIdx.x;
    ImgMeta* meta = a.meta + img;
    if (meta->error) {
        if (blockIdx.x == 0 && tid == 0 && a.out_lens) a.out_lens[img] = 0ull;
        return;
    }
    const unsigned long long seed = a.seed_src ? (*a.seed_src & 7ull) : a.seed_bits;
    const unsigned long long total_bytes =
        a.owned_mode == 1 ? (seed + meta->scan_bits) / 8
        : a.owned_mode == 2 ? (seed + meta->scan_bits + 7) / 8
        : a.n_bytes_override >= 0 ? (unsigned long long)a.n_bytes_override
                                  : (seed + meta->scan_bits + 7) / 8 - a.first_byte;
    const uint32_t or_first = a.or_first_src ? (uint32_t)(*a.or_first_src & 0xFF) : (uint32_t)a.or_first_byte;
    const uint32_t n_chunks = (uint32_t)((total_bytes + K4_CHUNK - 1) / K4_CHUNK);
    const unsigned long long hdr = (a.prepend_header ? meta->header_len : 0u) + (a.base_src ? *a.base_src : 0ull);
    uint8_t* out = a.out + (size_t)img * a.out_stride;
    if (n_chunks == 0) {  // nothing owned (possible for a shard); still terminate the file
        if (blockIdx.x == 0 && tid == 0) {
            unsigned long long len = hdr;
            if (a.append_eoi) out[len] = 0xFF, out[len + 1] = 0xD9, len += 2;
            meta->out_len = len;
            if (a.out_lens) a.out_lens[img] = len;
        }
        return;
    }
    const uint8_t* __restrict__ src = a.scan + (size_t)img * a.scan_img_stride_bytes + a.first_byte;
    const bool src_aligned = (reinterpret_cast<uintptr_t>(src) & 15) == 0;

    while (true) {
        __syncthreads();  // previous iteration's tiles are free
        if (tid == 0) s_chunk = atomicAdd(&a.ticket[img], 1u);
        __syncthreads();
        const uint32_t chunk = s_chunk;
        if (chunk >= n_chunks) return;
        const unsigned long long cbase = (unsigned long long)chunk * K4_CHUNK;
        const uint32_t cvalid = (uint32_t)min((unsigned long long)K4_CHUNK, total_bytes - cbase);

        // 1. stage the chunk (reads may run past cvalid inside the padded scan buffer; masked below)
#pragma unroll
        for (int i = 0; i < K4_CHUNK / 16 / K4_THREADS; i++) {
            const int j = i * K4_THREADS + tid;  // 16-byte unit of the chunk
            uint4 v = make_uint4(0, 0, 0, 0);
            if ((uint32_t)j * 16 < cvalid) {
                const uint8_t* p = src + cbase + (size_t)j * 16;
                if (src_aligned) v = *reinterpret_cast<const uint4*>(p);
                else {
                    uint32_t w[4];
#pragma unroll
                    for (int k = 0; k < 4; k++)
                        w[k] = (uint32_t)p[4 * k] | ((uint32_t)p[4 * k + 1] << 8) | ((uint32_t)p[4 * k + 2] << 16) |
                               ((uint32_t)p[4 * k + 3] << 24);
                    v = make_uint4(w[0], w[1], w[2], w[3]);
                }
            }
            *reinterpret_cast<uint4*>(s_in + (j >> 1) * K4_ROW + (j & 1) * 16) = v;
        }
        for (int i = tid; i < K4_OUT_WORDS / 4; i += K4_THREADS) reinterpret_cast<uint4*>(s_out)[i] = make_uint4(0, 0, 0, 0);
        __syncthreads();

        // 2. thread = 32 consecutive bytes
        uint32_t w[8];
        {
            const uint4 v0 = *reinterpret_cast<const uint4*>(s_in + tid * K4_ROW);
            const uint4 v1 = *reinterpret_cast<const uint4*>(s_in + tid * K4_ROW + 16);
            w[0] = v0.x, w[1] = v0.y, w[2] = v0.z, w[3] = v0.w, w[4] = v1.x, w[5] = v1.y, w[6] = v1.z, w[7] = v1.w;
        }
        const int tbase = tid * K4_BYTES_PER_THREAD;
        const int nvalid = max(0, min(K4_BYTES_PER_THREAD, (int)cvalid - tbase));
        if (chunk == 0 && tid == 0 && or_first) w[0] |= or_first;
        uint32_t nff = 0;
#pragma unroll
        for (int i = 0; i < 8; i++) {
            const int vw = nvalid - 4 * i;  // valid bytes in this word
            if (vw < 4) w[i] &= vw <= 0 ? 0u : (0xFFFFFFFFu >> (8 * (4 - vw)));  // bytes past the end: 0 (never 0xFF)
            nff += __popc(__vcmpeq4(w[i], 0xFFFFFFFFu)) >> 3;
        }
        uint32_t chunk_ff;
        const uint32_t ff_before = block_exclusive_scan<K4_THREADS>(nff, s_warp, &chunk_ff);
        // one thread walks the predecessors: measured against the warp-wide look-back (32 predecessors per round trip)
        // this is the faster one here, 0.60 vs 0.72 ms per 1024 frames -- the predecessor is almost always done already
        if (tid == 0) s_prefix = lookback_exclusive(a.lb_state + (size_t)img * a.max_chunks, (int)chunk, chunk_ff);
        __syncthreads();
        const unsigned long long gs = hdr + cbase + s_prefix;  // file offset of the chunk's first output byte
        const uint32_t n_out = cvalid + chunk_ff;
        const bool last = chunk == n_chunks - 1;
        if (gs + n_out + 2ull > a.out_stride) {
            // Error::FailedToWriteImageData: the image's slot of the output arena is too small
            if (tid == 0) {
                atomicCAS(&meta->error, 0, a.slot_err);
                if (last) {
                    meta->out_len = 0ull;
                    if (a.out_lens) a.out_lens[img] = 0ull;
                }
            }
            continue;
        }
        const uint32_t mis = (uint32_t)((reinterpret_cast<uintptr_t>(out) + gs) & 15);  // tile congruent to the file mod 16

        // 3. OR the bytes (little-endian words) into the zeroed tile
        uint32_t pos = mis + tbase + ff_before;
#pragma unroll
        for (int i = 0; i < 8; i++) {
            const int vw = nvalid - 4 * i;
            if (vw <= 0) break;
            const uint32_t eq = __vcmpeq4(w[i], 0xFFFFFFFFu);
            unsigned long long v;
            uint32_t adv;
            if (eq == 0u) {
                v = w[i];
                adv = (uint32_t)min(vw, 4);
            } else {
                // spread the bytes: after every 0xFF a 0x00 (already zero in the tile)
                v = 0ull;
                adv = 0;
#pragma unroll
                for (int b = 0; b < 4; b++) {
                    if (b < vw) {
                        const uint32_t byte = (w[i] >> (8 * b)) & 0xFFu;
                        v |= (unsigned long long)byte << (8 * adv);
                        adv += byte == 0xFFu ? 2u : 1u;
                    }
                }
            }
            // v holds up to 8 bytes; shifted to the byte lane of `pos` it spans up to 3 words
            const uint32_t sh = 8 * (pos & 3);
            const unsigned long long vs = v << sh;
            const uint32_t w0 = (uint32_t)vs, w1 = (uint32_t)(vs >> 32);
            const uint32_t w2 = sh ? (uint32_t)(v >> (64 - sh)) : 0u;
            if (w0) atomicOr(&s_out[pos >> 2], w0);
            if (w1) atomicOr(&s_out[(pos >> 2) + 1], w1);
            if (w2) atomicOr(&s_out[(pos >> 2) + 2], w2);
            pos += adv;
        }
        __syncthreads();

        // 4. tile -> file
        {
            const uint8_t* tile = reinterpret_cast<const uint8_t*>(s_out);
            uint8_t* dst = out + gs - mis;          // 16-byte aligned
            const uint32_t beg = mis, end = mis + n_out;
            const uint32_t abeg = (beg + 15) & ~15u, aend = end & ~15u;
            if (abeg <= aend) {
                for (uint32_t o = abeg + tid * 16; o < aend; o += K4_THREADS * 16)
                    *reinterpret_cast<uint4*>(dst + o) = *reinterpret_cast<const uint4*>(tile + o);
                if (tid < 16) {
                    const uint32_t o = beg + tid;
                    if (o < abeg && o < end) dst[o] = tile[o];
                } else if (tid < 32) {
                    const uint32_t o = aend + (tid - 16);
                    if (o >= abeg && o < end) dst[o] = tile[o];
                }
            } else {  // the whole chunk output lies inside one 16-byte unit
                if (tid < 16 && beg + tid < end) dst[beg + tid] = tile[beg + tid];
            }
        }
        if (last && tid == 0) {
            unsigned long long len = gs + n_out;
            if (a.append_eoi) out[len] = 0xFF, out[len + 1] = 0xD9, len += 2;
            meta->out_len = len;
            if (a.out_lens) a.out_lens[img] = len;
        }
    }
}

#endif  // !K4_GATHER

// K4, output-centric ("gather") form -- the one that is launched.  Same chunking, ticket, look-back and file layout as
// k4_stuff above; what differs is how a chunk's bytes reach the file.  k4_stuff pushes: every thread ORs its 32 input bytes,
// spread around their 0xFF bytes, into a zeroed shared tile (8 x 3 shared-memory atomics, 8-way bank conflicts), and the
// tile is copied out afterwards.  Here every thread PULLS one 16-byte unit of the FILE at a time:
//   1. the chunk is staged linearly; the thread that loads a 16-byte unit also derives its 16 0xFF flags (bitmap)
//   2. 0xFF count per 32-byte group (one popc), CTA scan -> s_ex[g] = stuffing zeros in front of group g; look-back
//   3. unit u of the file holds chunk-output offsets [16u - mis, 16u - mis + 16): find the group g with
//      32 g + s_ex[g] <= o (guess o / 32, step down), then the source byte inside the group from the group's flags;
//      two aligned 128-bit shared loads + a realignment give the 16 source bytes, the bitmap their 0xFF flags;
//      a unit without a stuffing zero (93 % of them) is done -- otherwise one byte-insertion per zero (PRMT);
//      one aligned 128-bit store (byte stores only in the first and last unit of the chunk).
// No zeroed tile, no atomics, no second pass: 5450 instead of 7000 warp-instructions per chunk (ncu), 28 k instead of
// 4.7 M shared-memory bank conflicts per launch, 0.41 instead of 0.51 ms per 1024 frames.
#ifndef K4G_MINB
#define K4G_MINB 8
#endif
// input bytes per chunk: 128 short of the 8 KB the CTA stages, so that the output (input + stuffing zeros + misalignment)
// still fits 512 file units = two full rounds of the 256 threads (8192 bytes of input made round three a 3-unit straggler)
#ifndef K4G_CHUNK_BYTES
#define K4G_CHUNK_BYTES (K4_THREADS_N * 32 - K4_THREADS_N / 2)
#endif
constexpr int K4G_CHUNK = K4G_CHUNK_BYTES;
static_assert(K4G_CHUNK % 32 == 0 && K4G_CHUNK <= K4_CHUNK, "whole 32-byte groups, at most one per thread");
#ifndef K4G_WARP_LB
#define K4G_WARP_LB 0
#endif
// inserts a zero byte at byte position pk (0..15) of the 16 bytes in r[0..3]; the bytes from pk on move up, byte 15 drops out
__device__ __forceinline__ void k4_insert_zero(uint32_t (&r)[4], uint32_t pk) {
    uint32_t o[4];
#pragma unroll
    for (int w = 0; w < 4; w++) {
        const int d = (int)pk - 4 * w;                     // position of the zero relative to this word
        const uint32_t dc = (uint32_t)max(d, 0);
        // output byte j of the word comes from pair byte 4 + j (j < d) or 3 + j (j >= d), pair = {r[w - 1], r[w]}
        const uint32_t sel = 0x7654u - (0x1111u & shl32(0xFFFFu, 4u * dc));
        const uint32_t v = __byte_perm(w ? r[w - 1] : 0u, r[w], sel);
        o[w] = v & ~shl32(0xFFu, 8u * (uint32_t)d);        // d outside 0..3: the shift clamps to 0, nothing is cleared
    }
#pragma unroll
    for (int w = 0; w < 4; w++) r[w] = o[w];
}
__global__ void __launch_bounds__(K4_THREADS, K4G_MINB) k4_stuff_gather(const K4Args a) {
    __shared__ __align__(16) uint8_t s_in[K4_CHUNK + 32];       // the chunk, linear; the 32 bytes behind it stay zero
    __shared__ __align__(4) uint16_t s_bm16[K4_CHUNK / 16 + 4];  // 0xFF flags, 16 per staged unit (= one u32 per 32-byte group)
    __shared__ uint32_t s_ex[K4_THREADS + 1];
    __shared__ uint32_t s_warp[K4_THREADS / 32 + 1];
    __shared__ unsigned long long s_prefix;
    __shared__ unsigned int s_chunk;
    const uint32_t* s_bm = reinterpret_cast<const uint32_t*>(s_bm16);
    pdl_wait();
    const int img = blockIdx.y, tid = threadIdx.x;
    ImgMeta* meta = a.meta + img;
    if (meta->error) {
        if (blockIdx.x == 0 && tid == 0 && a.out_lens) a.out_lens[img] = 0ull;
        return;
    }
    const unsigned long long seed = a.seed_src ? (*a.seed_src & 7ull) : a.seed_bits;
    const unsigned long long total_bytes =
        a.owned_mode == 1 ? (seed + meta->scan_bits) / 8
        : a.owned_mode == 2 ? (seed + meta->scan_bits + 7) / 8
        : a.n_bytes_override >= 0 ? (unsigned long long)a.n_bytes_override
                                  : (seed + meta->scan_bits + 7) / 8 - a.first_byte;
    const uint32_t or_first = a.or_first_src ? (uint32_t)(*a.or_first_src & 0xFF) : (uint32_t)a.or_first_byte;
    const uint32_t n_chunks = (uint32_t)((total_bytes + K4G_CHUNK - 1) / K4G_CHUNK);
    const unsigned long long hdr = (a.prepend_header ? meta->header_len : 0u) + (a.base_src ? *a.base_src : 0ull);
    uint8_t* out = a.out + (size_t)img * a.out_stride;
    if (n_chunks == 0) {  // nothing owned (possible for a shard); still terminate the file
        if (blockIdx.x == 0 && tid == 0) {
            unsigned long long len = hdr;
            if (a.append_eoi) out[len] = 0xFF, out[len + 1] = 0xD9, len += 2;
            meta->out_len = len;
            if (a.out_lens) a.out_lens[img] = len;
        }
        return;
    }
    const uint8_t* __restrict__ src = a.scan + (size_t)img * a.scan_img_stride_bytes + a.first_byte;
    const bool src_aligned = (reinterpret_cast<uintptr_t>(src) & 15) == 0;
    if (tid < 8) reinterpret_cast<uint32_t*>(s_in + K4_CHUNK)[tid] = 0u;
    if (tid < 4) s_bm16[K4_CHUNK / 16 + tid] = 0;
    constexpr int UPT = K4_CHUNK / 16 / K4_THREADS;  // 16-byte units staged per thread

    // The ticket and the bytes of the NEXT chunk are requested while the current one is scanned and written: a chunk
    // otherwise pays two dependent memory round trips (ticket, scan bytes) before its first instruction of real work.
    uint4 pre[UPT];  // raw units of `chunk` (aligned scans only)
    auto request = [&](uint32_t c) {
#pragma unroll
        for (int it = 0; it < UPT; it++) {
            const int j = it * K4_THREADS + tid;
            pre[it] = make_uint4(0, 0, 0, 0);
            if (src_aligned && c < n_chunks && (unsigned long long)c * K4G_CHUNK + (unsigned)j * 16u < total_bytes)
                pre[it] = *reinterpret_cast<const uint4*>(src + (unsigned long long)c * K4G_CHUNK + (size_t)j * 16);
        }
    };
    if (tid == 0) s_chunk = atomicAdd(&a.ticket[img], 1u);
    __syncthreads();
    uint32_t chunk = s_chunk;
    request(chunk);
    while (chunk < n_chunks) {
        __syncthreads();  // everybody has read s_chunk and is done with the previous chunk's staging area
        if (tid == 0) s_chunk = atomicAdd(&a.ticket[img], 1u);
        const unsigned long long cbase = (unsigned long long)chunk * K4G_CHUNK;
        const uint32_t cvalid = (uint32_t)min((unsigned long long)K4G_CHUNK, total_bytes - cbase);

        // 1. stage the chunk and its 0xFF flags (reads may run past cvalid inside the padded scan buffer: masked)
#pragma unroll
        for (int it = 0; it < UPT; it++) {
            const int j = it * K4_THREADS + tid;  // 16-byte unit of the chunk
            uint32_t w[4] = {pre[it].x, pre[it].y, pre[it].z, pre[it].w};
            const int nv = (int)cvalid - 16 * j;  // valid bytes of this unit
            if (nv > 0) {
                if (!src_aligned) {
                    const uint8_t* p = src + cbase + (size_t)j * 16;
#pragma unroll
                    for (int k = 0; k < 4; k++)
                        w[k] = (uint32_t)p[4 * k] | ((uint32_t)p[4 * k + 1] << 8) | ((uint32_t)p[4 * k + 2] << 16) |
                               ((uint32_t)p[4 * k + 3] << 24);
                }
                if (nv < 16) {
#pragma unroll
                    for (int k = 0; k < 4; k++) {
                        const int vw = nv - 4 * k;
                        if (vw < 4) w[k] &= vw <= 0 ? 0u : (0xFFFFFFFFu >> (8 * (4 - vw)));  // bytes past the end: 0 (never 0xFF)
                    }
                }
                if (j == 0 && chunk == 0) w[0] |= or_first;
            } else {
                w[0] = w[1] = w[2] = w[3] = 0u;
            }
            *reinterpret_cast<uint4*>(s_in + j * 16) = make_uint4(w[0], w[1], w[2], w[3]);
            uint32_t flags = 0u;
#pragma unroll
            for (int k = 0; k < 4; k++)  // 0xFF per matching byte -> one bit per byte
                flags |= (((__vcmpeq4(w[k], 0xFFFFFFFFu) & 0x08040201u) * 0x01010101u) >> 24) << (4 * k);
            s_bm16[j] = (uint16_t)flags;
        }
        __syncthreads();  // the chunk, its flags and the next ticket are visible
        const uint32_t next = s_chunk;
        request(next);

        // 2. stuffing zeros in front of every 32-byte group (warp scan + the warp totals, re-added by every warp: one
        //    barrier), and in front of the chunk (look-back)
        const int lane = tid & 31, wid = tid >> 5;
        const uint32_t nff = (uint32_t)__popc(s_bm[tid]);
        uint32_t inc = nff;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t t = __shfl_up_sync(0xffffffffu, inc, d);
            if (lane >= d) inc += t;
        }
        if (lane == 31) s_warp[wid] = inc;
        __syncthreads();
        uint32_t wbase = 0u, chunk_ff = 0u;
#pragma unroll
        for (int w = 0; w < K4_THREADS / 32; w++) {
            const uint32_t v = s_warp[w];
            if (w < wid) wbase += v;
            chunk_ff += v;
        }
        s_ex[tid] = wbase + inc - nff;
#if K4G_WARP_LB
        if (wid == 0) {
            unsigned long long* lb = a.lb_state + (size_t)img * a.max_chunks;
            if (lane == 0) lookback_publish_aggregate(lb, (int)chunk, chunk_ff);
            const unsigned long long ex = lookback_resolve_warp(lb, (int)chunk, chunk_ff);
            if (lane == 0) s_prefix = ex;
        }
#else
        if (tid == 0) s_prefix = lookback_exclusive(a.lb_state + (size_t)img * a.max_chunks, (int)chunk, chunk_ff);
#endif
        __syncthreads();
        const unsigned long long gs = hdr + cbase + s_prefix;  // file offset of the chunk's first output byte
        const uint32_t n_out = cvalid + chunk_ff;
        const bool last = chunk == n_chunks - 1;
        if (gs + n_out + 2ull > a.out_stride) {
            // Error::FailedToWriteImageData: the image's slot of the output arena is too small
            if (tid == 0) {
                atomicCAS(&meta->error, 0, a.slot_err);
                if (last) {
                    meta->out_len = 0ull;
                    if (a.out_lens) a.out_lens[img] = 0ull;
                }
            }
            chunk = next;
            continue;
        }
        const uint32_t mis = (uint32_t)((reinterpret_cast<uintptr_t>(out) + gs) & 15);
        uint8_t* dst16 = out + gs - mis;  // 16-byte aligned; unit u of the chunk is dst16 + 16 u
        const uint32_t n_units = (mis + n_out + 15u) >> 4;

        // 3. one 16-byte unit of the file per thread and round
        for (uint32_t u = tid; u < n_units; u += K4_THREADS) {
            const int o_lo = (int)(16u * u) - (int)mis;            // chunk-output offset of the unit's first byte
            const uint32_t o = (uint32_t)max(o_lo, 0);
            // group: the last one that starts at or before o
            uint32_t g = min(o >> 5, (uint32_t)K4_THREADS - 1u);
            uint32_t gstart = 32u * g + s_ex[g];
            while (gstart > o) {
                --g;
                gstart = 32u * g + s_ex[g];
            }
            const uint32_t op = o - gstart;                        // offset inside the group's output
            // source byte of the group that output offset op shows, or the stuffing zero in front of it (pending)
            uint32_t mm = s_bm[g], cnt = 0u;
            bool pending = false;
            while (mm) {
                const uint32_t z = (uint32_t)__ffs((int)mm) + cnt;  // output offset of the zero behind this 0xFF
                if (z > op) break;
                if (z == op) {
                    pending = true;
                    break;
                }
                ++cnt;
                mm &= mm - 1u;
            }
            const uint32_t i = 32u * g + (op - cnt);               // next source byte of the chunk
            // 16 source bytes from i on: two aligned 128-bit loads, word select, byte shift
            uint32_t r[4];
            {
                const uint4 lo = *reinterpret_cast<const uint4*>(s_in + (i & ~15u));
                const uint4 hi = *reinterpret_cast<const uint4*>(s_in + (i & ~15u) + 16);
                const uint32_t W[8] = {lo.x, lo.y, lo.z, lo.w, hi.x, hi.y, hi.z, hi.w};
                uint32_t X[6], Y[5];
#pragma unroll
                for (int k = 0; k < 6; k++) X[k] = (i & 8u) ? W[k + 2] : W[k];
#pragma unroll
                for (int k = 0; k < 5; k++) Y[k] = (i & 4u) ? X[k + 1] : X[k];
                const uint32_t bs = 8u * (i & 3u);
#pragma unroll
                for (int k = 0; k < 4; k++) r[k] = __funnelshift_r(Y[k], Y[k + 1], bs);
            }
            // their 0xFF flags; bit k of `ins`: a zero goes in front of source byte k of the window
            const uint32_t q = __funnelshift_r(s_bm[i >> 5], s_bm[(i >> 5) + 1], i & 31u);
            uint32_t ins = ((q << 1) | (pending ? 1u : 0u)) & 0xFFFFu;
            uint32_t c = 0u;
            while (ins) {
                const uint32_t pk = (uint32_t)__ffs((int)ins) - 1u + c;
                if (pk >= 16u) break;
                ins &= ins - 1u;
                k4_insert_zero(r, pk);
                ++c;
            }
            if (o_lo >= 0 && (uint32_t)o_lo + 16u <= n_out) {
                *reinterpret_cast<uint4*>(dst16 + 16u * u) = make_uint4(r[0], r[1], r[2], r[3]);
            } else {
                // first / last unit of the chunk: only the bytes of this chunk (the neighbours write the others)
                const uint32_t cntb = min((uint32_t)(o_lo + 16), n_out) - o;
                uint8_t* d = dst16 + 16u * u + (o - (uint32_t)o_lo);
#pragma unroll
                for (int k = 0; k < 16; k++)
                    if ((uint32_t)k < cntb) d[k] = (uint8_t)(r[k >> 2] >> (8 * (k & 3)));
            }
        }
        if (last && tid == 0) {
            unsigned long long len = gs + n_out;
            if (a.append_eoi) out[len] = 0xFF, out[len + 1] = 0xD9, len += 2;
            meta->out_len = len;
            if (a.out_lens) a.out_lens[img] = len;
        }
        chunk = next;
    }
}

// extracts the quantised DC of the last Y / Cb / Cr block (shard hand-over)
__global__ void k_last_dc(const int16_t* coef, uint32_t n_blocks, int ypm, int bpm, int16_t* out3) {
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        const size_t last_mcu = (size_t)(n_blocks / bpm - 1) * bpm;
        out3[0] = coef[(last_mcu + ypm - 1) * 64];
        out3[1] = coef[(last_mcu + ypm) * 64];
        out3[2] = coef[(last_mcu + ypm + 1) * 64];
    }
}

// =========================================== K5 ===========================================
// Packs the files of the strided output arena back to back (16-byte aligned starts) so that the
// host needs ONE device-to-host copy per batch.  k5_offsets: one CTA, exclusive scan of the
// aligned lengths; k5_copy: 128-bit copies.
__global__ void __launch_bounds__(1024) k5_offsets(unsigned long long* lens, int n,
                                                   unsigned long long dense_cap,
                                                   unsigned long long* offsets, ImgMeta* meta,
                                                   int chained, int* sticky_err) {
    __shared__ unsigned long long s_w[33];
    __shared__ unsigned long long s_run;
    // chained: offsets[0] already holds the end of the previous sub-batch (or 0)
    if (threadIdx.x == 0) s_run = chained ? offsets[0] : 0ull;
    __syncthreads();
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    for (int base = 0; base < n; base += 1024) {
        const int i = base + threadIdx.x;
        unsigned long long len = 0ull;
        if (i < n) {
            const int err = meta[i].error;
            if (err == 0) len = lens[i];
            else {
                lens[i] = 0ull;
                if (sticky_err) atomicCAS(sticky_err, 0, err);  // survives the reuse of this slot's meta
            }
        }
        const unsigned long long al = (len + 15ull) & ~15ull;
        unsigned long long inc = al;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const unsigned long long t = __shfl_up_sync(0xffffffffu, inc, d);
            if (lane >= d) inc += t;
        }
        if (lane == 31) s_w[wid] = inc;
        __syncthreads();
        if (wid == 0) {
            const unsigned long long w = s_w[lane];
            unsigned long long winc = w;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const unsigned long long t = __shfl_up_sync(0xffffffffu, winc, d);
                if (lane >= d) winc += t;
            }
            s_w[lane] = winc - w;
            if (lane == 31) s_w[32] = winc;
        }
        __syncthreads();
        const unsigned long long off = s_run + s_w[wid] + inc - al;
        if (i < n) {
            if (off + al > dense_cap) {  // dense arena too small: Error::FailedToWriteImageData
                if (len) {
                    atomicCAS(&meta[i].error, 0, DMMT_E_WRITE);
                    if (sticky_err) atomicCAS(sticky_err, 0, DMMT_E_WRITE);
                }
                lens[i] = 0ull;
            }
            offsets[i] = off;
        }
        __syncthreads();
        if (threadIdx.x == 0) s_run += s_w[32];
        __syncthreads();
    }
    if (threadIdx.x == 0) offsets[n] = s_run;
}

__global__ void __launch_bounds__(256) k5_copy(const uint8_t* __restrict__ out, size_t out_stride,
                                               const unsigned long long* __restrict__ lens,
                                               const unsigned long long* __restrict__ offsets,
                                               uint8_t* __restrict__ dense) {
    const int img = blockIdx.y;
    const unsigned long long nq = (lens[img] + 15ull) >> 4;
    const uint4* src = reinterpret_cast<const uint4*>(out + (size_t)img * out_stride);
    uint4* dst = reinterpret_cast<uint4*>(dense + offsets[img]);
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < nq;
         i += (unsigned long long)gridDim.x * blockDim.x)
        dst[i] = src[i];
}

}  // namespace

// ------------------------------------------------------------------------------------------
// host launchers
uint32_t k3_chunks(const Geom& g) { return (g.n_blocks + EB - 1) / EB; }
uint32_t tok_blocks_per_chunk() { return EB; }
cudaError_t launch_k2(const Geom& g, const int16_t* coef, size_t coef_img_stride, int n, unsigned int* hist,
                      ImgMeta* meta, const int16_t* seed_dc, const TokBuf& tb, cudaStream_t st) {
    K2Args a{coef, coef_img_stride, g.n_blocks, k3_chunks(g), g.ypm, g.bpm, hist, meta, seed_dc, tb};
    dim3 grid(a.n_chunks, n);
    k2_tokenize<<<grid, EB, 0, st>>>(a);
    return cudaGetLastError();
}

// Launch with the programmatic-stream-serialisation attribute (see pdl_wait in dmmt_common.cuh); DMMT_PDL=0 in the
// environment makes these plain launches (A/B, debugging).
namespace {
bool pdl_enabled() {
    static const bool on = [] {
        const char* e = getenv("DMMT_PDL");
        return !(e && e[0] == '0');
    }();
    return on;
}
template <typename... KArgs, typename... Args>
cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid, cfg.blockDim = block, cfg.dynamicSmemBytes = smem, cfg.stream = st;
    cudaLaunchAttribute at[2];
    int na = 0;
    if (pdl_enabled()) {
        at[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        at[na].val.programmaticStreamSerializationAllowed = 1;
        ++na;
    }
    // Experiment hook, off by default (DMMT_TAIL_PRIORITY=1): the kernels behind K1 at a higher priority than K1, so that
    // with two sub-batches in flight the entropy kernels of the first take the SM slots K1 of the second frees instead
    // of queueing behind its 139 000 CTAs.  Measured: 7.20 -> 7.33 ms per 1024 frames -- sharing the SMs costs K1 more
    // than the entropy kernels gain; K1 at the higher priority instead changes nothing (7.19 ms).
    static const int prio = [] {
        const char* e = getenv("DMMT_TAIL_PRIORITY");
        if (!e || e[0] != '1') return 0;
        int least = 0, greatest = 0;
        if (cudaDeviceGetStreamPriorityRange(&least, &greatest) != cudaSuccess) return 0;
        return greatest;   // numerically lower = more urgent; 0 is the default priority
    }();
    if (prio != 0) {
        at[na].id = cudaLaunchAttributePriority;
        at[na].val.priority = prio;
        ++na;
    }
    cfg.attrs = at, cfg.numAttrs = na;
    return cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...);
}
}  // namespace

cudaError_t launch_k2b(const Geom& g, const K2bHostArgs& h, int n, cudaStream_t st) {
    K2bArgs a;
    a.hist = h.hist;
    a.ghist = h.ghist;
    a.enc = h.enc;
    a.lens = h.lens;
    a.meta = h.meta;
    a.out = h.out;
    a.out_stride = h.out_stride;
    a.scan_cap_bits = h.scan_cap_bits;
    a.W = h.W;
    a.H = h.H;
    a.hr = g.hr;
    a.vr = g.vr;
    a.bits_per_channel = h.bits_per_channel;
    for (int i = 0; i < 64; i++) {
        a.qzz[0][i] = h.qtab_luma[zz_at(i)];
        a.qzz[1][i] = h.qtab_chroma[zz_at(i)];
    }
    a.n_stream_blocks = g.n_blocks;
    a.write_header = h.write_header;
    a.lcount = h.lcount;
    a.fix_dc = h.fix ? 1 : 0;
    a.fo = h.fix ? *h.fix : TileTok{};
    return launch_pdl(k2b_tables, dim3(n, 4), dim3(K2B_THREADS), 0, st, a);
}

cudaError_t launch_zero_scan(uint32_t* scan, size_t stride_words, const ImgMeta* meta, int n,
                             unsigned long long seed_bits, int blocks_per_image, cudaStream_t st,
                             const unsigned long long* seed_src) {
    return launch_pdl(k_zero_scan, dim3(blocks_per_image, n), dim3(256), 0, st, scan, stride_words, meta, seed_bits, seed_src);
}

uint32_t k4_max_chunks(size_t scan_cap_bytes) {
    constexpr size_t chunk = K4_GATHER ? K4G_CHUNK : K4_CHUNK;
    return (uint32_t)((scan_cap_bytes + chunk - 1) / chunk);
}

// SMs of the current device (queried once per device; every context of this library sets its device first)
static int sm_count() {
    static int cached[64] = {};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 1;
    if (!cached[dev]) {
        int sms = 0;
        if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0) sms = 1;
        cached[dev] = sms;
    }
    return cached[dev];
}

cudaError_t launch_k3(uint32_t n_chunks, uint32_t n_segs, int n, const TokBuf& tb, const EncTables* enc, ImgMeta* meta,
                      unsigned long long* lb_state, unsigned int* ticket, uint32_t* scan, size_t scan_stride_words,
                      unsigned long long seed_bits, int pad_ones, cudaStream_t st, const unsigned long long* seed_src) {
    K3Args a{n_chunks, n_segs, tb, enc, meta, lb_state, ticket, scan, scan_stride_words, seed_bits, pad_ones, seed_src,
             n_chunks * (uint32_t)n};
    // persistent CTAs, exactly as many as the device holds at once (a larger grid would run a second,
    // partly empty wave)
    // 37 KB of static + 16 KB of dynamic shared memory: above the 48 KB default, so both kernels opt in (per device)
    static bool opted[64] = {};
    int dev = 0;
    if (cudaError_t e = cudaGetDevice(&dev); e != cudaSuccess) return e;
    if (dev < 0 || dev >= 64 || !opted[dev]) {
        cudaError_t e = cudaFuncSetAttribute(k3_pack, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)K3_LUT_BYTES);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(k3_pack_tiles, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)K3_LUT_BYTES);
        if (e != cudaSuccess) return e;
        if (dev >= 0 && dev < 64) opted[dev] = true;
    }
    static int resident = 0;
    if (!resident) {
        int per_sm = 0;
        cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k3_pack, EB, K3_LUT_BYTES);
        if (e != cudaSuccess) return e;
        resident = (per_sm > 0 ? per_sm : 1) * sm_count();
    }
    if (n_segs) {
        // tile mode: the look-back runs over tiles (n_chunks := n_segs); as many CTAs as the device holds, but no more
        // than one warp per tile
        a.n_chunks = n_segs;
        a.n_items = n_segs * (uint32_t)n;
        static int resident_t = 0;
        if (!resident_t) {
            int per_sm = 0;
            cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k3_pack_tiles, EB, K3_LUT_BYTES);
            if (e != cudaSuccess) return e;
            resident_t = (per_sm > 0 ? per_sm : 1) * sm_count();
        }
        const uint32_t want = (a.n_items + EB / 32 - 1) / (EB / 32);
        const uint32_t grid_t = want < (uint32_t)resident_t ? want : (uint32_t)resident_t;
        return launch_pdl(k3_pack_tiles, dim3(grid_t ? grid_t : 1u), dim3(EB), K3_LUT_BYTES, st, a);
    }
    const uint32_t grid = a.n_items < (uint32_t)resident ? a.n_items : (uint32_t)resident;
    return launch_pdl(k3_pack, dim3(grid ? grid : 1u), dim3(EB), K3_LUT_BYTES, st, a);
}

cudaError_t launch_k4(const K4HostArgs& h, int n, uint32_t grid_chunks, cudaStream_t st) {
    K4Args a{h.scan, h.scan_stride_bytes, h.meta, h.lb_state, h.ticket, h.max_chunks, h.out, h.out_stride,
             h.out_lens, h.first_byte, h.n_bytes_override, h.seed_bits, h.prepend_header, h.append_eoi,
             h.or_first_byte, h.seed_src, h.owned_mode, h.or_first_src, h.base_src,
             h.base_src ? DMMT_E_WRITE : DMMT_E_OVERFLOW};
    // CTAs take chunks by ticket, so the grid only has to keep the device busy: about 8 CTAs per SM
    // over all images, never more than the chunks an image can have
    uint32_t per_image = (uint32_t)((sm_count() * (2048 / K4_THREADS) + n - 1) / n);
    if (per_image < 8) per_image = 8;
    if (grid_chunks == 0) grid_chunks = 1;
    if (per_image > grid_chunks) per_image = grid_chunks;
#if K4_GATHER
    return launch_pdl(k4_stuff_gather, dim3(per_image, n), dim3(K4_THREADS), 0, st, a);
#else
    return launch_pdl(k4_stuff, dim3(per_image, n), dim3(K4_THREADS), 0, st, a);
#endif
}

// ---- device-resident shard exchange helpers (dmmt_shard.cu) ------------------------------------
namespace {
__global__ void k_shard_widen(const int16_t* last_dc3, int* out4, const unsigned int* hist, long long* hist64,
                              const ImgMeta* meta, long long* bits_out, long long* err_out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (last_dc3 && i < 4) out4[i] = i < 3 ? (int)last_dc3[i] : 0;
    if (hist && i < 1024) hist64[i] = (long long)hist[i];
    if (bits_out && i == 0) *bits_out = (long long)meta->scan_bits;
    if (err_out && i == 0) *err_out = (long long)meta->error;
}
__global__ void k_shard_narrow_seed(const int* seed4, int16_t* seed3) {
    if (threadIdx.x < 3) seed3[threadIdx.x] = seed4 ? (int16_t)seed4[threadIdx.x] : (int16_t)0;
}
// trailing partial byte of the shard: {byte, number of valid leading bits} (0 bits on the last shard)
__global__ void k_shard_tail(const uint8_t* scan, const ImgMeta* meta, const unsigned long long* bit_offset, int is_last,
                             int* tail2) {
    const unsigned long long end = (*bit_offset & 7ull) + meta->scan_bits;
    const int nb = is_last ? 0 : (int)(end & 7ull);
    tail2[0] = nb ? (int)scan[end / 8] : 0;
    tail2[1] = nb;
}
// tail bits this shard must OR into its first byte: the previous shard's, plus those of earlier shards
// that did not complete a byte (a shard with fewer than 8 bits hands its predecessor's bits on)
__global__ void k_shard_prev_tail(const int* all_tail2, const long long* all_offs, const long long* all_bits, int rank,
                                  int* out) {
    int acc = 0;
    for (int r = 0; r < rank; r++) {
        const bool incomplete = ((all_offs[r] & 7) + all_bits[r]) < 8;
        acc = all_tail2[2 * r] | (incomplete ? acc : 0);
    }
    *out = acc;
}
// Stuffed size of the bytes this shard owns (what K4 will write), before K4 runs: header (first shard) + owned
// bytes + one 0x00 per 0xFF + EOI (last shard).  ctr[0] = 0xFF count, ctr[1] = finished CTAs; the last CTA
// writes the result and resets both, so the counters are zero again for the next encode.
__global__ void __launch_bounds__(256) k_shard_count_bytes(const uint8_t* __restrict__ scan, const ImgMeta* meta,
                                                           const unsigned long long* seed_src, int owned_mode,
                                                           const int* or_first_src, int is_first, int is_last,
                                                           unsigned long long* ctr, long long* n_bytes) {
    __shared__ uint32_t s_w[8];
    const unsigned long long seed = *seed_src & 7ull;
    const unsigned long long total = meta->error ? 0ull
                                   : owned_mode == 1 ? (seed + meta->scan_bits) / 8 : (seed + meta->scan_bits + 7) / 8;
    const unsigned long long units = (total + 15) / 16;
    uint32_t nff = 0;
    for (unsigned long long u = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; u < units;
         u += (unsigned long long)gridDim.x * blockDim.x) {
        uint4 v = reinterpret_cast<const uint4*>(scan)[u];
        uint32_t w[4] = {v.x, v.y, v.z, v.w};
        if (u == 0) w[0] |= (uint32_t)(*or_first_src & 0xFF);
        const long long left = (long long)(total - u * 16);  // valid bytes in this unit (may exceed 16)
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const long long vw = left - 4 * i;
            if (vw < 4) w[i] &= vw <= 0 ? 0u : (0xFFFFFFFFu >> (8 * (4 - (int)vw)));
            nff += __popc(__vcmpeq4(w[i], 0xFFFFFFFFu)) >> 3;
        }
    }
    nff = __reduce_add_sync(0xFFFFFFFFu, nff);
    if ((threadIdx.x & 31) == 0) s_w[threadIdx.x >> 5] = nff;
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t t = 0;
        for (int i = 0; i < 8; i++) t += s_w[i];
        if (t) atomicAdd(&ctr[0], (unsigned long long)t);
        __threadfence();
        if (atomicAdd(&ctr[1], 1ull) == gridDim.x - 1) {
            __threadfence();
            const unsigned long long ff = atomicExch(&ctr[0], 0ull);
            ctr[1] = 0ull;
            *n_bytes = meta->error ? 0ll
                                   : (long long)((is_first ? meta->header_len : 0u) + total + ff + (is_last ? 2u : 0u));
        }
    }
}
// the header K2b wrote in front of the shard's own output, copied to the start of the (peer) file
__global__ void k_shard_copy_header(const uint8_t* own_out, const ImgMeta* meta, uint8_t* file, size_t capacity) {
    const uint32_t n = meta->error ? 0u : meta->header_len;
    for (uint32_t i = threadIdx.x; i < n && i < capacity; i += blockDim.x) file[i] = own_out[i];
}
// {end offset of this shard's bytes in the file, device-side error}: the value of the closing all-gather,
// which is also what tells the destination rank that every peer's K4 has finished writing
__global__ void k_shard_result(const ImgMeta* meta, long long* out2) {
    out2[0] = (long long)meta->out_len;
    out2[1] = (long long)meta->error;
}
// ---- single-process exchange over peer memory (dmmt_encode_sharded): every shard reads the values of the other
// shards straight from their devices (peer access over NVLink), so an exchange is one tiny kernel per shard
// instead of a host round trip.  mode 0: out[i] = sum over shards of src_j[i] (i < elems: the 4 histograms);
// mode 1: out[j * elems + i] = src_j[i] (all-gather); mode 2 (elems == 1): all-gather, then out[n + j] = the
// exclusive prefix sum over the shards.
__global__ void k_peer_exchange(PeerPtrs srcs, int n, int elems, int mode, long long* out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (mode == 0) {
        if (i >= elems) return;
        long long acc = 0;
        for (int j = 0; j < n; j++) acc += static_cast<const long long*>(srcs.p[j])[i];
        out[i] = acc;
    } else {
        if (i < n * elems) out[i] = static_cast<const long long*>(srcs.p[i / elems])[i % elems];
        if (mode == 2 && i == 0) {
            long long run = 0;
            for (int j = 0; j < n; j++) {
                out[n + j] = run;
                run += *static_cast<const long long*>(srcs.p[j]);
            }
        }
    }
}
}  // namespace
cudaError_t launch_peer_exchange(const PeerPtrs& srcs, int n, int elems, int mode, long long* out, cudaStream_t st) {
    const int work = mode == 0 ? elems : n * elems;
    k_peer_exchange<<<(work + 255) / 256, 256, 0, st>>>(srcs, n, elems, mode, out);
    return cudaGetLastError();
}

// ---- mailbox exchange between PROCESSES (one per GPU): the values the shards of one image exchange ----------------
// Every rank owns a mailbox in its device memory, mapped into the other processes through CUDA IPC.  k_mailbox_post
// stores a rank's payload into row `rank` of slot `slot` of EVERY mailbox (its own and, over NVLink, the peers') and then
// releases the row's flag with the sequence number of the encode; k_mailbox_collect, one kernel later on the same
// stream, waits until all `world` flags of the slot carry that number and reduces / gathers the rows from LOCAL memory.
// No host round trip, no library call: an exchange costs two small launches and one NVLink store latency.
// A flag that does not arrive within MAILBOX_TIMEOUT_NS (a peer died) reports DMMT_E_NCCL instead of hanging the GPU.
namespace {
constexpr int MAILBOX_ROW_WORDS = 1024 + 8;          // payload (<= 1024 words of 8 bytes) + the flag word
constexpr int MAILBOX_FLAG = 1024;
constexpr unsigned long long MAILBOX_TIMEOUT_NS = 10ull * 1000 * 1000 * 1000;
__device__ __forceinline__ unsigned long long* mailbox_row(void* box, int world, int slot, int rank) {
    return static_cast<unsigned long long*>(box) + ((size_t)slot * world + rank) * MAILBOX_ROW_WORDS;
}
__global__ void __launch_bounds__(256) k_mailbox_post(PeerPtrs boxes, int world, int rank, int slot, unsigned long long seq,
                                                      const unsigned long long* __restrict__ src, int n_words) {
    unsigned long long* row = mailbox_row(const_cast<void*>(boxes.p[blockIdx.x]), world, slot, rank);  // CTA = destination
    for (int i = threadIdx.x; i < n_words; i += blockDim.x) row[i] = src[i];
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(row + MAILBOX_FLAG), "l"(seq) : "memory");
}
// mode 0: out[i] = sum over ranks (i < n_words); 1: out[r * n_words + i] (gather); 2 (n_words = 1): gather + exclusive
// prefix sums in out[world ..]
__global__ void __launch_bounds__(1024) k_mailbox_collect(void* box, int world, int slot, unsigned long long seq, int mode,
                                                          int n_words, long long* out, ImgMeta* meta) {
    __shared__ int s_timeout;
    if (threadIdx.x == 0) s_timeout = 0;
    __syncthreads();
    if (threadIdx.x < world) {
        const unsigned long long* flag = mailbox_row(box, world, slot, threadIdx.x) + MAILBOX_FLAG;
        unsigned long long t0 = 0, v;
        while (true) {
            asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(flag) : "memory");
            if (v >= seq) break;
            unsigned long long now;
            asm volatile("mov.u64 %0, %globaltimer;" : "=l"(now));
            if (!t0) t0 = now;
            if (now - t0 > MAILBOX_TIMEOUT_NS) {
                s_timeout = 1;
                break;
            }
            __nanosleep(100);
        }
    }
    __syncthreads();
    if (s_timeout) {
        if (threadIdx.x == 0 && meta) atomicCAS(&meta->error, 0, DMMT_E_NCCL);
        return;
    }
    const int i = threadIdx.x;
    if (mode == 0) {
        for (int k = i; k < n_words; k += blockDim.x) {
            long long acc = 0;
            for (int r = 0; r < world; r++) acc += (long long)__ldcg(mailbox_row(box, world, slot, r) + k);
            out[k] = acc;
        }
    } else {
        for (int k = i; k < world * n_words; k += blockDim.x)
            out[k] = (long long)__ldcg(mailbox_row(box, world, slot, k / n_words) + k % n_words);
        if (mode == 2 && i == 0) {
            long long run = 0;
            for (int r = 0; r < world; r++) {
                out[world + r] = run;
                run += (long long)__ldcg(mailbox_row(box, world, slot, r));
            }
        }
    }
}
}  // namespace
size_t mailbox_bytes(int world, int slots) { return (size_t)slots * world * MAILBOX_ROW_WORDS * 8; }
cudaError_t launch_mailbox_post(const PeerPtrs& boxes, int world, int rank, int slot, unsigned long long seq, const void* src,
                                int n_words, cudaStream_t st) {
    k_mailbox_post<<<world, 256, 0, st>>>(boxes, world, rank, slot, seq, static_cast<const unsigned long long*>(src), n_words);
    return cudaGetLastError();
}
cudaError_t launch_mailbox_collect(void* box, int world, int slot, unsigned long long seq, int mode, int n_words, long long* out,
                                   ImgMeta* meta, cudaStream_t st) {
    k_mailbox_collect<<<1, 1024, 0, st>>>(box, world, slot, seq, mode, n_words, out, meta);
    return cudaGetLastError();
}

cudaError_t launch_shard_count_bytes(const uint8_t* scan, const ImgMeta* meta, const unsigned long long* seed_src,
                                     int owned_mode, const int* or_first_src, int is_first, int is_last,
                                     unsigned long long* ctr2, long long* n_bytes, cudaStream_t st) {
    k_shard_count_bytes<<<sm_count() * 2, 256, 0, st>>>(scan, meta, seed_src, owned_mode, or_first_src, is_first, is_last, ctr2,
                                                 n_bytes);
    return cudaGetLastError();
}
cudaError_t launch_shard_copy_header(const uint8_t* own_out, const ImgMeta* meta, uint8_t* file, size_t capacity,
                                     cudaStream_t st) {
    k_shard_copy_header<<<1, 256, 0, st>>>(own_out, meta, file, capacity);
    return cudaGetLastError();
}
cudaError_t launch_shard_result(const ImgMeta* meta, long long* out2, cudaStream_t st) {
    k_shard_result<<<1, 1, 0, st>>>(meta, out2);
    return cudaGetLastError();
}

cudaError_t launch_shard_widen(const int16_t* last_dc3, int* out4, const unsigned int* hist, long long* hist64,
                               const ImgMeta* meta, long long* bits_out, cudaStream_t st, long long* err_out) {
    k_shard_widen<<<4, 256, 0, st>>>(last_dc3, out4, hist, hist64, meta, bits_out, err_out);
    return cudaGetLastError();
}
cudaError_t launch_shard_narrow_seed(const int* seed4, int16_t* seed3, cudaStream_t st) {
    k_shard_narrow_seed<<<1, 32, 0, st>>>(seed4, seed3);
    return cudaGetLastError();
}
cudaError_t launch_shard_tail(const uint8_t* scan, const ImgMeta* meta, const unsigned long long* bit_offset, int is_last,
                              int* tail2, cudaStream_t st) {
    k_shard_tail<<<1, 1, 0, st>>>(scan, meta, bit_offset, is_last, tail2);
    return cudaGetLastError();
}
cudaError_t launch_shard_prev_tail(const int* all_tail2, const long long* all_offs, const long long* all_bits, int rank,
                                   int* out, cudaStream_t st) {
    k_shard_prev_tail<<<1, 1, 0, st>>>(all_tail2, all_offs, all_bits, rank, out);
    return cudaGetLastError();
}

cudaError_t launch_last_dc(const Geom& g, const int16_t* coef, int16_t* d_out3, cudaStream_t st) {
    k_last_dc<<<1, 32, 0, st>>>(coef, g.n_blocks, g.ypm, g.bpm, d_out3);
    return cudaGetLastError();
}

cudaError_t launch_k5_compact(const uint8_t* out, size_t out_stride, const unsigned long long* lens,
                              int n, uint8_t* dense, unsigned long long dense_cap,
                              unsigned long long* offsets, ImgMeta* meta, int chained,
                              int* sticky_err, cudaStream_t st) {
    k5_offsets<<<1, 1024, 0, st>>>(const_cast<unsigned long long*>(lens), n, dense_cap, offsets, meta,
                                   chained, sticky_err);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    k5_copy<<<dim3(8, n), 256, 0, st>>>(out, out_stride, lens, offsets, dense);
    return cudaGetLastError();
}

}  // namespace dmmt
