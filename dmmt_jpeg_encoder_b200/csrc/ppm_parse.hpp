// ppm_parse.hpp -- host ingest of ASCII P3 PPM files: the token rules of the
// reference's reader, image/reader/ppm.rs:41-251, as a buffer tokenizer instead of a byte-at-a-time stream.
//
// Rules kept exactly (ppm.rs:41-78, tests/test_host_logic.py):
//   * `#` starts a comment ANYWHERE, even inside a token; it runs through the next '\n' and does NOT end the
//     token around it ("1#x\n2" is the token "12");
//   * tokens are separated by ASCII whitespace as Rust's u8::is_ascii_whitespace knows it: space, \t, \n, \x0C,
//     \r (NOT \x0B, which is an ordinary -- and therefore invalid -- token byte);
//   * every number is Rust's str::parse::<u16>: optional '+', at least one digit, digits only, value <= 65535;
//   * tokens: "P3", width, height, max value, then the samples; errors in file order: missing header token,
//     unparsable token, `n % 3 != 0` samples, pixel count != width * height, sample > max (color.rs:62-65).
//
// The common case -- digits separated by whitespace -- is a tight loop over the buffer; anything else in a token
// ('#', '+', garbage) takes the general path.  A sample section without any '#' is cut at whitespace into one
// piece per thread (the CLI's -t/--threads, cli.rs:104-109, which the GPU encoder has no other use for).
// Pure host C++, no CUDA: included by dmmt_host.hpp (PPMImageReader) and dmmt_api.cu (dmmt_ppm_parse).
#pragma once
#include <cstddef>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <new>
#include <string>
#include <thread>
#include <vector>
#if defined(__SSE2__)
#include <emmintrin.h>
#endif

namespace dmmt_ppm {

enum Status {
    OK = 0,
    MISSING_TOKEN = 1,      // detail = index of the header token (0 "P3 Header", 1 width, 2 height, 3 max value)
    BAD_TOKEN = 2,          // detail = 1..3 header token, 4 "Color Component Value"
    INCOMPLETE_PIXEL = 3,   // detail = n % 3
    SIZE_MISMATCH = 4,
    SAMPLE_ABOVE_MAX = 5,   // the reference panics (color.rs:62-65)
};

// malloc'd array of samples with single ownership (what the C ABI hands out, dmmt_free == free)
struct Samples {
    uint16_t* ptr = nullptr;
    size_t n = 0;
    Samples() = default;
    Samples(uint16_t* p, size_t count) : ptr(p), n(count) {}
    Samples(const Samples&) = delete;
    Samples& operator=(const Samples&) = delete;
    Samples(Samples&& o) noexcept : ptr(o.ptr), n(o.n) { o.ptr = nullptr, o.n = 0; }
    Samples& operator=(Samples&& o) noexcept {
        if (this != &o) {
            std::free(ptr);
            ptr = o.ptr, n = o.n, o.ptr = nullptr, o.n = 0;
        }
        return *this;
    }
    ~Samples() { std::free(ptr); }
    uint16_t* release() {
        uint16_t* p = ptr;
        ptr = nullptr, n = 0;
        return p;
    }
    const uint16_t* data() const { return ptr; }
    size_t size() const { return n; }
    bool empty() const { return n == 0; }
    const uint16_t* begin() const { return ptr; }
    const uint16_t* end() const { return ptr + n; }
    uint16_t operator[](size_t i) const { return ptr[i]; }
};

struct Result {
    Status status = OK;
    int detail = 0;
    uint16_t width = 0, height = 0, max_value = 0;
    Samples samples;  // interleaved R,G,B
};

inline bool is_ws(unsigned char c) { return c == ' ' || c == '\t' || c == '\n' || c == '\x0C' || c == '\r'; }

// General tokenizer (comments, any bytes).  Returns false at the end of the buffer.
inline bool next_token(const char*& p, const char* end, std::string& out) {
    out.clear();
    bool in_comment = false;
    while (p < end) {
        const char c = *p++;
        if (in_comment) {
            if (c == '\n') in_comment = false;
            continue;
        }
        if (c == '#') {
            in_comment = true;
            continue;
        }
        if (is_ws((unsigned char)c)) {
            if (!out.empty()) break;
        } else {
            out.push_back(c);
        }
    }
    return !out.empty();
}

// Rust's str::parse::<u16>
inline bool parse_u16(const std::string& t, uint16_t& v_out) {
    size_t i = (!t.empty() && t[0] == '+') ? 1 : 0;
    if (i >= t.size()) return false;
    uint32_t v = 0;
    for (; i < t.size(); i++) {
        if (t[i] < '0' || t[i] > '9') return false;
        v = v * 10 + (uint32_t)(t[i] - '0');
        if (v > 65535) return false;
    }
    v_out = (uint16_t)v;
    return true;
}

// byte classes of the fast path: 0-9 = the digit's value, kWs = whitespace, kOther = everything else
constexpr uint8_t kWs = 0x80, kOther = 0xFF;
struct ByteClass {
    uint8_t c[256];
    constexpr ByteClass() : c() {
        for (int i = 0; i < 256; i++) c[i] = kOther;
        for (int i = 0; i < 10; i++) c['0' + i] = (uint8_t)i;
        c[' '] = c['\t'] = c['\n'] = c['\x0C'] = c['\r'] = kWs;
    }
};
inline const ByteClass& byte_class() {
    static constexpr ByteClass k{};
    return k;
}

// One token by SWAR: `w` = the eight bytes at the token start (little endian: first character in the low byte).
// Up to five digits followed by whitespace -> value (may exceed 65535) and the digit count; anything else -> -1.
// Borrows and carries of the two 64-bit operations only travel towards LATER bytes, so everything up to the
// first non-digit is exact; the shift right-aligns the digits and three multiplies fold eight digit bytes.
inline int swar_token(uint64_t w, const uint8_t* cls, uint32_t& value) {
    const uint64_t d = w - 0x3030303030303030ull;
    const uint64_t nondigit = (d | (d + 0x7676767676767676ull)) & 0x8080808080808080ull;
    const int nd = nondigit ? (__builtin_ctzll(nondigit) >> 3) : 8;  // digits before the first non-digit
    if (nd < 1 || nd > 5 || cls[(uint8_t)(w >> (8 * nd))] != kWs) return -1;
    uint64_t x = d << (8 * (8 - nd));  // the nd digit values in the top bytes, zeros ("leading 0s") below
    x = (x & 0x0F0F0F0F0F0F0F0Full) * 2561 >> 8;
    x = (x & 0x00FF00FF00FF00FFull) * 6553601 >> 16;
    x = (x & 0x0000FFFF0000FFFFull) * 42949672960001ull >> 32;
    value = (uint32_t)x;
    return nd;
}

// Samples of [p, end) -> o[0 .. n) (room for (end - p) / 2 + 1 values); false on the first unparsable token.
// Random 1-3 digit numbers make every per-character branch unpredictable (2.3 ns per BYTE on the bench host) and
// a per-token loop is bound by the dependency "token length -> next load address" (8 ns per token).  So:
//   1. 64-byte blocks (SSE2): bit masks of the digit and whitespace bytes; the token starts are the digit bits
//      whose predecessor is not a digit, enumerated with ctz -- no dependency on the numbers themselves;
//   2. every token is one unaligned 64-bit load and swar_token(), no data-dependent branch;
//   3. a block with any other byte ('#', '+', garbage), a longer token, or the last bytes of the range fall back
//      to the token-at-a-time loop below, which also owns the general rules (comments inside tokens ...).
inline bool parse_samples(const char* p, const char* end, uint16_t* o, size_t& n_out) {
    const uint8_t* const cls = byte_class().c;
    uint16_t* const o0 = o;
    std::string tok;
    bool ok = true;
#if defined(__SSE2__)
    {
        bool carry = false;  // the byte before p is a digit of a token that is already parsed
        const __m128i k0 = _mm_set1_epi8('0'), k9 = _mm_set1_epi8(9), ksp = _mm_set1_epi8(' '), kht = _mm_set1_epi8('\t'),
                      knl = _mm_set1_epi8('\n'), kff = _mm_set1_epi8('\x0C'), kcr = _mm_set1_epi8('\r');
        while (end - p >= 64 + 8) {
            uint64_t D = 0, W = 0;
            for (int i = 0; i < 4; i++) {
                const __m128i c = _mm_loadu_si128(reinterpret_cast<const __m128i*>(p + 16 * i));
                const __m128i t = _mm_sub_epi8(c, k0);
                const __m128i dg = _mm_cmpeq_epi8(_mm_min_epu8(t, k9), t);  // (c - '0') <= 9 unsigned
                const __m128i ws = _mm_or_si128(_mm_or_si128(_mm_cmpeq_epi8(c, ksp), _mm_cmpeq_epi8(c, knl)),
                                                _mm_or_si128(_mm_or_si128(_mm_cmpeq_epi8(c, kht), _mm_cmpeq_epi8(c, kcr)),
                                                             _mm_cmpeq_epi8(c, kff)));
                D |= (uint64_t)(uint32_t)_mm_movemask_epi8(dg) << (16 * i);
                W |= (uint64_t)(uint32_t)_mm_movemask_epi8(ws) << (16 * i);
            }
            if ((D | W) != ~0ull) break;  // some other byte in this block: token-at-a-time from here
            uint64_t starts = D & ~((D << 1) | (carry ? 1ull : 0ull));
            const char* bad = nullptr;
            while (starts) {
                const int s0 = __builtin_ctzll(starts);
                starts &= starts - 1;
                uint64_t w;
                std::memcpy(&w, p + s0, 8);
                uint32_t v;
                if (swar_token(w, cls, v) < 0) {  // more than five digits: the general rules decide
                    bad = p + s0;
                    break;
                }
                if (v > 65535) {
                    n_out = (size_t)(o - o0);
                    return false;
                }
                *o++ = (uint16_t)v;
            }
            if (bad) {
                p = bad, carry = false;
                break;
            }
            carry = (D >> 63) != 0;
            p += 64;
        }
        if (carry)
            while (p < end && cls[(uint8_t)*p] <= 9) ++p;  // rest of the token that straddles the block boundary
    }
#endif
    while (p < end) {
        if (end - p >= 8) {
            uint64_t w;
            std::memcpy(&w, p, 8);
            uint32_t v;
            const int nd = swar_token(w, cls, v);
            if (nd > 0) {
                if (v > 65535) {
                    ok = false;
                    break;
                }
                *o++ = (uint16_t)v;
                p += nd + 1;  // the separator too
                continue;
            }
        }
        // scalar paths: the last bytes of the range, '#' / '+' / garbage in a token, more than five digits
        const uint8_t c = cls[(uint8_t)*p];
        if (c == kWs) {
            ++p;
            continue;
        }
        const char* start = p;
        if (c <= 9) {  // up to five digits, then whitespace or the end of the buffer
            uint32_t v = c;
            uint8_t dd = kWs;
            ++p;
            while (p < end && (dd = cls[(uint8_t)*p]) <= 9 && p - start < 5) {
                v = v * 10 + dd;
                ++p;
            }
            if (p == end || dd == kWs) {
                if (v > 65535) {
                    ok = false;
                    break;
                }
                *o++ = (uint16_t)v;
                continue;
            }
            p = start;  // '#' or '+' inside, a sixth digit (leading zeros), garbage: the general path decides
        }
        if (!next_token(p, end, tok)) break;  // only comments were left
        uint16_t v;
        if (!parse_u16(tok, v)) {
            ok = false;
            break;
        }
        *o++ = v;
    }
    n_out = (size_t)(o - o0);
    return ok;
}

inline Result parse(const char* buf, size_t len, unsigned threads = 1) {
    Result r;
    const char *p = buf, *end = buf + len;
    std::string tok;
    if (!next_token(p, end, tok) || tok != "P3") {
        r.status = MISSING_TOKEN, r.detail = 0;
        return r;
    }
    uint16_t* hdr[3] = {&r.width, &r.height, &r.max_value};
    for (int i = 0; i < 3; i++) {
        if (!next_token(p, end, tok)) {
            r.status = MISSING_TOKEN, r.detail = i + 1;
            return r;
        }
        if (!parse_u16(tok, *hdr[i])) {
            r.status = BAD_TOKEN, r.detail = i + 1;
            return r;
        }
    }
    const size_t rest = (size_t)(end - p);
    // pieces are parsed into raw malloc'd buffers sized for the densest possible text (one digit + one separator
    // per sample); one thread: that buffer, shrunk, IS the result; several: the pieces are copied once, in parallel
    struct Piece {
        const char *b = nullptr, *e = nullptr;
        uint16_t* v = nullptr;
        size_t n = 0, at = 0;
        uint16_t top = 0;  // largest sample of the piece (color.rs:62-65 check)
        bool ok = true, nomem = false;
    };
    if (threads > 64) threads = 64;
    if (threads < 2 || rest < (1u << 20) || std::memchr(p, '#', rest)) threads = 1;  // a comment could swallow a cut
    std::vector<Piece> piece(threads);
    for (unsigned t = 0; t < threads; t++) {
        const char* c = t ? p + rest / threads * t : p;
        if (t)
            while (c < end && !is_ws((unsigned char)*c)) ++c;  // cut at whitespace
        piece[t].b = c;
        if (t) piece[t - 1].e = c;
    }
    piece[threads - 1].e = end;
    auto run = [&](auto&& fn) {  // fn(t) on `threads` threads
        if (threads == 1) return fn(0u);
        std::vector<std::thread> pool;
        for (unsigned t = 1; t < threads; t++) pool.emplace_back(fn, t);
        fn(0u);
        for (auto& th : pool) th.join();
    };
    run([&](unsigned t) {
        Piece& q = piece[t];
        q.v = static_cast<uint16_t*>(std::malloc(((size_t)(q.e - q.b) / 2 + 1) * sizeof(uint16_t)));
        if (!q.v) {
            q.nomem = true;
            return;
        }
        q.ok = parse_samples(q.b, q.e, q.v, q.n);
        uint16_t top = 0;
        for (size_t i = 0; i < q.n; i++) top = q.v[i] > top ? q.v[i] : top;
        q.top = top;
    });
    struct FreePieces {
        std::vector<Piece>& v;
        ~FreePieces() {
            for (Piece& q : v) std::free(q.v);
        }
    } free_pieces{piece};
    bool ok = true;
    size_t n = 0;
    uint16_t top = 0;
    for (Piece& q : piece) {
        if (q.nomem) throw std::bad_alloc();
        ok = ok && q.ok, q.at = n, n += q.n, top = q.top > top ? q.top : top;
    }
    if (!ok) {
        r.status = BAD_TOKEN, r.detail = 4;
        return r;
    }
    if (n % 3) {
        r.status = INCOMPLETE_PIXEL, r.detail = (int)(n % 3);
        return r;
    }
    if (n / 3 != (size_t)r.width * r.height) {
        r.status = SIZE_MISMATCH;
        return r;
    }
    if (top > r.max_value) {
        r.status = SAMPLE_ABOVE_MAX;
        return r;
    }
    if (threads == 1) {
        uint16_t* shrunk = static_cast<uint16_t*>(std::realloc(piece[0].v, (n ? n : 1) * sizeof(uint16_t)));
        r.samples = Samples(shrunk ? shrunk : piece[0].v, n);
        piece[0].v = nullptr;
    } else {
        uint16_t* all = static_cast<uint16_t*>(std::malloc((n ? n : 1) * sizeof(uint16_t)));
        if (!all) throw std::bad_alloc();
        run([&](unsigned t) {
            if (piece[t].n) std::memcpy(all + piece[t].at, piece[t].v, piece[t].n * sizeof(uint16_t));
        });
        r.samples = Samples(all, n);
    }
    return r;
}

}  // namespace dmmt_ppm
