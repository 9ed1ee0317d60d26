// dmmt-jpeg-batch -- batch front-end of the B200 encode path (many files per process, packed output
// arena, direct file writes from pinned memory).  The reference converts ONE image per process (src/main.rs:5-12,
// src/lib.rs:59-77); this tool keeps its options and messages and widens the call:
//
//   dmmt-jpeg-batch [-p P444|P422|P420] [-q TABLE] [-b 8|16|32] [-t THREADS] <output_dir> <input.ppm>...
//
// P3 files are parsed on -t host threads (one file per thread, dmmt_ppm_parse rules), grouped by geometry and max
// value, staged in page-locked memory and encoded by dmmt_batch_encode_host (sub-batches pipelined over 3 streams:
// H2D || kernels || D2H); every <stem>.jpg is then written with one write(2) straight from the pinned output arena.
// Prints one line per failed input ("Conversion of <file> failed because of: ...") and "Converted N of M files".
#include <fcntl.h>
#include <unistd.h>

#include <atomic>
#include <cstdio>
#include <map>
#include <thread>
#include <tuple>

#include "dmmt_host.hpp"

using namespace dmmt_host;

namespace {
struct Input {
    std::string path, stem;
    Image image;
    std::string error;  // Display text of the reader's Error, empty = parsed
};

std::string stem_of(const std::string& p) {
    const size_t s = p.find_last_of('/');
    std::string f = s == std::string::npos ? p : p.substr(s + 1);
    const size_t d = f.find_last_of('.');
    return d == std::string::npos || d == 0 ? f : f.substr(0, d);
}

bool write_all(const std::string& path, const uint8_t* p, size_t n) {
    const int fd = ::open(path.c_str(), O_WRONLY | O_CREAT | O_TRUNC, 0644);
    if (fd < 0) return false;
    while (n) {
        const ssize_t w = ::write(fd, p, n);
        if (w <= 0) {
            ::close(fd);
            return false;
        }
        p += w, n -= (size_t)w;
    }
    return ::close(fd) == 0;
}
}  // namespace

int main(int argc, char** argv) {
    // the reference's options; the positionals become <output_dir> <input>...
    std::vector<const char*> opt_args{argv[0]};
    std::vector<std::string> pos;
    for (int i = 1; i < argc; i++) {
        const std::string s = argv[i];
        if (s.size() > 1 && s[0] == '-') {
            opt_args.push_back(argv[i]);
            if (s.find('=') == std::string::npos && s != "-h" && s != "--help" && i + 1 < argc) opt_args.push_back(argv[++i]);
        } else {
            pos.push_back(s);
        }
    }
    Arguments a;
    try {
        opt_args.push_back("in"), opt_args.push_back("out");  // placeholders for the two positionals the parser expects
        a = CLIParser().parse((int)opt_args.size(), opt_args.data());
    } catch (const UsageError& e) {
        std::fprintf(stderr, "error: %s\n\nUsage: dmmt-jpeg-batch [OPTIONS] <output_dir> <input.ppm>...\n(options as dmmt-jpeg-encoder)\n", e.what());
        return 2;
    }
    if (pos.size() < 2) {
        std::fprintf(stderr, "error: the following required arguments were not provided:\n  <output_dir>\n  <input.ppm>...\n");
        return 2;
    }
    const std::string out_dir = pos[0];
    std::vector<Input> in(pos.size() - 1);
    for (size_t i = 0; i < in.size(); i++) in[i].path = pos[i + 1], in[i].stem = stem_of(pos[i + 1]);

    // ---- ingest: one file per host thread
    std::atomic<size_t> next{0};
    auto worker = [&]() {
        for (size_t i; (i = next.fetch_add(1)) < in.size();) {
            try {
                std::ifstream f(in[i].path, std::ios::binary);
                if (!f) throw UnableToOpenInputFileForReading(in[i].path, os_error(errno));
                in[i].image = PPMImageReader(f, 1).read_image();
            } catch (const std::exception& e) {
                in[i].error = e.what();
            }
        }
    };
    std::vector<std::thread> pool;
    for (size_t t = 0; t < std::max<size_t>(1, std::min(a.number_of_threads, in.size())); t++) pool.emplace_back(worker);
    for (auto& t : pool) t.join();

    // ---- groups of equal geometry / max value -> one pipelined batch each
    std::map<std::tuple<int, int, int>, std::vector<size_t>> groups;
    size_t converted = 0;
    for (size_t i = 0; i < in.size(); i++) {
        if (!in[i].error.empty()) {
            std::fprintf(stderr, "Conversion of %s failed because of: %s\n", in[i].path.c_str(), in[i].error.c_str());
            continue;
        }
        groups[{in[i].image.width, in[i].image.height, in[i].image.max_value}].push_back(i);
    }
    dmmt_ctx* ctx = nullptr;
    int rc = dmmt_ctx_create(0, &ctx);
    if (rc != DMMT_OK) {
        std::fprintf(stderr, "thread 'main' panicked: dmmt_ctx_create: %s\n", dmmt_strerror(rc));
        return 101;
    }
    const dmmt_options o{(uint8_t)a.chroma_subsampling_preset, a.bits_per_channel, (uint8_t)a.quantization_table_preset};
    for (auto& [key, idx] : groups) {
        const auto [w, h, mx] = key;
        const bool u8 = mx <= 255;
        const size_t n = idx.size(), px = (size_t)w * h * 3, img_bytes = px * (u8 ? 1 : 2);
        void *h_in = nullptr, *h_out = nullptr;
        dmmt_batch* b = nullptr;
        const int sub = (int)std::min<size_t>(n, std::max<size_t>(1, (256u << 20) / img_bytes));  // <= 256 MB of pixels per slot
        rc = dmmt_batch_create(ctx, (uint16_t)w, (uint16_t)h, u8 ? DMMT_RGB_U8 : DMMT_RGB_U16, (uint16_t)mx, &o, sub, 3, &b);
        size_t cap = n * (px / 2 + 4096);  // generous for photographic content; grown on DMMT_E_WRITE
        if (rc == DMMT_OK) rc = dmmt_host_alloc(n * img_bytes, &h_in);
        if (rc == DMMT_OK) {
            for (size_t k = 0; k < n; k++) {  // samples -> pinned staging (u8 when the max value allows: a quarter of the f32 bytes)
                const auto& s = in[idx[k]].image.samples;
                if (u8) {
                    uint8_t* d = static_cast<uint8_t*>(h_in) + k * img_bytes;
                    for (size_t i = 0; i < px; i++) d[i] = (uint8_t)s[i];
                } else {
                    std::memcpy(static_cast<uint8_t*>(h_in) + k * img_bytes, s.data(), img_bytes);
                }
            }
        }
        std::vector<uint64_t> offs(n), lens(n);
        for (int attempt = 0; rc == DMMT_OK && attempt < 4; attempt++) {
            rc = dmmt_host_alloc(cap, &h_out);
            if (rc != DMMT_OK) break;
            rc = dmmt_batch_encode_host(b, h_in, (int)n, static_cast<uint8_t*>(h_out), cap, offs.data(), lens.data());
            if (rc != DMMT_E_OVERFLOW && rc != DMMT_E_WRITE) break;
            dmmt_host_free(h_out), h_out = nullptr;
            if (attempt == 3) break;
            if (rc == DMMT_E_WRITE) cap *= 4;  // the host arena was too small
            // denser than 128 B of scan per block: the worst case always fits (the reference encodes any input)
            rc = rc == DMMT_E_OVERFLOW ? dmmt_batch_set_scan_capacity(b, dmmt_batch_worst_case_scan_bytes(b)) : DMMT_OK;
        }
        for (size_t k = 0; k < n; k++) {
            const std::string path = out_dir + "/" + in[idx[k]].stem + ".jpg";
            if (rc != DMMT_OK) {
                std::fprintf(stderr, "Conversion of %s failed because of: %s\n", in[idx[k]].path.c_str(),
                             rc == DMMT_E_SYMBOL ? HuffmanSymbolNotPresentInTranslator().what() : dmmt_strerror(rc));
            } else if (!write_all(path, static_cast<const uint8_t*>(h_out) + offs[k], (size_t)lens[k])) {
                std::fprintf(stderr, "Conversion of %s failed because of: %s\n", in[idx[k]].path.c_str(),
                             UnableToOpenOutputFileForWriting(path, os_error(errno)).what());
            } else {
                converted++;
            }
        }
        dmmt_host_free(h_out), dmmt_host_free(h_in);
        dmmt_batch_destroy(b);
    }
    dmmt_ctx_destroy(ctx);
    std::printf("Converted %zu of %zu files\n", converted, in.size());
    return 0;
}
