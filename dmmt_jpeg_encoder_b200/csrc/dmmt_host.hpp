// dmmt_host.hpp -- C++ host side above the C ABI: the reference crate's public encode interface
// (same names, argument meaning and error behaviour), with JpegImageWriter::write_image's body
// replaced by ONE call into libdmmt_cuda (dmmt_encode).  Citations: /root/reference/src.
//
//   Error (+ Display texts)                       error.rs:4-98
//   Image<f32>                                    image.rs:7-11
//   PPMImageReader::read_image                    image/reader/ppm.rs:9-251
//   ChromaSubsamplingPreset                       image/subsampling.rs:11-46
//   QuantizationTablePreset (+ CLI spellings)     image/writer/jpeg/quantization_tables.rs:232-284
//   JpegTransformationOptions / JpegImageWriter   image/writer/jpeg.rs:25-75
//   Arguments / CLIParser                         lib.rs:33-40, cli.rs:12-180
//   convert_ppm_to_jpeg                           lib.rs:59-77
//
// There is no CPU encode path here: without a CUDA device write_image fails (DMMT_E_NODEVICE).
#pragma once

#include <cerrno>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <istream>
#include <ostream>
#include <stdexcept>
#include <string>
#include <thread>
#include <vector>

#include "../../include/dmmt_cuda.h"
#include "ppm_parse.hpp"

namespace dmmt_host {

// ------------------------------------------------------------------------------------ errors
// crate::Error; what() is the reference's Display text (error.rs:25-98)
struct Error : std::runtime_error {
    using std::runtime_error::runtime_error;
};
inline Error PPMFileDoesNotContainRequiredToken(const char* t) {
    return Error(std::string("Expected token '") + t + "' not found in PPM file");
}
inline Error ParsingOfTokenFailed(const char* t) { return Error(std::string("Parsing of token '") + t + "' failed"); }
inline Error IncompletePixelParsed(size_t n) {
    return Error("Incomplete pixel parsed. Expected 3 components, but got " + std::to_string(n) + ".");
}
inline Error MismatchOfSizeBetweenHeaderAndValues() {
    return Error("Nubmer of pixels do not match the size, provided in header");  // sic, error.rs:46
}
inline Error UnableToOpenInputFileForReading(const std::string& p, const std::string& why) {
    return Error("Unable to open input file '" + p + "' for reading: " + why);
}
inline Error UnableToOpenOutputFileForWriting(const std::string& p, const std::string& why) {
    return Error("Unable to open output file '" + p + "' for writing: " + why);
}
inline Error FailedToWriteImageData() { return Error("Failed to write image data"); }
inline Error HuffmanSymbolNotPresentInTranslator() { return Error("Huffman symbol '?' not present in cuda translator"); }
// conditions on which the reference panics instead of returning an Error
struct Panic : std::logic_error {
    using std::logic_error::logic_error;
};

constexpr const char* P3_HEADER_TOKEN_NAME = "P3 Header";
constexpr const char* WIDTH_HEADER_TOKEN_NAME = "Width Header";
constexpr const char* HEIGHT_HEADER_TOKEN_NAME = "Height Header";
constexpr const char* MAX_VALUE_HEADER_TOKEN_NAME = "Max Value Header";
constexpr const char* COLOR_COMPONENT_VALUE_TOKEN_NAME = "Color Component Value";

// ------------------------------------------------------------------------------------ enums
enum class ChromaSubsamplingPreset : uint8_t { P444 = DMMT_P444, P422 = DMMT_P422, P420 = DMMT_P420 };
enum class QuantizationTablePreset : uint8_t {
    Specification = 0, Flat, MSSIMKodakTuned, PSNRHVSNKodakTuned, DCTunePerceptualOptimization, AVisualDetectionModel,
    AnImprovedDetectionModel
};

// ------------------------------------------------------------------------------------ image
// Image<f32> (image.rs:7-11).  The reader keeps the raw samples + max value: the device performs the
// identical IEEE `v as f32 / max as f32` (color.rs:45-53), so `dots` is only materialised on demand.
struct Image {
    uint16_t width = 0, height = 0;
    uint16_t max_value = 0;            // 0: `dots` is authoritative
    dmmt_ppm::Samples samples;         // interleaved R,G,B (u16, as the reader delivers them)
    std::vector<float> dots;           // normalised f32 RGB
    const std::vector<float>& normalised() {
        if (dots.empty() && max_value) {
            dots.resize(samples.size());
            for (size_t i = 0; i < samples.size(); i++) dots[i] = (float)samples[i] / (float)max_value;
        }
        return dots;
    }
};

// ASCII P3 reader with the reference's token rules (ppm.rs:41-78): `#` starts a comment that runs
// through the next newline ANYWHERE (even inside a token), tokens split on ASCII whitespace
// (space, \t, \n, \x0C, \r -- not \x0B), every number must parse as u16.  The stream is read in one go
// and tokenised from memory (ppm_parse.hpp): a 3840x2160 file takes ~0.1 s (one thread) instead of the
// 1.1 s of a byte-at-a-time reader -- the GPU encodes that frame in 0.1 ms, so ingest IS the CLI's run time.
class PPMImageReader {
   public:
    explicit PPMImageReader(std::istream& reader, size_t threads = 1) : reader_(reader), threads_(threads) {}
    Image read_image() {
        const std::string text = slurp(reader_);
        dmmt_ppm::Result r = dmmt_ppm::parse(text.data(), text.size(), (unsigned)(threads_ ? threads_ : 1));
        static const char* const kTok[5] = {P3_HEADER_TOKEN_NAME, WIDTH_HEADER_TOKEN_NAME, HEIGHT_HEADER_TOKEN_NAME,
                                            MAX_VALUE_HEADER_TOKEN_NAME, COLOR_COMPONENT_VALUE_TOKEN_NAME};
        switch (r.status) {
            case dmmt_ppm::OK: break;
            case dmmt_ppm::MISSING_TOKEN: throw PPMFileDoesNotContainRequiredToken(kTok[r.detail]);
            case dmmt_ppm::BAD_TOKEN: throw ParsingOfTokenFailed(kTok[r.detail]);
            case dmmt_ppm::INCOMPLETE_PIXEL: throw IncompletePixelParsed((size_t)r.detail);
            case dmmt_ppm::SIZE_MISMATCH: throw MismatchOfSizeBetweenHeaderAndValues();
            case dmmt_ppm::SAMPLE_ABOVE_MAX: throw Panic("color component exceeds the max value (color.rs:62-65)");
        }
        Image im;
        im.width = r.width, im.height = r.height, im.max_value = r.max_value;
        im.samples = std::move(r.samples);
        return im;
    }

   private:
    std::istream& reader_;
    size_t threads_;
    static std::string slurp(std::istream& in) {
        std::string s;
        const std::istream::pos_type at = in.tellg();
        if (at != std::istream::pos_type(-1) && in.seekg(0, std::ios::end)) {  // seekable: one read of the right size
            const std::istream::pos_type end = in.tellg();
            in.seekg(at);
            if (end > at) {
                s.resize((size_t)(end - at));
                in.read(&s[0], (std::streamsize)s.size());
                s.resize((size_t)in.gcount());
                return s;
            }
        }
        in.clear();
        char buf[1 << 16];
        while (in.read(buf, sizeof buf) || in.gcount() > 0) s.append(buf, (size_t)in.gcount());
        return s;
    }
};

// ------------------------------------------------------------------------------------ writer
struct JpegTransformationOptions {
    ChromaSubsamplingPreset chroma_subsampling_preset = ChromaSubsamplingPreset::P420;
    uint8_t bits_per_channel = 8;
    QuantizationTablePreset quantization_table_preset = QuantizationTablePreset::Specification;
};

// one CUDA context per process, created on first use (the reference's ThreadPool has no role here)
inline dmmt_ctx* default_context() {
    static dmmt_ctx* ctx = nullptr;
    if (!ctx) {
        const int rc = dmmt_ctx_create(0, &ctx);
        if (rc != DMMT_OK) throw Panic(std::string("dmmt_ctx_create: ") + dmmt_strerror(rc));
    }
    return ctx;
}

class JpegImageWriter {
   public:
    // JpegImageWriter::new(writer, &image, &options, &threadpool) (jpeg.rs:48-62); the thread pool of the
    // reference only fans the DCT out (transformer.rs:126-148) and has no counterpart
    JpegImageWriter(std::ostream& writer, Image& image, const JpegTransformationOptions& options, dmmt_ctx* ctx = nullptr)
        : writer_(writer), image_(image), options_(options), ctx_(ctx) {}
    void write_image() {
        dmmt_image im{};
        im.width = image_.width, im.height = image_.height;
        std::vector<uint8_t> u8;
        if (image_.max_value) {
            im.max_value = image_.max_value;
            if (image_.max_value <= 255) {
                u8.assign(image_.samples.begin(), image_.samples.end());
                im.fmt = DMMT_RGB_U8, im.pixels = u8.data();
            } else {
                im.fmt = DMMT_RGB_U16, im.pixels = image_.samples.data();
            }
        } else {
            im.max_value = 1, im.fmt = DMMT_RGB_F32_NORM, im.pixels = image_.dots.data();
        }
        dmmt_options o{(uint8_t)options_.chroma_subsampling_preset, options_.bits_per_channel,
                       (uint8_t)options_.quantization_table_preset};
        uint8_t* jpeg = nullptr;
        size_t len = 0;
        const int rc = dmmt_encode(ctx_ ? ctx_ : default_context(), &im, &o, &jpeg, &len);
        if (rc == DMMT_E_SYMBOL) throw HuffmanSymbolNotPresentInTranslator();
        if (rc == DMMT_E_WRITE) throw FailedToWriteImageData();
        if (rc != DMMT_OK) {
            std::string msg = std::string("dmmt_encode: ") + dmmt_strerror(rc);
            if (rc == DMMT_E_CUDA) msg += std::string(": ") + dmmt_last_cuda_error();
            throw Panic(msg);
        }
        writer_.write(reinterpret_cast<const char*>(jpeg), (std::streamsize)len);
        dmmt_free(jpeg);
        if (!writer_) throw FailedToWriteImageData();
        writer_.flush();
    }

   private:
    std::ostream& writer_;
    Image& image_;
    JpegTransformationOptions options_;
    dmmt_ctx* ctx_;
};

// ------------------------------------------------------------------------------------ CLI
struct Arguments {
    std::string input_file, output_file;
    uint8_t bits_per_channel = 8;
    ChromaSubsamplingPreset chroma_subsampling_preset = ChromaSubsamplingPreset::P420;
    size_t number_of_threads = 1;
    QuantizationTablePreset quantization_table_preset = QuantizationTablePreset::Specification;
};

// clap usage errors print to stderr and exit with status 2 (cli.rs:28-31)
struct UsageError : std::runtime_error {
    using std::runtime_error::runtime_error;
};

class CLIParser {
   public:
    static const char* usage() {
        return "Usage: dmmt-jpeg-encoder [OPTIONS] <input_file> <output_file>\n\n"
               "Arguments:\n  <input_file>   Path to PPM imput file\n  <output_file>  Path to JPEG output file\n\n"
               "Options:\n"
               "  -b, --bits_per_channel <BITS>              Bits per color channel [default: 8] [possible values: 8, 16, 32]\n"
               "  -p, --chroma_subsampling_preset <PRESET>   Chroma subsampling preset [default: P420] [possible values: P444, P422, P420]\n"
               "  -t, --threads <THREADS>                    Number of Threads\n"
               "  -q, --quantization_table <TABLE>           Quantization table preset [default: Specification]\n"
               "  -h, --help                                 Print help\n";
    }
    Arguments parse(int argc, const char* const* argv) const {
        Arguments a;
        const unsigned hc = std::thread::hardware_concurrency();
        a.number_of_threads = hc ? hc : 1;  // cli.rs:104-109,178-180
        std::vector<std::string> pos;
        for (int i = 1; i < argc; i++) {
            std::string s = argv[i], val;
            auto value = [&](const char* name) {
                if (!val.empty()) return val;
                if (i + 1 >= argc) throw UsageError(std::string("a value is required for '") + name + "' but none was supplied");
                return std::string(argv[++i]);
            };
            const size_t eq = s.rfind("--", 0) == 0 ? s.find('=') : std::string::npos;
            if (eq != std::string::npos) val = s.substr(eq + 1), s = s.substr(0, eq);
            if (s == "-h" || s == "--help") throw UsageError("help");
            else if (s == "-b" || s == "--bits_per_channel") {
                const std::string v = value("--bits_per_channel <BITS>");
                if (v != "8" && v != "16" && v != "32") throw UsageError("invalid value '" + v + "' for '--bits_per_channel <BITS>'\n  [possible values: 8, 16, 32]");
                a.bits_per_channel = (uint8_t)std::stoi(v);
            } else if (s == "-p" || s == "--chroma_subsampling_preset") {
                const std::string v = value("--chroma_subsampling_preset <PRESET>");
                if (v == "P444") a.chroma_subsampling_preset = ChromaSubsamplingPreset::P444;
                else if (v == "P422") a.chroma_subsampling_preset = ChromaSubsamplingPreset::P422;
                else if (v == "P420") a.chroma_subsampling_preset = ChromaSubsamplingPreset::P420;
                else throw UsageError("invalid value '" + v + "' for '--chroma_subsampling_preset <PRESET>'\n  [possible values: P444, P422, P420]");
            } else if (s == "-t" || s == "--threads") {
                const std::string v = value("--threads <THREADS>");
                if (v.empty() || v.find_first_not_of("0123456789") != std::string::npos) throw UsageError("invalid value '" + v + "' for '--threads <THREADS>': invalid digit found in string");
                a.number_of_threads = (size_t)std::stoull(v);
            } else if (s == "-q" || s == "--quantization_table") {
                a.quantization_table_preset = table(value("--quantization_table <TABLE>"));
            } else if (s.size() > 1 && s[0] == '-') {
                throw UsageError("unexpected argument '" + s + "' found");
            } else {
                pos.push_back(s);
            }
        }
        if (pos.size() < 2) throw UsageError("the following required arguments were not provided:\n  <input_file>\n  <output_file>");
        if (pos.size() > 2) throw UsageError("unexpected argument '" + pos[2] + "' found");
        a.input_file = pos[0], a.output_file = pos[1];
        return a;
    }

   private:
    // spellings + aliases of quantization_tables.rs:260-283 (numeric aliases skip 3 and 5)
    static QuantizationTablePreset table(const std::string& v) {
        using Q = QuantizationTablePreset;
        if (v == "Specification" || v == "Spec" || v == "Default" || v == "0") return Q::Specification;
        if (v == "Flat" || v == "1") return Q::Flat;
        if (v == "MSSIM-Kodak-Tuned" || v == "2") return Q::MSSIMKodakTuned;
        if (v == "PSNR-HVS-N-Kodak-Tuned" || v == "4") return Q::PSNRHVSNKodakTuned;
        if (v == "DCTune-Perceptual-Optimization" || v == "6") return Q::DCTunePerceptualOptimization;
        if (v == "A-visual-detection-model" || v == "7") return Q::AVisualDetectionModel;
        if (v == "An-improved-detection-model" || v == "8") return Q::AnImprovedDetectionModel;
        throw UsageError("invalid value '" + v + "' for '--quantization_table <TABLE>'");
    }
};

// Display of std::io::Error: "<strerror> (os error <errno>)"
inline std::string os_error(int e) { return std::string(std::strerror(e)) + " (os error " + std::to_string(e) + ")"; }

// lib.rs:59-77: opens the input, creates/truncates the output (before parsing), reads the P3 image,
// writes the JPEG
inline void convert_ppm_to_jpeg(const Arguments& arguments) {
    std::ifstream in(arguments.input_file, std::ios::binary);
    if (!in) throw UnableToOpenInputFileForReading(arguments.input_file, os_error(errno));
    std::ofstream out(arguments.output_file, std::ios::binary | std::ios::trunc);
    if (!out) throw UnableToOpenOutputFileForWriting(arguments.output_file, os_error(errno));
    PPMImageReader reader(in, arguments.number_of_threads);  // -t/--threads (cli.rs:104-109) drives the P3 ingest
    Image image = reader.read_image();
    JpegTransformationOptions options{arguments.chroma_subsampling_preset, arguments.bits_per_channel,
                                      arguments.quantization_table_preset};
    JpegImageWriter(out, image, options).write_image();
}

}  // namespace dmmt_host
