// dmmt-jpeg-encoder -- drop-in for the reference CLI (src/main.rs:5-12): parses the same flags
// (src/cli.rs), converts one P3 PPM to a baseline JPEG on the GPU, prints
// "Conversion successful" / "Conversion failed because of: <Display of Error>" and exits with
// status 0 in both cases; usage errors exit with status 2 like clap.
#include <cstdio>

#include "dmmt_host.hpp"

int main(int argc, char** argv) {
    using namespace dmmt_host;
    Arguments arguments;
    try {
        arguments = CLIParser().parse(argc, argv);
    } catch (const UsageError& e) {
        if (std::string(e.what()) == "help") {
            std::fputs(CLIParser::usage(), stdout);
            return 0;
        }
        std::fprintf(stderr, "error: %s\n\n%s", e.what(), CLIParser::usage());
        return 2;
    }
    try {
        convert_ppm_to_jpeg(arguments);
        std::puts("Conversion successful");
    } catch (const Error& e) {
        std::fprintf(stderr, "Conversion failed because of: %s\n", e.what());
    } catch (const Panic& e) {  // the reference panics here (exit status 101)
        std::fprintf(stderr, "thread 'main' panicked: %s\n", e.what());
        return 101;
    }
    return 0;
}
