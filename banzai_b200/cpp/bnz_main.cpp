// bnz — command-line front end over the B200 encoder, flag- and exit-code-compatible with the
// reference CLI (bnz/src/main.rs): same options (:32-59, 183-257), same defaults for the output
// path (:268-283) and for removing the input (:292-309), same exit codes (:11-14).
#include <cstdio>
#include <cstring>
#include <fstream>
#include <iostream>
#include <string>

#include "banzai.hpp"

namespace {

enum { SUCCESS = 0, ERR_ARGS = 1, ERR_FILESYSTEM = 2, ERR_OUTPUT = 3 };   // main.rs:11-14

const char *TAGLINE = "bnz (banzai_b200): bzip2 encoder with banzai's interface, running on NVIDIA B200";
const char *VERSION = "version b200-r1 (stream-compatible with banzai alpha 0.3.1)";

[[noreturn]] void die_args(const std::string &msg)
{
    std::fprintf(stderr, "%s\n", msg.c_str());
    std::exit(ERR_ARGS);
}

[[noreturn]] void die_fs(const std::string &msg)
{
    std::fprintf(stderr, "[filesystem error] %s\n", msg.c_str());
    std::exit(ERR_FILESYSTEM);
}

[[noreturn]] void die_synopsis()
{
    std::fprintf(stderr, "%s\n   'bnz --help' lists the options, 'bnz --info' describes the program\n%s\n", TAGLINE, VERSION);
    std::exit(ERR_ARGS);
}

[[noreturn]] void die_help()
{
    std::fprintf(stderr,
                 "%s\n\n"
                 "  usage: bnz [options] <input_path>\n\n"
                 "  options:\n"
                 "     --output <path.bz2>    write the stream to this file\n"
                 "     --stdout, -c           write the stream to standard output\n"
                 "     --keep, -k             keep the input file\n"
                 "     --remove, -r           remove the input file\n"
                 "     -1 .. -9               block size in 100 kB units (default -9)\n"
                 "     --fast / --best        same as -1 / -9\n"
                 "     --verbose, -v          accepted for compatibility\n\n"
                 "  commands:\n"
                 "     --help  --info  --version\n\n"
                 "  '-' as the input path reads standard input. Without --output/--stdout the result\n"
                 "  goes to '<input_path>.bz2' (or to stdout when reading stdin). The input file is\n"
                 "  removed only when no output was named, unless --keep/--remove say otherwise.\n\n"
                 "%s\n",
                 TAGLINE, VERSION);
    std::exit(SUCCESS);
}

[[noreturn]] void die_info()
{
    std::fprintf(stderr,
                 "%s\n\n"
                 "The Burrows-Wheeler transform is a cyclic prefix-doubling radix sort, and RLE1, MTF,\n"
                 "Huffman modelling and bit packing are CUDA kernels for sm_100a; the output is\n"
                 "byte-identical to banzai 0.3.1 at every level. There is no CPU fallback.\n\n%s\n",
                 TAGLINE, VERSION);
    std::exit(SUCCESS);
}

struct Invocation {
    enum In { IN_NONE, IN_FILE, IN_STDIN } in = IN_NONE;
    enum Out { OUT_NONE, OUT_FILE, OUT_STDOUT } out = OUT_NONE;
    std::string in_path, out_path;
    int keep = -1;        // -1 unspecified
    int level = 9;
    bool verbose = false;

    void set_input(In kind, const std::string &p)
    {
        if (in != IN_NONE) die_args("Only one input may be specified");
        in = kind;
        in_path = p;
    }
    void set_output(Out kind, const std::string &p)
    {
        if (out == OUT_NONE) {
            out = kind;
            out_path = p;
            return;
        }
        if (out == OUT_STDOUT && kind == OUT_STDOUT) return;   // main.rs:154-158
        die_args("Only one output may be specified");
    }
};

}  // namespace

int main(int argc, char **argv)
{
    if (argc <= 1) die_synopsis();
    Invocation inv;
    enum { ANY, NOARGS, OUTPATH } expect = ANY;
    for (int i = 1; i < argc; i++) {
        std::string a = argv[i];
        if (expect == OUTPATH) {
            if (!a.empty() && a[0] == '-') die_args("Argument '--output' requires a file path");
            inv.set_output(Invocation::OUT_FILE, a);
            expect = ANY;
        } else if (expect == ANY && a.rfind("--", 0) == 0) {
            if (a == "--help") die_help();
            else if (a == "--version") { std::fprintf(stderr, "%s\n", VERSION); return SUCCESS; }
            else if (a == "--info") die_info();
            else if (a == "--verbose") inv.verbose = true;
            else if (a == "--keep") inv.keep = 1;
            else if (a == "--remove") inv.keep = 0;
            else if (a == "--fast") inv.level = 1;
            else if (a == "--best") inv.level = 9;
            else if (a == "--output") expect = OUTPATH;
            else if (a == "--stdout") inv.set_output(Invocation::OUT_STDOUT, "");
            else if (a == "--") expect = NOARGS;
            else die_args("Unrecognised argument " + a);
        } else if (expect == ANY && !a.empty() && a[0] == '-') {
            if (a == "-") {
                inv.set_input(Invocation::IN_STDIN, "");
            } else {
                for (size_t k = 1; k < a.size(); k++) {
                    char c = a[k];
                    if (c == 'c') inv.set_output(Invocation::OUT_STDOUT, "");
                    else if (c == 'k') inv.keep = 1;
                    else if (c == 'r') inv.keep = 0;
                    else if (c == 'v') inv.verbose = true;
                    else if (c >= '1' && c <= '9') inv.level = c - '0';
                    else die_args(std::string("Flag '") + c + "' is not valid");
                }
            }
        } else {
            inv.set_input(Invocation::IN_FILE, a);
        }
    }
    if (inv.in == Invocation::IN_NONE) die_args("An input must be specified");

    std::ifstream inf;
    std::istream *reader = &std::cin;
    if (inv.in == Invocation::IN_FILE) {
        inf.open(inv.in_path, std::ios::binary);
        if (!inf) die_fs("cannot open " + inv.in_path + ": " + std::strerror(errno));
        reader = &inf;
    }
    std::ofstream outf;
    std::ostream *writer = &std::cout;
    std::string out_path;
    if (inv.out == Invocation::OUT_FILE) out_path = inv.out_path;
    else if (inv.out == Invocation::OUT_NONE && inv.in == Invocation::IN_FILE) out_path = inv.in_path + ".bz2";
    if (!out_path.empty()) {
        outf.open(out_path, std::ios::binary | std::ios::trunc);
        if (!outf) die_fs("cannot create " + out_path + ": " + std::strerror(errno));
        writer = &outf;
    }

    try {
        banzai::Context ctx(0);
        banzai::encode(ctx, *reader, *writer, inv.level);
    } catch (const std::exception &e) {
        std::fprintf(stderr, "error during compression: %s\n", e.what());
        return ERR_OUTPUT;
    }
    if (outf.is_open()) {
        // the reference flushes its BufWriter and propagates the error before the input may be removed
        // (main.rs:287-309): a close that fails (quota, NFS) must not cost the only copy of the data
        outf.close();
        if (!outf) {
            std::fprintf(stderr, "error during compression: cannot write %s: %s\n", out_path.c_str(), std::strerror(errno));
            return ERR_OUTPUT;
        }
    } else {
        std::cout.flush();
        if (!std::cout) {
            std::fprintf(stderr, "error during compression: cannot write to stdout\n");
            return ERR_OUTPUT;
        }
    }

    const bool keep = inv.keep >= 0 ? inv.keep == 1 : inv.out != Invocation::OUT_NONE;   // main.rs:292-300
    if (!keep && inv.in == Invocation::IN_FILE) {
        if (std::remove(inv.in_path.c_str()) != 0) {
            std::fprintf(stderr, "error deleting input file: %s\n", std::strerror(errno));
            return ERR_OUTPUT;
        }
    }
    return SUCCESS;
}
