// banzai.hpp — C++ host side over the C ABI (include/banzai_b200.h), mirroring the reference's
// public surface so callers of the Rust crate find the same two functions:
//
//   banzai::encode(reader, writer, level) -> bytes consumed     reference lib/lib.rs:84-132
//   banzai::encode_file(in_path, out_path) -> bytes consumed    reference lib/lib.rs:141-153
//
// Errors: the reference returns io::Result and panics on level outside 1..=9 (lib/lib.rs:89).
// Here I/O and device errors throw std::runtime_error, a bad level throws std::invalid_argument.
#pragma once
#include <algorithm>
#include <cstdint>
#include <fstream>
#include <istream>
#include <iterator>
#include <ostream>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/banzai_b200.h"

namespace banzai {

class Context {
  public:
    explicit Context(int n_gpus = 1)
    {
        int rc = bnz_ctx_create(&ctx_, n_gpus);
        if (rc != BNZ_OK) throw std::runtime_error(std::string("banzai_b200: ") + bnz_strerror(rc));
    }
    ~Context() { bnz_ctx_destroy(ctx_); }
    Context(const Context &) = delete;
    Context &operator=(const Context &) = delete;
    bnz_ctx *get() const { return ctx_; }

  private:
    bnz_ctx *ctx_ = nullptr;
};

// whole-buffer encode; returns the finished .bz2 stream
inline std::vector<uint8_t> encode_bytes(Context &ctx, const uint8_t *data, size_t len, int level)
{
    if (level < 1 || level > 9) throw std::invalid_argument("level must be in 1..=9");
    uint8_t *out = nullptr;
    size_t out_len = 0, consumed = 0;
    int rc = bnz_encode(ctx.get(), data, len, level, &out, &out_len, &consumed);
    if (rc != BNZ_OK)
        throw std::runtime_error(std::string(bnz_strerror(rc)) + ": " + bnz_last_error(ctx.get()));
    std::vector<uint8_t> res(out, out + out_len);
    bnz_free(ctx.get(), out);
    return res;
}

// banzai::encode(reader, BufWriter, level) -> usize   (lib/lib.rs:84)
// Streams: the reader fills pinned windows in place while the previous window is on the GPU, and
// finished stream bytes go to the writer as they appear (bnz_stream_*), so pipes and inputs
// larger than memory work like they do with the reference's refill loop (lib/rle.rs:43-91).
inline size_t encode(Context &ctx, std::istream &reader, std::ostream &writer, int level)
{
    if (level < 1 || level > 9) throw std::invalid_argument("level must be in 1..=9");
    auto err = [&](int rc) {
        return std::runtime_error(std::string(bnz_strerror(rc)) + ": " + bnz_last_error(ctx.get()));
    };
    auto sink = [](void *user, const uint8_t *data, size_t len) -> int {
        std::ostream *w = static_cast<std::ostream *>(user);
        w->write(reinterpret_cast<const char *>(data), (std::streamsize)len);
        return w->good() ? 0 : 1;
    };
    bnz_stream *s = nullptr;
    int rc = bnz_stream_open(ctx.get(), level, sink, &writer, &s);
    if (rc != BNZ_OK) throw err(rc);
    struct Closer {
        bnz_stream *s;
        ~Closer() { bnz_stream_close(s); }
    } closer{ s };
    for (;;) {
        uint8_t *buf = nullptr;
        size_t cap = 0;
        rc = bnz_stream_reserve(s, &buf, &cap);
        if (rc != BNZ_OK) throw err(rc);
        reader.read(reinterpret_cast<char *>(buf), (std::streamsize)std::min<size_t>(cap, (size_t)8 << 20));
        const size_t got = (size_t)reader.gcount();
        if (reader.bad()) throw std::runtime_error("read error");
        if (got) {
            rc = bnz_stream_commit(s, got);
            if (rc != BNZ_OK) throw err(rc);
        }
        if (reader.eof() || got == 0) break;
    }
    size_t consumed = 0;
    rc = bnz_stream_finish(s, &consumed);
    if (rc != BNZ_OK) throw err(rc);
    writer.flush();                                     // out.rs:22-28 close() flushes
    if (!writer) throw std::runtime_error("write error");
    return consumed;
}

// banzai::encode_file(in_path, out_path) -> usize, level 9   (lib/lib.rs:141-153)
inline size_t encode_file(Context &ctx, const std::string &in_path, const std::string &out_path)
{
    std::ifstream inf(in_path, std::ios::binary);
    if (!inf) throw std::runtime_error("cannot open " + in_path);
    std::ofstream outf(out_path, std::ios::binary | std::ios::trunc);
    if (!outf) throw std::runtime_error("cannot create " + out_path);
    return encode(ctx, inf, outf, 9);
}

}  // namespace banzai
