// banzai.hpp — C++ host side over the C ABI (include/banzai_b200.h), mirroring the reference's
// public surface so callers of the Rust crate find the same two functions:
//
//   banzai::encode(reader, writer, level) -> bytes consumed     reference lib/lib.rs:84-132
//   banzai::encode_file(in_path, out_path) -> bytes consumed    reference lib/lib.rs:141-153
//
// Errors: the reference returns io::Result and panics on level outside 1..=9 (lib/lib.rs:89).
// Here I/O and device errors throw std::runtime_error, a bad level throws std::invalid_argument.
#pragma once
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
inline size_t encode(Context &ctx, std::istream &reader, std::ostream &writer, int level)
{
    if (level < 1 || level > 9) throw std::invalid_argument("level must be in 1..=9");
    std::vector<uint8_t> data((std::istreambuf_iterator<char>(reader)), std::istreambuf_iterator<char>());
    if (reader.bad()) throw std::runtime_error("read error");
    std::vector<uint8_t> out = encode_bytes(ctx, data.data(), data.size(), level);
    writer.write(reinterpret_cast<const char *>(out.data()), (std::streamsize)out.size());
    writer.flush();                                     // out.rs:22-28 close() flushes
    if (!writer) throw std::runtime_error("write error");
    return data.size();
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
