// pipeline.cu — host side of libbanzai_b200.so: context, device arenas, stage exports and the
// bnz_encode orchestration (the B200 replacement of the block loop in lib/lib.rs:101-126).
#include "../../include/banzai_b200.h"
#include "common.cuh"
#include "kernels.h"

#include <algorithm>
#include <string>
#include <vector>
#include <string.h>
#include <stdlib.h>

using namespace bnz;

// ---------------------------------------------------------------------------------------
// context
// ---------------------------------------------------------------------------------------

struct DevBuf {
    void *p = nullptr;
    size_t cap = 0;
    cudaError_t ensure(size_t bytes)
    {
        if (bytes <= cap) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
        size_t want = bytes + bytes / 8 + 256;
        cudaError_t e = cudaMalloc(&p, want);
        if (e != cudaSuccess) {
            e = cudaMalloc(&p, bytes);
            want = bytes;
        }
        if (e == cudaSuccess) cap = want;
        return e;
    }
    void release()
    {
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
    }
    template <class X> X *as() const { return reinterpret_cast<X *>(p); }
};

struct Device {
    int id = 0;
    int sm_count = 0;
    cudaStream_t stream = nullptr;
    // arenas (grown on demand, kept across calls)
    DevBuf in, rle, bwt, blk_off, blk_len, ptr, has_byte, bwt_stats, counters, ws_rec, ws_rank;
    uint32_t launches = 0;
};

struct bnz_ctx {
    std::vector<Device> devs;
    std::string err;
    bnz_stats stats;
    int radix_bits = 8;
    int ctas_per_sm = 0;
};

#define CK(ctx, call)                                                                          \
    do {                                                                                       \
        cudaError_t e__ = (call);                                                              \
        if (e__ != cudaSuccess) {                                                              \
            (ctx)->err = std::string(#call) + ": " + cudaGetErrorString(e__);                  \
            return (e__ == cudaErrorMemoryAllocation) ? BNZ_ENOMEM : BNZ_ECUDA;                \
        }                                                                                      \
    } while (0)

static int fail(bnz_ctx *ctx, int code, const std::string &msg)
{
    if (ctx) ctx->err = msg;
    return code;
}

extern "C" const char *bnz_strerror(int code)
{
    switch (code) {
    case BNZ_OK: return "ok";
    case BNZ_EINVAL: return "invalid argument (level must be 1..=9)";
    case BNZ_ECUDA: return "CUDA error or no usable sm_100a device";
    case BNZ_ENOMEM: return "out of memory";
    case BNZ_EINTERNAL: return "internal error";
    default: return "unknown error";
    }
}

extern "C" const char *bnz_last_error(const bnz_ctx *ctx) { return ctx ? ctx->err.c_str() : ""; }

extern "C" int bnz_ctx_create_on(bnz_ctx **out, const int *device_ids, int n_devices)
{
    if (!out || !device_ids || n_devices <= 0) return BNZ_EINVAL;
    *out = nullptr;
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count <= 0) return BNZ_ECUDA;
    bnz_ctx *ctx = new bnz_ctx();
    memset(&ctx->stats, 0, sizeof ctx->stats);
    for (int i = 0; i < n_devices; i++) {
        if (device_ids[i] < 0 || device_ids[i] >= count) {
            delete ctx;
            return BNZ_EINVAL;
        }
        Device d;
        d.id = device_ids[i];
        cudaDeviceProp prop;
        if (cudaSetDevice(d.id) != cudaSuccess || cudaGetDeviceProperties(&prop, d.id) != cudaSuccess ||
            cudaStreamCreateWithFlags(&d.stream, cudaStreamNonBlocking) != cudaSuccess) {
            delete ctx;
            return BNZ_ECUDA;
        }
        if (prop.major < 10) {      // kernels are built for sm_100a only; fail loudly
            delete ctx;
            return BNZ_ECUDA;
        }
        d.sm_count = prop.multiProcessorCount;
        ctx->devs.push_back(d);
    }
    *out = ctx;
    return BNZ_OK;
}

extern "C" int bnz_ctx_create(bnz_ctx **out, int n_gpus)
{
    if (!out || n_gpus < 0) return BNZ_EINVAL;
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count <= 0) return BNZ_ECUDA;
    if (n_gpus == 0 || n_gpus > count) n_gpus = count;
    std::vector<int> ids(n_gpus);
    for (int i = 0; i < n_gpus; i++) ids[i] = i;
    return bnz_ctx_create_on(out, ids.data(), n_gpus);
}

extern "C" void bnz_ctx_destroy(bnz_ctx *ctx)
{
    if (!ctx) return;
    for (Device &d : ctx->devs) {
        cudaSetDevice(d.id);
        if (d.stream) cudaStreamSynchronize(d.stream);
        for (DevBuf *b : { &d.in, &d.rle, &d.bwt, &d.blk_off, &d.blk_len, &d.ptr, &d.has_byte,
                           &d.bwt_stats, &d.counters, &d.ws_rec, &d.ws_rank })
            b->release();
        if (d.stream) cudaStreamDestroy(d.stream);
    }
    delete ctx;
}

extern "C" int bnz_ctx_set(bnz_ctx *ctx, const char *key, long value)
{
    if (!ctx || !key) return BNZ_EINVAL;
    if (!strcmp(key, "bwt_radix_bits")) {
        if (value != 8 && value != 10) return BNZ_EINVAL;
        ctx->radix_bits = (int)value;
        return BNZ_OK;
    }
    if (!strcmp(key, "bwt_ctas_per_sm")) {
        if (value < 0 || value > 8) return BNZ_EINVAL;
        ctx->ctas_per_sm = (int)value;
        return BNZ_OK;
    }
    return BNZ_EINVAL;
}

extern "C" int bnz_get_stats(const bnz_ctx *ctx, bnz_stats *out)
{
    if (!ctx || !out) return BNZ_EINVAL;
    *out = ctx->stats;
    return BNZ_OK;
}

extern "C" void *bnz_host_alloc(size_t bytes)
{
    void *p = nullptr;
    if (cudaHostAlloc(&p, bytes ? bytes : 1, cudaHostAllocPortable) != cudaSuccess) return nullptr;
    return p;
}
extern "C" void bnz_host_free(void *p)
{
    if (p) cudaFreeHost(p);
}
extern "C" void *bnz_device_alloc(bnz_ctx *ctx, size_t bytes)
{
    if (!ctx) return nullptr;
    void *p = nullptr;
    if (cudaSetDevice(ctx->devs[0].id) != cudaSuccess) return nullptr;
    if (cudaMalloc(&p, bytes ? bytes : 1) != cudaSuccess) return nullptr;
    return p;
}
extern "C" void bnz_device_free(bnz_ctx *ctx, void *p)
{
    if (!ctx || !p) return;
    cudaSetDevice(ctx->devs[0].id);
    cudaFree(p);
}
extern "C" int bnz_memcpy_h2d(bnz_ctx *ctx, void *d_dst, const void *h_src, size_t bytes)
{
    if (!ctx) return BNZ_EINVAL;
    Device &d = ctx->devs[0];
    CK(ctx, cudaSetDevice(d.id));
    CK(ctx, cudaMemcpyAsync(d_dst, h_src, bytes, cudaMemcpyHostToDevice, d.stream));
    CK(ctx, cudaStreamSynchronize(d.stream));
    return BNZ_OK;
}
extern "C" int bnz_memcpy_d2h(bnz_ctx *ctx, void *h_dst, const void *d_src, size_t bytes)
{
    if (!ctx) return BNZ_EINVAL;
    Device &d = ctx->devs[0];
    CK(ctx, cudaSetDevice(d.id));
    CK(ctx, cudaMemcpyAsync(h_dst, d_src, bytes, cudaMemcpyDeviceToHost, d.stream));
    CK(ctx, cudaStreamSynchronize(d.stream));
    return BNZ_OK;
}

extern "C" size_t bnz_max_compressed_size(size_t in_len)
{
    // worst case: RLE1 expands 4 -> 5, every MTF symbol costs <= 17 bits, plus per-block tables
    return in_len + in_len / 2 + (in_len / 80000 + 2) * 24576 + 4096;
}

// ---------------------------------------------------------------------------------------
// BWT stage on one device (device pointers in, device pointers out)
// ---------------------------------------------------------------------------------------

static int run_bwt_device(bnz_ctx *ctx, Device &d, const uint8_t *d_rle, uint8_t *d_bwt,
                          const uint64_t *d_blk_off, const uint32_t *d_blk_len, uint32_t n_blocks,
                          uint32_t max_len, uint32_t *d_ptr, uint8_t *d_has_byte, BwtStats *d_stats)
{
    if (n_blocks == 0) return BNZ_OK;
    int per_sm = 0;
    CK(ctx, bwt_max_ctas(ctx->radix_bits, &per_sm));
    if (per_sm <= 0) return fail(ctx, BNZ_ECUDA, "bwt kernel does not fit on an SM");
    if (ctx->ctas_per_sm > 0) per_sm = std::min(per_sm, ctx->ctas_per_sm);
    int grid = (int)std::min<uint64_t>((uint64_t)n_blocks, (uint64_t)d.sm_count * per_sm);
    size_t stride = ((size_t)max_len + 15) & ~(size_t)15;
    CK(ctx, d.ws_rec.ensure((size_t)grid * 2 * stride * sizeof(uint64_t)));
    CK(ctx, d.ws_rank.ensure((size_t)grid * stride * sizeof(uint32_t)));
    CK(ctx, d.counters.ensure(256));
    CK(ctx, cudaMemsetAsync(d.counters.p, 0, 256, d.stream));

    BwtArgs a;
    a.rle = d_rle;
    a.bwt = d_bwt;
    a.blk_off = d_blk_off;
    a.blk_len = d_blk_len;
    a.ptr = d_ptr;
    a.has_byte = d_has_byte;
    a.stats = d_stats;
    a.next_block = d.counters.as<uint32_t>();
    a.n_blocks = n_blocks;
    a.ws_rec = d.ws_rec.as<uint64_t>();
    a.ws_rank = d.ws_rank.as<uint32_t>();
    a.ws_stride = stride;
    CK(ctx, bwt_launch(a, ctx->radix_bits, grid, d.stream));
    d.launches++;
    return BNZ_OK;
}

extern "C" int bnz_stage_bwt(bnz_ctx *ctx, const uint8_t *blocks, const uint64_t *blk_off,
                             const uint32_t *blk_len, size_t n_blocks, int level, uint8_t *bwt_out,
                             uint32_t *ptr_out, uint8_t *has_byte_out,
                             bnz_bwt_block_stats *stats_out)
{
    if (!ctx || level < 1 || level > 9) return BNZ_EINVAL;
    if (n_blocks == 0) return BNZ_OK;
    if (!blocks || !blk_off || !blk_len || !bwt_out || !ptr_out || !has_byte_out) return BNZ_EINVAL;
    Device &d = ctx->devs[0];
    CK(ctx, cudaSetDevice(d.id));
    size_t total = 0;
    uint32_t max_len = 0;
    for (size_t b = 0; b < n_blocks; b++) {
        if (blk_len[b] == 0 || blk_len[b] > (uint32_t)(100000 * level)) return BNZ_EINVAL;
        total = std::max<size_t>(total, blk_off[b] + blk_len[b]);
        max_len = std::max(max_len, blk_len[b]);
    }
    CK(ctx, d.rle.ensure(total));
    CK(ctx, d.bwt.ensure(total));
    CK(ctx, d.blk_off.ensure(n_blocks * sizeof(uint64_t)));
    CK(ctx, d.blk_len.ensure(n_blocks * sizeof(uint32_t)));
    CK(ctx, d.ptr.ensure(n_blocks * sizeof(uint32_t)));
    CK(ctx, d.has_byte.ensure(n_blocks * 256));
    CK(ctx, d.bwt_stats.ensure(n_blocks * sizeof(BwtStats)));
    CK(ctx, cudaMemcpyAsync(d.rle.p, blocks, total, cudaMemcpyHostToDevice, d.stream));
    CK(ctx, cudaMemcpyAsync(d.blk_off.p, blk_off, n_blocks * sizeof(uint64_t), cudaMemcpyHostToDevice, d.stream));
    CK(ctx, cudaMemcpyAsync(d.blk_len.p, blk_len, n_blocks * sizeof(uint32_t), cudaMemcpyHostToDevice, d.stream));
    cudaEvent_t e0, e1;
    CK(ctx, cudaEventCreate(&e0));
    CK(ctx, cudaEventCreate(&e1));
    CK(ctx, cudaEventRecord(e0, d.stream));
    int rc = run_bwt_device(ctx, d, d.rle.as<uint8_t>(), d.bwt.as<uint8_t>(), d.blk_off.as<uint64_t>(),
                            d.blk_len.as<uint32_t>(), (uint32_t)n_blocks, max_len, d.ptr.as<uint32_t>(),
                            d.has_byte.as<uint8_t>(), d.bwt_stats.as<BwtStats>());
    if (rc != BNZ_OK) return rc;
    CK(ctx, cudaEventRecord(e1, d.stream));
    CK(ctx, cudaMemcpyAsync(bwt_out, d.bwt.p, total, cudaMemcpyDeviceToHost, d.stream));
    CK(ctx, cudaMemcpyAsync(ptr_out, d.ptr.p, n_blocks * sizeof(uint32_t), cudaMemcpyDeviceToHost, d.stream));
    CK(ctx, cudaMemcpyAsync(has_byte_out, d.has_byte.p, n_blocks * 256, cudaMemcpyDeviceToHost, d.stream));
    std::vector<BwtStats> st(n_blocks);
    CK(ctx, cudaMemcpyAsync(st.data(), d.bwt_stats.p, n_blocks * sizeof(BwtStats), cudaMemcpyDeviceToHost, d.stream));
    CK(ctx, cudaStreamSynchronize(d.stream));
    float ms = 0;
    CK(ctx, cudaEventElapsedTime(&ms, e0, e1));
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);

    bnz_stats &s = ctx->stats;
    memset(&s, 0, sizeof s);
    s.n_blocks = (uint32_t)n_blocks;
    s.n_devices = 1;
    s.kernel_launches = 1;
    s.bwt_radix_bits = (uint32_t)ctx->radix_bits;
    s.bwt_ms = ms;
    for (size_t b = 0; b < n_blocks; b++) {
        s.bwt_n += st[b].n;
        s.bwt_sum_active += st[b].sum_active;
        s.bwt_sum_active_passes += st[b].sum_active_passes;
        s.bwt_rounds_total += st[b].rounds;
        s.bwt_max_rounds = std::max(s.bwt_max_rounds, st[b].rounds);
        s.bwt_tied_blocks += st[b].tied;
        if (stats_out) {
            stats_out[b].n = st[b].n;
            stats_out[b].rounds = st[b].rounds;
            stats_out[b].tied = st[b].tied;
            stats_out[b].pad = 0;
            stats_out[b].sum_active = st[b].sum_active;
            stats_out[b].sum_active_passes = st[b].sum_active_passes;
        }
    }
    s.bwt_algorithmic_bytes = 9 * s.bwt_n + 16 * s.bwt_sum_active_passes + 36 * s.bwt_sum_active;
    return BNZ_OK;
}

// ---------------------------------------------------------------------------------------
// not yet implemented entry points (filled in by later milestones)
// ---------------------------------------------------------------------------------------

extern "C" int bnz_encode(bnz_ctx *ctx, const uint8_t *, size_t, int, uint8_t **, size_t *, size_t *)
{
    return fail(ctx, BNZ_EINTERNAL, "bnz_encode: not implemented yet");
}
extern "C" void bnz_free(bnz_ctx *, uint8_t *p) { free(p); }
extern "C" int bnz_encode_device(bnz_ctx *ctx, const void *, size_t, int, void *, size_t, size_t *)
{
    return fail(ctx, BNZ_EINTERNAL, "bnz_encode_device: not implemented yet");
}
extern "C" int bnz_encode_file(bnz_ctx *ctx, const char *, const char *, size_t *)
{
    return fail(ctx, BNZ_EINTERNAL, "bnz_encode_file: not implemented yet");
}
extern "C" int bnz_stage_rle1(bnz_ctx *ctx, const uint8_t *, size_t, int, uint64_t *, uint64_t *, uint64_t *,
                              uint32_t *, uint32_t *, size_t, uint8_t *, size_t, size_t *)
{
    return fail(ctx, BNZ_EINTERNAL, "bnz_stage_rle1: not implemented yet");
}
extern "C" int bnz_stage_mtf(bnz_ctx *ctx, const uint8_t *, const uint64_t *, const uint32_t *, const uint8_t *,
                             size_t, uint16_t *, uint32_t *, uint32_t *, uint32_t *)
{
    return fail(ctx, BNZ_EINTERNAL, "bnz_stage_mtf: not implemented yet");
}
extern "C" int bnz_stage_huffman(bnz_ctx *ctx, const uint16_t *, const uint64_t *, const uint32_t *,
                                 const uint32_t *, const uint32_t *, size_t, uint8_t *, size_t, uint64_t *,
                                 uint8_t *, uint32_t *)
{
    return fail(ctx, BNZ_EINTERNAL, "bnz_stage_huffman: not implemented yet");
}
