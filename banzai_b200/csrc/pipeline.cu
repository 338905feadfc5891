// pipeline.cu — host side of libbanzai_b200.so: context, device arenas, stage exports and the
// bnz_encode orchestration (the B200 replacement of the block loop in lib/lib.rs:101-126).
#include "../../include/banzai_b200.h"
#include "common.cuh"
#include "kernels.h"

#include <algorithm>
#include <atomic>
#include <condition_variable>
#include <deque>
#include <mutex>
#include <string>
#include <thread>
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

struct PinBuf {                      // grow-only pinned host staging
    void *p = nullptr;
    size_t cap = 0;
    cudaError_t ensure(size_t bytes)
    {
        if (bytes <= cap) return cudaSuccess;
        if (p) cudaFreeHost(p);
        p = nullptr;
        cap = 0;
        size_t want = bytes + bytes / 4 + 4096;
        cudaError_t e = cudaHostAlloc(&p, want, cudaHostAllocPortable | cudaHostAllocMapped);
        if (e == cudaSuccess) cap = want;
        return e;
    }
    void release()
    {
        if (p) cudaFreeHost(p);
        p = nullptr;
        cap = 0;
    }
    template <class X> X *as() const { return reinterpret_cast<X *>(p); }
};

struct Device {
    int id = 0;
    int sm_count = 0;
    cudaStream_t stream = nullptr;
    cudaStream_t stream2 = nullptr;     // side stream: block CRCs run beside the sort
    cudaStream_t stream3[3] = {};       // low-priority streams: MTF of finished blocks fills the sort's tail
    // arenas (grown on demand, kept across calls)
    DevBuf in, rle, bwt, blk_off, blk_len, ptr, has_byte, bwt_stats, counters, ws_rec, ws_rank, ws_ctl, ws_hist, bwt_score, bwt_order;
    DevBuf ch_lasthead, ch_meta, ch_restsum, ch_oin, ch_P, ch_tiles, rle_blocks, crc_acc;
    DevBuf seg_base, seg_list, seg_cnt, seg_state, num_names, syms, sym_off, sym_len, freqs, mtf_ids, mtf_cseg;
    DevBuf lens, codes, tf, num_tables, num_sel, span_base, hdr, hdr_bits, crc, blk_bits, blk_bitoff,
        total_bits, out;
    cudaEvent_t ev[16] = {};
    PinBuf h_P, h_oin, h_acc, h_mtf, h_done;   // h_done: per-block completion flags the sort writes (mapped)
    bool crc_tables = false;
    uint32_t launches = 0;
};

struct bnz_ctx {
    std::vector<Device> devs;
    std::string err;
    bnz_stats stats;
    int radix_bits = 8;
    int ctas_per_sm = 0;
    int bwt_cluster = -1;         // CTAs per bzip2 block (-1: auto, 0/1: single-CTA kernel)
    int bwt_threads = 512;
    int bwt_lpt = 0;                   // longest-predicted-first work queue (measured: no robust gain, off)
    std::vector<uint32_t> last_scores; // predictor output of the last sort (debug / tests)
    int bwt_cluster_below = 400;       // auto mode: cluster kernel when a device gets fewer blocks than this
    // cached pinned output buffer handed to the caller by bnz_encode / returned by bnz_free
    uint8_t *out_cache = nullptr;
    size_t out_cache_cap = 0;
    bool out_cache_lent = false;
    uint8_t *out_big = nullptr;         // streaming-batch output (ordinary host memory, kept across calls)
    size_t out_big_cap = 0;
    bool out_big_lent = false;
    size_t max_batch_bytes = (size_t)3 << 30;   // inputs above this are encoded in streaming batches
    size_t stream_window_bytes = (size_t)512 << 20;   // bnz_stream_*: input bytes per pipeline window
    int open_streams = 0;
    int crc_low_prio = 1;              // block CRCs on the low-priority stream (they would delay the start of the sort)
    int mtf_groups = 2;
    int mtf_overlap = 70;              // percent of a device's blocks whose MTF may run beside the sort (0: off)
};

// worker threads of a multi-device encode record their error text in their own string
static thread_local std::string *t_err_sink = nullptr;
static void set_err(bnz_ctx *ctx, const std::string &msg)
{
    if (t_err_sink) *t_err_sink = msg;
    else if (ctx) ctx->err = msg;
}

#define CK(ctx, call)                                                                          \
    do {                                                                                       \
        cudaError_t e__ = (call);                                                              \
        if (e__ != cudaSuccess) {                                                              \
            set_err((ctx), std::string(#call) + ": " + cudaGetErrorString(e__));               \
            return (e__ == cudaErrorMemoryAllocation) ? BNZ_ENOMEM : BNZ_ECUDA;                \
        }                                                                                      \
    } while (0)

static int fail(bnz_ctx *ctx, int code, const std::string &msg)
{
    set_err(ctx, msg);
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
    case BNZ_EIO: return "I/O error (sink or file)";
    default: return "unknown error";
    }
}

extern "C" const char *bnz_last_error(const bnz_ctx *ctx) { return ctx ? ctx->err.c_str() : ""; }

extern "C" void bnz_ctx_destroy(bnz_ctx *ctx);

extern "C" int bnz_ctx_create_on(bnz_ctx **out, const int *device_ids, int n_devices)
{
    if (!out || !device_ids || n_devices <= 0) return BNZ_EINVAL;
    *out = nullptr;
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count <= 0) return BNZ_ECUDA;
    for (int i = 0; i < n_devices; i++)
        if (device_ids[i] < 0 || device_ids[i] >= count) return BNZ_EINVAL;
    bnz_ctx *ctx = new bnz_ctx();
    memset(&ctx->stats, 0, sizeof ctx->stats);
    ctx->devs.reserve(n_devices);
    for (int i = 0; i < n_devices; i++) {
        // a device id may repeat: every entry is an independent lane (own streams and arenas)
        ctx->devs.emplace_back();
        Device &d = ctx->devs.back();
        d.id = device_ids[i];
        cudaDeviceProp prop;
        int prio_lo = 0, prio_hi = 0;
        bool ok = cudaSetDevice(d.id) == cudaSuccess && cudaGetDeviceProperties(&prop, d.id) == cudaSuccess &&
                  prop.major >= 10 &&        // kernels are built for sm_100a only; fail loudly
                  cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi) == cudaSuccess &&
                  // the sort's persistent CTAs must win every SM slot over the work that fills its tail
                  cudaStreamCreateWithPriority(&d.stream, cudaStreamNonBlocking, prio_hi) == cudaSuccess &&
                  cudaStreamCreateWithPriority(&d.stream2, cudaStreamNonBlocking, prio_hi) == cudaSuccess &&
                  cudaStreamCreateWithPriority(&d.stream3[0], cudaStreamNonBlocking, prio_lo) == cudaSuccess &&
                  cudaStreamCreateWithPriority(&d.stream3[1], cudaStreamNonBlocking, prio_lo) == cudaSuccess &&
                  cudaStreamCreateWithPriority(&d.stream3[2], cudaStreamNonBlocking, prio_lo) == cudaSuccess;
        for (cudaEvent_t &e : d.ev)
            if (ok && cudaEventCreate(&e) != cudaSuccess) ok = false;
        if (!ok) {
            bnz_ctx_destroy(ctx);
            return BNZ_ECUDA;
        }
        d.sm_count = prop.multiProcessorCount;
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
                           &d.bwt_stats, &d.counters, &d.ws_rec, &d.ws_rank, &d.ws_ctl, &d.ws_hist, &d.bwt_score, &d.bwt_order, &d.ch_lasthead, &d.ch_meta,
                           &d.ch_restsum, &d.ch_oin, &d.ch_P, &d.ch_tiles, &d.rle_blocks, &d.crc_acc, &d.seg_base,
                           &d.seg_list, &d.seg_cnt, &d.seg_state, &d.num_names, &d.syms, &d.sym_off,
                           &d.sym_len, &d.freqs, &d.mtf_ids, &d.mtf_cseg, &d.lens, &d.codes, &d.tf, &d.num_tables, &d.num_sel,
                           &d.span_base, &d.hdr, &d.hdr_bits, &d.crc, &d.blk_bits, &d.blk_bitoff,
                           &d.total_bits, &d.out })
            b->release();
        for (cudaEvent_t e : d.ev)
            if (e) cudaEventDestroy(e);
        d.h_P.release();
        d.h_oin.release();
        d.h_acc.release();
        d.h_done.release();
        d.h_mtf.release();
        if (d.stream) cudaStreamDestroy(d.stream);
        if (d.stream2) cudaStreamDestroy(d.stream2);
        for (cudaStream_t st : d.stream3)
            if (st) cudaStreamDestroy(st);
    }
    if (ctx->out_cache) cudaFreeHost(ctx->out_cache);
    free(ctx->out_big);
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
    if (!strcmp(key, "bwt_cluster")) {
        if (value < -1 || value > BWT_CLUSTER_MAX) return BNZ_EINVAL;
        ctx->bwt_cluster = (int)value;
        return BNZ_OK;
    }
    if (!strcmp(key, "max_batch_bytes")) {
        if (value < (1 << 20)) return BNZ_EINVAL;
        ctx->max_batch_bytes = (size_t)value;
        return BNZ_OK;
    }
    if (!strcmp(key, "crc_low_prio")) {
        ctx->crc_low_prio = value != 0;
        return BNZ_OK;
    }
    if (!strcmp(key, "mtf_groups")) {
        if (value < 1 || value > 16) return BNZ_EINVAL;
        ctx->mtf_groups = (int)value;
        return BNZ_OK;
    }
    if (!strcmp(key, "mtf_overlap")) {
        if (value < 0 || value > 95) return BNZ_EINVAL;
        ctx->mtf_overlap = (int)value;
        return BNZ_OK;
    }
    if (!strcmp(key, "stream_window_bytes")) {
        if (value < (1 << 16)) return BNZ_EINVAL;
        ctx->stream_window_bytes = (size_t)value;
        return BNZ_OK;
    }
    if (!strcmp(key, "bwt_lpt")) {
        if (value < 0 || value > 2) return BNZ_EINVAL;
        ctx->bwt_lpt = (int)value;       // 0 off, 1 longest first, 2 light blocks last
        return BNZ_OK;
    }
    if (!strcmp(key, "bwt_cluster_below")) {
        if (value < 0) return BNZ_EINVAL;
        ctx->bwt_cluster_below = (int)value;
        return BNZ_OK;
    }
    if (!strcmp(key, "bwt_threads")) {
        if (value != 512 && value != 1024) return BNZ_EINVAL;
        ctx->bwt_threads = (int)value;
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
                          uint32_t max_len, uint32_t *d_ptr, uint8_t *d_has_byte, BwtStats *d_stats,
                          uint32_t *d_done = nullptr, bool *done_armed = nullptr)
{
    if (done_armed) *done_armed = false;
    if (n_blocks == 0) return BNZ_OK;
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
    a.ws_ctl = nullptr;
    a.ws_hist = nullptr;
    a.order = nullptr;
    a.done = nullptr;

    // auto: many blocks -> one persistent CTA per block (best aggregate throughput);
    // few blocks -> one cluster per block so that every SM has work and the randomly accessed
    // arrays stay in L2 (measured crossover ~400 blocks per device, tools/bwt_blocks_sweep.py)
    int C = ctx->bwt_cluster;
    if (C < 0) C = (n_blocks >= (uint32_t)ctx->bwt_cluster_below) ? 0 : (n_blocks <= 40 ? 16 : 8);
    if (C > 1) {
        int max_clusters = 0;
        CK(ctx, bwtc_max_clusters(ctx->bwt_threads, C, &max_clusters));
        if (max_clusters <= 0) return fail(ctx, BNZ_ECUDA, "bwt cluster shape cannot be scheduled");
        if (ctx->ctas_per_sm > 0) max_clusters = std::min(max_clusters, ctx->ctas_per_sm * d.sm_count / C);
        int n_clusters = (int)std::min<uint64_t>((uint64_t)n_blocks, (uint64_t)std::max(1, max_clusters));
        size_t stride = (((size_t)max_len + 15) & ~(size_t)15) + (size_t)BWT_CLUSTER_MAX * 4096;
        CK(ctx, d.ws_rec.ensure((size_t)n_clusters * 2 * stride * sizeof(uint64_t)));
        CK(ctx, d.ws_rank.ensure((size_t)n_clusters * stride * sizeof(uint32_t)));
        CK(ctx, d.ws_ctl.ensure((size_t)n_clusters * BWT_CTL_BYTES));
        CK(ctx, cudaMemsetAsync(d_has_byte, 0, (size_t)n_blocks * 256, d.stream));
        a.ws_rec = d.ws_rec.as<uint64_t>();
        a.ws_rank = d.ws_rank.as<uint32_t>();
        a.ws_stride = stride;
        a.ws_ctl = d.ws_ctl.p;
        CK(ctx, bwtc_launch(a, ctx->bwt_threads, C, n_clusters, d.stream));
        d.launches++;
        return BNZ_OK;
    }

    int per_sm = 0;
    CK(ctx, bwt_max_ctas(ctx->radix_bits, &per_sm));
    if (per_sm <= 0) return fail(ctx, BNZ_ECUDA, "bwt kernel does not fit on an SM");
    if (ctx->ctas_per_sm > 0) per_sm = std::min(per_sm, ctx->ctas_per_sm);
    int grid = (int)std::min<uint64_t>((uint64_t)n_blocks, (uint64_t)d.sm_count * per_sm);
    if (ctx->bwt_lpt && n_blocks > (uint32_t)grid) {
        // blocks differ several-fold in sort cost (doubling rounds); with a plain index-order queue
        // the last wave's long blocks leave most SMs idle.  Predict, then schedule longest first.
        CK(ctx, d.bwt_score.ensure((size_t)n_blocks * 4));
        CK(ctx, d.bwt_order.ensure((size_t)n_blocks * 4));
        CK(ctx, bwt_predict_launch(d_rle, d_blk_off, d_blk_len, n_blocks, d.bwt_score.as<uint32_t>(), d.stream));
        d.launches++;
        std::vector<uint32_t> score(n_blocks), order(n_blocks);
        CK(ctx, cudaMemcpyAsync(score.data(), d.bwt_score.p, (size_t)n_blocks * 4, cudaMemcpyDeviceToHost, d.stream));
        CK(ctx, cudaStreamSynchronize(d.stream));
        ctx->last_scores = score;
        if (ctx->bwt_lpt == 1) {
            // full longest-first order
            for (uint32_t i = 0; i < n_blocks; i++) order[i] = i;
            std::stable_sort(order.begin(), order.end(), [&](uint32_t x, uint32_t y) { return score[x] > score[y]; });
        } else {
            // "light tail": keep the natural (type-interleaved) order, but move the `grid` cheapest
            // blocks to the end of the queue so that the last wave consists of short blocks
            std::vector<uint32_t> by(n_blocks);
            for (uint32_t i = 0; i < n_blocks; i++) by[i] = i;
            std::stable_sort(by.begin(), by.end(), [&](uint32_t x, uint32_t y) { return score[x] < score[y]; });
            std::vector<uint8_t> tail(n_blocks, 0);
            for (int i = 0; i < grid; i++) tail[by[i]] = 1;
            uint32_t k = 0;
            for (uint32_t i = 0; i < n_blocks; i++) if (!tail[i]) order[k++] = i;
            for (uint32_t i = 0; i < n_blocks; i++) if (tail[i]) order[k++] = i;
        }
        CK(ctx, cudaMemcpyAsync(d.bwt_order.p, order.data(), (size_t)n_blocks * 4, cudaMemcpyHostToDevice, d.stream));
        CK(ctx, cudaStreamSynchronize(d.stream));      // `order` is a stack vector
        a.order = d.bwt_order.as<uint32_t>();
    }
    size_t stride = ((size_t)max_len + 15) & ~(size_t)15;
    CK(ctx, d.ws_rec.ensure((size_t)grid * 2 * stride * sizeof(uint64_t)));
    CK(ctx, d.ws_rank.ensure((size_t)grid * stride * sizeof(uint32_t)));
    CK(ctx, d.ws_hist.ensure((size_t)grid * BWT_HIST_WORDS * 4));
    a.ws_hist = d.ws_hist.as<uint32_t>();
    a.ws_rec = d.ws_rec.as<uint64_t>();
    a.ws_rank = d.ws_rank.as<uint32_t>();
    a.ws_stride = stride;
    if (d_done && n_blocks > (uint32_t)grid) {          // per-block completion flags (the queue has a tail)
        a.done = d_done;
        if (done_armed) *done_armed = true;
    }
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
    // device layout: every block image 16-byte aligned (what the pipeline guarantees the kernels)
    std::vector<uint64_t> doff(n_blocks);
    size_t total = 0;
    uint32_t max_len = 0;
    for (size_t b = 0; b < n_blocks; b++) {
        if (blk_len[b] == 0 || blk_len[b] > (uint32_t)(100000 * level)) return BNZ_EINVAL;
        doff[b] = total;
        total += ((size_t)blk_len[b] + 15) & ~(size_t)15;
        max_len = std::max(max_len, blk_len[b]);
    }
    CK(ctx, d.rle.ensure(total + 16));
    CK(ctx, d.bwt.ensure(total + 16));
    CK(ctx, d.blk_off.ensure(n_blocks * sizeof(uint64_t)));
    CK(ctx, d.blk_len.ensure(n_blocks * sizeof(uint32_t)));
    CK(ctx, d.ptr.ensure(n_blocks * sizeof(uint32_t)));
    CK(ctx, d.has_byte.ensure(n_blocks * 256));
    CK(ctx, d.bwt_stats.ensure(n_blocks * sizeof(BwtStats)));
    CK(ctx, cudaMemsetAsync(d.rle.p, 0, total + 16, d.stream));
    for (size_t b = 0; b < n_blocks; b++)
        CK(ctx, cudaMemcpyAsync(d.rle.as<uint8_t>() + doff[b], blocks + blk_off[b], blk_len[b], cudaMemcpyHostToDevice, d.stream));
    CK(ctx, cudaMemcpyAsync(d.blk_off.p, doff.data(), n_blocks * sizeof(uint64_t), cudaMemcpyHostToDevice, d.stream));
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
    for (size_t b = 0; b < n_blocks; b++)
        CK(ctx, cudaMemcpyAsync(bwt_out + blk_off[b], d.bwt.as<uint8_t>() + doff[b], blk_len[b], cudaMemcpyDeviceToHost, d.stream));
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
        s.bwt_cyc_build += st[b].cyc_build;
        s.bwt_cyc_radix += st[b].cyc_radix;
        s.bwt_cyc_rerank += st[b].cyc_rerank;
        if (stats_out) {
            stats_out[b].n = st[b].n;
            stats_out[b].rounds = st[b].rounds;
            stats_out[b].tied = st[b].tied;
            stats_out[b].pad = b < ctx->last_scores.size() ? ctx->last_scores[b] : 0;
            stats_out[b].sum_active = st[b].sum_active;
            stats_out[b].sum_active_passes = st[b].sum_active_passes;
            stats_out[b].cycles = st[b].cyc_build + st[b].cyc_radix + st[b].cyc_rerank;
        }
    }
    s.bwt_algorithmic_bytes = 9 * s.bwt_n + 16 * s.bwt_sum_active_passes + 36 * s.bwt_sum_active;
    return BNZ_OK;
}

// ---------------------------------------------------------------------------------------
// RLE1 + cuts + CRC on one device.  d_in: device copy of the input, h_in: host copy (the cut
// walk reads <= 2 KiB of it per block).  Leaves the RLE1 images in d.rle and fills `blocks`
// and `crcs`.
// ---------------------------------------------------------------------------------------

// K1 part 1 on one device: chunk tables -> host -> cut walk.  Leaves P / o_in in d.h_P / d.h_oin
// (host, pinned) and in d.ch_P / d.ch_oin (device).
static int rle_plan(bnz_ctx *ctx, Device &d, const uint8_t *d_in, const uint8_t *h_in, uint64_t N, int level,
                    std::vector<RleBlock> &blocks, bool final = true, uint64_t *consumed = nullptr)
{
    blocks.clear();
    if (consumed) *consumed = 0;
    if (N == 0) return BNZ_OK;
    const uint64_t n_chunks = (N + RLE_CHUNK - 1) / RLE_CHUNK;
    CK(ctx, d.ch_lasthead.ensure(n_chunks * 8));
    CK(ctx, d.ch_meta.ensure(n_chunks * 4));
    CK(ctx, d.ch_restsum.ensure(n_chunks * 4));
    CK(ctx, d.ch_oin.ensure(n_chunks * 8));
    CK(ctx, d.ch_P.ensure((n_chunks + 1) * 8));
    CK(ctx, d.ch_tiles.ensure(rle_scan_tiles(n_chunks) * 16 + 64));
    CK(ctx, rle_summary_launch(d_in, N, n_chunks, d.ch_lasthead.as<uint64_t>(), d.ch_meta.as<uint32_t>(),
                               d.ch_restsum.as<uint32_t>(), d.ch_oin.as<uint64_t>(), d.ch_P.as<uint64_t>(),
                               d.ch_tiles.as<uint64_t>(), d.stream));
    d.launches += 4;
    CK(ctx, d.h_P.ensure((n_chunks + 1) * 8));
    CK(ctx, d.h_oin.ensure(n_chunks * 8));
    CK(ctx, cudaMemcpyAsync(d.h_P.p, d.ch_P.p, (n_chunks + 1) * 8, cudaMemcpyDeviceToHost, d.stream));
    CK(ctx, cudaMemcpyAsync(d.h_oin.p, d.ch_oin.p, n_chunks * 8, cudaMemcpyDeviceToHost, d.stream));
    CK(ctx, cudaStreamSynchronize(d.stream));
    if (rle_walk_cuts(h_in, N, level, d.h_P.as<uint64_t>(), d.h_oin.as<uint64_t>(), n_chunks, blocks, final, consumed) != 0)
        return fail(ctx, BNZ_EINTERNAL, "RLE1 cut walk failed");
    return BNZ_OK;
}

// K1 part 2 + K2 on one device for a contiguous range of blocks.  `blocks` holds the range with
// rle_off rebased to 0; in_base / oin_base / P_base are device pointers indexed by GLOBAL input
// position / chunk (rebased by the caller when only a sub-range is resident).
static int rle_emit_shard(bnz_ctx *ctx, Device &d, const uint8_t *in_base, uint64_t N, const uint64_t *oin_base,
                          const uint64_t *P_base, const std::vector<RleBlock> &blocks, std::vector<uint32_t> *crcs,
                          uint64_t *rle_total)
{
    if (crcs) crcs->clear();
    *rle_total = 0;
    const size_t nb = blocks.size();
    if (nb == 0) return BNZ_OK;
    if (!d.crc_tables) {
        CK(ctx, crc_upload_tables());
        d.crc_tables = true;
    }
    const uint64_t total = blocks.back().rle_off + ((blocks.back().n + 15) & ~15ull);
    *rle_total = total;
    const uint64_t c_begin = blocks.front().s / RLE_CHUNK;
    const uint64_t c_end = (blocks.back().c + RLE_CHUNK - 1) / RLE_CHUNK;
    CK(ctx, d.rle_blocks.ensure(nb * sizeof(RleBlock)));
    CK(ctx, d.crc_acc.ensure(nb * 4));
    CK(ctx, d.crc.ensure(nb * 4));
    CK(ctx, d.rle.ensure(total));
    CK(ctx, cudaMemcpyAsync(d.rle_blocks.p, blocks.data(), nb * sizeof(RleBlock), cudaMemcpyHostToDevice, d.stream));
    CK(ctx, cudaMemsetAsync(d.crc_acc.p, 0, nb * 4, d.stream));
    CK(ctx, cudaEventRecord(d.ev[9], d.stream));
    CK(ctx, rle_emit_launch(in_base, N, c_begin, c_end, oin_base, P_base, d.rle_blocks.as<RleBlock>(), (uint32_t)nb,
                            d.rle.as<uint8_t>(), d.stream));
    // K2 on the side stream; d.ev[10] marks d.crc complete
    cudaStream_t crc_st = ctx->crc_low_prio ? d.stream3[2] : d.stream2;
    CK(ctx, cudaStreamWaitEvent(crc_st, d.ev[9], 0));
    CK(ctx, crc_launch(in_base, N, c_begin, c_end, d.rle_blocks.as<RleBlock>(), (uint32_t)nb, d.crc_acc.as<uint32_t>(),
                       d.crc.as<uint32_t>(), crc_st));
    CK(ctx, cudaEventRecord(d.ev[10], crc_st));
    d.launches += 3;
    if (crcs) {
        CK(ctx, d.h_acc.ensure(nb * 4));
        CK(ctx, cudaStreamWaitEvent(d.stream, d.ev[10], 0));
        CK(ctx, cudaMemcpyAsync(d.h_acc.p, d.crc.p, nb * 4, cudaMemcpyDeviceToHost, d.stream));
        CK(ctx, cudaStreamSynchronize(d.stream));
        crcs->assign(d.h_acc.as<uint32_t>(), d.h_acc.as<uint32_t>() + nb);
    }
    return BNZ_OK;
}

static int run_rle_device(bnz_ctx *ctx, Device &d, const uint8_t *d_in, const uint8_t *h_in, uint64_t N,
                          int level, std::vector<RleBlock> &blocks, std::vector<uint32_t> &crcs,
                          uint64_t *rle_total)
{
    crcs.clear();
    *rle_total = 0;
    int rc = rle_plan(ctx, d, d_in, h_in, N, level, blocks);
    if (rc != BNZ_OK || blocks.empty()) return rc;
    return rle_emit_shard(ctx, d, d_in, N, d.ch_oin.as<uint64_t>(), d.ch_P.as<uint64_t>(), blocks, &crcs, rle_total);
}

extern "C" int bnz_stage_rle1(bnz_ctx *ctx, const uint8_t *in, size_t in_len, int level, uint64_t *blk_in_off,
                              uint64_t *blk_in_len, uint64_t *blk_rle_off, uint32_t *blk_rle_len,
                              uint32_t *blk_crc, size_t max_blocks, uint8_t *rle_out, size_t rle_cap,
                              size_t *n_blocks)
{
    if (!ctx || level < 1 || level > 9 || !n_blocks) return BNZ_EINVAL;
    *n_blocks = 0;
    if (in_len == 0) return BNZ_OK;
    Device &d = ctx->devs[0];
    CK(ctx, cudaSetDevice(d.id));
    CK(ctx, d.in.ensure(in_len + 16));
    CK(ctx, cudaMemcpyAsync(d.in.p, in, in_len, cudaMemcpyHostToDevice, d.stream));
    std::vector<RleBlock> blocks;
    std::vector<uint32_t> crcs;
    uint64_t total = 0;
    int rc = run_rle_device(ctx, d, d.in.as<uint8_t>(), in, in_len, level, blocks, crcs, &total);
    if (rc != BNZ_OK) return rc;
    if (blocks.size() > max_blocks) return fail(ctx, BNZ_EINVAL, "max_blocks too small");
    std::vector<uint8_t> tmp(total);
    CK(ctx, cudaMemcpyAsync(tmp.data(), d.rle.p, total, cudaMemcpyDeviceToHost, d.stream));
    CK(ctx, cudaStreamSynchronize(d.stream));
    uint64_t off = 0;
    for (size_t b = 0; b < blocks.size(); b++) {
        if (off + blocks[b].n > rle_cap) return fail(ctx, BNZ_EINVAL, "rle_cap too small");
        blk_in_off[b] = blocks[b].s;
        blk_in_len[b] = blocks[b].c - blocks[b].s;
        blk_rle_off[b] = off;
        blk_rle_len[b] = blocks[b].n;
        blk_crc[b] = crcs[b];
        memcpy(rle_out + off, tmp.data() + blocks[b].rle_off, blocks[b].n);
        off += blocks[b].n;
    }
    *n_blocks = blocks.size();
    return BNZ_OK;
}

// ---------------------------------------------------------------------------------------
// MTF / Huffman stages on one device (device-resident batch)
// ---------------------------------------------------------------------------------------

struct Batch {                      // host description of the blocks resident on a device
    std::vector<uint64_t> blk_off;  // byte offset of each block image (16-byte aligned)
    std::vector<uint32_t> blk_len;  // n
    std::vector<uint64_t> sym_off;  // element offset of each block's symbols
    std::vector<uint32_t> seg_base; // [nb + 1]
    std::vector<uint32_t> span_base;// [nb + 1]
    uint64_t bytes_total = 0;       // size of the rle / bwt / idx arrays
    uint64_t syms_total = 0;        // elements in the syms array
    uint32_t max_len = 0;
    void build()
    {
        const size_t nb = blk_len.size();
        sym_off.resize(nb);
        seg_base.resize(nb + 1);
        span_base.resize(nb + 1);
        uint64_t so = 0;
        uint32_t sg = 0, sp = 0;
        const uint32_t gps = huff_groups_per_span();
        max_len = 0;
        for (size_t b = 0; b < nb; b++) {
            sym_off[b] = so;
            so += ((uint64_t)blk_len[b] + 1 + 15) & ~15ull;
            seg_base[b] = sg;
            sg += (blk_len[b] + MTF_SEG - 1) / MTF_SEG;
            span_base[b] = sp;
            uint32_t groups = (blk_len[b] + 1 + 49) / 50;            // upper bound: m <= n + 1
            sp += (groups + gps - 1) / gps;
            max_len = std::max(max_len, blk_len[b]);
        }
        seg_base[nb] = sg;
        span_base[nb] = sp;
        syms_total = so;
    }
};

template <class X>
static cudaError_t upload(DevBuf &buf, const std::vector<X> &v, cudaStream_t st)
{
    cudaError_t e = buf.ensure(v.size() * sizeof(X) + 16);
    if (e != cudaSuccess) return e;
    return cudaMemcpyAsync(buf.p, v.data(), v.size() * sizeof(X), cudaMemcpyHostToDevice, st);
}

static int upload_batch(bnz_ctx *ctx, Device &d, const Batch &bt)
{
    CK(ctx, upload(d.blk_off, bt.blk_off, d.stream));
    CK(ctx, upload(d.blk_len, bt.blk_len, d.stream));
    CK(ctx, upload(d.sym_off, bt.sym_off, d.stream));
    CK(ctx, upload(d.seg_base, bt.seg_base, d.stream));
    CK(ctx, upload(d.span_base, bt.span_base, d.stream));
    return BNZ_OK;
}

// bwt bytes in d_bwt -> symbols in d.syms (+ sym_len, num_names, freqs); d_idx is scratch of
// the same size/layout as d_bwt.
static int mtf_ensure(bnz_ctx *ctx, Device &d, const Batch &bt)
{
    const uint32_t nb = (uint32_t)bt.blk_len.size();
    const uint32_t segs = bt.seg_base[nb];
    CK(ctx, d.seg_list.ensure((size_t)segs * 256));
    CK(ctx, d.seg_cnt.ensure((size_t)segs * 4));
    CK(ctx, d.seg_state.ensure((size_t)segs * 256));
    CK(ctx, d.num_names.ensure((size_t)nb * 4));
    CK(ctx, d.syms.ensure(bt.syms_total * 2));
    CK(ctx, d.sym_len.ensure((size_t)nb * 4));
    CK(ctx, d.freqs.ensure((size_t)nb * 258 * 4));
    CK(ctx, d.mtf_ids.ensure((size_t)nb * 4));
    CK(ctx, d.mtf_cseg.ensure(((size_t)nb + 64) * 4));
    CK(ctx, d.h_mtf.ensure(((size_t)nb * 2 + 64) * 4));
    return BNZ_OK;
}

// MTF of a list of blocks of the batch (any subset, any order) on stream `st`.  `ids_used` /
// `lists_used` count what earlier lists of the same batch took from the id / segment-prefix
// arrays (every block is listed once per batch, so nothing is overwritten while in use).
static int run_mtf_list(bnz_ctx *ctx, Device &d, const Batch &bt, const uint8_t *d_bwt, uint8_t *d_idx,
                        const uint8_t *d_has_byte, const uint32_t *ids, uint32_t n_list, uint32_t &ids_used,
                        uint32_t &lists_used, cudaStream_t st)
{
    if (n_list == 0) return BNZ_OK;
    const uint32_t nb = (uint32_t)bt.blk_len.size();
    if (ids_used + n_list > nb || lists_used >= 64) return fail(ctx, BNZ_EINTERNAL, "MTF list bookkeeping");
    uint32_t *h_ids = d.h_mtf.as<uint32_t>() + ids_used;
    uint32_t *h_cseg = d.h_mtf.as<uint32_t>() + nb + ids_used + lists_used;
    uint32_t segs = 0;
    for (uint32_t k = 0; k < n_list; k++) {
        h_ids[k] = ids[k];
        h_cseg[k] = segs;
        segs += bt.seg_base[ids[k] + 1] - bt.seg_base[ids[k]];
    }
    h_cseg[n_list] = segs;
    uint32_t *d_ids = d.mtf_ids.as<uint32_t>() + ids_used;
    uint32_t *d_cseg = d.mtf_cseg.as<uint32_t>() + ids_used + lists_used;
    CK(ctx, cudaMemcpyAsync(d_ids, h_ids, (size_t)n_list * 4, cudaMemcpyHostToDevice, st));
    CK(ctx, cudaMemcpyAsync(d_cseg, h_cseg, ((size_t)n_list + 1) * 4, cudaMemcpyHostToDevice, st));
    ids_used += n_list;
    lists_used++;
    MtfArgs a;
    a.bwt = d_bwt;
    a.idx = d_idx;
    a.blk_off = d.blk_off.as<uint64_t>();
    a.blk_len = d.blk_len.as<uint32_t>();
    a.has_byte = d_has_byte;
    a.n_blocks = n_list;
    a.ids = d_ids;
    a.cseg_base = d_cseg;
    a.seg_base = d.seg_base.as<uint32_t>();
    a.total_segs = segs;
    a.seg_list = d.seg_list.as<uint8_t>();
    a.seg_cnt = d.seg_cnt.as<uint32_t>();
    a.seg_state = d.seg_state.as<uint8_t>();
    a.num_names = d.num_names.as<uint32_t>();
    a.syms = d.syms.as<uint16_t>();
    a.sym_off = d.sym_off.as<uint64_t>();
    a.sym_len = d.sym_len.as<uint32_t>();
    a.freqs = d.freqs.as<uint32_t>();
    CK(ctx, mtf_launch(a, st, &d.launches));
    return BNZ_OK;
}

static int run_mtf_device(bnz_ctx *ctx, Device &d, const Batch &bt, const uint8_t *d_bwt, uint8_t *d_idx,
                          const uint8_t *d_has_byte)
{
    int rc = mtf_ensure(ctx, d, bt);
    if (rc != BNZ_OK) return rc;
    std::vector<uint32_t> all(bt.blk_len.size());
    for (uint32_t b = 0; b < all.size(); b++) all[b] = b;
    uint32_t ids_used = 0, lists_used = 0;
    return run_mtf_list(ctx, d, bt, d_bwt, d_idx, d_has_byte, all.data(), (uint32_t)all.size(), ids_used, lists_used, d.stream);
}

static size_t hdr_stride_words(int level)
{
    const size_t smax = ((size_t)100000 * level + 1 + 49) / 50;
    const size_t bits = 512 + smax + 6 * (5 + 258 * 33);
    return ((bits + 31) / 32 + 3) & ~(size_t)3;
}

static void fill_huff_args(HuffArgs &a, Device &d, const Batch &bt, int level)
{
    const uint32_t nb = (uint32_t)bt.blk_len.size();
    memset(&a, 0, sizeof a);
    a.syms = d.syms.as<uint16_t>();
    a.sym_off = d.sym_off.as<uint64_t>();
    a.sym_len = d.sym_len.as<uint32_t>();
    a.num_names = d.num_names.as<uint32_t>();
    a.freqs = d.freqs.as<uint32_t>();
    a.n_blocks = nb;
    a.lens = d.lens.as<uint8_t>();
    a.codes = d.codes.as<uint32_t>();
    a.tf = d.tf.as<uint32_t>();
    a.num_tables = d.num_tables.as<uint32_t>();
    a.num_sel = d.num_sel.as<uint32_t>();
    a.selectors = nullptr;
    a.sel_stride = 0;
    a.span_base = d.span_base.as<uint32_t>();
    a.hdr = d.hdr.as<uint32_t>();
    a.hdr_stride = hdr_stride_words(level);
    a.hdr_bits = d.hdr_bits.as<uint32_t>();
    a.crc = d.crc.as<uint32_t>();
    a.ptr = d.ptr.as<uint32_t>();
    a.has_byte = d.has_byte.as<uint8_t>();
    a.blk_bits = d.blk_bits.as<uint64_t>();
    a.blk_bitoff = d.blk_bitoff.as<uint64_t>();
    a.total_bits = d.total_bits.as<uint64_t>();
    a.out_words = d.out.as<uint32_t>();
}

// modelling + tables + headers + bit offsets; leaves total bits (bit_base + sum) in *total_bits_host
static int run_huff_model_device(bnz_ctx *ctx, Device &d, const Batch &bt, int level, int with_block_header,
                                 uint64_t bit_base, uint64_t fixed_stride_bits, HuffArgs &a)
{
    const uint32_t nb = (uint32_t)bt.blk_len.size();
    const size_t tsz = (size_t)nb * HUFF_MAX_TABLES * HUFF_MAX_SYMS;
    CK(ctx, d.lens.ensure(tsz));
    CK(ctx, d.codes.ensure(tsz * 4));
    CK(ctx, d.tf.ensure(tsz * 4));
    CK(ctx, d.num_tables.ensure((size_t)nb * 4));
    CK(ctx, d.num_sel.ensure((size_t)nb * 4));
    CK(ctx, d.hdr.ensure((size_t)nb * hdr_stride_words(level) * 4));
    CK(ctx, d.hdr_bits.ensure((size_t)nb * 4));
    CK(ctx, d.blk_bits.ensure((size_t)nb * 8));
    CK(ctx, d.blk_bitoff.ensure((size_t)nb * 8));
    CK(ctx, d.total_bits.ensure(64));
    CK(ctx, cudaMemsetAsync(d.tf.p, 0, tsz * 4, d.stream));
    fill_huff_args(a, d, bt, level);
    a.with_block_header = with_block_header;
    a.bit_base = bit_base;
    a.fixed_stride_bits = fixed_stride_bits;
    CK(ctx, huff_launch(a, bt.span_base[nb], d.stream, &d.launches));
    return BNZ_OK;
}

// ---------------------------------------------------------------------------------------
// stage exports: MTF and Huffman
// ---------------------------------------------------------------------------------------

extern "C" int bnz_stage_mtf(bnz_ctx *ctx, const uint8_t *bwt, const uint64_t *blk_off, const uint32_t *blk_len,
                             const uint8_t *has_byte, size_t n_blocks, uint16_t *syms_out, uint32_t *sym_len,
                             uint32_t *num_syms, uint32_t *freqs_out)
{
    if (!ctx) return BNZ_EINVAL;
    if (n_blocks == 0) return BNZ_OK;
    if (!bwt || !blk_off || !blk_len || !has_byte || !syms_out || !sym_len || !num_syms || !freqs_out)
        return BNZ_EINVAL;
    Device &d = ctx->devs[0];
    CK(ctx, cudaSetDevice(d.id));
    Batch bt;
    bt.blk_off.resize(n_blocks);
    bt.blk_len.assign(blk_len, blk_len + n_blocks);
    uint64_t off = 0;
    for (size_t b = 0; b < n_blocks; b++) {
        if (blk_len[b] == 0 || blk_len[b] > 900000) return BNZ_EINVAL;
        bt.blk_off[b] = off;
        off += ((uint64_t)blk_len[b] + 15) & ~15ull;
    }
    bt.bytes_total = off;
    bt.build();
    CK(ctx, d.bwt.ensure(off));
    CK(ctx, d.rle.ensure(off));
    CK(ctx, d.has_byte.ensure(n_blocks * 256));
    for (size_t b = 0; b < n_blocks; b++)
        CK(ctx, cudaMemcpyAsync(d.bwt.as<uint8_t>() + bt.blk_off[b], bwt + blk_off[b], blk_len[b],
                                cudaMemcpyHostToDevice, d.stream));
    CK(ctx, cudaMemcpyAsync(d.has_byte.p, has_byte, n_blocks * 256, cudaMemcpyHostToDevice, d.stream));
    int rc = upload_batch(ctx, d, bt);
    if (rc != BNZ_OK) return rc;
    rc = run_mtf_device(ctx, d, bt, d.bwt.as<uint8_t>(), d.rle.as<uint8_t>(), d.has_byte.as<uint8_t>());
    if (rc != BNZ_OK) return rc;
    std::vector<uint32_t> nn(n_blocks);
    CK(ctx, cudaMemcpyAsync(sym_len, d.sym_len.p, n_blocks * 4, cudaMemcpyDeviceToHost, d.stream));
    CK(ctx, cudaMemcpyAsync(nn.data(), d.num_names.p, n_blocks * 4, cudaMemcpyDeviceToHost, d.stream));
    CK(ctx, cudaMemcpyAsync(freqs_out, d.freqs.p, n_blocks * 258 * 4, cudaMemcpyDeviceToHost, d.stream));
    CK(ctx, cudaStreamSynchronize(d.stream));
    for (size_t b = 0; b < n_blocks; b++) {
        num_syms[b] = nn[b] + 2;
        CK(ctx, cudaMemcpyAsync(syms_out + blk_off[b] + b, d.syms.as<uint16_t>() + bt.sym_off[b],
                                (size_t)sym_len[b] * 2, cudaMemcpyDeviceToHost, d.stream));
    }
    CK(ctx, cudaStreamSynchronize(d.stream));
    return BNZ_OK;
}

extern "C" int bnz_stage_huffman(bnz_ctx *ctx, const uint16_t *syms, const uint64_t *sym_off,
                                 const uint32_t *sym_len, const uint32_t *num_syms, const uint32_t *freqs,
                                 size_t n_blocks, uint8_t *bits_out, size_t out_stride, uint64_t *bit_len,
                                 uint8_t *tables_out, uint32_t *num_tables)
{
    if (!ctx) return BNZ_EINVAL;
    if (n_blocks == 0) return BNZ_OK;
    if (!syms || !sym_off || !sym_len || !num_syms || !freqs || !bits_out || !bit_len || !tables_out ||
        !num_tables || (out_stride & 3))
        return BNZ_EINVAL;
    Device &d = ctx->devs[0];
    CK(ctx, cudaSetDevice(d.id));
    Batch bt;
    bt.blk_off.resize(n_blocks);
    bt.blk_len.resize(n_blocks);
    std::vector<uint32_t> nn(n_blocks);
    uint64_t off = 0;
    for (size_t b = 0; b < n_blocks; b++) {
        if (sym_len[b] < 1 || sym_len[b] > 900001 || num_syms[b] < 3 || num_syms[b] > 258) return BNZ_EINVAL;
        bt.blk_len[b] = sym_len[b] - 1;        // m <= n + 1 bookkeeping
        if (bt.blk_len[b] == 0) bt.blk_len[b] = 1;
        bt.blk_off[b] = off;
        off += ((uint64_t)bt.blk_len[b] + 15) & ~15ull;
        nn[b] = num_syms[b] - 2;
    }
    bt.build();
    int rc = upload_batch(ctx, d, bt);
    if (rc != BNZ_OK) return rc;
    CK(ctx, d.syms.ensure(bt.syms_total * 2));
    CK(ctx, d.sym_len.ensure(n_blocks * 4));
    CK(ctx, d.num_names.ensure(n_blocks * 4));
    CK(ctx, d.freqs.ensure(n_blocks * 258 * 4));
    for (size_t b = 0; b < n_blocks; b++)
        CK(ctx, cudaMemcpyAsync(d.syms.as<uint16_t>() + bt.sym_off[b], syms + sym_off[b], (size_t)sym_len[b] * 2,
                                cudaMemcpyHostToDevice, d.stream));
    CK(ctx, cudaMemcpyAsync(d.sym_len.p, sym_len, n_blocks * 4, cudaMemcpyHostToDevice, d.stream));
    CK(ctx, cudaMemcpyAsync(d.num_names.p, nn.data(), n_blocks * 4, cudaMemcpyHostToDevice, d.stream));
    CK(ctx, cudaMemcpyAsync(d.freqs.p, freqs, n_blocks * 258 * 4, cudaMemcpyHostToDevice, d.stream));
    CK(ctx, d.out.ensure(n_blocks * out_stride + 64));
    CK(ctx, cudaMemsetAsync(d.out.p, 0, n_blocks * out_stride + 64, d.stream));
    HuffArgs a;
    rc = run_huff_model_device(ctx, d, bt, 9, 0, 0, (uint64_t)out_stride * 8, a);
    if (rc != BNZ_OK) return rc;
    a.out_words = d.out.as<uint32_t>();
    CK(ctx, huff_pack_launch(a, d.stream, &d.launches));
    std::vector<uint64_t> bb(n_blocks);
    CK(ctx, cudaMemcpyAsync(bb.data(), d.blk_bits.p, n_blocks * 8, cudaMemcpyDeviceToHost, d.stream));
    CK(ctx, cudaMemcpyAsync(num_tables, d.num_tables.p, n_blocks * 4, cudaMemcpyDeviceToHost, d.stream));
    CK(ctx, cudaMemcpyAsync(tables_out, d.lens.p, n_blocks * HUFF_MAX_TABLES * HUFF_MAX_SYMS,
                            cudaMemcpyDeviceToHost, d.stream));
    CK(ctx, cudaMemcpyAsync(bits_out, d.out.p, n_blocks * out_stride, cudaMemcpyDeviceToHost, d.stream));
    CK(ctx, cudaStreamSynchronize(d.stream));
    for (size_t b = 0; b < n_blocks; b++) {
        bit_len[b] = bb[b];
        if ((bb[b] + 7) / 8 > out_stride) return fail(ctx, BNZ_EINVAL, "out_stride too small");
    }
    return BNZ_OK;
}

// ---------------------------------------------------------------------------------------
// bnz_encode: the whole path
// ---------------------------------------------------------------------------------------

static void put_bits_host(uint8_t *buf, uint64_t bitpos, uint64_t value, int nbits)   // MSB first
{
    for (int i = nbits - 1; i >= 0; i--, bitpos++)
        if ((value >> i) & 1) buf[bitpos >> 3] |= (uint8_t)(0x80u >> (bitpos & 7));
}

// One device's share of a bnz_encode call: a contiguous range of blocks.
struct Shard {
    Device *d = nullptr;
    std::vector<RleBlock> blocks;      // rle_off rebased to this device's rle buffer
    std::vector<uint32_t> crcs;
    std::vector<BwtStats> bst;
    Batch bt;
    HuffArgs ha;
    uint64_t block_bits = 0;           // sum of the shard's block bit lengths
    uint64_t bit_base = 0;             // global bit offset of the shard's first block
    int rc = BNZ_OK;
    std::string err;
    // lanes on one GPU: the persistent BWT kernels must not share the SMs, so lane g's sort waits
    // for lane g-1's (event recorded on that lane's stream; the flag orders the host threads)
    Shard *bwt_after = nullptr;
    std::atomic<int> bwt_recorded{0};
    Shard() = default;
    Shard(const Shard &o) { d = o.d; }
};

// K1 emit .. K7 + headers for one shard; ends with a host sync that yields block_bits.
// in_base/oin_base/P_base: see rle_emit_shard.
static int shard_model(bnz_ctx *ctx, Shard &sh, const uint8_t *in_base, uint64_t N, const uint64_t *oin_base,
                       const uint64_t *P_base, int level)
{
    Device &d = *sh.d;
    uint64_t rle_total = 0;
    int rc = rle_emit_shard(ctx, d, in_base, N, oin_base, P_base, sh.blocks, nullptr, &rle_total);
    if (rc != BNZ_OK) return rc;
    CK(ctx, cudaEventRecord(d.ev[2], d.stream));
    const uint32_t nb = (uint32_t)sh.blocks.size();
    Batch &bt = sh.bt;
    bt.blk_off.resize(nb);
    bt.blk_len.resize(nb);
    for (uint32_t b = 0; b < nb; b++) {
        bt.blk_off[b] = sh.blocks[b].rle_off;
        bt.blk_len[b] = sh.blocks[b].n;
    }
    bt.bytes_total = rle_total;
    bt.build();
    rc = upload_batch(ctx, d, bt);
    if (rc != BNZ_OK) return rc;

    // K3/K4
    CK(ctx, d.bwt.ensure(rle_total));
    CK(ctx, d.ptr.ensure((size_t)nb * 4));
    CK(ctx, d.has_byte.ensure((size_t)nb * 256));
    CK(ctx, d.bwt_stats.ensure((size_t)nb * sizeof(BwtStats)));
    if (sh.bwt_after) {
        while (sh.bwt_after->bwt_recorded.load(std::memory_order_acquire) == 0) std::this_thread::yield();
        if (sh.bwt_after->bwt_recorded.load() > 0) CK(ctx, cudaStreamWaitEvent(d.stream, sh.bwt_after->d->ev[3], 0));
        CK(ctx, cudaEventRecord(d.ev[2], d.stream));       // RLE stage ends where the sort may start
    }
    // MTF arenas and the completion flags are set up before the sort is launched (allocation
    // would synchronise with it)
    rc = mtf_ensure(ctx, d, bt);
    if (rc != BNZ_OK) return rc;
    CK(ctx, d.h_done.ensure((size_t)nb * 4));
    volatile uint32_t *h_done = d.h_done.as<uint32_t>();
    memset(d.h_done.p, 0, (size_t)nb * 4);
    uint32_t *d_done = nullptr;
    CK(ctx, cudaHostGetDevicePointer((void **)&d_done, d.h_done.p, 0));
    CK(ctx, cudaEventRecord(d.ev[12], d.stream));
    bool armed = false;
    rc = run_bwt_device(ctx, d, d.rle.as<uint8_t>(), d.bwt.as<uint8_t>(), d.blk_off.as<uint64_t>(),
                        d.blk_len.as<uint32_t>(), nb, bt.max_len, d.ptr.as<uint32_t>(), d.has_byte.as<uint8_t>(),
                        d.bwt_stats.as<BwtStats>(), ctx->mtf_overlap > 0 ? d_done : nullptr, &armed);
    if (rc != BNZ_OK) {
        sh.bwt_recorded.store(-1, std::memory_order_release);
        return rc;
    }
    CK(ctx, cudaEventRecord(d.ev[3], d.stream));
    sh.bwt_recorded.store(1, std::memory_order_release);

    // K5 (the RLE1 images are dead once their block is sorted: their buffer holds the MTF index
    // bytes).  The one-CTA-per-block sort ends in a long tail (blocks differ 5x in cost and only
    // ~4 fit per CTA), so the MTF of the blocks that finish first runs beside it: the sort raises a
    // host-visible flag per finished block, and as soon as a leading group of blocks is complete
    // this thread queues its MTF on a low-priority stream, whose CTAs get the SM slots the sort
    // leaves empty.
    std::vector<uint32_t> list;
    std::vector<uint8_t> taken(nb, 0);
    uint32_t ids_used = 0, lists_used = 0;
    if (armed) {
        for (cudaStream_t st : d.stream3) CK(ctx, cudaStreamWaitEvent(st, d.ev[12], 0));
        const uint32_t budget = (uint32_t)((uint64_t)nb * (uint32_t)ctx->mtf_overlap / 100);   // blocks that may go beside the sort
        const uint32_t step = std::max<uint32_t>(32, budget / (uint32_t)ctx->mtf_groups);
        uint32_t n_over = 0;
        int g = 0;
        list.reserve(nb);
        while (n_over < budget) {
            for (uint32_t b = 0; b < nb && list.size() < step; b++)
                if (!taken[b] && h_done[b]) {
                    taken[b] = 1;
                    list.push_back(b);
                }
            if (list.size() >= step) {
                std::atomic_thread_fence(std::memory_order_acquire);
                rc = run_mtf_list(ctx, d, bt, d.bwt.as<uint8_t>(), d.rle.as<uint8_t>(), d.has_byte.as<uint8_t>(), list.data(),
                                  (uint32_t)list.size(), ids_used, lists_used, d.stream3[g++ % 3]);
                if (rc != BNZ_OK) return rc;
                n_over += (uint32_t)list.size();
                list.clear();
                continue;
            }
            if (cudaEventQuery(d.ev[3]) != cudaErrorNotReady) break;       // the sort ended (or failed)
            std::this_thread::yield();
        }
        for (int k = 0; k < 3; k++) CK(ctx, cudaEventRecord(d.ev[13 + k], d.stream3[k]));
    }
    // everything not queued beside the sort follows it on the main stream
    for (uint32_t b = 0; b < nb; b++)
        if (!taken[b]) list.push_back(b);      // (a partly gathered list is already in `list`)
    rc = run_mtf_list(ctx, d, bt, d.bwt.as<uint8_t>(), d.rle.as<uint8_t>(), d.has_byte.as<uint8_t>(), list.data(),
                      (uint32_t)list.size(), ids_used, lists_used, d.stream);
    if (rc != BNZ_OK) return rc;
    if (armed)
        for (int k = 0; k < 3; k++) CK(ctx, cudaStreamWaitEvent(d.stream, d.ev[13 + k], 0));
    CK(ctx, cudaEventRecord(d.ev[4], d.stream));

    // K6/K7 + headers + block bit lengths (the headers need the block CRCs from the side stream)
    CK(ctx, cudaStreamWaitEvent(d.stream, d.ev[10], 0));
    rc = run_huff_model_device(ctx, d, bt, level, 1, 0, 0, sh.ha);
    if (rc != BNZ_OK) return rc;
    sh.bst.resize(nb);
    sh.crcs.resize(nb);
    CK(ctx, cudaMemcpyAsync(sh.crcs.data(), d.crc.p, (size_t)nb * 4, cudaMemcpyDeviceToHost, d.stream));
    CK(ctx, cudaMemcpyAsync(&sh.block_bits, d.total_bits.p, 8, cudaMemcpyDeviceToHost, d.stream));
    CK(ctx, cudaMemcpyAsync(sh.bst.data(), d.bwt_stats.p, (size_t)nb * sizeof(BwtStats), cudaMemcpyDeviceToHost, d.stream));
    CK(ctx, cudaEventRecord(d.ev[5], d.stream));
    CK(ctx, cudaStreamSynchronize(d.stream));
    return BNZ_OK;
}

// K8 for one shard once its global bit offset is known.  The shard's bits land in d.out such
// that d.out word 0 is global word (bit_base / 32).
static int shard_pack(bnz_ctx *ctx, Shard &sh, size_t *out_bytes)
{
    Device &d = *sh.d;
    const uint64_t local_base = sh.bit_base & 31;
    const size_t bytes = (size_t)((local_base + sh.block_bits + 31) / 32) * 4;
    *out_bytes = bytes;
    CK(ctx, d.out.ensure(bytes + 256));
    CK(ctx, cudaMemsetAsync(d.out.p, 0, ((bytes + 127) & ~(size_t)63), d.stream));
    sh.ha.out_words = d.out.as<uint32_t>();
    sh.ha.bit_base = local_base;
    CK(ctx, huff_rescan_launch(sh.ha, d.stream, &d.launches));
    CK(ctx, huff_pack_launch(sh.ha, d.stream, &d.launches));
    CK(ctx, cudaEventRecord(d.ev[6], d.stream));
    return BNZ_OK;
}

static void add_stats(bnz_stats &st, const Shard &sh)
{
    st.n_blocks += (uint32_t)sh.blocks.size();
    for (const BwtStats &b : sh.bst) {
        st.bwt_n += b.n;
        st.bwt_sum_active += b.sum_active;
        st.bwt_sum_active_passes += b.sum_active_passes;
        st.bwt_rounds_total += b.rounds;
        st.bwt_max_rounds = std::max(st.bwt_max_rounds, b.rounds);
        st.bwt_tied_blocks += b.tied;
        st.bwt_cyc_build += b.cyc_build;
        st.bwt_cyc_radix += b.cyc_radix;
        st.bwt_cyc_rerank += b.cyc_rerank;
    }
    st.bwt_algorithmic_bytes = 9 * st.bwt_n + 16 * st.bwt_sum_active_passes + 36 * st.bwt_sum_active;
    st.kernel_launches += sh.d->launches;
}

static void finish_stats(bnz_ctx *ctx, std::vector<Shard> &shards, bool have_d2h)
{
    // per batch: max over the shards (they run concurrently); batches add up
    bnz_stats &st = ctx->stats;
    float h2d = 0, rle = 0, bwt = 0, mtf = 0, huff = 0, pack = 0, d2h = 0, total = 0;
    for (Shard &sh : shards) {
        if (sh.blocks.empty()) continue;
        Device &d = *sh.d;
        cudaSetDevice(d.id);
        auto el = [&](int a, int b) { float ms = 0; cudaEventElapsedTime(&ms, d.ev[a], d.ev[b]); return ms; };
        h2d = std::max(h2d, el(0, 1));
        rle = std::max(rle, el(1, 2));
        bwt = std::max(bwt, el(2, 3));
        mtf = std::max(mtf, el(3, 4));
        huff = std::max(huff, el(4, 5));
        pack = std::max(pack, el(5, 6));
        if (have_d2h) d2h = std::max(d2h, el(6, 7));
        total = std::max(total, el(0, have_d2h ? 7 : 6));
    }
    st.h2d_ms += h2d;
    st.rle_ms += rle;
    st.bwt_ms += bwt;
    st.mtf_ms += mtf;
    st.huff_ms += huff;
    st.pack_ms += pack;
    st.d2h_ms += d2h;
    st.total_ms += total;
    st.n_devices = std::max<uint32_t>(st.n_devices, (uint32_t)shards.size());
    st.bwt_radix_bits = (uint32_t)ctx->radix_bits;
}

static uint32_t fold_stream_crc(const std::vector<uint32_t> &crcs)       // lib.rs:108
{
    uint32_t s = 0;
    for (uint32_t c : crcs) s = c ^ ((s << 1) | (s >> 31));
    return s;
}

static int ensure_out_cache(bnz_ctx *ctx, size_t nbytes)
{
    if (ctx->out_cache_cap < nbytes + 16) {
        if (ctx->out_cache) cudaFreeHost(ctx->out_cache);
        ctx->out_cache = nullptr;
        ctx->out_cache_cap = 0;
        size_t want = nbytes + nbytes / 4 + 4096;
        CK(ctx, cudaHostAlloc((void **)&ctx->out_cache, want, cudaHostAllocPortable));
        ctx->out_cache_cap = want;
    }
    return BNZ_OK;
}

// contiguous block ranges with ~equal RLE1 bytes per device
static std::vector<uint32_t> split_blocks(const std::vector<RleBlock> &blocks, size_t n_dev)
{
    std::vector<uint32_t> cut(n_dev + 1, 0);
    uint64_t total = 0;
    for (const RleBlock &b : blocks) total += b.n;
    uint64_t acc = 0;
    size_t g = 1;
    for (uint32_t i = 0; i < blocks.size() && g < n_dev; i++) {
        acc += blocks[i].n;
        while (g < n_dev && acc * n_dev >= total * g) cut[g++] = i + 1;
    }
    for (; g <= n_dev; g++) cut[g] = (uint32_t)blocks.size();
    return cut;
}

// One batch of the path: the blocks that can be cut from h_in[0, N).  d_in0: optional device copy
// already resident on device 0.  `final`: no input follows (otherwise the trailing incomplete
// block is left for the next batch; *consumed tells where it starts).  `bit_base`: bit offset of
// the batch's first block in the stream.  Leaves every shard's bits in its device's d.out and
// returns the layout; the callers move the bytes.
static int encode_all(bnz_ctx *ctx, const uint8_t *h_in, const uint8_t *d_in0, size_t N, int level,
                      std::vector<Shard> &shards, std::vector<uint32_t> &crcs, uint64_t *total_bits,
                      bool final = true, uint64_t bit_base = 32, uint64_t *consumed = nullptr)
{
    Device &d0 = ctx->devs[0];
    bnz_stats &st = ctx->stats;
    for (Device &d : ctx->devs) d.launches = 0;
    CK(ctx, cudaSetDevice(d0.id));
    CK(ctx, cudaEventRecord(d0.ev[0], d0.stream));
    const uint8_t *d_in = d_in0;
    if (!d_in) {
        CK(ctx, d0.in.ensure(N + 64));
        CK(ctx, cudaMemcpyAsync(d0.in.p, h_in, N, cudaMemcpyHostToDevice, d0.stream));
        d_in = d0.in.as<uint8_t>();
        st.h2d_bytes += N;
    }
    CK(ctx, cudaEventRecord(d0.ev[1], d0.stream));

    std::vector<RleBlock> blocks;
    int rc = rle_plan(ctx, d0, d_in, h_in, N, level, blocks, final, consumed);
    if (rc != BNZ_OK) return rc;
    if (blocks.empty()) {               // (non-final batch shorter than one block)
        shards.clear();
        *total_bits = bit_base;
        return BNZ_OK;
    }

    CK(ctx, cudaEventRecord(d0.ev[8], d0.stream));       // input + chunk tables resident on device 0
    const size_t n_dev = std::min(ctx->devs.size(), std::max<size_t>(1, blocks.size()));
    std::vector<uint32_t> cut = split_blocks(blocks, n_dev);
    shards = std::vector<Shard>(n_dev);
    for (size_t g = 0; g < n_dev; g++) {
        Shard &sh = shards[g];
        sh.d = &ctx->devs[g];
        if (g > 0 && ctx->devs[g].id == ctx->devs[g - 1].id) sh.bwt_after = &shards[g - 1];
        sh.blocks.assign(blocks.begin() + cut[g], blocks.begin() + cut[g + 1]);
        const uint64_t off0 = sh.blocks.empty() ? 0 : sh.blocks.front().rle_off;
        for (RleBlock &b : sh.blocks) b.rle_off -= off0;
    }

    const uint64_t *h_P = d0.h_P.as<uint64_t>(), *h_oin = d0.h_oin.as<uint64_t>();
    auto work = [&](size_t g) -> int {
        Shard &sh = shards[g];
        Device &d = *sh.d;
        if (sh.blocks.empty()) {
            sh.bwt_recorded.store(-1, std::memory_order_release);
            return BNZ_OK;
        }
        CK(ctx, cudaSetDevice(d.id));
        if (d.id == d0.id) {
            // same physical GPU (lane 0, or an extra lane that overlaps its stages with the other
            // lanes' kernels): the input and the chunk tables are already resident
            if (g != 0) {
                CK(ctx, cudaStreamWaitEvent(d.stream, d0.ev[8], 0));
                CK(ctx, cudaEventRecord(d.ev[0], d.stream));
                CK(ctx, cudaEventRecord(d.ev[1], d.stream));
            }
            return shard_model(ctx, sh, d_in, N, d0.ch_oin.as<uint64_t>(), d0.ch_P.as<uint64_t>(), level);
        }
        // other devices: make their input range and chunk tables resident
        CK(ctx, cudaEventRecord(d.ev[0], d.stream));
        const uint64_t c0 = sh.blocks.front().s / RLE_CHUNK;
        const uint64_t c1 = (sh.blocks.back().c + RLE_CHUNK - 1) / RLE_CHUNK;
        const uint64_t a = c0 ? c0 * RLE_CHUNK - 16 : 0;
        const uint64_t b = std::min<uint64_t>(N, c1 * RLE_CHUNK + 16);
        CK(ctx, d.in.ensure(b - a + 64));
        CK(ctx, cudaMemcpyAsync(d.in.p, h_in + a, b - a, cudaMemcpyHostToDevice, d.stream));
        CK(ctx, d.ch_oin.ensure((c1 - c0 + 1) * 8));
        CK(ctx, d.ch_P.ensure((c1 - c0 + 2) * 8));
        CK(ctx, cudaMemcpyAsync(d.ch_oin.p, h_oin + c0, (c1 - c0) * 8, cudaMemcpyHostToDevice, d.stream));
        CK(ctx, cudaMemcpyAsync(d.ch_P.p, h_P + c0, (c1 - c0 + 1) * 8, cudaMemcpyHostToDevice, d.stream));
        CK(ctx, cudaEventRecord(d.ev[1], d.stream));
        return shard_model(ctx, sh, d.in.as<uint8_t>() - a, N, d.ch_oin.as<uint64_t>() - c0, d.ch_P.as<uint64_t>() - c0, level);
    };

    if (n_dev == 1) {
        rc = work(0);
        if (rc != BNZ_OK) return rc;
    } else {
        std::vector<std::thread> th;
        for (size_t g = 0; g < n_dev; g++)
            th.emplace_back([&, g]() {
                t_err_sink = &shards[g].err;
                shards[g].rc = work(g);
                if (shards[g].bwt_recorded.load() == 0) shards[g].bwt_recorded.store(-1, std::memory_order_release);
                t_err_sink = nullptr;
            });
        for (std::thread &t : th) t.join();
        for (Shard &sh : shards)
            if (sh.rc != BNZ_OK) {
                ctx->err = sh.err;
                return sh.rc;
            }
    }

    // bit offsets of the shards (blocks are concatenated at bit granularity, lib.rs:101-126 + out.rs)
    uint64_t bits = bit_base;
    for (Shard &sh : shards) {
        sh.bit_base = bits;
        bits += sh.block_bits;
        crcs.insert(crcs.end(), sh.crcs.begin(), sh.crcs.end());
    }
    *total_bits = bits;
    for (Shard &sh : shards) add_stats(st, sh);
    return BNZ_OK;
}

// pack every shard of a batch at its bit phase and copy it to host memory `o` (the stream buffer,
// byte 0 = stream byte 0).  `stream_start`: the batch begins right after the 32-bit stream header,
// so its first word is not shared with earlier data.  (The streaming front end passes a buffer
// that starts at stream byte `o_first_byte`, a multiple of 4.)
static int pack_and_download(bnz_ctx *ctx, std::vector<Shard> &shards, uint8_t *o, bool stream_start,
                             size_t o_first_byte = 0)
{
    std::vector<uint32_t> first_word(shards.size(), 0);
    for (size_t g = 0; g < shards.size(); g++) {
        Shard &sh = shards[g];
        if (sh.blocks.empty()) continue;
        Device &d = *sh.d;
        CK(ctx, cudaSetDevice(d.id));
        size_t bytes = 0;
        int rc = shard_pack(ctx, sh, &bytes);
        if (rc != BNZ_OK) return rc;
        const size_t w0 = (size_t)(sh.bit_base >> 5) * 4 - o_first_byte;
        if (g == 0 && stream_start) {
            CK(ctx, cudaMemcpyAsync(o + w0, d.out.p, bytes, cudaMemcpyDeviceToHost, d.stream));
        } else {
            CK(ctx, cudaMemcpyAsync(&first_word[g], d.out.p, 4, cudaMemcpyDeviceToHost, d.stream));
            if (bytes > 4)
                CK(ctx, cudaMemcpyAsync(o + w0 + 4, d.out.as<uint8_t>() + 4, bytes - 4, cudaMemcpyDeviceToHost, d.stream));
        }
        CK(ctx, cudaEventRecord(d.ev[7], d.stream));
        ctx->stats.d2h_bytes += bytes;
    }
    for (Shard &sh : shards) {
        if (sh.blocks.empty()) continue;
        CK(ctx, cudaSetDevice(sh.d->id));
        CK(ctx, cudaStreamSynchronize(sh.d->stream));
    }
    // merge the words shared with the previous shard / batch
    for (size_t g = 0; g < shards.size(); g++) {
        if (shards[g].blocks.empty() || (g == 0 && stream_start)) continue;
        uint8_t *w = o + ((size_t)(shards[g].bit_base >> 5) * 4 - o_first_byte);
        const uint8_t *f = reinterpret_cast<const uint8_t *>(&first_word[g]);
        if ((shards[g].bit_base & 31) == 0) memcpy(w, f, 4);
        else for (int k = 0; k < 4; k++) w[k] |= f[k];
    }
    return BNZ_OK;
}

extern "C" int bnz_encode(bnz_ctx *ctx, const uint8_t *in, size_t in_len, int level, uint8_t **out,
                          size_t *out_len, size_t *consumed)
{
    if (!ctx || !out || !out_len) return BNZ_EINVAL;
    *out = nullptr;
    *out_len = 0;
    if (consumed) *consumed = 0;
    if (level < 1 || level > 9) return fail(ctx, BNZ_EINVAL, "level must be in 1..=9 (lib/lib.rs:89)");
    if (in_len && !in) return BNZ_EINVAL;
    if (ctx->out_cache_lent || ctx->out_big_lent) return fail(ctx, BNZ_EINVAL, "previous output not released with bnz_free");
    memset(&ctx->stats, 0, sizeof ctx->stats);
    ctx->stats.in_bytes = in_len;

    uint64_t total_bits = 32;
    std::vector<uint32_t> crcs;
    uint8_t *o = nullptr;
    size_t nbytes = 0;

    if (in_len <= ctx->max_batch_bytes) {
        // ---- one batch: the stream is assembled in the context's pinned buffer
        std::vector<Shard> shards;
        if (in_len > 0) {
            int rc = encode_all(ctx, in, nullptr, in_len, level, shards, crcs, &total_bits);
            if (rc != BNZ_OK) return rc;
        }
        nbytes = (size_t)((total_bits + 80 + 7) / 8);
        int rc = ensure_out_cache(ctx, nbytes + 8);
        if (rc != BNZ_OK) return rc;
        o = ctx->out_cache;
        if (in_len > 0) {
            rc = pack_and_download(ctx, shards, o, true);
            if (rc != BNZ_OK) return rc;
            const size_t written = (size_t)((total_bits + 31) / 32) * 4;
            if (written < nbytes + 8) memset(o + written, 0, nbytes + 8 - written);
            finish_stats(ctx, shards, true);
        } else {
            memset(o, 0, nbytes + 8);
        }
        ctx->out_cache_lent = true;
    } else {
        // ---- streaming batches (inputs larger than one device-resident batch): every batch runs
        // the whole pipeline on the blocks that are complete inside its window; the trailing
        // partial block is re-read by the next batch.  The stream grows in an ordinary host buffer
        // sized for the worst case up front: untouched pages cost nothing, and nothing is ever
        // copied or cleared in bulk.
        size_t pos = 0, win = ctx->max_batch_bytes;
        const size_t cap = bnz_max_compressed_size(in_len) + 64;
        if (ctx->out_big_cap < cap) {             // (kept across calls: its pages stay faulted in)
            free(ctx->out_big);
            ctx->out_big = static_cast<uint8_t *>(malloc(cap));
            ctx->out_big_cap = ctx->out_big ? cap : 0;
        }
        o = ctx->out_big;
        if (!o) return fail(ctx, BNZ_ENOMEM, "output buffer");
        memset(o, 0, 64);
        auto grow = [&](size_t need) -> bool { return need <= cap; };
        bool first = true;
        while (pos < in_len) {
            const size_t len = std::min(win, in_len - pos);
            const bool final = pos + len == in_len;
            std::vector<Shard> shards;
            uint64_t used = 0, bits_after = total_bits;
            int rc = encode_all(ctx, in + pos, nullptr, len, level, shards, crcs, &bits_after, final, total_bits, &used);
            if (rc != BNZ_OK) return rc;
            if (shards.empty()) {               // window shorter than one block: widen it
                if (final) break;
                win *= 2;
                continue;
            }
            if (!grow((size_t)((bits_after + 80 + 7) / 8) + 16)) return fail(ctx, BNZ_EINTERNAL, "output bound exceeded");
            rc = pack_and_download(ctx, shards, o, first);
            if (rc != BNZ_OK) return rc;
            // the bytes behind the last (word-rounded) shard must be zero for the next OR-merge / footer
            memset(o + (size_t)((bits_after + 31) / 32) * 4, 0, 32);
            finish_stats(ctx, shards, true);
            first = false;
            total_bits = bits_after;
            pos += final ? len : (size_t)used;
        }
        nbytes = (size_t)((total_bits + 80 + 7) / 8);
        if (!grow(nbytes + 16)) return fail(ctx, BNZ_EINTERNAL, "output bound exceeded");
        ctx->out_big_lent = true;
    }
    // stream header (lib.rs:18-22), footer (lib.rs:66-70), zero padding (out.rs:22-28)
    o[0] = 0x42; o[1] = 0x5A; o[2] = 0x68; o[3] = (uint8_t)('0' + level);
    put_bits_host(o, total_bits, 0x177245385090ull, 48);
    put_bits_host(o, total_bits + 48, fold_stream_crc(crcs), 32);
    ctx->stats.out_bytes = nbytes;
    *out = o;
    *out_len = nbytes;
    if (consumed) *consumed = in_len;
    return BNZ_OK;
}

extern "C" void bnz_free(bnz_ctx *ctx, uint8_t *p)
{
    if (!ctx || !p) return;
    if (p == ctx->out_cache) ctx->out_cache_lent = false;
    if (p == ctx->out_big) ctx->out_big_lent = false;
}

extern "C" int bnz_encode_device(bnz_ctx *ctx, const void *d_in, const uint8_t *h_in, size_t in_len, int level,
                                 void *d_out, size_t d_out_cap, size_t *out_len)
{
    if (!ctx || !out_len || !d_out) return BNZ_EINVAL;
    *out_len = 0;
    if (level < 1 || level > 9) return fail(ctx, BNZ_EINVAL, "level must be in 1..=9 (lib/lib.rs:89)");
    if (in_len == 0 || !d_in || !h_in) return BNZ_EINVAL;
    for (Device &dv : ctx->devs)
        if (dv.id != ctx->devs[0].id) return fail(ctx, BNZ_EINVAL, "bnz_encode_device needs a single-GPU context");
    memset(&ctx->stats, 0, sizeof ctx->stats);
    ctx->stats.in_bytes = in_len;
    Device &d = ctx->devs[0];
    uint64_t total_bits = 32;
    std::vector<uint32_t> crcs;
    std::vector<Shard> shards;
    int rc = encode_all(ctx, h_in, (const uint8_t *)d_in, in_len, level, shards, crcs, &total_bits);
    if (rc != BNZ_OK) return rc;
    const size_t nbytes = (size_t)((total_bits + 80 + 7) / 8);
    if (nbytes + 8 > d_out_cap) return fail(ctx, BNZ_EINVAL, "d_out_cap too small");
    uint8_t *dst = static_cast<uint8_t *>(d_out);

    // every lane packs at its bit phase and copies device-to-device; a word shared by two lanes
    // is merged through the host (4 bytes)
    std::vector<uint32_t> first_word(shards.size(), 0);
    for (size_t g = 0; g < shards.size(); g++) {
        Shard &sh = shards[g];
        if (sh.blocks.empty()) continue;
        Device &dl = *sh.d;
        size_t bytes = 0;
        rc = shard_pack(ctx, sh, &bytes);
        if (rc != BNZ_OK) return rc;
        const size_t w0 = (size_t)(sh.bit_base >> 5) * 4;
        if (g == 0) {
            CK(ctx, cudaMemcpyAsync(dst + w0, dl.out.p, bytes, cudaMemcpyDeviceToDevice, dl.stream));
        } else {
            CK(ctx, cudaMemcpyAsync(&first_word[g], dl.out.p, 4, cudaMemcpyDeviceToHost, dl.stream));
            if (bytes > 4)
                CK(ctx, cudaMemcpyAsync(dst + w0 + 4, dl.out.as<uint8_t>() + 4, bytes - 4, cudaMemcpyDeviceToDevice, dl.stream));
        }
    }
    for (Shard &sh : shards)
        if (!sh.blocks.empty()) CK(ctx, cudaStreamSynchronize(sh.d->stream));
    for (size_t g = 1; g < shards.size(); g++) {
        if (shards[g].blocks.empty()) continue;
        uint8_t *w = dst + (size_t)(shards[g].bit_base >> 5) * 4;
        uint32_t cur = 0;
        if (shards[g].bit_base & 31) {
            CK(ctx, cudaMemcpyAsync(&cur, w, 4, cudaMemcpyDeviceToHost, d.stream));
            CK(ctx, cudaStreamSynchronize(d.stream));
        }
        cur |= first_word[g];
        CK(ctx, cudaMemcpyAsync(w, &cur, 4, cudaMemcpyHostToDevice, d.stream));
        CK(ctx, cudaStreamSynchronize(d.stream));
    }
    // header, and the footer patched over the last partial byte
    uint8_t tail[16] = { 0 };
    const uint64_t tb = total_bits & 7;
    const size_t last = (size_t)(total_bits >> 3);
    uint8_t lastbyte = 0;
    if (tb) {
        CK(ctx, cudaMemcpyAsync(&lastbyte, dst + last, 1, cudaMemcpyDeviceToHost, d.stream));
        CK(ctx, cudaStreamSynchronize(d.stream));
    }
    tail[0] = lastbyte;
    put_bits_host(tail, tb, 0x177245385090ull, 48);
    put_bits_host(tail, tb + 48, fold_stream_crc(crcs), 32);
    const uint8_t head[4] = { 0x42, 0x5A, 0x68, (uint8_t)('0' + level) };
    CK(ctx, cudaMemcpyAsync(dst, head, 4, cudaMemcpyHostToDevice, d.stream));
    CK(ctx, cudaMemcpyAsync(dst + last, tail, nbytes - last, cudaMemcpyHostToDevice, d.stream));
    for (Shard &sh : shards)
        if (!sh.blocks.empty()) CK(ctx, cudaEventRecord(sh.d->ev[7], sh.d->stream));
    CK(ctx, cudaStreamSynchronize(d.stream));
    finish_stats(ctx, shards, false);
    ctx->stats.out_bytes = nbytes;
    *out_len = nbytes;
    return BNZ_OK;
}

// ---------------------------------------------------------------------------------------
// streaming front end: the BufRead -> BufWriter shape of banzai::encode
// (lib/lib.rs:84-132, refill loop lib/rle.rs:43-91) without holding the input or the stream in
// memory.  The caller fills pinned windows (reserve/commit); a worker thread runs each full
// window through encode_all while the caller reads the next one; finished stream bytes go back
// to the caller's thread, which hands them to the sink.  A window's trailing partial block is
// carried into the headroom in front of the next window, so the bytes are those of one
// bnz_encode over the whole input.
// ---------------------------------------------------------------------------------------

struct bnz_stream {
    bnz_ctx *ctx = nullptr;
    int level = 9;
    bnz_sink_fn sink = nullptr;
    void *user = nullptr;
    size_t window = 0;            // new input bytes per window
    size_t head = 0;              // headroom >= the longest input one block can consume
    PinBuf in[2];
    int cur = 0;                  // window being filled by the caller
    size_t fill = 0;              // its new bytes: in[cur][head, head + fill)
    size_t total_in = 0;
    bool header_sent = false, finished = false, started = false;

    std::thread th;
    std::mutex mu;
    std::condition_variable cv;
    bool job_ready = false, busy = false, quit = false;
    int job_buf = 0;
    size_t job_len = 0;
    bool job_final = false;
    int rc = BNZ_OK;              // first failure (worker or sink)

    static constexpr int NOUT = 2;
    PinBuf out[NOUT];
    bool out_free[NOUT] = { true, true };
    std::deque<std::pair<int, size_t>> ready;    // (out buffer, bytes) in stream order

    // worker-private
    size_t tail_len = 0;          // unencoded bytes of the previous window, in front of the next
    uint64_t total_bits = 32;
    std::vector<uint32_t> crcs;
    uint8_t carry[4] = { 0, 0, 0, 0 };           // the stream's last, partially filled 32-bit word
};

static int stream_job(bnz_stream *s, int b, size_t len, bool fin)
{
    bnz_ctx *ctx = s->ctx;
    const size_t N = s->tail_len + len;
    const uint8_t *base = N ? s->in[b].as<uint8_t>() + s->head - s->tail_len : nullptr;
    std::vector<Shard> shards;
    uint64_t used = 0, bits_after = s->total_bits;
    if (N > 0) {
        int rc = encode_all(ctx, base, nullptr, N, s->level, shards, s->crcs, &bits_after, fin, s->total_bits, &used);
        if (rc != BNZ_OK) return rc;
    }
    if (fin) used = N;
    else if (shards.empty()) return fail(ctx, BNZ_EINTERNAL, "stream window shorter than one block");
    const size_t tail = N - (size_t)used;
    if (tail > s->head) return fail(ctx, BNZ_EINTERNAL, "stream tail exceeds the headroom");
    if (tail) memcpy(s->in[b ^ 1].as<uint8_t>() + s->head - tail, base + used, tail);
    s->tail_len = tail;

    // stream bytes of this window: words [total_bits/32, ...) ; the first word may be shared
    const uint64_t w_first = s->total_bits >> 5;
    const uint64_t end_bits = fin ? bits_after + 80 : bits_after;
    const size_t span = (size_t)(((end_bits + 31) >> 5) - w_first) * 4 + 16;
    int ob = -1;
    {
        std::unique_lock<std::mutex> lk(s->mu);
        s->cv.wait(lk, [&] { return s->quit || s->rc != BNZ_OK || s->out_free[0] || s->out_free[1]; });
        if (s->quit || s->rc != BNZ_OK) return s->rc;
        ob = s->out_free[0] ? 0 : 1;
        s->out_free[ob] = false;
    }
    CK(ctx, s->out[ob].ensure(span));
    uint8_t *o = s->out[ob].as<uint8_t>();
    memcpy(o, s->carry, 4);
    if (!shards.empty()) {
        int rc = pack_and_download(ctx, shards, o, false, (size_t)w_first * 4);
        if (rc != BNZ_OK) return rc;
        finish_stats(ctx, shards, true);
    }
    const size_t written = (size_t)(((bits_after + 31) >> 5) - w_first) * 4;
    memset(o + std::max<size_t>(written, 4), 0, span - std::max<size_t>(written, 4));
    size_t emit;
    if (fin) {
        const uint64_t rel = bits_after - w_first * 32;
        put_bits_host(o, rel, 0x177245385090ull, 48);                    // lib.rs:66-70
        put_bits_host(o, rel + 48, fold_stream_crc(s->crcs), 32);
        emit = (size_t)((rel + 80 + 7) / 8);                                // out.rs:22-28 zero padding
        ctx->stats.out_bytes = (size_t)w_first * 4 + emit;
    } else {
        emit = (size_t)((bits_after >> 5) - w_first) * 4;
        memset(s->carry, 0, 4);
        if (bits_after & 31) memcpy(s->carry, o + emit, 4);
    }
    s->total_bits = bits_after;
    {
        std::lock_guard<std::mutex> lk(s->mu);
        s->ready.emplace_back(ob, emit);
    }
    s->cv.notify_all();
    return BNZ_OK;
}

static void stream_worker(bnz_stream *s)
{
    std::unique_lock<std::mutex> lk(s->mu);
    for (;;) {
        s->cv.wait(lk, [&] { return s->job_ready || s->quit; });
        if (s->quit) return;
        const int b = s->job_buf;
        const size_t len = s->job_len;
        const bool fin = s->job_final;
        s->job_ready = false;
        s->busy = true;
        int rc = s->rc;
        lk.unlock();
        if (rc == BNZ_OK) rc = stream_job(s, b, len, fin);
        lk.lock();
        s->busy = false;
        if (rc != BNZ_OK && s->rc == BNZ_OK) s->rc = rc;
        s->cv.notify_all();
    }
}

// caller's thread, lock held: hand finished stream bytes to the sink
static int stream_drain(bnz_stream *s, std::unique_lock<std::mutex> &lk)
{
    while (!s->ready.empty() && s->rc == BNZ_OK) {
        const std::pair<int, size_t> c = s->ready.front();
        s->ready.pop_front();
        lk.unlock();
        int e = 0;
        if (!s->header_sent) {
            const uint8_t hdr[4] = { 0x42, 0x5A, 0x68, (uint8_t)('0' + s->level) };       // lib.rs:18-22
            e = s->sink(s->user, hdr, 4);
            s->header_sent = true;
        }
        if (!e && c.second) e = s->sink(s->user, s->out[c.first].as<uint8_t>(), c.second);
        lk.lock();
        s->out_free[c.first] = true;
        if (e) {
            s->rc = BNZ_EIO;
            s->ctx->err = "sink failed";
        }
        s->cv.notify_all();
    }
    return s->rc;
}

static int stream_submit(bnz_stream *s, bool fin)
{
    bnz_ctx *ctx = s->ctx;
    if (!fin) CK(ctx, s->in[s->cur ^ 1].ensure(s->head + s->window));      // receives this window's tail
    std::unique_lock<std::mutex> lk(s->mu);
    for (;;) {
        int rc = stream_drain(s, lk);
        if (rc != BNZ_OK) return rc;
        if (!s->busy && !s->job_ready) break;
        s->cv.wait(lk);
    }
    if (!s->started) {
        s->th = std::thread(stream_worker, s);
        s->started = true;
    }
    s->job_buf = s->cur;
    s->job_len = s->fill;
    s->job_final = fin;
    s->job_ready = true;
    s->cur ^= 1;
    s->fill = 0;
    s->cv.notify_all();
    return BNZ_OK;
}

extern "C" int bnz_stream_open(bnz_ctx *ctx, int level, bnz_sink_fn sink, void *user, bnz_stream **out)
{
    if (!ctx || !sink || !out) return BNZ_EINVAL;
    *out = nullptr;
    if (level < 1 || level > 9) return fail(ctx, BNZ_EINVAL, "level must be in 1..=9 (lib/lib.rs:89)");
    if (ctx->open_streams) return fail(ctx, BNZ_EINVAL, "the context already has an open stream");
    bnz_stream *s = new bnz_stream();
    s->ctx = ctx;
    s->level = level;
    s->sink = sink;
    s->user = user;
    // one block consumes at most 255 input bytes per 5 RLE1 bytes (lib/rle.rs:211-223)
    s->head = (((size_t)100000 * level / 5 + 1) * 255 + 4096 + 4095) & ~(size_t)4095;
    s->window = std::max(ctx->stream_window_bytes, s->head);
    memset(&ctx->stats, 0, sizeof ctx->stats);
    ctx->open_streams++;
    *out = s;
    return BNZ_OK;
}

extern "C" int bnz_stream_reserve(bnz_stream *s, uint8_t **buf, size_t *cap)
{
    if (!s || !buf || !cap || s->finished) return BNZ_EINVAL;
    bnz_ctx *ctx = s->ctx;
    if (s->fill == s->window) {
        int rc = stream_submit(s, false);
        if (rc != BNZ_OK) return rc;
    }
    PinBuf &w = s->in[s->cur];
    if (w.cap < s->head + s->fill + 1) {
        // first window: grow geometrically so that small inputs do not pin a whole window
        const size_t room = std::min(s->window, std::max<size_t>((size_t)4 << 20, s->fill * 4));
        PinBuf nw;
        CK(ctx, nw.ensure(s->head + room));
        if (s->fill) memcpy(nw.as<uint8_t>() + s->head, w.as<uint8_t>() + s->head, s->fill);
        w.release();
        w = nw;
    }
    *buf = w.as<uint8_t>() + s->head + s->fill;
    *cap = std::min(w.cap - s->head, s->window) - s->fill;
    return BNZ_OK;
}

extern "C" int bnz_stream_commit(bnz_stream *s, size_t n)
{
    if (!s || s->finished) return BNZ_EINVAL;
    PinBuf &w = s->in[s->cur];
    if (n > std::min(w.cap > s->head ? w.cap - s->head : 0, s->window) - s->fill) return BNZ_EINVAL;
    s->fill += n;
    s->total_in += n;
    if (s->fill == s->window) return stream_submit(s, false);      // start the window right away
    if (s->started) {                                              // pass on whatever is finished
        std::unique_lock<std::mutex> lk(s->mu);
        return stream_drain(s, lk);
    }
    return BNZ_OK;
}

extern "C" int bnz_stream_write(bnz_stream *s, const uint8_t *data, size_t len)
{
    if (!s || (len && !data)) return BNZ_EINVAL;
    while (len) {
        uint8_t *p = nullptr;
        size_t cap = 0;
        int rc = bnz_stream_reserve(s, &p, &cap);
        if (rc != BNZ_OK) return rc;
        const size_t n = std::min(cap, len);
        memcpy(p, data, n);
        rc = bnz_stream_commit(s, n);
        if (rc != BNZ_OK) return rc;
        data += n;
        len -= n;
    }
    return BNZ_OK;
}

extern "C" int bnz_stream_finish(bnz_stream *s, size_t *consumed)
{
    if (!s || s->finished) return BNZ_EINVAL;
    if (consumed) *consumed = 0;
    int rc = stream_submit(s, true);
    if (rc != BNZ_OK) return rc;
    s->finished = true;
    std::unique_lock<std::mutex> lk(s->mu);
    for (;;) {
        rc = stream_drain(s, lk);
        if (rc != BNZ_OK) return rc;
        if (!s->busy && !s->job_ready && s->ready.empty()) break;
        s->cv.wait(lk);
    }
    s->ctx->stats.in_bytes = s->total_in;
    if (consumed) *consumed = s->total_in;
    return BNZ_OK;
}

extern "C" void bnz_stream_close(bnz_stream *s)
{
    if (!s) return;
    if (s->started) {
        {
            std::lock_guard<std::mutex> lk(s->mu);
            s->quit = true;
        }
        s->cv.notify_all();
        s->th.join();
    }
    for (PinBuf &b : s->in) b.release();
    for (PinBuf &b : s->out) b.release();
    s->ctx->open_streams--;
    delete s;
}

static int file_sink(void *user, const uint8_t *data, size_t len)
{
    return fwrite(data, 1, len, static_cast<FILE *>(user)) == len ? 0 : 1;
}

extern "C" int bnz_encode_file(bnz_ctx *ctx, const char *in_path, const char *out_path, size_t *consumed)
{
    if (!ctx || !in_path || !out_path) return BNZ_EINVAL;
    if (consumed) *consumed = 0;
    FILE *f = fopen(in_path, "rb");
    if (!f) return fail(ctx, BNZ_EIO, std::string("cannot open ") + in_path);
    FILE *g = fopen(out_path, "wb");
    if (!g) {
        fclose(f);
        return fail(ctx, BNZ_EIO, std::string("cannot create ") + out_path);
    }
    bnz_stream *s = nullptr;
    int rc = bnz_stream_open(ctx, 9, file_sink, g, &s);                    // lib.rs:152: level 9
    while (rc == BNZ_OK) {
        uint8_t *p = nullptr;
        size_t cap = 0;
        rc = bnz_stream_reserve(s, &p, &cap);
        if (rc != BNZ_OK) break;
        const size_t got = fread(p, 1, std::min<size_t>(cap, (size_t)8 << 20), f);
        if (got == 0) {
            if (ferror(f)) rc = fail(ctx, BNZ_EIO, std::string("read error on ") + in_path);
            break;
        }
        rc = bnz_stream_commit(s, got);
    }
    if (rc == BNZ_OK) rc = bnz_stream_finish(s, consumed);
    bnz_stream_close(s);
    fclose(f);
    if (fclose(g) != 0 && rc == BNZ_OK) rc = fail(ctx, BNZ_EIO, "short write");
    return rc;
}
