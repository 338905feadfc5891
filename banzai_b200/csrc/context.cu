// context.cu — context, tunables, error text and memory helpers of libbanzai_b200.so
#include "host.h"

thread_local std::string *t_err_sink = nullptr;
void set_err(bnz_ctx *ctx, const std::string &msg)
{
    if (t_err_sink) *t_err_sink = msg;
    else if (ctx) ctx->err = msg;
}


int fail(bnz_ctx *ctx, int code, const std::string &msg)
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
    case BNZ_EVERIFY: return "self-verification failed";
    default: return "unknown error";
    }
}

extern "C" const char *bnz_last_error(const bnz_ctx *ctx) { return ctx ? ctx->err.c_str() : ""; }

// streams, events of one lane on device `id` (arenas grow on demand)
bool device_init(Device &d, int id)
{
    d.id = id;
    cudaDeviceProp prop;
    int prio_lo = 0, prio_hi = 0;
    bool ok = cudaSetDevice(d.id) == cudaSuccess && cudaGetDeviceProperties(&prop, d.id) == cudaSuccess &&
              prop.major == 10 && prop.minor == 0 &&   // the library holds sm_100a SASS only (no PTX): fail loudly elsewhere
              cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi) == cudaSuccess &&
              // the sort's persistent CTAs must win every SM slot over the work that fills its tail
              cudaStreamCreateWithPriority(&d.stream, cudaStreamNonBlocking, prio_hi) == cudaSuccess &&
              cudaStreamCreateWithPriority(&d.stream2, cudaStreamNonBlocking, prio_hi) == cudaSuccess &&
              cudaStreamCreateWithPriority(&d.stream3[0], cudaStreamNonBlocking, prio_lo) == cudaSuccess &&
              cudaStreamCreateWithPriority(&d.stream3[1], cudaStreamNonBlocking, prio_lo) == cudaSuccess &&
              cudaStreamCreateWithPriority(&d.stream3[2], cudaStreamNonBlocking, prio_lo) == cudaSuccess;
    for (cudaEvent_t &e : d.ev)
        if (ok && cudaEventCreate(&e) != cudaSuccess) ok = false;
    if (ok) d.sm_count = prop.multiProcessorCount;
    return ok;
}

void device_release(Device &d)
{
    cudaSetDevice(d.id);
    if (d.stream) cudaStreamSynchronize(d.stream);
    for (DevBuf *b : { &d.in, &d.rle, &d.bwt, &d.blk_off, &d.blk_len, &d.ptr, &d.has_byte,
                       &d.bwt_stats, &d.counters, &d.ws_rec, &d.ws_rank, &d.ws_ctl, &d.ws_hist, &d.ws_rec2, &d.ws_rank2, &d.ws_defer, &d.v_marks, &d.v_lfl, &d.v_flags, &d.sel, &d.ch_lasthead, &d.ch_meta,
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
    d.h_tiles.release();
    if (d.stream) cudaStreamDestroy(d.stream);
    if (d.stream2) cudaStreamDestroy(d.stream2);
    for (cudaStream_t st : d.stream3)
        if (st) cudaStreamDestroy(st);
}

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
        if (!device_init(ctx->devs.back(), device_ids[i])) {
            bnz_ctx_destroy(ctx);
            return BNZ_ECUDA;
        }
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
    for (Device &d : ctx->devs) device_release(d);
    for (Device *a : ctx->aux) {
        device_release(*a);
        delete a;
    }
    if (ctx->out_cache) cudaFreeHost(ctx->out_cache);
    free(ctx->out_big);
    delete ctx;
}

extern "C" int bnz_ctx_set(bnz_ctx *ctx, const char *key, long value)
{
    if (!ctx || !key) return BNZ_EINVAL;
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
    if (!strcmp(key, "piece_blocks_per_sm_x16")) {
        if (value < 1 || value > 64) return BNZ_EINVAL;
        ctx->piece_blocks_per_sm_x16 = (int)value;
        return BNZ_OK;
    }
    if (!strcmp(key, "h2d_pieces")) {
        if (value < 2 || value > 8) return BNZ_EINVAL;
        ctx->h2d_pieces = (int)value;
        return BNZ_OK;
    }
    if (!strcmp(key, "h2d_overlap")) {
        if (value < 0 || value > 2) return BNZ_EINVAL;
        ctx->h2d_overlap = (int)value;      // 0 off | 1 auto | 2 forced even for small inputs (tests)
        return BNZ_OK;
    }
    if (!strcmp(key, "huff_literal")) {
        ctx->huff_literal = value != 0;
        return BNZ_OK;
    }
    if (!strcmp(key, "verify")) {
        ctx->verify = value != 0;
        return BNZ_OK;
    }
    if (!strcmp(key, "verify_corrupt")) {
        if (value < 0 || value > 4) return BNZ_EINVAL;
        ctx->verify_corrupt = (int)value;
        return BNZ_OK;
    }
    if (!strcmp(key, "reuse_input")) {
        ctx->reuse_input = value != 0;
        for (Device &d : ctx->devs) d.tag(nullptr, 0, 0, 0);
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
    if (!strcmp(key, "bwt_cluster_below")) {
        if (value < 0) return BNZ_EINVAL;
        ctx->bwt_cluster_below = (int)value;
        return BNZ_OK;
    }
    if (!strcmp(key, "bwt_periodic")) {
        if (value != 0 && value != 1) return BNZ_EINVAL;
        ctx->bwt_periodic = (int)value;
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
