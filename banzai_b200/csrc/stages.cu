// stages.cu — the stages on one device (device pointers in, device pointers out) and the stage
// exports of the C ABI (parity seams mirroring rle::rle_one, bwt::bwt, mtf::mtf_and_rle,
// huffman::encode).
#include "host.h"

// ---------------------------------------------------------------------------------------
// BWT stage on one device (device pointers in, device pointers out)
// ---------------------------------------------------------------------------------------

int run_bwt_device(bnz_ctx *ctx, Device &d, const uint8_t *d_rle, uint8_t *d_bwt,
                          const uint64_t *d_blk_off, const uint32_t *d_blk_len, uint32_t n_blocks,
                          uint32_t max_len, uint32_t *d_ptr, uint8_t *d_has_byte, BwtStats *d_stats,
                          uint32_t *d_done, bool *done_armed, uint32_t *d_marks)
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
    a.done = nullptr;
    a.marks = d_marks;
    a.defer_list = a.defer_count = nullptr;
    a.blk_list = a.n_blocks_dev = nullptr;

    // auto: many blocks -> one persistent CTA per block (best aggregate throughput);
    // few blocks -> one cluster per block so that every SM has work and the randomly accessed
    // arrays stay in L2 (measured crossover ~115 blocks per device with the final kernels of round 2,
    // profiles/r2_experiments/cluster_vs_one_cta_by_blocks.txt)
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
        if (d_done && n_blocks > (uint32_t)n_clusters) {     // per-block completion flags (the queue has a tail)
            a.done = d_done;
            if (done_armed) *done_armed = true;
        }
        // blocks with a long periodic run go to a follow-up launch of the one-CTA kernel (bwt_common.cuh: Period;
        // the launch is unconditional — its block count lives in device memory — and exits at once when the
        // list is empty, so no host round trip is needed)
        // Only when the clusters need more than one wave: a single wave of clusters does the ~19 rounds of such a
        // block in about the time one CTA alone needs for its two (measured, 2 blocks: 7.2 against 12.2 ms).
        const bool defer = ctx->bwt_periodic && max_len >= 32768u /* PERIOD_MIN_N */ && n_blocks > (uint32_t)n_clusters;
        int grid2 = 0;
        size_t stride2 = 0;
        if (defer) {
            int per_sm = 0;
            CK(ctx, bwt_max_ctas(&per_sm));                     // (also sets the kernel's shared-memory attribute)
            if (per_sm <= 0) return fail(ctx, BNZ_ECUDA, "bwt kernel does not fit on an SM");
            grid2 = (int)std::min<uint64_t>((uint64_t)n_blocks, (uint64_t)d.sm_count);
            stride2 = ((size_t)max_len + 8191) & ~(size_t)8191;
            CK(ctx, d.ws_defer.ensure((size_t)n_blocks * sizeof(uint32_t)));
            CK(ctx, d.ws_rec2.ensure((size_t)grid2 * 3 * stride2 * sizeof(uint64_t)));
            CK(ctx, d.ws_rank2.ensure((size_t)grid2 * stride2 * sizeof(uint32_t)));
            CK(ctx, d.ws_hist.ensure((size_t)grid2 * BWT_HIST_WORDS * 4));
            a.defer_list = d.ws_defer.as<uint32_t>();
            a.defer_count = d.counters.as<uint32_t>() + 2;
        }
        CK(ctx, bwtc_launch(a, ctx->bwt_threads, C, n_clusters, d.stream));
        d.launches++;
        if (defer) {
            BwtArgs b = a;
            b.blk_list = a.defer_list;
            b.n_blocks_dev = a.defer_count;
            b.defer_list = b.defer_count = nullptr;
            b.next_block = d.counters.as<uint32_t>() + 1;
            b.ws_ctl = nullptr;
            b.ws_hist = d.ws_hist.as<uint32_t>();
            b.ws_rec = d.ws_rec2.as<uint64_t>();
            b.ws_rank = d.ws_rank2.as<uint32_t>();
            b.ws_stride = stride2;
            CK(ctx, bwt_launch(b, grid2, d.stream, true));
            d.launches++;
        }
        return BNZ_OK;
    }

    int per_sm = 0;
    CK(ctx, bwt_max_ctas(&per_sm));
    if (per_sm <= 0) return fail(ctx, BNZ_ECUDA, "bwt kernel does not fit on an SM");
    if (ctx->ctas_per_sm > 0) per_sm = std::min(per_sm, ctx->ctas_per_sm);
    int grid = (int)std::min<uint64_t>((uint64_t)n_blocks, (uint64_t)d.sm_count * per_sm);
    size_t stride = ((size_t)max_len + 8191) & ~(size_t)8191;    // whole 8192-record buckets (apply_ranks_bucketed)
    CK(ctx, d.ws_rec.ensure((size_t)grid * 3 * stride * sizeof(uint64_t)));
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
    const bool defer = ctx->bwt_periodic && max_len >= 32768u /* PERIOD_MIN_N */;
    if (defer) {
        CK(ctx, d.ws_defer.ensure((size_t)n_blocks * sizeof(uint32_t)));
        a.defer_list = d.ws_defer.as<uint32_t>();
        a.defer_count = d.counters.as<uint32_t>() + 2;
    }
    CK(ctx, bwt_launch(a, grid, d.stream));
    d.launches++;
    if (defer) {
        // the blocks with a long periodic run, if any (the count lives in device memory: no host round
        // trip; an empty list costs one launch of CTAs that exit at once); same workspace, stream order
        BwtArgs b = a;
        b.blk_list = a.defer_list;
        b.n_blocks_dev = a.defer_count;
        b.defer_list = b.defer_count = nullptr;
        b.next_block = d.counters.as<uint32_t>() + 1;
        CK(ctx, bwt_launch(b, grid, d.stream, true));
        d.launches++;
    }
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
    cudaEvent_t e0 = d.ev[0], e1 = d.ev[1];                 // the device's own timing events (idle in a stage call)
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

    bnz_stats &s = ctx->stats;
    memset(&s, 0, sizeof s);
    s.n_blocks = (uint32_t)n_blocks;
    s.n_devices = 1;
    s.kernel_launches = 1;
    s.bwt_radix_bits = 8;
    s.bwt_ms = ms;
    for (size_t b = 0; b < n_blocks; b++) {
        s.bwt_n += st[b].n;
        s.bwt_sum_active += st[b].sum_active;
        s.bwt_sum_active_passes += st[b].sum_active_passes;
        s.bwt_sum_tile += st[b].sum_tile;
        s.bwt_cyc_tile += st[b].cyc_tile;
        s.bwt_cyc_final += st[b].cyc_final;
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
            stats_out[b].period = st[b].period;
            stats_out[b].sum_tile = st[b].sum_tile;
            stats_out[b].sum_active = st[b].sum_active;
            stats_out[b].sum_active_passes = st[b].sum_active_passes;
            stats_out[b].cycles = st[b].cyc_build + st[b].cyc_radix + st[b].cyc_rerank + st[b].cyc_tile;
        }
    }
    s.bwt_algorithmic_bytes = bwt_algorithmic_bytes(s);
    return BNZ_OK;
}

// ---------------------------------------------------------------------------------------
// RLE1 + cuts + CRC on one device.  d_in: device copy of the input, h_in: host copy (the cut
// walk reads <= 2 KiB of it per block).  Leaves the RLE1 images in d.rle and fills `blocks`
// and `crcs`.
// ---------------------------------------------------------------------------------------

// K1 part 1 on one device: chunk tables -> host -> cut walk.  Leaves P / o_in in d.h_P / d.h_oin
// (host, pinned) and in d.ch_P / d.ch_oin (device).
int rle_plan(bnz_ctx *ctx, Device &d, const uint8_t *d_in, const uint8_t *h_in, uint64_t N, int level,
             std::vector<RleBlock> &blocks, bool final, uint64_t *consumed)
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
int rle_emit_shard(bnz_ctx *ctx, Device &d, const uint8_t *in_base, uint64_t N, const uint64_t *oin_base,
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

// The host part of K1 on its own (no device needed): the sequential cut chain over chunk tables
// the caller supplies.  Lets the cut rule (lib/rle.rs:121-240, SURVEY A-Q1) be checked on a CPU.
extern "C" int bnz_host_cut_chain(const uint8_t *in, size_t in_len, int level, const uint64_t *P,
                                  const uint64_t *o_in, size_t n_chunks, int final, uint64_t *blk_in_off,
                                  uint64_t *blk_in_len, uint32_t *blk_rle_len, size_t max_blocks, size_t *n_blocks,
                                  size_t *consumed)
{
    if (!n_blocks || level < 1 || level > 9 || (in_len && (!in || !P || !o_in))) return BNZ_EINVAL;
    *n_blocks = 0;
    if (consumed) *consumed = 0;
    if (in_len == 0) return BNZ_OK;
    if (n_chunks != (in_len + RLE_CHUNK - 1) / RLE_CHUNK) return BNZ_EINVAL;
    std::vector<RleBlock> blocks;
    uint64_t used = 0;
    if (rle_walk_cuts(in, in_len, level, P, o_in, n_chunks, blocks, final != 0, &used) != 0) return BNZ_EINTERNAL;
    if (blocks.size() > max_blocks) return BNZ_EINVAL;
    for (size_t b = 0; b < blocks.size(); b++) {
        if (blk_in_off) blk_in_off[b] = blocks[b].s;
        if (blk_in_len) blk_in_len[b] = blocks[b].c - blocks[b].s;
        if (blk_rle_len) blk_rle_len[b] = blocks[b].n;
    }
    *n_blocks = blocks.size();
    if (consumed) *consumed = (size_t)used;
    return BNZ_OK;
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


template <class X>
static cudaError_t upload(DevBuf &buf, const std::vector<X> &v, cudaStream_t st)
{
    cudaError_t e = buf.ensure(v.size() * sizeof(X) + 16);
    if (e != cudaSuccess) return e;
    return cudaMemcpyAsync(buf.p, v.data(), v.size() * sizeof(X), cudaMemcpyHostToDevice, st);
}

int upload_batch(bnz_ctx *ctx, Device &d, const Batch &bt)
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
int mtf_ensure(bnz_ctx *ctx, Device &d, const Batch &bt)
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
int run_mtf_list(bnz_ctx *ctx, Device &d, const Batch &bt, const uint8_t *d_bwt, uint8_t *d_idx,
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
int run_huff_model_device(bnz_ctx *ctx, Device &d, const Batch &bt, int level, int with_block_header,
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
    if (ctx->huff_literal) {
        // the reference's modelling loop, literally (one selector byte per 50-symbol group)
        a.sel_stride = (size_t)(100000 * level + 1 + 49) / 50 + 16;
        CK(ctx, d.sel.ensure((size_t)nb * a.sel_stride));
        a.sel_out = d.sel.as<uint8_t>();
        a.selectors = d.sel.as<uint8_t>();
        HuffArgs run = a;
        CK(ctx, huff_launch_literal(run, bt.span_base[nb], d.stream, &d.launches));
        return BNZ_OK;
    }
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
