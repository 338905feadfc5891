// encode.cu — bnz_encode / bnz_encode_device: the B200 replacement of the block loop in
// lib/lib.rs:101-126 (one batch of blocks per call of encode_all, sharded over the devices).
#include "host.h"

// ---------------------------------------------------------------------------------------
// bnz_encode: the whole path
// ---------------------------------------------------------------------------------------

void put_bits_host(uint8_t *buf, uint64_t bitpos, uint64_t value, int nbits)   // MSB first
{
    for (int i = nbits - 1; i >= 0; i--, bitpos++)
        if ((value >> i) & 1) buf[bitpos >> 3] |= (uint8_t)(0x80u >> (bitpos & 7));
}

// One device's share of a bnz_encode call: a contiguous range of blocks.
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
    const bool verify = ctx->verify != 0;
    if (verify) {
        CK(ctx, d.v_marks.ensure((size_t)nb * VERIFY_MARKS * 4));
        CK(ctx, d.v_lfl.ensure((size_t)rle_total * 4 + 64));
        CK(ctx, d.v_flags.ensure((size_t)nb * 4));
        CK(ctx, cudaMemsetAsync(d.v_flags.p, 0, (size_t)nb * 4, d.stream));
        CK(ctx, cudaMemsetAsync(d.v_marks.p, 0xff, (size_t)nb * VERIFY_MARKS * 4, d.stream));
    }
    // (verify: the MTF must not overwrite the RLE1 images while they are being checked, so nothing
    // runs beside the sort)
    rc = run_bwt_device(ctx, d, d.rle.as<uint8_t>(), d.bwt.as<uint8_t>(), d.blk_off.as<uint64_t>(),
                        d.blk_len.as<uint32_t>(), nb, bt.max_len, d.ptr.as<uint32_t>(), d.has_byte.as<uint8_t>(),
                        d.bwt_stats.as<BwtStats>(), (ctx->mtf_overlap > 0 && !verify) ? d_done : nullptr, &armed,
                        verify ? d.v_marks.as<uint32_t>() : nullptr);
    if (rc != BNZ_OK) return rc;
    CK(ctx, cudaEventRecord(d.ev[3], d.stream));
    if (verify) {
        VerifyArgs va;
        va.in_base = in_base;
        va.rle = d.rle.as<uint8_t>();
        va.bwt = d.bwt.as<uint8_t>();
        va.ptr = d.ptr.as<uint32_t>();
        va.blk_off = d.blk_off.as<uint64_t>();
        va.blk_len = d.blk_len.as<uint32_t>();
        va.blocks = d.rle_blocks.as<RleBlock>();
        va.n_blocks = nb;
        va.lfl = d.v_lfl.as<uint32_t>();
        va.marks = d.v_marks.as<uint32_t>();
        va.stats = d.bwt_stats.as<BwtStats>();
        va.flags = d.v_flags.as<uint32_t>();
        va.corrupt = ctx->verify_corrupt >= 1 && ctx->verify_corrupt <= 3 ? ctx->verify_corrupt : 0;
        CK(ctx, verify_launch(va, d.stream, &d.launches));
        sh.vflags.assign(nb, 0);
        CK(ctx, cudaMemcpyAsync(sh.vflags.data(), d.v_flags.p, (size_t)nb * 4, cudaMemcpyDeviceToHost, d.stream));
    }

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
    for (uint32_t b = 0; b < (uint32_t)sh.vflags.size(); b++)
        if (sh.vflags[b])
            return fail(ctx, BNZ_EVERIFY, std::string("block at input offset ") + std::to_string(sh.blocks[b].s) +
                                              ((sh.vflags[b] & VERIFY_BAD_RLE) ? ": the RLE1 image does not decode to the input" : "") +
                                              ((sh.vflags[b] & VERIFY_BAD_BWT) ? ": the inverse BWT is not the RLE1 image" : ""));
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
        st.bwt_sum_tile += b.sum_tile;
        st.bwt_cyc_tile += b.cyc_tile;
        st.bwt_cyc_final += b.cyc_final;
        st.bwt_rounds_total += b.rounds;
        st.bwt_max_rounds = std::max(st.bwt_max_rounds, b.rounds);
        st.bwt_tied_blocks += b.tied;
        st.bwt_cyc_build += b.cyc_build;
        st.bwt_cyc_radix += b.cyc_radix;
        st.bwt_cyc_rerank += b.cyc_rerank;
    }
    st.bwt_algorithmic_bytes = bwt_algorithmic_bytes(st);
    st.kernel_launches += sh.d->launches;
}

void finish_stats(bnz_ctx *ctx, std::vector<Shard> &shards, bool have_d2h)
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
    uint32_t busy = 0;
    for (const Shard &sh : shards) busy += sh.blocks.empty() ? 0u : 1u;
    st.n_devices = std::max<uint32_t>(st.n_devices, busy);
    st.bwt_radix_bits = 8;
}

uint32_t fold_stream_crc(const std::vector<uint32_t> &crcs)       // lib.rs:108
{
    uint32_t s = 0;
    for (uint32_t c : crcs) s = c ^ ((s << 1) | (s >> 31));
    return s;
}

static int ensure_out_cache(bnz_ctx *ctx, size_t nbytes, bool exact = false)
{
    if (ctx->out_cache_cap < nbytes + 16) {
        if (ctx->out_cache) cudaFreeHost(ctx->out_cache);
        ctx->out_cache = nullptr;
        ctx->out_cache_cap = 0;
        size_t want = exact ? nbytes + 4096 : nbytes + nbytes / 4 + 4096;
        CK(ctx, cudaHostAlloc((void **)&ctx->out_cache, want, cudaHostAllocPortable));
        ctx->out_cache_cap = want;
    }
    return BNZ_OK;
}


// ---------------------------------------------------------------------------------------
// "verify", host part: the cut chain must cover in[0, covered) without gap or overlap, and the
// block CRCs the device computed must equal an independent recomputation on the host cores
// (plain table-driven CRC-32/BZIP2, MSB first, lib/crc32.rs:31-48).
// ---------------------------------------------------------------------------------------
static uint32_t host_crc32_bzip2(const uint8_t *p, size_t n)
{
    static uint32_t tab[256];
    static std::once_flag once;
    std::call_once(once, [] {
        for (uint32_t b = 0; b < 256; b++) {
            uint32_t c = b << 24;
            for (int k = 0; k < 8; k++) c = (c & 0x80000000u) ? (c << 1) ^ 0x04C11DB7u : (c << 1);
            tab[b] = c;
        }
    });
    uint32_t crc = 0xFFFFFFFFu;
    for (size_t i = 0; i < n; i++) crc = (crc << 8) ^ tab[(crc >> 24) ^ p[i]];
    return ~crc;
}

int verify_host(bnz_ctx *ctx, const uint8_t *h_in, uint64_t covered, const std::vector<Shard> &shards)
{
    struct Item { uint64_t s, c; uint32_t crc; };
    std::vector<Item> items;
    for (const Shard &sh : shards)
        for (size_t b = 0; b < sh.blocks.size(); b++) items.push_back({ sh.blocks[b].s, sh.blocks[b].c, sh.crcs[b] });
    uint64_t at = 0;
    for (const Item &it : items) {
        if (it.s != at || it.c <= it.s) return fail(ctx, BNZ_EVERIFY, "the block cuts do not tile the input");
        at = it.c;
    }
    if (at != covered) return fail(ctx, BNZ_EVERIFY, "the block cuts do not cover the input");
    std::atomic<size_t> next{0};
    std::atomic<long long> bad{-1};
    const unsigned nt = std::max(1u, std::min<unsigned>(std::thread::hardware_concurrency(), (unsigned)items.size()));
    std::vector<std::thread> th;
    for (unsigned t = 0; t < nt; t++)
        th.emplace_back([&] {
            for (size_t k; (k = next.fetch_add(1)) < items.size();) {
                uint32_t c = host_crc32_bzip2(h_in + items[k].s, (size_t)(items[k].c - items[k].s));
                if (ctx->verify_corrupt == 4 && k == 0) c ^= 1u;           // test hook
                if (c != items[k].crc) bad.store((long long)k);
            }
        });
    for (std::thread &t : th) t.join();
    if (bad.load() >= 0)
        return fail(ctx, BNZ_EVERIFY, "block CRC differs from the host recomputation at input offset " +
                                          std::to_string(items[(size_t)bad.load()].s));
    return BNZ_OK;
}

// ---------------------------------------------------------------------------------------
// Several devices, one stream (block independence: lib/lib.rs:101-126).
// Every device uploads ONLY its own 1/G byte range and builds the chunk tables of that range; the
// ranges are coupled by two scalars (the last run head before the range and the cost prefix at its
// start), which the device threads exchange through host memory between the three table steps - no
// collective, nothing is uploaded twice, no device waits for another one's bytes.  One thread then
// walks the sequential cut chain over the stitched tables; a device owns the blocks that START in
// its range and fetches the few bytes (and table entries) by which its last block reaches into
// the next range.  From there on the devices are independent until the bit offsets are known.
// ---------------------------------------------------------------------------------------
namespace {
struct HostBarrier {
    std::mutex m;
    std::condition_variable cv;
    const int n;
    int waiting = 0;
    unsigned gen = 0;
    explicit HostBarrier(int n_) : n(n_) {}
    void wait()
    {
        std::unique_lock<std::mutex> lk(m);
        const unsigned g = gen;
        if (++waiting == n) {
            waiting = 0;
            gen++;
            cv.notify_all();
        } else {
            cv.wait(lk, [&] { return gen != g; });
        }
    }
};
}  // namespace

static int encode_sharded(bnz_ctx *ctx, const uint8_t *h_in, size_t N, int level, std::vector<Shard> &shards,
                          std::vector<uint32_t> &crcs, uint64_t *total_bits, bool final, uint64_t bit_base,
                          uint64_t *consumed)
{
    bnz_stats &st = ctx->stats;
    if (consumed) *consumed = 0;
    const uint64_t n_chunks = (N + RLE_CHUNK - 1) / RLE_CHUNK;
    const size_t G_all = ctx->devs.size();
    const uint64_t per = (n_chunks + G_all - 1) / G_all;                 // chunks per device
    const size_t G = (size_t)((n_chunks + per - 1) / per);               // devices that get a range
    const uint64_t blk = (uint64_t)100000 * level;
    const uint64_t slop = ((blk + blk / 4 + 4095) / 4096) * 4096;        // look-ahead uploaded with the range
    const uint64_t max_reach = 51 * blk + 2 * RLE_CHUNK;                 // the longest input one block can swallow

    Device &d0 = ctx->devs[0];
    CK(ctx, cudaSetDevice(d0.id));
    CK(ctx, d0.h_P.ensure((n_chunks + 1) * 8));                          // stitched tables (pinned, portable)
    CK(ctx, d0.h_oin.ensure(n_chunks * 8));
    uint64_t *h_P = d0.h_P.as<uint64_t>(), *h_oin = d0.h_oin.as<uint64_t>();

    shards = std::vector<Shard>(G);
    std::vector<uint64_t> head_max(G, 0), cost_sum(G, 0);
    std::vector<RleBlock> all_blocks;
    std::atomic<bool> abort_all{false};
    int walk_rc = 0;
    uint64_t used = 0;
    HostBarrier bar((int)G);
    std::mutex bits_mu;
    std::condition_variable bits_cv;
    std::vector<char> bits_known(G, 0);

    auto work = [&](size_t g) -> int {
        Shard &sh = shards[g];
        Device &d = ctx->devs[g];
        sh.d = &d;
        d.launches = 0;
        const uint64_t ca = g * per, cb = std::min(n_chunks, (g + 1) * per);
        const uint64_t nt = (cb - ca + rle_scan_tile_chunks() - 1) / rle_scan_tile_chunks();
        const uint64_t a = ca ? ca * RLE_CHUNK - 16 : 0;                 // first resident input byte
        uint64_t b = std::min<uint64_t>(N, cb * RLE_CHUNK + slop);       // one past the last resident byte
        const uint8_t *in_base = nullptr;
        uint64_t *lasthead = nullptr, *oin = nullptr, *P = nullptr, *tile_head = nullptr, *tile_sum = nullptr;
        uint32_t *meta = nullptr, *restsum = nullptr;
        int rc = BNZ_OK;
        bool resident = false;
        uint64_t uploaded = 0;
        auto step = [&](auto &&fn) {
            if (rc == BNZ_OK && !abort_all.load()) {
                rc = fn();
                if (rc != BNZ_OK) abort_all.store(true);
            }
        };
        // ---- own byte range -> device; chunk summaries and tile maxima
        step([&]() -> int {
            CK(ctx, cudaSetDevice(d.id));
            const uint64_t reach = (cb - ca) + max_reach / RLE_CHUNK + 4;      // chunks the tables may have to cover
            CK(ctx, d.in.ensure((cb * RLE_CHUNK - a) + max_reach + 4096));
            CK(ctx, d.ch_lasthead.ensure((cb - ca) * 8));
            CK(ctx, d.ch_meta.ensure((cb - ca) * 4));
            CK(ctx, d.ch_restsum.ensure((cb - ca) * 4));
            CK(ctx, d.ch_oin.ensure(reach * 8));
            CK(ctx, d.ch_P.ensure((reach + 1) * 8));
            CK(ctx, d.ch_tiles.ensure(nt * 16 + 64));
            CK(ctx, d.h_tiles.ensure(nt * 16 + 64));
            in_base = d.in.as<uint8_t>() - a;
            lasthead = d.ch_lasthead.as<uint64_t>() - ca;
            meta = d.ch_meta.as<uint32_t>() - ca;
            restsum = d.ch_restsum.as<uint32_t>() - ca;
            oin = d.ch_oin.as<uint64_t>() - ca;
            P = d.ch_P.as<uint64_t>() - ca;
            tile_head = d.ch_tiles.as<uint64_t>();
            tile_sum = tile_head + nt;
            CK(ctx, cudaEventRecord(d.ev[0], d.stream));
            if (ctx->reuse_input && d.holds(h_in, N, a) && d.tag_b >= b) {
                b = d.tag_b;                                          // resident since the previous call
                resident = true;
            } else {
                CK(ctx, cudaMemcpyAsync(d.in.p, h_in + a, b - a, cudaMemcpyHostToDevice, d.stream));
            }
            CK(ctx, cudaEventRecord(d.ev[1], d.stream));
            CK(ctx, rle_tables_heads_launch(in_base, N, ca, ca, cb, lasthead, meta, restsum, tile_head, d.stream));
            CK(ctx, cudaMemcpyAsync(d.h_tiles.p, tile_head, nt * 8, cudaMemcpyDeviceToHost, d.stream));
            CK(ctx, cudaStreamSynchronize(d.stream));
            d.launches += 2;
            for (uint64_t t = 0; t < nt; t++) head_max[g] = std::max(head_max[g], d.h_tiles.as<uint64_t>()[t]);
            return BNZ_OK;
        });
        bar.wait();
        // ---- run offsets at the chunk starts; tile cost sums
        step([&]() -> int {
            uint64_t carry_head = 0;
            for (size_t q = 0; q < g; q++) carry_head = std::max(carry_head, head_max[q]);
            CK(ctx, rle_tables_oin_launch(ca, ca, cb, carry_head, lasthead, meta, restsum, tile_head, oin, tile_sum, d.stream));
            CK(ctx, cudaMemcpyAsync(d.h_tiles.p, tile_sum, nt * 8, cudaMemcpyDeviceToHost, d.stream));
            CK(ctx, cudaStreamSynchronize(d.stream));
            d.launches += 1;
            for (uint64_t t = 0; t < nt; t++) cost_sum[g] += d.h_tiles.as<uint64_t>()[t];
            return BNZ_OK;
        });
        bar.wait();
        // ---- cost prefix; this range's slice of the stitched host tables
        step([&]() -> int {
            uint64_t carry_sum = 0;
            for (size_t q = 0; q < g; q++) carry_sum += cost_sum[q];
            CK(ctx, rle_tables_p_launch(ca, ca, cb, carry_sum, meta, restsum, oin, tile_sum, P, d.stream));
            CK(ctx, cudaMemcpyAsync(h_P + ca, P + ca, (cb - ca + 1) * 8, cudaMemcpyDeviceToHost, d.stream));
            CK(ctx, cudaMemcpyAsync(h_oin + ca, oin + ca, (cb - ca) * 8, cudaMemcpyDeviceToHost, d.stream));
            CK(ctx, cudaStreamSynchronize(d.stream));
            d.launches += 1;
            return BNZ_OK;
        });
        bar.wait();
        // ---- the sequential cut chain, once; a device owns the blocks that start in its range
        if (g == 0 && !abort_all.load()) {
            walk_rc = rle_walk_cuts(h_in, N, level, h_P, h_oin, n_chunks, all_blocks, final, &used);
            if (walk_rc == 0) {
                size_t i = 0;
                for (size_t q = 0; q < G; q++) {
                    const uint64_t end = std::min<uint64_t>(N, (q + 1) * per * RLE_CHUNK);
                    size_t j = i;
                    while (j < all_blocks.size() && all_blocks[j].s < end) j++;
                    shards[q].blocks.assign(all_blocks.begin() + i, all_blocks.begin() + j);
                    const uint64_t off0 = (j > i) ? all_blocks[i].rle_off : 0;
                    for (RleBlock &bk : shards[q].blocks) bk.rle_off -= off0;
                    i = j;
                }
            } else {
                abort_all.store(true);
            }
        }
        bar.wait();
        // ---- the part of the last block that reaches into the next range; then the block pipeline
        step([&]() -> int {
            if (sh.blocks.empty()) return BNZ_OK;
            const uint64_t c_end = (sh.blocks.back().c + RLE_CHUNK - 1) / RLE_CHUNK;
            const uint64_t e = std::min<uint64_t>(N, c_end * RLE_CHUNK + 16);
            if (e - a > d.in.cap || c_end - ca + 1 > d.ch_P.cap / 8) return fail(ctx, BNZ_EINTERNAL, "block reaches beyond the reserved look-ahead");
            if (e > b) {
                CK(ctx, cudaMemcpyAsync(d.in.as<uint8_t>() + (b - a), h_in + b, e - b, cudaMemcpyHostToDevice, d.stream));
                                b = e;
            }
            if (c_end > cb) {
                CK(ctx, cudaMemcpyAsync(oin + cb, h_oin + cb, (c_end - cb) * 8, cudaMemcpyHostToDevice, d.stream));
                CK(ctx, cudaMemcpyAsync(P + cb, h_P + cb, (c_end - cb + 1) * 8, cudaMemcpyHostToDevice, d.stream));
            }
            return shard_model(ctx, sh, in_base, N, oin, P, level);
        });
        // ---- the shard's bit phase is known as soon as the shards before it have been modelled: pack
        // and download it right away, under the sort of the devices that are still busy
        {
            std::unique_lock<std::mutex> lk(bits_mu);
            bits_known[g] = 1;
            bits_cv.notify_all();
            bits_cv.wait(lk, [&] {
                if (abort_all.load()) return true;
                for (size_t q = 0; q < g; q++)
                    if (!bits_known[q]) return false;
                return true;
            });
        }
        step([&]() -> int {
            if (!ctx->early_out.o || sh.blocks.empty()) return BNZ_OK;
            uint64_t bits = bit_base;
            bool first = true;                                   // no shard before this one holds blocks
            for (size_t q = 0; q < g; q++) {
                bits += shards[q].block_bits;
                if (!shards[q].blocks.empty()) first = false;
            }
            sh.bit_base = bits;
            size_t bytes = 0;
            int prc = shard_pack(ctx, sh, &bytes);
            if (prc != BNZ_OK) return prc;
            const size_t w0 = (size_t)(sh.bit_base >> 5) * 4;
            if (w0 + bytes + 64 > ctx->early_out.cap) return BNZ_OK;     // does not fit: packed after the join
            uint8_t *o = ctx->early_out.o;
            if (first && bit_base == 32) {
                CK(ctx, cudaMemcpyAsync(o + w0, d.out.p, bytes, cudaMemcpyDeviceToHost, d.stream));
            } else {
                CK(ctx, cudaMemcpyAsync(&sh.first_word, d.out.p, 4, cudaMemcpyDeviceToHost, d.stream));
                if (bytes > 4)
                    CK(ctx, cudaMemcpyAsync(o + w0 + 4, d.out.as<uint8_t>() + 4, bytes - 4, cudaMemcpyDeviceToHost, d.stream));
            }
            CK(ctx, cudaEventRecord(d.ev[7], d.stream));
            CK(ctx, cudaStreamSynchronize(d.stream));
            sh.d2h_bytes = bytes;
            sh.packed = true;
            return BNZ_OK;
        });
        if (!resident) uploaded += std::min<uint64_t>(N, cb * RLE_CHUNK + slop) - a;
        sh.h2d_bytes = uploaded;
        if (ctx->reuse_input && rc == BNZ_OK) d.tag(h_in, N, a, b);
        else d.tag(nullptr, 0, 0, 0);
        return rc;
    };

    std::vector<std::thread> th;
    for (size_t g = 0; g < G; g++)
        th.emplace_back([&, g]() {
            t_err_sink = &shards[g].err;
            shards[g].rc = work(g);
            if (shards[g].rc != BNZ_OK) {
                std::lock_guard<std::mutex> lk(bits_mu);
                abort_all.store(true);
                bits_cv.notify_all();
            }
            t_err_sink = nullptr;
        });
    for (std::thread &t : th) t.join();
    if (walk_rc != 0) return fail(ctx, BNZ_EINTERNAL, "RLE1 cut walk failed");
    for (Shard &sh : shards)
        if (sh.rc != BNZ_OK) {
            ctx->err = sh.err;
            return sh.rc;
        }
    if (consumed) *consumed = used;
    for (Shard &sh : shards) st.h2d_bytes += sh.h2d_bytes;
    if (all_blocks.empty()) {           // (non-final batch shorter than one block)
        shards.clear();
        *total_bits = bit_base;
        return BNZ_OK;
    }

    // bit offsets of the shards (blocks are concatenated at bit granularity, lib.rs:101-126 + out.rs)
    uint64_t bits = bit_base;
    for (Shard &sh : shards) {
        sh.bit_base = bits;
        bits += sh.block_bits;
        crcs.insert(crcs.end(), sh.crcs.begin(), sh.crcs.end());
    }
    *total_bits = bits;
    for (Shard &sh : shards)
        if (!sh.blocks.empty()) add_stats(st, sh);
    return BNZ_OK;
}

// One GPU, host input, many blocks: the input is uploaded and cut in pieces.  The first piece is small
// (7/16 of a block per SM, ~60 MB at level 9: on the device after ~1 ms); its blocks are sorted
// (lane 0, cluster kernel) while the rest is still on the PCIe bus.
// The chunk tables of a later piece continue the earlier ones (the scan carries are kept per tile),
// the host walk is simply repeated over the input so far (1 us per block) and must reproduce the
// earlier pieces' blocks; the new blocks run on a further lane of the same device, whose sort CTAs
// take the SM slots the earlier launches leave and free.  A lane that finishes early runs its MTF,
// Huffman and packing under the later lanes' sort.  *handled = false: the input is too small for
// this, nothing was done.
static int encode_pieces(bnz_ctx *ctx, const uint8_t *h_in, size_t N, int level, std::vector<Shard> &shards,
                         std::vector<uint32_t> &crcs, uint64_t *total_bits, uint64_t bit_base, bool *handled)
{
    *handled = false;
    Device &d0 = ctx->devs[0];
    const uint64_t n_chunks = (N + RLE_CHUNK - 1) / RLE_CHUNK;
    const uint64_t tile = rle_scan_tile_chunks();
    const uint64_t blk = (uint64_t)100000 * level;
    // piece 0: piece_blocks_per_sm_x16 / 16 blocks per SM (rounded up to whole scan tiles): large enough to
    // keep every SM busy until the second piece has arrived, small enough that its sort is over by
    // then, so that the RLE kernels of the next piece find free SM slots (measured, 1 GiB: 6..9
    // sixteenths give 150.4-151.2 ms end to end, 17 gives 153.0, 28 gives 167.0).
    const uint64_t slots = (uint64_t)d0.sm_count * (uint64_t)ctx->piece_blocks_per_sm_x16 / 16;
    uint64_t c0 = ((slots * blk / RLE_CHUNK + tile - 1) / tile) * tile;
    int K = ctx->h2d_pieces;
    if (ctx->h2d_overlap == 2) {                              // forced (tests): one scan tile, whatever the size
        c0 = tile;
        if (c0 * RLE_CHUNK * 2 > N) return BNZ_OK;
    } else if (c0 * RLE_CHUNK * 4 > N || N < ((size_t)256 << 20)) {
        return BNZ_OK;                                        // too little to hide: the copy is short, extra lanes cost latency
    }
    // piece ends in chunks (multiples of the scan tile, the last one = n_chunks): the rest is split evenly
    std::vector<uint64_t> cend;
    cend.push_back(c0);
    for (int k = 1; k < K; k++) {
        uint64_t e = c0 + (n_chunks - c0) * (uint64_t)k / (uint64_t)(K - 1);
        e = (k == K - 1) ? n_chunks : (e / tile) * tile;
        if (e > cend.back()) cend.push_back(e);
    }
    if (cend.back() != n_chunks) cend.push_back(n_chunks);
    K = (int)cend.size();
    while ((int)ctx->aux.size() < K - 1) {
        Device *a = new Device();
        if (!device_init(*a, d0.id)) {
            device_release(*a);
            delete a;
            return fail(ctx, BNZ_ECUDA, "extra lane");
        }
        ctx->aux.push_back(a);
    }
    auto lane = [&](int k) -> Device & { return k == 0 ? d0 : *ctx->aux[k - 1]; };
    *handled = true;
    bnz_stats &st = ctx->stats;
    for (int k = 0; k < K; k++) lane(k).launches = 0;
    CK(ctx, cudaSetDevice(d0.id));
    CK(ctx, d0.in.ensure(N + 64));
    CK(ctx, d0.ch_lasthead.ensure(n_chunks * 8));
    CK(ctx, d0.ch_meta.ensure(n_chunks * 4));
    CK(ctx, d0.ch_restsum.ensure(n_chunks * 4));
    CK(ctx, d0.ch_oin.ensure(n_chunks * 8));
    CK(ctx, d0.ch_P.ensure((n_chunks + 1) * 8));
    CK(ctx, d0.ch_tiles.ensure(rle_scan_tiles(n_chunks) * 16 + 64));
    CK(ctx, d0.h_P.ensure((n_chunks + 1) * 8));
    CK(ctx, d0.h_oin.ensure(n_chunks * 8));
    uint8_t *d_in = d0.in.as<uint8_t>();
    uint64_t *h_P = d0.h_P.as<uint64_t>(), *h_oin = d0.h_oin.as<uint64_t>();

    // all copies are queued at once, each on its lane's stream (the copy engine takes them in order);
    // a piece's bytes end one look-ahead page behind its last chunk
    auto bytes_end = [&](int k) -> uint64_t { return k == K - 1 ? N : std::min<uint64_t>(N, cend[k] * RLE_CHUNK + 4096); };
    for (int k = 0; k < K; k++) {
        Device &d = lane(k);
        const uint64_t a = k ? bytes_end(k - 1) : 0, b = bytes_end(k);
        CK(ctx, cudaEventRecord(d.ev[0], d.stream));
        CK(ctx, cudaMemcpyAsync(d_in + a, h_in + a, b - a, cudaMemcpyHostToDevice, d.stream));
        CK(ctx, cudaEventRecord(d.ev[1], d.stream));
    }
    st.h2d_bytes += N;
    if (ctx->reuse_input) d0.tag(h_in, N, 0, N);          // (the whole input ends up resident in d0.in)

    shards = std::vector<Shard>(K);
    std::vector<std::thread> th;
    std::vector<RleBlock> prev;                              // blocks cut so far
    int rc = BNZ_OK;
    for (int k = 0; k < K && rc == BNZ_OK; k++) {
        Device &d = lane(k);
        shards[k].d = &d;
        const uint64_t ca = k ? cend[k - 1] : 0, cb = cend[k];
        const bool last = k == K - 1;
        auto plan = [&]() -> int {
            CK(ctx, rle_summary_range_launch(d_in, N, n_chunks, ca, cb, d0.ch_lasthead.as<uint64_t>(),
                                             d0.ch_meta.as<uint32_t>(), d0.ch_restsum.as<uint32_t>(), d0.ch_oin.as<uint64_t>(),
                                             d0.ch_P.as<uint64_t>(), d0.ch_tiles.as<uint64_t>(), d.stream));
            d.launches += 4;
            CK(ctx, cudaMemcpyAsync(h_P + ca, d0.ch_P.as<uint64_t>() + ca, (cb + 1 - ca) * 8, cudaMemcpyDeviceToHost, d.stream));
            CK(ctx, cudaMemcpyAsync(h_oin + ca, d0.ch_oin.as<uint64_t>() + ca, (cb - ca) * 8, cudaMemcpyDeviceToHost, d.stream));
            CK(ctx, cudaStreamSynchronize(d.stream));
            std::vector<RleBlock> all;
            uint64_t used = 0;
            if (rle_walk_cuts(h_in, last ? N : cb * RLE_CHUNK, level, h_P, h_oin, cb, all, last, &used) != 0 ||
                all.size() < prev.size())
                return fail(ctx, BNZ_EINTERNAL, "RLE1 cut walk failed");
            for (size_t b = 0; b < prev.size(); b++)
                if (all[b].s != prev[b].s || all[b].c != prev[b].c || all[b].n != prev[b].n || all[b].rle_off != prev[b].rle_off)
                    return fail(ctx, BNZ_EINTERNAL, "cut chain of the earlier pieces is not a prefix of the longer one");
            shards[k].blocks.assign(all.begin() + prev.size(), all.end());
            prev.swap(all);
            if (!shards[k].blocks.empty()) {
                const uint64_t off0 = shards[k].blocks.front().rle_off;
                for (RleBlock &b : shards[k].blocks) b.rle_off -= off0;
            }
            return BNZ_OK;
        };
        rc = plan();
        if (rc != BNZ_OK) break;
        th.emplace_back([&, k]() {
            Shard &sh = shards[k];
            t_err_sink = &sh.err;
            if (cudaSetDevice(sh.d->id) != cudaSuccess) {            // a new thread starts on device 0
                sh.err = "cudaSetDevice";
                sh.rc = BNZ_ECUDA;
            } else {
                sh.rc = sh.blocks.empty() ? BNZ_OK
                                          : shard_model(ctx, sh, d_in, N, d0.ch_oin.as<uint64_t>(), d0.ch_P.as<uint64_t>(), level);
            }
            t_err_sink = nullptr;
        });
    }
    for (std::thread &t : th) t.join();
    if (rc != BNZ_OK) return rc;
    for (Shard &sh : shards)
        if (sh.rc != BNZ_OK) {
            ctx->err = sh.err;
            return sh.rc;
        }

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

// One batch of the path: the blocks that can be cut from h_in[0, N).  d_in0: optional device copy
// already resident on device 0.  `final`: no input follows (otherwise the trailing incomplete
// block is left for the next batch; *consumed tells where it starts).  `bit_base`: bit offset of
// the batch's first block in the stream.  Leaves every shard's bits in its device's d.out and
// returns the layout; the callers move the bytes.
static int encode_all_impl(bnz_ctx *ctx, const uint8_t *h_in, const uint8_t *d_in0, size_t N, int level,
                           std::vector<Shard> &shards, std::vector<uint32_t> &crcs, uint64_t *total_bits, bool final,
                           uint64_t bit_base, uint64_t *consumed);

int encode_all(bnz_ctx *ctx, const uint8_t *h_in, const uint8_t *d_in0, size_t N, int level,
               std::vector<Shard> &shards, std::vector<uint32_t> &crcs, uint64_t *total_bits, bool final,
               uint64_t bit_base, uint64_t *consumed)
{
    uint64_t used = 0;
    int rc = encode_all_impl(ctx, h_in, d_in0, N, level, shards, crcs, total_bits, final, bit_base, &used);
    if (consumed) *consumed = used;
    if (rc == BNZ_OK && ctx->verify && !shards.empty()) rc = verify_host(ctx, h_in, final ? (uint64_t)N : used, shards);
    return rc;
}

static int encode_all_impl(bnz_ctx *ctx, const uint8_t *h_in, const uint8_t *d_in0, size_t N, int level,
                           std::vector<Shard> &shards, std::vector<uint32_t> &crcs, uint64_t *total_bits, bool final,
                           uint64_t bit_base, uint64_t *consumed)
{
    Device &d0 = ctx->devs[0];
    bnz_stats &st = ctx->stats;
    if (ctx->h2d_overlap && ctx->devs.size() == 1 && !d_in0 && final && N >= ((size_t)16 << 20) &&
        !(ctx->reuse_input && d0.holds(h_in, N, 0) && d0.tag_b == N)) {
        bool handled = false;
        int rc = encode_pieces(ctx, h_in, N, level, shards, crcs, total_bits, bit_base, &handled);
        if (rc != BNZ_OK || handled) {
            if (handled && consumed) *consumed = N;
            return rc;
        }
    }
    if (ctx->devs.size() > 1) {
        if (d_in0) return fail(ctx, BNZ_EINVAL, "device-resident input needs a single-GPU context");
        int rc = encode_sharded(ctx, h_in, N, level, shards, crcs, total_bits, final, bit_base, consumed);
        return rc;
    }
    d0.launches = 0;
    CK(ctx, cudaSetDevice(d0.id));
    CK(ctx, cudaEventRecord(d0.ev[0], d0.stream));
    const uint8_t *d_in = d_in0;
    if (!d_in) {
        if (!(ctx->reuse_input && d0.holds(h_in, N, 0) && d0.tag_b == N)) {
            CK(ctx, d0.in.ensure(N + 64));
            CK(ctx, cudaMemcpyAsync(d0.in.p, h_in, N, cudaMemcpyHostToDevice, d0.stream));
            st.h2d_bytes += N;
            if (ctx->reuse_input) d0.tag(h_in, N, 0, N);
        }
        d_in = d0.in.as<uint8_t>();
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
    shards = std::vector<Shard>(1);
    Shard &sh = shards[0];
    sh.d = &d0;
    sh.blocks.swap(blocks);
    rc = shard_model(ctx, sh, d_in, N, d0.ch_oin.as<uint64_t>(), d0.ch_P.as<uint64_t>(), level);
    if (rc != BNZ_OK) return rc;
    sh.bit_base = bit_base;
    *total_bits = bit_base + sh.block_bits;
    crcs.insert(crcs.end(), sh.crcs.begin(), sh.crcs.end());
    add_stats(st, sh);
    return BNZ_OK;
}

// pack every shard of a batch at its bit phase and copy it to host memory `o` (the stream buffer,
// byte 0 = stream byte 0).  `stream_start`: the batch begins right after the 32-bit stream header,
// so its first word is not shared with earlier data.  (The streaming front end passes a buffer
// that starts at stream byte `o_first_byte`, a multiple of 4.)
int pack_and_download(bnz_ctx *ctx, std::vector<Shard> &shards, uint8_t *o, bool stream_start,
                      size_t o_first_byte)
{
    size_t first = 0;                                        // the first shard that holds blocks
    while (first < shards.size() && shards[first].blocks.empty()) first++;
    for (size_t g = 0; g < shards.size(); g++) {
        Shard &sh = shards[g];
        if (sh.blocks.empty()) continue;
        if (sh.packed) {                                     // (done by the shard's own thread, encode_sharded)
            ctx->stats.d2h_bytes += sh.d2h_bytes;
            continue;
        }
        Device &d = *sh.d;
        CK(ctx, cudaSetDevice(d.id));
        size_t bytes = 0;
        int rc = shard_pack(ctx, sh, &bytes);
        if (rc != BNZ_OK) return rc;
        const size_t w0 = (size_t)(sh.bit_base >> 5) * 4 - o_first_byte;
        if (g == first && stream_start) {
            CK(ctx, cudaMemcpyAsync(o + w0, d.out.p, bytes, cudaMemcpyDeviceToHost, d.stream));
        } else {
            CK(ctx, cudaMemcpyAsync(&sh.first_word, d.out.p, 4, cudaMemcpyDeviceToHost, d.stream));
            if (bytes > 4)
                CK(ctx, cudaMemcpyAsync(o + w0 + 4, d.out.as<uint8_t>() + 4, bytes - 4, cudaMemcpyDeviceToHost, d.stream));
        }
        CK(ctx, cudaEventRecord(d.ev[7], d.stream));
        ctx->stats.d2h_bytes += bytes;
    }
    for (Shard &sh : shards) {
        if (sh.blocks.empty() || sh.packed) continue;
        CK(ctx, cudaSetDevice(sh.d->id));
        CK(ctx, cudaStreamSynchronize(sh.d->stream));
    }
    // merge the words shared with the previous shard / batch
    for (size_t g = 0; g < shards.size(); g++) {
        if (shards[g].blocks.empty() || (g == first && stream_start)) continue;
        uint8_t *w = o + ((size_t)(shards[g].bit_base >> 5) * 4 - o_first_byte);
        const uint8_t *f = reinterpret_cast<const uint8_t *>(&shards[g].first_word);
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

    if (in_len <= ctx->max_batch_bytes * ctx->devs.size()) {      // (the limit is per device)
        // ---- one batch: the stream is assembled in the context's pinned buffer
        std::vector<Shard> shards;
        if (in_len > 0) {
            if (ctx->devs.size() > 1) {
                // several devices: a shard is downloaded by its own thread as soon as its bit phase is known,
                // so the stream buffer must exist up front (sized for incompressible input; a shard that
                // does not fit is packed after the exact size is known; the buffer is kept across calls)
                int rc0 = ensure_out_cache(ctx, in_len + in_len / 8 + ((size_t)64 << 20), true);
                if (rc0 != BNZ_OK) return rc0;
                ctx->early_out.o = ctx->out_cache;
                ctx->early_out.cap = ctx->out_cache_cap;
            }
            int rc = encode_all(ctx, in, nullptr, in_len, level, shards, crcs, &total_bits);
            ctx->early_out.o = nullptr;
            if (rc != BNZ_OK) return rc;
        }
        nbytes = (size_t)((total_bits + 80 + 7) / 8);
        if (ctx->out_cache_cap < nbytes + 8 + 16)               // the buffer is about to be replaced: nothing in it counts
            for (Shard &sh : shards) sh.packed = false;
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
        size_t pos = 0, win = ctx->max_batch_bytes * ctx->devs.size();
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
    if (ctx->devs.size() != 1) return fail(ctx, BNZ_EINVAL, "bnz_encode_device needs a single-GPU context");
    // the RLE1 and CRC kernels read the input with 16-byte vector loads
    if (((uintptr_t)d_in & 15) != 0) return fail(ctx, BNZ_EINVAL, "d_in must be 16-byte aligned");
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
