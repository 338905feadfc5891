// stream.cu — streaming front end (bnz_stream_*) and bnz_encode_file.
#include "host.h"

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
