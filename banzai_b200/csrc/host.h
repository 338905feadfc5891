// host.h — internal declarations shared by the host-side translation units of libbanzai_b200.so
// (context.cu, stages.cu, encode.cu, stream.cu).  Nothing here is part of the C ABI.
#pragma once
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
    DevBuf in, rle, bwt, blk_off, blk_len, ptr, has_byte, bwt_stats, counters, ws_rec, ws_rank, ws_ctl, ws_hist, ws_rec2, ws_rank2, ws_defer;
    DevBuf v_marks, v_lfl, v_flags;     // self-verification (verify.cu)
    DevBuf sel;                         // huff_literal: selectors [nb][sel_stride]
    DevBuf ch_lasthead, ch_meta, ch_restsum, ch_oin, ch_P, ch_tiles, rle_blocks, crc_acc;
    DevBuf seg_base, seg_list, seg_cnt, seg_state, num_names, syms, sym_off, sym_len, freqs, mtf_ids, mtf_cseg;
    DevBuf lens, codes, tf, num_tables, num_sel, span_base, hdr, hdr_bits, crc, blk_bits, blk_bitoff,
        total_bits, out;
    cudaEvent_t ev[16] = {};
    PinBuf h_P, h_oin, h_acc, h_mtf, h_done, h_tiles;   // h_done: per-block completion flags the sort writes (mapped)
    bool crc_tables = false;
    uint32_t launches = 0;
    // what d.in holds (ctx->reuse_input): bytes [tag_a, tag_b) of the host buffer tag_ptr[0, tag_len)
    const void *tag_ptr = nullptr;
    size_t tag_len = 0;
    uint64_t tag_a = 0, tag_b = 0;
    bool holds(const void *p, size_t n, uint64_t a) const { return tag_ptr && tag_ptr == p && tag_len == n && tag_a == a; }
    void tag(const void *p, size_t n, uint64_t a, uint64_t b) { tag_ptr = p; tag_len = n; tag_a = a; tag_b = b; }
};

struct bnz_ctx {
    std::vector<Device> devs;
    std::string err;
    bnz_stats stats;
    int ctas_per_sm = 0;
    int bwt_cluster = -1;         // CTAs per bzip2 block (-1: auto, 0/1: single-CTA kernel)
    int bwt_threads = 512;
    int bwt_cluster_below = 128;       // auto mode: cluster kernel when a device gets fewer blocks than this (measured crossover ~115
                                       // level-9 blocks with the packed-key kernels: 137 blocks 22.1 vs 25.6 ms, profiles/r2_experiments)
    int bwt_periodic = 1;              // closed-form order of periodic runs (bwt_common.cuh: Period); 0 = plain doubling
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
    std::vector<Device *> aux;         // further lanes on the first device (created on demand): the blocks of later input pieces
    int h2d_pieces = 3;                // pieces the input of one call is uploaded in (first piece: see piece_blocks_per_sm_x16)
    int piece_blocks_per_sm_x16 = 7;   // size of the first piece in blocks per SM (x 1/16): measured best 6..9 (profiles/r2_experiments)
    int h2d_overlap = 1;               // one device, host input: upload in two pieces, sort the first while the second arrives
    int crc_low_prio = 1;              // block CRCs on the low-priority stream (they would delay the start of the sort)
    // where the shards of the current bnz_encode call may be downloaded as soon as their bit phase is
    // known (several devices, one batch); o == nullptr: the caller packs and downloads afterwards
    struct { uint8_t *o = nullptr; size_t cap = 0; } early_out;
    int huff_literal = 0;              // run huffman::encode's 4-round table refinement literally (default: its closed form)
    int verify = 0;                    // check every block on the device (and the CRCs on the host) before emitting
    int verify_corrupt = 0;            // test hook: damage an intermediate of block 0 behind the sort's back
    int reuse_input = 0;               // the caller re-encodes the SAME host buffer: keep its device copy resident (benchmarks)
    int mtf_groups = 2;
    int mtf_overlap = 70;              // percent of a device's blocks whose MTF may run beside the sort (0: off)
};


// worker threads of a multi-device encode record their error text in their own string
extern thread_local std::string *t_err_sink;
bool device_init(Device &d, int id);
void device_release(Device &d);
void set_err(bnz_ctx *ctx, const std::string &msg);
int fail(bnz_ctx *ctx, int code, const std::string &msg);

#define CK(ctx, call)                                                                          \
    do {                                                                                       \
        cudaError_t e__ = (call);                                                              \
        if (e__ != cudaSuccess) {                                                              \
            set_err((ctx), std::string(#call) + ": " + cudaGetErrorString(e__));               \
            return (e__ == cudaErrorMemoryAllocation) ? BNZ_ENOMEM : BNZ_ECUDA;                \
        }                                                                                      \
    } while (0)

// ---------------------------------------------------------------------------------------
// stages on one device (stages.cu)
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

int run_bwt_device(bnz_ctx *ctx, Device &d, const uint8_t *d_rle, uint8_t *d_bwt, const uint64_t *d_blk_off,
                   const uint32_t *d_blk_len, uint32_t n_blocks, uint32_t max_len, uint32_t *d_ptr,
                   uint8_t *d_has_byte, BwtStats *d_stats, uint32_t *d_done = nullptr, bool *done_armed = nullptr,
                   uint32_t *d_marks = nullptr);
int rle_plan(bnz_ctx *ctx, Device &d, const uint8_t *d_in, const uint8_t *h_in, uint64_t N, int level,
             std::vector<RleBlock> &blocks, bool final = true, uint64_t *consumed = nullptr);
int rle_emit_shard(bnz_ctx *ctx, Device &d, const uint8_t *in_base, uint64_t N, const uint64_t *oin_base,
                   const uint64_t *P_base, const std::vector<RleBlock> &blocks, std::vector<uint32_t> *crcs,
                   uint64_t *rle_total);
int upload_batch(bnz_ctx *ctx, Device &d, const Batch &bt);
int mtf_ensure(bnz_ctx *ctx, Device &d, const Batch &bt);
int run_mtf_list(bnz_ctx *ctx, Device &d, const Batch &bt, const uint8_t *d_bwt, uint8_t *d_idx,
                 const uint8_t *d_has_byte, const uint32_t *ids, uint32_t n_list, uint32_t &ids_used,
                 uint32_t &lists_used, cudaStream_t st);
int run_huff_model_device(bnz_ctx *ctx, Device &d, const Batch &bt, int level, int with_block_header,
                          uint64_t bit_base, uint64_t fixed_stride_bits, HuffArgs &a);

// ---------------------------------------------------------------------------------------
// the whole path (encode.cu)
// ---------------------------------------------------------------------------------------

struct Shard {
    Device *d = nullptr;
    std::vector<RleBlock> blocks;      // rle_off rebased to this device's rle buffer
    std::vector<uint32_t> crcs;
    std::vector<uint32_t> vflags;      // verify: per block VERIFY_BAD_* bits
    std::vector<BwtStats> bst;
    Batch bt;
    HuffArgs ha;
    uint64_t block_bits = 0;           // sum of the shard's block bit lengths
    uint64_t bit_base = 0;             // global bit offset of the shard's first block
    int rc = BNZ_OK;
    std::string err;
    uint64_t h2d_bytes = 0;            // input bytes this shard's device received
    bool packed = false;               // already packed at its bit phase and copied to the host (sharded path)
    uint32_t first_word = 0;           // its first 32-bit word, when that word is shared with the previous shard
    size_t d2h_bytes = 0;
    Shard() = default;
    Shard(const Shard &o) { d = o.d; }
};

// K1 emit .. K7 + headers for one shard; ends with a host sync that yields block_bits.

// SURVEY §8(d) algorithmic bytes of the BWT sort: 9 per byte of the block (text read, initial 8-byte
// key) + per record and round 16 per radix pass through HBM (read + write of the 8-byte record;
// P_r = 0 for a record that is sorted inside shared memory) + 36 (key/list record write and read,
// the rank gathers, the rank write).
inline uint64_t bwt_algorithmic_bytes(const bnz_stats &s)
{
    return 9 * s.bwt_n + 16 * s.bwt_sum_active_passes + 36 * s.bwt_sum_active;
}

void put_bits_host(uint8_t *buf, uint64_t bitpos, uint64_t value, int nbits);      // MSB first
uint32_t fold_stream_crc(const std::vector<uint32_t> &crcs);                       // lib.rs:108
void finish_stats(bnz_ctx *ctx, std::vector<Shard> &shards, bool have_d2h);
int encode_all(bnz_ctx *ctx, const uint8_t *h_in, const uint8_t *d_in0, size_t N, int level,
               std::vector<Shard> &shards, std::vector<uint32_t> &crcs, uint64_t *total_bits,
               bool final = true, uint64_t bit_base = 32, uint64_t *consumed = nullptr);
int pack_and_download(bnz_ctx *ctx, std::vector<Shard> &shards, uint8_t *o, bool stream_start,
                      size_t o_first_byte = 0);
