/* banzai_b200.h — C ABI of the B200-native bzip2 encoder core.
 *
 * Drop-in boundary for the block-compression path of jgbyrne/banzai v0.3.1.  Every entry
 * point cites the reference interface it replaces (paths relative to the reference repo).
 * Plain pointers and sizes only; no C++/torch types.  The library fails loudly
 * (BNZ_ECUDA) when no CUDA device / sm_100a kernel image is usable: there is no CPU
 * fallback.
 *
 * Threading: a bnz_ctx is not thread-safe; use one per host thread.
 */
#ifndef BANZAI_B200_H
#define BANZAI_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define BNZ_API __attribute__((visibility("default")))
#else
#define BNZ_API
#endif

typedef struct bnz_ctx bnz_ctx;

enum {
    BNZ_OK = 0,
    BNZ_EINVAL = 1,     /* bad argument; level outside 1..=9 (the reference panics: lib/lib.rs:19,89) */
    BNZ_ECUDA = 2,      /* CUDA runtime / driver error, or no usable device */
    BNZ_ENOMEM = 3,     /* host or device allocation failed */
    BNZ_EINTERNAL = 4,  /* internal invariant violated (reference: assert!/panic! sites) */
    BNZ_EIO = 5,        /* a sink callback or a file operation failed (reference: io::Error via `?`) */
    BNZ_EVERIFY = 6     /* "verify": a block failed the self-check; nothing was emitted */
};

/* ---- context -------------------------------------------------------------------------
 * The reference has no context (single-threaded, lib/lib.rs:84); the context owns what a
 * GPU path needs across calls: per-device streams, device arenas, pinned staging. */

/* n_gpus: number of visible CUDA devices to shard blocks over (0 = all visible). */
BNZ_API int bnz_ctx_create(bnz_ctx **out, int n_gpus);
/* explicit device ordinals (e.g. {LOCAL_RANK} for one-process-per-GPU launches); an ordinal may
 * repeat, each entry is then an independent lane on that GPU */
BNZ_API int bnz_ctx_create_on(bnz_ctx **out, const int *device_ids, int n_devices);
BNZ_API void bnz_ctx_destroy(bnz_ctx *ctx);
BNZ_API const char *bnz_strerror(int code);
/* human-readable detail of the last failing call on this context ("" if none) */
BNZ_API const char *bnz_last_error(const bnz_ctx *ctx);

/* tunables (call before encoding). keys:
 *   "bwt_cluster"        -1 auto | 0 one CTA per block | 2..16 CTAs (one cluster) per block
 *   "bwt_cluster_below"  auto mode: cluster kernel when a device gets fewer blocks than this (128)
 *   "bwt_threads"        512 | 1024, cluster kernel
 *   "bwt_periodic"       1 (default): blocks with a long periodic run (zero pages, "abab...", repeated
 *                        records) are ordered in closed form after the first rounds instead of by ~log2(n)
 *                        doubling rounds — the reference's SA-IS has no bad case there either (README.md:7);
 *                        the cluster kernel leaves such blocks to the one-CTA kernel.  0: plain doubling
 *   "bwt_ctas_per_sm"    0 = auto
 *   "huff_literal"       1: run the 4-round table refinement of huffman::encode literally on the
 *                        device (per-group cost and argmin over all tables, rebuild, selectors);
 *                        0 (default): its closed form for this reference (lib/huffman.rs:399-460
 *                        zeroes the tables before rounds 1..3, SURVEY A-Q10) — same bits, ~10 ms
 *                        less per GiB
 *   "verify"             1: self-verification (the reference has no decoder, README.md:9; its safety
 *                        net is the libbz2 round trip of fuzz/fuzz_targets/round_trip.rs): before a
 *                        stream is returned, every block's RLE1 image is decoded back to its input
 *                        bytes and its (BWT, origPtr) is inverted back to the RLE1 image on the
 *                        device, and the cut chain and the block CRCs are re-derived on the host;
 *                        BNZ_EVERIFY on any mismatch.  Covers RLE1, BWT and CRC, not the entropy stage.
 *   "reuse_input"        1: the caller encodes the SAME host buffer again and again (benchmarks);
 *                        every device keeps its copy of the input resident, the next call of the
 *                        identical (pointer, length) skips the upload
 *   "mtf_overlap"        percent of a device's blocks whose MTF may run beside the sort, in the SM
 *                        slots its tail leaves empty (70; 0 = strictly after the sort)
 *   "mtf_groups"         number of block lists the overlapped MTF is issued in (2)
 *   "h2d_overlap"        1 (default): one GPU, host input of >= 256 MiB: upload in "h2d_pieces" (3)
 *                        pieces and sort each while the next arrives | 0 single upload | 2 forced (tests)
 *   "h2d_pieces"         pieces of that upload (3; one lane of the GPU per piece)
 *   "piece_blocks_per_sm_x16"  size of the first piece in blocks per SM, in sixteenths (7)
 *   "crc_low_prio"       1: block CRCs on the low-priority stream (default) | 0: normal side stream
 *   "max_batch_bytes"    inputs above this (3 GiB) are encoded in batches so that device memory
 *                        stays bounded
 *   "stream_window_bytes" bnz_stream_*: input bytes per pinned window (512 MiB)
 * None of them changes the bytes of the stream. */
BNZ_API int bnz_ctx_set(bnz_ctx *ctx, const char *key, long value);

/* ---- the hot path --------------------------------------------------------------------
 * Replaces `banzai::encode(reader, BufWriter, level) -> io::Result<usize>`
 * (lib/lib.rs:84-132) on a whole in-memory input: the caller's shim does
 * read_to_end -> bnz_encode -> write_all -> flush.
 *   in/in_len : input bytes in HOST memory (pinned memory from bnz_host_alloc is fastest)
 *   level     : 1..=9, block size = level * 100 000 (lib/rle.rs:121)
 *   *out      : complete .bz2 stream, byte-identical to the reference's; allocated by the
 *               library, release with bnz_free
 *   *consumed : input bytes encoded (== in_len on success; lib/lib.rs:119,131)
 */
BNZ_API int bnz_encode(bnz_ctx *ctx, const uint8_t *in, size_t in_len, int level,
                       uint8_t **out, size_t *out_len, size_t *consumed);
BNZ_API void bnz_free(bnz_ctx *ctx, uint8_t *p);

/* Same computation with the input already resident in device memory of the context's
 * first device and the stream left on the device (used to time the kernels without
 * PCIe).  h_in is the host mirror of the same bytes: the sequential block-cut walk
 * (lib/rle.rs capacity rule) reads <= 2 KiB of it per block; nothing is uploaded.
 * Needs a context with exactly one device.  d_in must be 16-byte aligned (the kernels read it
 * with vector loads; memory from bnz_device_alloc is) — BNZ_EINVAL otherwise.  d_out_cap must be
 * at least the stream size + 8 bytes (the stream is written in whole words):
 * bnz_max_compressed_size(in_len) always suffices. */
BNZ_API int bnz_encode_device(bnz_ctx *ctx, const void *d_in, const uint8_t *h_in, size_t in_len,
                              int level, void *d_out, size_t d_out_cap, size_t *out_len);
BNZ_API size_t bnz_max_compressed_size(size_t in_len);

/* `banzai::encode_file(in_path, out_path)` (lib/lib.rs:141-153): level 9, returns bytes
 * encoded through *consumed.  Streams the file through bnz_stream_* (bounded memory). */
BNZ_API int bnz_encode_file(bnz_ctx *ctx, const char *in_path, const char *out_path,
                            size_t *consumed);

/* ---- streaming front end ---------------------------------------------------------------
 * The reader/writer shape of `banzai::encode` (lib/lib.rs:84-132; refill loop
 * lib/rle.rs:43-91; OutputStream lib/out.rs:7-104) for pipes and inputs larger than host
 * or device memory.  The caller fills pinned windows in place (reserve -> read into *buf ->
 * commit); while a full window is on the GPU the caller keeps reading the next one.  Finished
 * stream bytes are handed to `sink` in order, always on the caller's thread and only inside
 * bnz_stream_commit / _write / _finish.  The bytes written are exactly those of one bnz_encode
 * over the concatenated input, whatever the commit sizes and the window size
 * (bnz_ctx_set "stream_window_bytes", default 512 MiB; at least one block's worth is enforced).
 * One open stream per context; the context must not be used otherwise until bnz_stream_close.
 *   sink      : returns 0 on success; anything else aborts the stream with BNZ_EIO
 *   reserve   : (*buf, *cap) = writable space for the next input bytes (cap >= 1)
 *   commit    : n <= cap bytes were placed at *buf
 *   write     : reserve + memcpy + commit for callers that already hold the bytes
 *   finish    : encodes what is left, writes the footer (lib/lib.rs:66-70) and the zero padding
 *               (lib/out.rs:22-28); *consumed = total input bytes (lib/lib.rs:131)
 *   close     : releases the stream (also valid without finish: abandons the output) */
typedef struct bnz_stream bnz_stream;
typedef int (*bnz_sink_fn)(void *user, const uint8_t *data, size_t len);
BNZ_API int bnz_stream_open(bnz_ctx *ctx, int level, bnz_sink_fn sink, void *user, bnz_stream **out);
BNZ_API int bnz_stream_reserve(bnz_stream *s, uint8_t **buf, size_t *cap);
BNZ_API int bnz_stream_commit(bnz_stream *s, size_t n);
BNZ_API int bnz_stream_write(bnz_stream *s, const uint8_t *data, size_t len);
BNZ_API int bnz_stream_finish(bnz_stream *s, size_t *consumed);
BNZ_API void bnz_stream_close(bnz_stream *s);

/* pinned host buffers for inputs (H2D at full PCIe rate) */
BNZ_API void *bnz_host_alloc(size_t bytes);
BNZ_API void bnz_host_free(void *p);
/* raw device buffers on the context's first device (for bnz_encode_device callers) */
BNZ_API void *bnz_device_alloc(bnz_ctx *ctx, size_t bytes);
BNZ_API void bnz_device_free(bnz_ctx *ctx, void *p);
BNZ_API int bnz_memcpy_h2d(bnz_ctx *ctx, void *d_dst, const void *h_src, size_t bytes);
BNZ_API int bnz_memcpy_d2h(bnz_ctx *ctx, void *h_dst, const void *d_src, size_t bytes);

/* ---- measurement ----------------------------------------------------------------------
 * Filled by the last bnz_encode / bnz_encode_device on this context.  Times are CUDA-event
 * milliseconds on the library's own stream of device 0 (max over devices for *_ms_max). */
typedef struct bnz_stats {
    uint64_t in_bytes, out_bytes;
    uint32_t n_blocks;
    uint32_t n_devices;
    uint32_t kernel_launches;        /* kernels launched by this library during the call */
    uint32_t bwt_radix_bits;
    float total_ms;                  /* first kernel/copy -> last, device 0 */
    float h2d_ms, d2h_ms;
    float rle_ms, crc_ms, bwt_ms, mtf_ms, huff_ms, pack_ms;
    /* BWT sort accounting: sum over blocks */
    uint64_t bwt_n;                  /* sum of n */
    uint64_t bwt_sum_active;         /* sum over blocks/rounds of records sorted */
    uint64_t bwt_sum_active_passes;  /* sum of records sorted x radix passes executed */
    uint32_t bwt_max_rounds;
    uint32_t bwt_tied_blocks;
    uint64_t bwt_rounds_total;
    uint64_t bwt_algorithmic_bytes;  /* 9n + sum over rounds a_r * (16 P_r + 36), P_r = 0 for records sorted
                                        inside shared memory   */
    uint64_t bwt_cyc_build, bwt_cyc_radix, bwt_cyc_rerank;   /* SM cycles per phase, summed over blocks */
    uint64_t h2d_bytes, d2h_bytes;
    uint64_t bwt_sum_tile;           /* of bwt_sum_active: records sorted inside shared memory (no HBM pass) */
    uint64_t bwt_cyc_tile;           /* SM cycles of that path */
    uint64_t bwt_cyc_final;          /* SM cycles of the sort's last pass (bwt[rank[i]] = S[i-1]) */
} bnz_stats;
BNZ_API int bnz_get_stats(const bnz_ctx *ctx, bnz_stats *out);

/* ---- stage-level exports (parity seams; mirror the reference's private stage fns) -----
 * All take HOST buffers describing a batch of independent blocks and run only the named
 * kernels on the context's first device.
 */

/* `rle::rle_one` applied repeatedly as `encode` does (lib/rle.rs:102, lib/lib.rs:101-126):
 * block cuts, RLE1 bytes and block CRCs for a whole input.
 *   blk_in_off[b], blk_in_len[b]  : consumed input range of block b
 *   blk_rle_off[b], blk_rle_len[b]: its RLE1 image inside rle_out (concatenated)
 *   blk_crc[b]                    : CRC-32/BZIP2 of the consumed input (lib/crc32.rs:31)
 * Capacity: max_blocks entries / rle_cap bytes; *n_blocks returns the count. */
BNZ_API int bnz_stage_rle1(bnz_ctx *ctx, const uint8_t *in, size_t in_len, int level,
                           uint64_t *blk_in_off, uint64_t *blk_in_len, uint64_t *blk_rle_off,
                           uint32_t *blk_rle_len, uint32_t *blk_crc, size_t max_blocks,
                           uint8_t *rle_out, size_t rle_cap, size_t *n_blocks);

/* The host half of that stage alone, no device involved: the sequential cut chain
 * (`encode` calling `rle_one` block after block, lib/lib.rs:101-126 with the capacity rule of
 * lib/rle.rs:121-240) over chunk tables supplied by the caller.  For every 1 KiB chunk c of the
 * input: o_in[c] = run offset of its first byte (number of equal bytes directly before it),
 * P[c] = RLE1 bytes emitted for in[0 .. 1024 c) when no block is ever cut (P has n_chunks + 1
 * entries).  final == 0: more input follows, a block that only ends with the data is not
 * reported.  *consumed = input bytes covered by the reported blocks. */
BNZ_API int bnz_host_cut_chain(const uint8_t *in, size_t in_len, int level, const uint64_t *P,
                               const uint64_t *o_in, size_t n_chunks, int final,
                               uint64_t *blk_in_off, uint64_t *blk_in_len, uint32_t *blk_rle_len,
                               size_t max_blocks, size_t *n_blocks, size_t *consumed);

/* `bwt::bwt` (lib/bwt.rs:526) on n_blocks independent blocks stored back to back:
 * block b = blocks[blk_off[b] .. blk_off[b] + blk_len[b]).  Outputs use the same layout.
 * has_byte is [n_blocks][256]. max block length = 100000*level. */
typedef struct bnz_bwt_block_stats {
    uint32_t n, rounds, tied, period;   /* period: of the periodic run the sort used, 0 = none */
    uint64_t sum_active, sum_active_passes;
    uint64_t cycles;              /* SM cycles the block occupied its CTA / cluster */
    uint64_t sum_tile;            /* of sum_active: records sorted inside shared memory */
} bnz_bwt_block_stats;
BNZ_API int bnz_stage_bwt(bnz_ctx *ctx, const uint8_t *blocks, const uint64_t *blk_off,
                          const uint32_t *blk_len, size_t n_blocks, int level, uint8_t *bwt_out,
                          uint32_t *ptr_out, uint8_t *has_byte_out,
                          bnz_bwt_block_stats *stats_out /* may be NULL */);

/* `mtf::mtf_and_rle` (lib/mtf.rs:14) on a batch of BWT blocks (same layout as above).
 * syms_out: u16 symbols of block b at syms_out + sym_off[b] where sym_off[b] = blk_off[b] + b
 * (each block may emit up to n+1 symbols); sym_len[b] = m; num_syms[b]; freqs [n_blocks][258]. */
BNZ_API int bnz_stage_mtf(bnz_ctx *ctx, const uint8_t *bwt, const uint64_t *blk_off,
                          const uint32_t *blk_len, const uint8_t *has_byte, size_t n_blocks,
                          uint16_t *syms_out, uint32_t *sym_len, uint32_t *num_syms,
                          uint32_t *freqs_out);

/* `huffman::encode` (lib/huffman.rs:313) on a batch of MTF blocks: block b's symbols at
 * syms + sym_off[b], length sym_len[b].  Emits, per block, the bits huffman::encode writes
 * to a fresh byte-aligned writer: bits_out + bit_byte_off[b] (zero padded), bit_len[b] bits.
 * tables_out [n_blocks][6][258] code lengths, num_tables[b]; selectors are all in the stream.
 * Per-block output capacity = out_stride bytes. */
BNZ_API int bnz_stage_huffman(bnz_ctx *ctx, const uint16_t *syms, const uint64_t *sym_off,
                              const uint32_t *sym_len, const uint32_t *num_syms,
                              const uint32_t *freqs, size_t n_blocks, uint8_t *bits_out,
                              size_t out_stride, uint64_t *bit_len, uint8_t *tables_out,
                              uint32_t *num_tables);

#ifdef __cplusplus
}
#endif
#endif /* BANZAI_B200_H */
