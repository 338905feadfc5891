// kernels.h — internal interface between the host pipeline (context.cu, stages.cu, encode.cu, stream.cu; shared declarations in host.h) and the kernel
// translation units.  Nothing here is part of the public C ABI (include/banzai_b200.h).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stddef.h>
#include <vector>

#define RLE_CHUNK 1024
#define MTF_SEG 1024
#define HUFF_MAX_SYMS 258
#define HUFF_MAX_TABLES 6
#define HUFF_REFINEMENTS 4        /* huffman.rs:307 */
#define BWT_CLUSTER_MAX 16
#define BWT_CTL_BYTES (3 * 16 * 256 * 4 + 2 * 16 * 4 + 16 * 2 * 4 + 2 * 4 * 4 + 4 + 7 * 4)

namespace bnz {

// ---------------------------------------------------------------- K1/K2 RLE1 + CRC (rle1.cu)

struct RleBlock {                 // one bzip2 block as cut by the host walk
    uint64_t s;                   // first input byte
    uint64_t c;                   // one past the last consumed input byte
    uint64_t e0;                  // end of the (possibly truncated) run the block starts in, <= c
    uint64_t P_e0;                // cost prefix P at e0
    uint64_t rle_off;             // offset of the block's RLE1 image in the rle buffer
    uint32_t u0;                  // output bytes of the first run: g(e0 - s)
    uint32_t n;                   // RLE1 length of the block
};

cudaError_t crc_upload_tables();
uint32_t crc_finalize(uint32_t acc, uint64_t len);
size_t rle_scan_tiles(uint64_t n_chunks);
uint64_t rle_scan_tile_chunks();
cudaError_t rle_summary_range_launch(const uint8_t *d_in, uint64_t N, uint64_t n_chunks_total, uint64_t c_first,
                                     uint64_t c_last, uint64_t *d_lasthead, uint32_t *d_meta, uint32_t *d_restsum,
                                     uint64_t *d_oin, uint64_t *d_P, uint64_t *d_tiles, cudaStream_t st);
cudaError_t rle_tables_heads_launch(const uint8_t *d_in, uint64_t N, uint64_t c_base, uint64_t c_first, uint64_t c_last,
                                    uint64_t *d_lasthead, uint32_t *d_meta, uint32_t *d_restsum, uint64_t *d_tile_head,
                                    cudaStream_t st);
cudaError_t rle_tables_oin_launch(uint64_t c_base, uint64_t c_first, uint64_t c_last, uint64_t carry_head,
                                  const uint64_t *d_lasthead, const uint32_t *d_meta, const uint32_t *d_restsum,
                                  const uint64_t *d_tile_head, uint64_t *d_oin, uint64_t *d_tile_sum, cudaStream_t st);
cudaError_t rle_tables_p_launch(uint64_t c_base, uint64_t c_first, uint64_t c_last, uint64_t carry_sum,
                                const uint32_t *d_meta, const uint32_t *d_restsum, const uint64_t *d_oin,
                                const uint64_t *d_tile_sum, uint64_t *d_P, cudaStream_t st);
cudaError_t rle_summary_launch(const uint8_t *d_in, uint64_t N, uint64_t n_chunks, uint64_t *d_lasthead,
                               uint32_t *d_meta, uint32_t *d_restsum, uint64_t *d_oin, uint64_t *d_P,
                               uint64_t *d_tiles /* 2 * rle_scan_tiles(n_chunks) words of scratch */,
                               cudaStream_t st);
cudaError_t rle_emit_launch(const uint8_t *d_in, uint64_t N, uint64_t c_begin, uint64_t c_end,
                            const uint64_t *d_oin, const uint64_t *d_P, const RleBlock *d_blocks,
                            uint32_t n_blocks, uint8_t *d_out, cudaStream_t st);
cudaError_t crc_launch(const uint8_t *d_in, uint64_t N, uint64_t c_begin, uint64_t c_end, const RleBlock *d_blocks,
                       uint32_t n_blocks, uint32_t *d_crc_acc, uint32_t *d_crc, cudaStream_t st);
int rle_walk_cuts(const uint8_t *in, uint64_t N, int level, const uint64_t *P, const uint64_t *o_in,
                  uint64_t n_chunks, std::vector<RleBlock> &blocks, bool final, uint64_t *consumed);

// ---------------------------------------------------------------- K3/K4 BWT (bwt_sort.cu)

struct BwtStats {                 // per bzip2 block, written by the sort kernel
    uint32_t n;                   // RLE1 length of the block
    uint32_t rounds;              // sort rounds executed (incl. the 5-byte round)
    uint32_t tied;                // 1 if the block had identical rotations (period | n)
    uint32_t period;              // period of the block's periodic run if the sort used one (bwt_common.cuh), else 0
    uint64_t sum_active;          // sum over rounds of records sorted (a_r)
    uint64_t sum_active_passes;   // sum over rounds of (records sorted through HBM) * radix passes executed (P_r)
    uint64_t sum_tile;            // records (summed over rounds) that were sorted inside shared memory (no HBM pass)
    uint64_t cyc_build, cyc_radix, cyc_rerank, cyc_tile;   // SM cycles spent per phase (thread 0's clock64)
    uint64_t cyc_final;           // ... and in the last pass (bwt[rank[i]] = S[i-1]); one-CTA kernel only
};

constexpr int BWT_HIST_WORDS = 5 * 1024;
struct BwtArgs {
    const uint8_t *rle;           // RLE1 bytes of all blocks (block b at rle + blk_off[b])
    uint8_t *bwt;                 // BWT bytes, same layout
    const uint64_t *blk_off;      // [n_blocks] byte offset of each block
    const uint32_t *blk_len;      // [n_blocks] n of each block
    uint32_t *ptr;                // [n_blocks] origPtr out
    uint8_t *has_byte;            // [n_blocks][256] presence flags out
    BwtStats *stats;              // [n_blocks] or nullptr
    uint32_t *next_block;         // work-queue counter (zeroed by the host)
    uint32_t n_blocks;
    uint64_t *ws_rec;             // per CTA: 3 (one-CTA kernel) | 2 (cluster kernel) * ws_stride records
    uint32_t *ws_rank;            // per CTA: ws_stride ranks
    size_t ws_stride;             // >= max block length (+ cluster slack), multiple of 16
    uint32_t *ws_hist;            // one-CTA kernel: per CTA BWT_HIST_WORDS words (per-pass digit histograms)
    void *ws_ctl;                 // cluster kernel only: per cluster BWT_CTL_BYTES of control state
    uint32_t *done;               // optional [n_blocks], host-mapped: set to 1 (release.sys) when a block's outputs are complete
    uint32_t *marks;              // optional [n_blocks][VERIFY_MARKS]: row (sorted position) of the rotations 0, 4096, 8192, ... (verify.cu)
    // Blocks with a long periodic run (bwt_common.cuh: Period) are left by the first launch (either kernel) to a
    // follow-up launch of the one-CTA kernel's PERIODIC instantiation, which finishes them in two rounds: the
    // first launch appends them to defer_list / *defer_count, the follow-up launch takes its blocks from
    // blk_list[0 .. *n_blocks_dev).  nullptr: not used.
    uint32_t *defer_list, *defer_count;
    const uint32_t *blk_list, *n_blocks_dev;
};

size_t bwt_smem_bytes();
cudaError_t bwt_max_ctas(int *ctas_per_sm);
// periodic: the instantiation for the blocks of a.blk_list (blocks with a periodic run, left by the first launch)
cudaError_t bwt_launch(const BwtArgs &a, int grid, cudaStream_t stream, bool periodic = false);
// cluster-cooperative variant (bwt_cluster.cu)
size_t bwtc_smem_bytes(int threads);
cudaError_t bwtc_max_clusters(int threads, int C, int *n_clusters);
cudaError_t bwtc_launch(const BwtArgs &a, int threads, int C, int n_clusters, cudaStream_t stream);

// ---------------------------------------------------------------- self-verification (verify.cu)

#define VERIFY_SPACING 4096       /* text positions per backward walk */
#define VERIFY_MARKS 224          /* >= ceil(900000 / VERIFY_SPACING) + 1 */
#define VERIFY_BAD_RLE 1u
#define VERIFY_BAD_BWT 2u
struct VerifyArgs {
    const uint8_t *in_base;       // input bytes, indexed by GLOBAL input position
    uint8_t *rle;                 // RLE1 images (block b at rle + blocks[b].rle_off == rle + blk_off[b])
    uint8_t *bwt;                 // BWT bytes, same layout
    uint32_t *ptr;                // [n_blocks] origPtr
    const uint64_t *blk_off;
    const uint32_t *blk_len;
    const RleBlock *blocks;       // [n_blocks] as cut by the host walk
    uint32_t n_blocks;
    uint32_t *lfl;                // scratch: one word per byte of the rle buffer
    const uint32_t *marks;        // [n_blocks][VERIFY_MARKS] from the sort kernel
    const BwtStats *stats;        // [n_blocks] (tied flag)
    uint32_t *flags;              // [n_blocks] out: VERIFY_BAD_* (zeroed by the host)
    int corrupt;                  // test hook: 1 flip a BWT byte, 2 shift origPtr, 3 flip an RLE1 byte (block 0)
};
cudaError_t verify_launch(const VerifyArgs &a, cudaStream_t st, uint32_t *launches);

// ---------------------------------------------------------------- K5 MTF + RLE2 (mtf.cu)

struct MtfArgs {
    const uint8_t *bwt;           // BWT bytes of all blocks (block b at bwt + blk_off[b])
    uint8_t *idx;                 // scratch: one MTF index byte per position, same layout
    const uint64_t *blk_off;      // [n_blocks]
    const uint32_t *blk_len;      // [n_blocks]
    const uint8_t *has_byte;      // [n_blocks][256]
    uint32_t n_blocks;            // blocks handled by this launch (a list: any subset of the batch, any order)
    const uint32_t *ids;          // [n_blocks] list position -> block id; every per-block array is indexed by block id
    const uint32_t *cseg_base;    // [n_blocks + 1] running segment count over the list
    const uint32_t *seg_base;     // [batch blocks + 1] first global segment of each block
    uint32_t total_segs;          // segments in this launch (= cseg_base[n_blocks])
    uint8_t *seg_list;            // [total_segs][256] distinct bytes, newest first
    uint32_t *seg_cnt;            // [total_segs]
    uint8_t *seg_state;           // [total_segs][256] recency list at the segment start
    uint32_t *num_names;          // [n_blocks] out
    uint16_t *syms;               // symbols out (block b at syms + sym_off[b])
    const uint64_t *sym_off;      // [n_blocks] element offsets, room for blk_len + 1 each
    uint32_t *sym_len;            // [n_blocks] out: m (incl. EOB)
    uint32_t *freqs;              // [n_blocks][258] out
};
cudaError_t mtf_launch(const MtfArgs &a, cudaStream_t st, uint32_t *launches);

// ---------------------------------------------------------------- K6-K8 Huffman + packing (huffman.cu)

struct HuffArgs {
    const uint16_t *syms;         // MTF symbols (block b at syms + sym_off[b])
    const uint64_t *sym_off;      // [n_blocks]
    const uint32_t *sym_len;      // [n_blocks] m
    const uint32_t *num_names;    // [n_blocks]; num_syms = num_names + 2
    const uint32_t *freqs;        // [n_blocks][258]
    uint32_t n_blocks;
    uint8_t *lens;                // [n_blocks][6][258] code lengths (initial tables, then final)
    uint32_t *codes;              // [n_blocks][6][258] (len << 24) | code
    uint32_t *tf;                 // [n_blocks][6][258] table_freqs (zeroed by the host)
    uint32_t *num_tables;         // [n_blocks]
    uint32_t *num_sel;            // [n_blocks]
    const uint8_t *selectors;     // nullptr: every selector is 0 (SURVEY A-Q10)
    uint8_t *sel_out;             // literal loop: the assign kernel records the table of every group here
    int literal;                  // literal loop: table_freqs are used as they are (no closed-form fold)
    size_t sel_stride;
    const uint32_t *span_base;    // [n_blocks + 1] first assign-CTA of each block
    uint32_t *hdr;                // [n_blocks][hdr_stride] header bits as MSB-first words
    size_t hdr_stride;
    uint32_t *hdr_bits;           // [n_blocks]
    int with_block_header;        // 1: magic/crc/ptr/symmap precede the huffman part
    const uint32_t *crc;          // [n_blocks]
    const uint32_t *ptr;          // [n_blocks]
    const uint8_t *has_byte;      // [n_blocks][256]
    uint64_t *blk_bits;           // [n_blocks] bits of each block
    uint64_t *blk_bitoff;         // [n_blocks] bit offset of each block in out_words
    uint64_t bit_base;            // bit offset of the first block
    uint64_t fixed_stride_bits;   // != 0: block b starts at b * fixed_stride_bits (stage tests)
    uint64_t *total_bits;         // bit_base + sum of blk_bits
    uint32_t *out_words;          // output stream (zeroed by the host)
};
cudaError_t huff_launch(const HuffArgs &a, uint32_t total_spans, cudaStream_t st, uint32_t *launches);
cudaError_t huff_launch_literal(HuffArgs a, uint32_t total_spans, cudaStream_t st, uint32_t *launches);
cudaError_t huff_pack_launch(const HuffArgs &a, cudaStream_t st, uint32_t *launches);
cudaError_t huff_rescan_launch(const HuffArgs &a, cudaStream_t st, uint32_t *launches);
uint32_t huff_groups_per_span();

}  // namespace bnz
