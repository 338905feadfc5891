// kernels.h — internal interface between the host pipeline (pipeline.cu) and the kernel
// translation units.  Nothing here is part of the public C ABI (include/banzai_b200.h).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stddef.h>

namespace bnz {

// ---------------------------------------------------------------- K3/K4 BWT (bwt_sort.cu)

struct BwtStats {                 // per bzip2 block, written by the sort kernel
    uint32_t n;                   // RLE1 length of the block
    uint32_t rounds;              // sort rounds executed (incl. the 5-byte round)
    uint32_t tied;                // 1 if the block had identical rotations (period | n)
    uint32_t pad;
    uint64_t sum_active;          // sum over rounds of records sorted (a_r)
    uint64_t sum_active_passes;   // sum over rounds of a_r * radix passes executed (P_r)
};

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
    uint64_t *ws_rec;             // per CTA: 2 * ws_stride records
    uint32_t *ws_rank;            // per CTA: ws_stride ranks
    size_t ws_stride;             // >= max block length, multiple of 2
};

size_t bwt_smem_bytes(int bits);
int bwt_passes(int bits);
cudaError_t bwt_max_ctas(int bits, int *ctas_per_sm);
cudaError_t bwt_launch(const BwtArgs &a, int bits, int grid, cudaStream_t stream);

}  // namespace bnz
