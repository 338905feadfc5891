// bwt_common.cuh — record layout and the warp-level digit matching shared by the two BWT sort
// kernels (bwt_sort.cu: one CTA per block; bwt_cluster.cu: one thread-block cluster per block).
#pragma once
#include "common.cuh"

namespace bnz {
namespace bwtk {

// sort records are 64-bit [ key:40 | idx:20 ] (n <= 900 000 < 2^20); ranks are u32, bit 31 = final
constexpr int TILE = 4096;           // records per tile = 32 KB
constexpr int BITS = 8;              // radix digit
constexpr int BINS = 1 << BITS;
constexpr int KEY_BITS = 40;
constexpr int PASSES = KEY_BITS / BITS;
constexpr int IDX_BITS = 20;
constexpr u32 IDX_MASK = (1u << IDX_BITS) - 1u;
constexpr u32 RANK_MASK = (1u << 20) - 1u;
constexpr u32 DONE = 0x80000000u;

__device__ __forceinline__ u32 digit_of(u64 rec, int pass) { return (u32)(rec >> (IDX_BITS + pass * BITS)) & (u32)(BINS - 1); }

// warp-wide "which lanes hold my digit": BITS ballots (match.any costs ~1000 cycles on this part,
// tools/micro/match_bench.cu).  peers = AND over the digit's bits of XNOR(ballot(bit), my bit).
__device__ __forceinline__ u32 match_digit(u32 d)
{
    u32 peers = 0xffffffffu;
#pragma unroll
    for (int b = 0; b < BITS; b++) {
        asm("{\n"
            ".reg .pred p;\n"
            ".reg .b32 t, v;\n"
            "and.b32 t, %1, %2;\n"
            "setp.ne.u32 p, t, 0;\n"
            "vote.sync.ballot.b32 v, p, 0xffffffff;\n"
            "@!p not.b32 v, v;\n"
            "and.b32 %0, %0, v;\n"
            "}\n"
            : "+r"(peers)
            : "r"(d), "r"(1u << b));
    }
    return peers;
}

}  // namespace bwtk
}  // namespace bnz
