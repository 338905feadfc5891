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

// ---- blocks with a long periodic run (zero pages after RLE1, "abab...", a repeated record) ----------
// The reference's SA-IS has no bad case for them (README.md:7); prefix doubling needs log2(n) rounds,
// because the rotations i, i + p, i + 2p, ... of a run with period p agree until the run ends.  But their
// order is known in closed form.  Let S[x] == S[x + p] for s <= x < e - p.  For s <= i and i + p < e the
// rotations i and i + p agree on their first e - i - p bytes and then continue like the rotations e - p
// and e (mod n) — the same two rotations for every i.  They differ (else all of S would be cyclically
// p-periodic, and they would differ within n - (e - i - p) bytes, else rotation i == rotation i + p).
// So with asc = (rotation e - p < rotation e): rotation i < rotation i + p iff asc, for every such i, and by
// transitivity ANY two rotations a < b of [s, e) with a == b (mod p) are ordered by `asc`.  A group of the
// doubling whose members all lie in [s, e) and in one class mod p is therefore finished by sorting it
// by index — no rank[idx + h] is needed (bwt_sort.cu: refine_tile, refine_periodic_big).
struct Period {
    u32 p;               // period of the run, 0 = the block has none (or too short a one)
    u32 s, e;            // S[x] == S[x + p] for s <= x < e - p
    u32 asc;             // rotation i sorts before rotation i + p
    u64 magic;           // ceil(2^44 / p): x / p for x < 2^20 is (x * magic) >> 44 (exact for p < 2^11)
    __device__ __forceinline__ u32 div(u32 x) const { return (u32)(((u64)x * magic) >> 44); }
    __device__ __forceinline__ u32 cls(u32 x) const { return x - div(x) * p; }
};
constexpr u32 PERIOD_MAX = 2040;         // periods looked for (the class is kept in 11 bits)
constexpr u32 PERIOD_WIN = 32;           // bytes of the candidate test
constexpr u32 PERIOD_MIN_N = 32768;      // shorter blocks are not worth the test

// Finds the smallest p <= PERIOD_MAX with S[m .. m+32) == S[m+p .. m+p+32) at the middle m of the block,
// then the maximal run [s, e) around m with that period, and the order of the rotations e - p and e.
// *out (shared memory) is valid for all threads on return; out->p == 0 unless the run covers at least
// half of the block.  `sh` = 4 words of shared scratch.  Called by all NT threads of the CTA.
// Cost: the candidate test reads ~2 KB (every block); the rest streams S twice (candidates only).
template <int NT>
__device__ void detect_period(const u8 *__restrict__ S, u32 n, u32 *sh, Period *out)
{
    const u32 tid = threadIdx.x;
    if (tid == 0) {
        out->p = 0;
        sh[0] = 0xffffffffu;
    }
    __syncthreads();
    if (n < PERIOD_MIN_N) return;
    const u32 m = n / 2;
    for (u32 d = tid + 1; d <= PERIOD_MAX; d += NT) {
        bool eq = true;
        for (u32 j = 0; j < PERIOD_WIN; j++)
            if (S[m + j] != S[m + d + j]) {
                eq = false;
                break;
            }
        if (eq) atomicMin(&sh[0], d);
    }
    __syncthreads();
    const u32 p = sh[0];
    __syncthreads();
    if (p == 0xffffffffu) return;

    // the violations S[x] != S[x + p] nearest to m on both sides
    if (tid == 0) {
        sh[1] = 0;
        sh[2] = n - p;
    }
    __syncthreads();
    u32 lo_v = 0, hi_v = n - p;
#pragma unroll 4
    for (u32 x = tid; x + p < n; x += NT) {
        if (S[x] != S[x + p]) {
            if (x < m) lo_v = max(lo_v, x + 1);
            else hi_v = min(hi_v, x);
        }
    }
    lo_v = __reduce_max_sync(0xffffffffu, lo_v);
    hi_v = __reduce_min_sync(0xffffffffu, hi_v);
    if ((tid & 31u) == 0) {
        atomicMax(&sh[1], lo_v);
        atomicMin(&sh[2], hi_v);
    }
    __syncthreads();
    const u32 s = sh[1], e = sh[2] + p;
    __syncthreads();
    if (e - s < n / 2) return;

    // first byte in which the rotations e - p and e (mod n) differ
    const u32 a0 = e - p, b0 = (e == n) ? 0u : e;
    if (tid == 0) sh[0] = 0xffffffffu;
    __syncthreads();
    u32 found = 0xffffffffu;
    for (u32 base = 0; base < n; base += NT) {
        const u32 u = base + tid;
        if (u < n) {
            u32 xa = a0 + u, xb = b0 + u;
            if (xa >= n) xa -= n;
            if (xb >= n) xb -= n;
            if (S[xa] != S[xb]) atomicMin(&sh[0], u);
        }
        __syncthreads();
        found = sh[0];
        __syncthreads();
        if (found != 0xffffffffu) break;
    }
    if (found == 0xffffffffu) return;    // all of S is cyclically p-periodic: identical rotations, the tie rule decides
    if (tid == 0) {
        u32 xa = a0 + found, xb = b0 + found;
        if (xa >= n) xa -= n;
        if (xb >= n) xb -= n;
        out->s = s;
        out->e = e;
        out->asc = S[xa] < S[xb] ? 1u : 0u;
        out->magic = ((1ull << 44) + p - 1) / p;
        out->p = p;
    }
    __syncthreads();
}

}  // namespace bwtk
}  // namespace bnz
