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

// ---- round-0 keys: as many symbols as fit KEY_BITS ------------------------------------------------
// A block that uses (nearly) all 256 byte values gets k = 5 raw bytes.  A block with a smaller alphabet
// (text: 55-90 symbols) gets MORE symbols into the same 40 bits: bytes are replaced by their dense codes
// 0..sigma-1 (order preserving) and the key is the base-sigma number c0 c1 ... c(k-1) with the largest k
// such that sigma^k <= 2^40 (k = 6 for sigma <= 101, 7 for <= 52, 8 for <= 32); what is left of the 40
// bits holds the NEXT symbol coarsened to L = floor(2^40 / sigma^k) levels (text, sigma = 56: 35 levels,
// nearly a seventh symbol).  That is sound: the doubling only needs ranks that are consistent with the
// true order and whose ties imply equal h-prefixes (h = k); extra information merely splits more groups.
// The first round then already separates what differs within k symbols, fewer rotations stay active and
// the doubling continues from h = k (measured: DESIGN.md §4).
struct KeyCode {
    u32 sigma, k, L;
    u8 code[256];        // dense code of every present byte
    u8 code2[256];       // the code coarsened to L levels
};

// present[256] (0/1, cleared by the caller) := the bytes of S; *kc := the packing.  All NT >= 256 threads.
template <int NT>
__device__ void build_alphabet(const u8 *__restrict__ S, u32 n, u8 *present, KeyCode *kc, u32 *sh)
{
    const u32 tid = threadIdx.x;
    const u32 *S32 = reinterpret_cast<const u32 *>(S);
    for (u32 i = tid * 4; i < n; i += NT * 4) {
        const u32 w = __ldg(S32 + (i >> 2));
#pragma unroll
        for (int j = 0; j < 4; j++)
            if (i + j < n) present[(w >> (8 * j)) & 0xffu] = 1;
    }
    __syncthreads();
    u32 bal = 0;
    if (tid < 256) {
        bal = __ballot_sync(0xffffffffu, present[tid] != 0);
        if ((tid & 31u) == 0) sh[tid >> 5] = __popc(bal);
    }
    __syncthreads();
    u32 sigma = 0;
#pragma unroll
    for (int w = 0; w < 8; w++) sigma += sh[w];
    const u32 k = sigma > 101 ? 5u : sigma > 52 ? 6u : sigma > 32 ? 7u : 8u;
    u64 pw = 1;
    for (u32 j = 0; j < k; j++) pw *= sigma;
    const u64 room = (1ull << KEY_BITS) / pw;
    const u32 L = room < (u64)sigma ? (u32)room : sigma;      // levels of the coarse next symbol (1: none)
    if (tid < 256) {
        u32 before = 0;
        for (u32 w = 0; w < (tid >> 5); w++) before += sh[w];
        const u32 c = before + __popc(bal & ((1u << (tid & 31u)) - 1u));
        kc->code[tid] = (u8)c;                                // (sigma = 256: the identity)
        kc->code2[tid] = (u8)((c * L) / sigma);
    }
    if (tid == 0) {
        kc->sigma = sigma;
        kc->k = k;
        kc->L = L;
    }
    __syncthreads();
}

// key of rotation i.  S is 16-byte aligned and padded to 16 bytes (aligned 32-bit loads cover the window).
__device__ __forceinline__ u64 raw_key5(const u8 *__restrict__ S, u32 n, u32 i)
{
    const u32 *S32 = reinterpret_cast<const u32 *>(S);
    u64 key = 0;
    if (i + 8 <= n) {
        const u32 w0 = __ldg(S32 + (i >> 2)), w1 = __ldg(S32 + (i >> 2) + 1);
        const u32 sh = (i & 3) * 8;
        const u32 lo = __funnelshift_r(w0, w1, sh);            // bytes i .. i+3 (little endian)
        const u32 b4 = (w1 >> sh) & 0xffu;                     // byte i+4
        key = ((u64)__byte_perm(lo, 0, 0x0123) << 8) | b4;
    } else {
        u32 q = i;
        for (int j = 0; j < 5; j++) {
            key = (key << 8) | S[q];
            q = (q + 1 == n) ? 0 : q + 1;
        }
    }
    return key;
}
__device__ __forceinline__ u64 packed_key(const u8 *__restrict__ S, u32 n, u32 i, const KeyCode &kc)
{
    const u32 *S32 = reinterpret_cast<const u32 *>(S);
    const u32 sigma = kc.sigma, k = kc.k;
    u64 key = 0;
    if (i + 12 <= n) {
        const u32 w0 = __ldg(S32 + (i >> 2)), w1 = __ldg(S32 + (i >> 2) + 1), w2 = __ldg(S32 + (i >> 2) + 2);
        const u32 sh = (i & 3) * 8;
        const u32 lo = __funnelshift_r(w0, w1, sh);            // bytes i .. i+3
        const u32 hi = __funnelshift_r(w1, w2, sh);            // bytes i+4 .. i+7
        u32 nx = (w2 >> sh) & 0xffu;                           // byte i+8: the symbol after the k-th if k = 8
#pragma unroll
        for (int j = 0; j < 8; j++) {
            const u32 b = ((j < 4 ? lo : hi) >> (8 * (j & 3))) & 0xffu;
            if ((u32)j < k) key = key * sigma + kc.code[b];
            else if ((u32)j == k) nx = b;
        }
        key = key * kc.L + kc.code2[nx];
    } else {
        u32 q = i;
        for (u32 j = 0; j < k; j++) {
            key = key * sigma + kc.code[S[q]];
            q = (q + 1 == n) ? 0 : q + 1;
        }
        key = key * kc.L + kc.code2[S[q]];
    }
    return key;
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
