// verify.cu — self-verification of the block pipeline on the device.
//
// The reference has no decoder (README.md:9); its only safety net is the libbz2 round trip of
// fuzz/fuzz_targets/round_trip.rs:8-22.  With bnz_ctx_set("verify", 1) every block is checked
// before anything is emitted:
//   V1  the RLE1 image decodes (bzip2's run rule: four equal bytes, then a repeat count) to
//       exactly the input bytes the block claims to have consumed (lib/rle.rs:102-253);
//   V2  the inverse Burrows-Wheeler transform of (bwt, origPtr) is the RLE1 image
//       (lib/bwt.rs:526-756): LF[i] = C[bwt[i]] + #{j < i : bwt[j] = bwt[i]} is built by a
//       stable counting sort, then the text is walked backwards from the rows of the rotations
//       0, 4096, 8192, ... (the sort kernels note them down: "marks"), 4096 steps per thread,
//       every chain also checking that it arrives at the row the previous mark names;
//   V3  (host, encode.cu) the cut chain covers the input without gap or overlap and the block
//       CRCs are recomputed on the host cores by an independent table-driven CRC-32/BZIP2.
// The entropy stage (MTF, Huffman, bit packing) is not decoded here.
#include "common.cuh"
#include "kernels.h"

namespace bnz {
namespace verify {

constexpr int T = 512;
constexpr int NW = T / 32;
constexpr int K = 8;
constexpr int TILE = T * K;

// lanes of the warp that hold the same byte (8 ballots; match.any is ~1000 cycles on this part)
__device__ __forceinline__ u32 match_byte(u32 d)
{
    u32 peers = 0xffffffffu;
#pragma unroll
    for (int b = 0; b < 8; b++) {
        const u32 v = __ballot_sync(0xffffffffu, (d >> b) & 1u);
        peers &= ((d >> b) & 1u) ? v : ~v;
    }
    return peers;
}

// ------------------------------------------------------------------ V2a: LF mapping, one CTA per block
// lfl[i] = (bwt[i] << 24) | LF[i]
__global__ void __launch_bounds__(T) verify_lf_kernel(const u8 *__restrict__ bwt, const u64 *__restrict__ blk_off,
                                                     const u32 *__restrict__ blk_len, u32 *__restrict__ lfl)
{
    __shared__ u32 hist[256];          // byte counts, then the running cursor C[byte] + bytes seen so far
    __shared__ u32 wcnt[NW][256];      // per warp: bytes of this tile (then: exclusive prefix over the warps)
    __shared__ u32 scratch[40];
    const u32 b = blockIdx.x, tid = threadIdx.x, lane = lane_id(), w = warp_id();
    const u8 *L = bwt + blk_off[b];
    u32 *out = lfl + blk_off[b];       // (block images are 16-byte aligned, so the offsets serve for u32 too)
    const u32 n = blk_len[b];

    for (int i = tid; i < 256; i += T) hist[i] = 0;
    __syncthreads();
    for (u32 base = 0; base < n; base += T) {
        const u32 i = base + tid;
        const u32 d = (i < n) ? L[i] : 0x100u;
        const u32 peers = match_byte(d & 0xffu) & __ballot_sync(0xffffffffu, i < n);
        if (i < n && lane == (u32)(31 - __clz(peers))) atomicAdd(&hist[d], (u32)__popc(peers));
    }
    __syncthreads();
    {
        const u32 v = (tid < 256) ? hist[tid] : 0;
        u32 tot;
        const u32 ex = block_excl_sum<T>(v, scratch, &tot);
        if (tid < 256) hist[tid] = ex;
    }
    __syncthreads();

    for (u32 base = 0; base < n; base += TILE) {
        for (int i = lane; i < 256; i += 32) wcnt[w][i] = 0;
        __syncwarp();
        u32 dk[K], rk[K];
#pragma unroll
        for (int k = 0; k < K; k++) {
            const u32 i = base + w * (K * 32) + k * 32 + lane;
            const bool valid = i < n;
            dk[k] = valid ? L[i] : 0x100u;
            const u32 peers = match_byte(dk[k] & 0xffu) & __ballot_sync(0xffffffffu, valid);
            u32 before = 0;
            const u32 leader = valid ? (u32)(31 - __clz(peers)) : 0u;
            if (valid && lane == leader) {
                before = wcnt[w][dk[k]];
                wcnt[w][dk[k]] = before + __popc(peers);
            }
            __syncwarp();
            before = __shfl_sync(0xffffffffu, before, leader);
            rk[k] = before + __popc(peers & lanemask_lt());
        }
        __syncthreads();
        if (tid < 256) {                // exclusive prefix over the warps; advance the cursor
            u32 run = hist[tid];
#pragma unroll
            for (int ww = 0; ww < NW; ww++) {
                const u32 v = wcnt[ww][tid];
                wcnt[ww][tid] = run;
                run += v;
            }
            hist[tid] = run;
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < K; k++) {
            const u32 i = base + w * (K * 32) + k * 32 + lane;
            if (i < n) out[i] = (dk[k] << 24) | (wcnt[w][dk[k]] + rk[k]);
        }
        __syncthreads();
    }
}

// ------------------------------------------------------------------ V2b: backward walks, one thread per 4096 text positions
__global__ void __launch_bounds__(VERIFY_MARKS) verify_chain_kernel(const u8 *__restrict__ rle, const u64 *__restrict__ blk_off,
                                                                   const u32 *__restrict__ blk_len, const u32 *__restrict__ ptr,
                                                                   const u32 *__restrict__ lfl, const u32 *__restrict__ marks,
                                                                   const BwtStats *__restrict__ stats, u32 *__restrict__ flags)
{
    const u32 b = blockIdx.x, c = threadIdx.x;
    const u32 n = blk_len[b];
    const u8 *S = rle + blk_off[b];
    const u32 *lf = lfl + blk_off[b];
    const u32 *mk = marks + (size_t)b * VERIFY_MARKS;
    const u32 p0 = ptr[b];
    u32 bad = 0;
    if (p0 >= n) {
        if (c == 0) atomicOr(&flags[b], VERIFY_BAD_BWT);
        return;
    }
    if (stats[b].tied) {
        // identical rotations: their rows were never told apart, walk the whole text in one chain
        if (c != 0) return;
        u32 row = p0;
        for (u32 t = n; t-- > 0;) {
            const u32 v = lf[row];
            bad |= (v >> 24) != S[t];
            row = v & 0xffffffu;
        }
        bad |= row != p0;
    } else {
        const u32 lo = c * VERIFY_SPACING;
        if (lo >= n) return;
        const u32 hi = min(n, lo + VERIFY_SPACING);
        if (c == 0) bad |= mk[0] != p0;                        // origPtr is the row of rotation 0
        u32 row = (hi == n) ? p0 : mk[c + 1];                    // row of rotation hi (rotation n = rotation 0)
        if (row >= n) {
            bad = 1;
        } else {
            for (u32 t = hi; t-- > lo;) {
                const u32 v = lf[row];                           // bwt[row(t+1)] = S[t], LF(row(t+1)) = row(t)
                bad |= (v >> 24) != S[t];
                row = v & 0xffffffu;
            }
            bad |= row != mk[c];
        }
    }
    if (bad) atomicOr(&flags[b], VERIFY_BAD_BWT);
}

// ------------------------------------------------------------------ V1: RLE1 image -> input bytes, one thread per block
__global__ void __launch_bounds__(64) verify_rle_kernel(const u8 *__restrict__ in_base, const u8 *__restrict__ rle,
                                                       const RleBlock *__restrict__ blocks, u32 n_blocks,
                                                       u32 *__restrict__ flags)
{
    const u32 b = blockIdx.x * 64 + threadIdx.x;
    if (b >= n_blocks) return;
    const RleBlock bk = blocks[b];
    const u8 *S = rle + bk.rle_off;
    const u8 *in = in_base + bk.s;
    const u64 len = bk.c - bk.s;
    u64 i = 0;                          // decoded bytes so far
    u32 run = 0, last = 0x100u;
    bool bad = false;
    for (u32 j = 0; j < bk.n && !bad; j++) {
        const u32 x = S[j];
        if (run == 4) {                 // x is a repeat count
            if (i + x > len) { bad = true; break; }
            for (u32 r = 0; r < x; r++) bad |= in[i + r] != last;
            i += x;
            run = 0;
            last = 0x100u;              // the next byte starts a new run, whatever its value
            continue;
        }
        if (i >= len || in[i] != x) { bad = true; break; }
        i++;
        run = (x == last) ? run + 1 : 1;
        last = x;
    }
    if (bad || i != len) atomicOr(&flags[b], VERIFY_BAD_RLE);
}

// test hook: flip one byte / origPtr of block 0 behind the sort's back (bnz_ctx_set "verify_corrupt")
__global__ void verify_corrupt_kernel(u8 *rle, u8 *bwt, u32 *ptr, const u64 *blk_off, const u32 *blk_len, int what)
{
    const u32 n = blk_len[0];
    if (what == 1) bwt[blk_off[0] + n / 2] ^= 1u;
    else if (what == 2) ptr[0] = (ptr[0] + 1) % n;
    else if (what == 3) rle[blk_off[0] + n / 3] ^= 0x20u;
}

}  // namespace verify

cudaError_t verify_launch(const VerifyArgs &a, cudaStream_t st, uint32_t *launches)
{
    if (a.n_blocks == 0) return cudaSuccess;
    if (a.corrupt) verify::verify_corrupt_kernel<<<1, 1, 0, st>>>(a.rle, a.bwt, a.ptr, a.blk_off, a.blk_len, a.corrupt);
    verify::verify_rle_kernel<<<(a.n_blocks + 63) / 64, 64, 0, st>>>(a.in_base, a.rle, a.blocks, a.n_blocks, a.flags);
    verify::verify_lf_kernel<<<a.n_blocks, verify::T, 0, st>>>(a.bwt, a.blk_off, a.blk_len, a.lfl);
    verify::verify_chain_kernel<<<a.n_blocks, VERIFY_MARKS, 0, st>>>(a.rle, a.blk_off, a.blk_len, a.ptr, a.lfl, a.marks,
                                                                      a.stats, a.flags);
    if (launches) *launches += 3;
    return cudaGetLastError();
}

}  // namespace bnz
