// rle1.cu — K1 (RLE1: run detection, global prefix sums, block cuts, emission) and K2 (block
// CRC-32/BZIP2).  Replaces rle::rle_one (reference lib/rle.rs:102-253) and crc32::checksum
// (lib/crc32.rs:31-48) for a whole input at once.
//
// Model (proven equal to the reference's 2-byte-hop loop by
// tests/test_oracle.py): RLE1 is greedy per maximal run, 255 input bytes per token, restarted
// at every block start.  Let o(i) = number of equal bytes immediately before i (offset of i in
// its maximal run) and r = o mod 255.  Consuming byte i costs
//        need(r) = 1 (r < 3) | 2 (r == 3: 4th copy + count byte) | 0 (r >= 4)
// output bytes, and a block consumes bytes while the running cost stays <= M = 100000*L - 1.
// With P(x) = sum_{i<x} need(r(i)) (one max-scan for o, one sum-scan for P), the cost of a
// block that starts at s inside a run ending at e0 is g(e0-s) + P(x) - P(e0), g(t) = P of a
// fresh run of length t.  The device produces P and o at 1 KiB chunk granularity; the host
// walks the (inherently sequential) cut chain with a binary search + <= 2 KiB byte scan per
// block; a second kernel then emits every block's RLE1 bytes position-parallel.
#include <string.h>

#include "common.cuh"
#include "kernels.h"

namespace bnz {
namespace rle {

constexpr int CH = RLE_CHUNK;        // 1024 bytes per chunk = 32 lanes x 32 bytes
constexpr int WPB = 8;               // warps (chunks) per CTA
constexpr u32 NOBYTE = 0x100u;
constexpr int STAGE_BYTES = 1312;   // emission staging per chunk: 1 + 1280 + 1 bytes, padded (+4 for the word reads)

__host__ __device__ __forceinline__ u32 f_of(u32 r) { return r < 4 ? r : 5u; }
__host__ __device__ __forceinline__ u32 need_of(u32 r) { return r < 3 ? 1u : (r == 3 ? 2u : 0u); }

struct LaneBytes {
    u32 w[8];          // 32 bytes, little endian
    u32 valid;         // valid bytes in this lane (0..32)
    u32 prev;          // byte before this lane's first byte (NOBYTE if none)
    u32 next;          // byte after this lane's last valid byte (NOBYTE if none)
    __device__ __forceinline__ u32 at(int j) const { return (w[j >> 2] >> ((j & 3) * 8)) & 0xffu; }
};

// warp-cooperative load of chunk c: lane l owns bytes [x_c + 32 l, x_c + 32 l + 32)
__device__ __forceinline__ void load_chunk(const u8 *__restrict__ in, u64 N, u64 xc, LaneBytes &lb)
{
    const u32 lane = lane_id();
    const u64 p0 = xc + (u64)lane * 32;
    if (p0 + 32 <= N) {
        const uint4 *p = reinterpret_cast<const uint4 *>(in + p0);
        uint4 a = __ldg(p), b = __ldg(p + 1);
        lb.w[0] = a.x; lb.w[1] = a.y; lb.w[2] = a.z; lb.w[3] = a.w;
        lb.w[4] = b.x; lb.w[5] = b.y; lb.w[6] = b.z; lb.w[7] = b.w;
        lb.valid = 32;
    } else {
#pragma unroll
        for (int k = 0; k < 8; k++) lb.w[k] = 0;
        lb.valid = p0 < N ? (u32)(N - p0) : 0;
        for (u32 j = 0; j < lb.valid; j++) lb.w[j >> 2] |= (u32)in[p0 + j] << ((j & 3) * 8);
    }
    u32 last = lb.valid ? lb.at((int)lb.valid - 1) : NOBYTE;
    u32 first = lb.valid ? lb.at(0) : NOBYTE;
    u32 up = __shfl_up_sync(0xffffffffu, last, 1);
    u32 down = __shfl_down_sync(0xffffffffu, first, 1);
    if (lane == 0) up = (xc > 0) ? (u32)in[xc - 1] : NOBYTE;
    if (lane == 31) down = (xc + CH < N) ? (u32)in[xc + CH] : NOBYTE;
    lb.prev = up;
    lb.next = (lb.valid == 32) ? down : NOBYTE;
}

// bit j set <=> byte j of this lane starts a new maximal run
__device__ __forceinline__ u32 head_mask(const LaneBytes &lb)
{
    u32 hm = 0;
    u32 p = lb.prev;
#pragma unroll
    for (int j = 0; j < 32; j++) {
        u32 b = lb.at(j);
        if ((u32)j < lb.valid && b != p) hm |= 1u << j;
        p = b;
    }
    return hm;
}

// ------------------------------------------------------------------ kernel 1: chunk summaries
__global__ void __launch_bounds__(WPB * 32) rle_summary_kernel(const u8 *__restrict__ in, u64 N, u64 c_first,
                                                              u64 n_chunks, u64 *__restrict__ lasthead,
                                                              u32 *__restrict__ meta, u32 *__restrict__ restsum)
{
    // chunks [c_first, n_chunks)
    const u64 c = c_first + (u64)blockIdx.x * WPB + warp_id();
    if (c >= n_chunks) return;
    const u32 lane = lane_id();
    const u64 xc = c * CH;
    LaneBytes lb;
    load_chunk(in, N, xc, lb);
    const u32 hm = head_mask(lb);

    // last head at or before each lane start (1-based chunk-local position, 0 = none)
    u32 mine = hm ? lane * 32 + (31 - __clz(hm)) + 1 : 0;
    u32 inc = warp_incl_max(mine);
    u32 before = __shfl_up_sync(0xffffffffu, inc, 1);
    if (lane == 0) before = 0;
    const u32 last_local = __shfl_sync(0xffffffffu, inc, 31);

    // lead = length of the chunk's first run = first head at local position >= 1
    u32 hm0 = (lane == 0) ? (hm & ~1u) : hm;
    u32 bal = __ballot_sync(0xffffffffu, hm0 != 0);
    const u32 chunk_len = (u32)min((u64)CH, N - xc);
    u32 lead = chunk_len;
    if (bal) {
        int fl = __ffs(bal) - 1;
        u32 fh = __shfl_sync(0xffffffffu, hm0, fl);
        lead = fl * 32 + (__ffs(fh) - 1);
    }
    const bool cont = !(__shfl_sync(0xffffffffu, hm, 0) & 1u);   // byte 0 continues the previous run

    // cost of the positions after the lead run (their run offsets are chunk-local)
    u32 sum = 0;
    {
        u32 pos = lane * 32;
        u32 o = (before > 0) ? pos - (before - 1) : 0;
        u32 r = o % 255u;
#pragma unroll
        for (int j = 0; j < 32; j++) {
            if ((u32)j < lb.valid) {
                if (hm & (1u << j)) r = 0;
                if (pos + j >= lead) sum += need_of(r);
                r = (r == 254) ? 0 : r + 1;
            }
        }
    }
    sum = __reduce_add_sync(0xffffffffu, sum);
    if (lane == 0) {
        lasthead[c] = last_local ? xc + last_local : 0;     // 1-based global position
        meta[c] = lead | (cont ? 0x80000000u : 0u);
        restsum[c] = sum;
    }
}

// ------------------------------------------------------------------ kernel 2: scans over chunks
// single CTA: o_in[c] = run offset of the chunk's first byte, P[c] = cost prefix at the chunk
// start, P[n_chunks] = total.
constexpr int ST = 1024;
constexpr int SI = 8;                 // chunks per thread and tile
constexpr int STILE = ST * SI;        // chunks per CTA

// block-wide (ST threads) exclusive scans; *total = aggregate of the whole CTA
__device__ __forceinline__ u64 block_excl_max64(u64 v, u64 *sh, u64 *total)
{
    const u32 lane = lane_id(), w = warp_id();
    u64 inc = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        u64 t = __shfl_up_sync(0xffffffffu, inc, d);
        if (lane >= (u32)d) inc = max(inc, t);
    }
    u64 ex = __shfl_up_sync(0xffffffffu, inc, 1);
    if (lane == 0) ex = 0;
    if (lane == 31) sh[w] = inc;
    __syncthreads();
    if (w == 0) {
        u64 vi = sh[lane];
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            u64 t = __shfl_up_sync(0xffffffffu, vi, d);
            if (lane >= (u32)d) vi = max(vi, t);
        }
        u64 ve = __shfl_up_sync(0xffffffffu, vi, 1);
        if (lane == 0) ve = 0;
        sh[lane] = ve;
        if (lane == 31) sh[32] = vi;
    }
    __syncthreads();
    const u64 r = max(sh[w], ex);
    *total = sh[32];
    __syncthreads();
    return r;
}
__device__ __forceinline__ u64 block_excl_sum64(u64 v, u64 *sh, u64 *total)
{
    const u32 lane = lane_id(), w = warp_id();
    const u64 si = warp_incl_sum64(v);
    if (lane == 31) sh[w] = si;
    __syncthreads();
    if (w == 0) {
        u64 x = sh[lane];
        u64 xi = warp_incl_sum64(x);
        sh[lane] = xi - x;
        if (lane == 31) sh[32] = xi;
    }
    __syncthreads();
    const u64 r = sh[w] + si - v;
    *total = sh[32];
    __syncthreads();
    return r;
}

// The chunk tables are two chained scans over the chunks: a running maximum (position of the last
// run head so far -> run offset o_in at every chunk start), then a running sum of the chunk costs
// (-> P).  One CTA per tile of STILE chunks, three launches: tile maxima; o_in + tile cost sums
// (the carry is the maximum over the earlier tiles, at most a few hundred values); P.
// Tile t covers the chunks [c_base + t * STILE, c_base + (t + 1) * STILE): c_base is the first chunk
// of the range this device owns (0 on a single device), so ranges need no tile alignment.
__global__ void __launch_bounds__(ST) rle_scan_heads_kernel(const u64 *__restrict__ lasthead, u64 n_chunks, u64 c_base,
                                                           u32 tile0, u64 *__restrict__ tile_head)
{
    __shared__ u64 sh[40];
    const u32 tile = tile0 + blockIdx.x;
    const u64 c0 = c_base + (u64)tile * STILE + (u64)threadIdx.x * SI;
    u64 agg = 0;
#pragma unroll
    for (int k = 0; k < SI; k++)
        if (c0 + k < n_chunks) agg = max(agg, lasthead[c0 + k]);
    u64 tot;
    block_excl_max64(agg, sh, &tot);
    if (threadIdx.x == 0) tile_head[tile] = tot;
}

// cost of chunk c given the run offset at its start
__device__ __forceinline__ u64 chunk_cost(u32 mt, u32 rs, u64 oin)
{
    const u32 lead = mt & 0x7fffffffu;
    const u32 r_in = (u32)(oin % 255u);
    const u32 t = r_in + lead;
    return 5u * (t / 255u) + f_of(t % 255u) - f_of(r_in) + rs;
}

__global__ void __launch_bounds__(ST) rle_scan_oin_kernel(const u64 *__restrict__ lasthead, const u32 *__restrict__ meta,
                                                         const u32 *__restrict__ restsum, u64 n_chunks, u64 c_base,
                                                         u32 tile0, u64 carry_in, const u64 *__restrict__ tile_head,
                                                         u64 *__restrict__ o_in, u64 *__restrict__ tile_sum)
{
    __shared__ u64 sh[40];
    const u32 t = tile0 + blockIdx.x;
    u64 tot;
    // carry: last run head in the earlier tiles (carry_in: in the ranges of the devices before this one)
    u64 cm = carry_in;
    for (u32 q = threadIdx.x; q < t; q += ST) cm = max(cm, tile_head[q]);
    block_excl_max64(cm, sh, &tot);
    const u64 carry_head = tot;

    const u64 c0 = c_base + (u64)t * STILE + (u64)threadIdx.x * SI;
    u64 lh[SI];
    u64 agg = 0;
#pragma unroll
    for (int k = 0; k < SI; k++) {
        lh[k] = (c0 + k < n_chunks) ? lasthead[c0 + k] : 0;
        agg = max(agg, lh[k]);
    }
    u64 hprev = max(block_excl_max64(agg, sh, &tot), carry_head);
    u64 tsum = 0;
#pragma unroll
    for (int k = 0; k < SI; k++) {
        if (c0 + k < n_chunks) {
            const u32 mt = meta[c0 + k];
            u64 oin = 0;
            if (mt & 0x80000000u) oin = (c0 + k) * CH - (hprev - 1);
            o_in[c0 + k] = oin;
            tsum += chunk_cost(mt, restsum[c0 + k], oin);
        }
        hprev = max(hprev, lh[k]);
    }
    block_excl_sum64(tsum, sh, &tot);
    if (threadIdx.x == 0) tile_sum[t] = tot;
}

__global__ void __launch_bounds__(ST) rle_scan_p_kernel(const u32 *__restrict__ meta, const u32 *__restrict__ restsum,
                                                       const u64 *__restrict__ o_in, u64 n_chunks, u64 c_base,
                                                       u32 tile0, u32 n_tiles, u64 carry_in,
                                                       const u64 *__restrict__ tile_sum, u64 *__restrict__ P)
{
    __shared__ u64 sh[40];
    const u32 t = tile0 + blockIdx.x;
    u64 tot;
    u64 cs = 0;
    for (u32 q = threadIdx.x; q < t; q += ST) cs += tile_sum[q];
    block_excl_sum64(cs, sh, &tot);
    const u64 carry_sum = tot + carry_in;           // carry_in = P at c_base (cost of the earlier devices' ranges)

    const u64 c0 = c_base + (u64)t * STILE + (u64)threadIdx.x * SI;
    u64 sv[SI];
    u64 tsum = 0;
#pragma unroll
    for (int k = 0; k < SI; k++) {
        sv[k] = (c0 + k < n_chunks) ? chunk_cost(meta[c0 + k], restsum[c0 + k], o_in[c0 + k]) : 0;
        tsum += sv[k];
    }
    u64 pre = carry_sum + block_excl_sum64(tsum, sh, &tot);
#pragma unroll
    for (int k = 0; k < SI; k++) {
        if (c0 + k < n_chunks) P[c0 + k] = pre;
        pre += sv[k];
    }
    if (t + 1 == n_tiles && threadIdx.x == 0) P[n_chunks] = carry_sum + tot;
}

// ------------------------------------------------------------------ kernel 3: emission
__device__ __forceinline__ u64 g_of(u64 t) { return 5ull * (t / 255ull) + f_of((u32)(t % 255ull)); }

__global__ void __launch_bounds__(WPB * 32) rle_emit_kernel(const u8 *__restrict__ in, u64 N, u64 c_begin, u64 n_chunks,
                                                           const u64 *__restrict__ o_in, const u64 *__restrict__ P,
                                                           const RleBlock *__restrict__ blocks, u32 n_blocks,
                                                           u8 *__restrict__ out)
{
    // chunks [c_begin, n_chunks) of the GLOBAL input; `in`, `o_in`, `P` are indexed globally (the
    // host passes rebased pointers when only a sub-range is resident on this device)
    const u64 c = c_begin + (u64)blockIdx.x * WPB + warp_id();
    if (c >= n_chunks) return;
    const u32 lane = lane_id();
    const u64 xc = c * CH;
    LaneBytes lb;
    load_chunk(in, N, xc, lb);
    const u32 hm = head_mask(lb);

    u32 mine = hm ? lane * 32 + (31 - __clz(hm)) + 1 : 0;
    u32 inc = warp_incl_max(mine);
    u32 before = __shfl_up_sync(0xffffffffu, inc, 1);
    if (lane == 0) before = 0;

    // run offset (mod 255) of this lane's first byte
    const u64 oin = o_in[c];
    const u32 pos0 = lane * 32;
    u32 r0 = (before > 0) ? (pos0 - (before - 1)) % 255u : (u32)((oin + pos0) % 255ull);

    // pass 1: this lane's cost, then the warp prefix -> P at the lane start
    u32 cost = 0;
    bool long_run = false;              // some byte of this lane is the 4th or a later byte of its run
    {
        u32 r = r0;
#pragma unroll
        for (int j = 0; j < 32; j++) {
            if ((u32)j < lb.valid) {
                if (hm & (1u << j)) r = 0;
                cost += need_of(r);
                long_run |= r >= 3;
                r = (r == 254) ? 0 : r + 1;
            }
        }
    }
    u32 cinc = warp_incl_sum(cost);
    u64 Pi = P[c] + (cinc - cost);

    // block that owns the chunk start: last block with s <= xc
    u32 lo = 0, hi = n_blocks;          // invariant: blocks[lo].s <= xc < blocks[hi].s
    while (hi - lo > 1) {
        u32 mid = (lo + hi) >> 1;
        if (blocks[mid].s <= xc) lo = mid; else hi = mid;
    }
    u32 k = lo;
    RleBlock bk = blocks[k];

    // pass 2: emit.  A chunk that lies inside one block writes one contiguous range of that
    // block's image (tokens are consecutive; only a count byte at either edge may belong to the
    // neighbouring chunk), so its bytes are collected in shared memory and stored as aligned words.
    __shared__ __align__(16) u8 stage_all[WPB][STAGE_BYTES];
    u8 *stg = stage_all[warp_id()];
    const bool simple = bk.s <= xc && xc + CH <= bk.c;        // warp-uniform
    u64 W = 0;                                                // image position of stg[0]
    if (simple) {
        const u64 before0 = (xc < bk.e0) ? g_of(xc - bk.s) : bk.u0 + (P[c] - bk.P_e0);
        W = before0 - 1;                                      // (wraps for the first byte of a block: only differences are used)
    }
    u32 wlo = 0xffffffffu, whi = 0;                           // written range, relative to W
    // no run reaches four bytes anywhere in the chunk (and the chunk lies behind the block's first
    // run): every byte is a literal, the image is the input shifted — copy the words
    const bool literal = simple && xc >= bk.e0 && !__any_sync(0xffffffffu, long_run);
    if (literal) {
        W += 1;                                               // stg[0] is the chunk's first byte
        u32 *stg32w = reinterpret_cast<u32 *>(stg);
#pragma unroll
        for (int k = 0; k < 8; k++) stg32w[lane * 8 + k] = lb.w[k];
        wlo = 0;
        whi = CH;
    }
    u32 r = r0;
    u64 i = xc + pos0;
#pragma unroll 4
    for (int j = 0; j < 32 && !literal; j++, i++) {
        if ((u32)j >= lb.valid) break;
        if (hm & (1u << j)) r = 0;
        while (i >= bk.c && k + 1 < n_blocks) { k++; bk = blocks[k]; }
        if (i < bk.s || i >= bk.c) {        // byte belongs to a block owned by another device
            Pi += need_of(r);
            r = (r == 254) ? 0 : r + 1;
            continue;
        }
        const u32 b = lb.at(j);
        const u32 nb = (j + 1 < 32) ? (((u32)(j + 1) < lb.valid) ? lb.at(j + 1) : NOBYTE) : lb.next;
        u32 rb;            // offset inside the block-relative token
        u64 before_out;    // block-relative output bytes accounted before this byte
        if (i < bk.e0) {
            u64 ob = i - bk.s;
            rb = (u32)(ob % 255ull);
            before_out = g_of(ob);
        } else {
            rb = r;
            before_out = bk.u0 + (Pi - bk.P_e0);
        }
        const bool last_tok = (rb == 254) || (i + 1 == bk.c) || (nb != b);
        if (simple) {
            const u32 q = (u32)(before_out - W);
            if (rb < 4) {
                stg[q] = (u8)b;
                wlo = min(wlo, q);
                whi = max(whi, q + 1);
                if (rb == 3 && last_tok) {
                    stg[q + 1] = 0;
                    whi = max(whi, q + 2);
                }
            } else if (last_tok) {
                stg[q - 1] = (u8)(rb - 3);
                wlo = min(wlo, q - 1);
                whi = max(whi, q);
            }
        } else {
            u8 *o = out + bk.rle_off;
            if (rb < 4) {
                o[before_out] = (u8)b;
                if (rb == 3 && last_tok) o[before_out + 1] = 0;
            } else if (last_tok) {
                o[before_out - 1] = (u8)(rb - 3);
            }
        }
        Pi += need_of(r);
        r = (r == 254) ? 0 : r + 1;
    }
    if (!simple) return;
    wlo = __reduce_min_sync(0xffffffffu, wlo);
    whi = __reduce_max_sync(0xffffffffu, whi);
    if (wlo >= whi) return;                                   // the chunk sits inside one long token
    __syncwarp();
    u8 *g = out + bk.rle_off + (W + wlo);                     // first image byte of this chunk
    const u32 nbytes = whi - wlo;
    const u32 head = min(nbytes, (u32)((4 - ((uintptr_t)g & 3)) & 3));
    if (lane < head) g[lane] = stg[wlo + lane];
    const u32 nwords = (nbytes - head) >> 2;
    const u32 s0 = wlo + head;                                // stage offset of the first full word
    const u32 *stg32 = reinterpret_cast<const u32 *>(stg);
    u32 *g32 = reinterpret_cast<u32 *>(g + head);
    for (u32 q = lane; q < nwords; q += 32) {
        const u32 so = s0 + 4 * q;
        const u32 w0 = stg32[so >> 2], w1 = stg32[(so >> 2) + 1];
        g32[q] = __funnelshift_r(w0, w1, (so & 3) * 8);
    }
    const u32 done = head + 4 * nwords;
    if (lane < nbytes - done) g[done + lane] = stg[wlo + done + lane];
}

// ------------------------------------------------------------------ K2: CRC-32/BZIP2
// MSB-first CRC, poly 0x04C11DB7.  raw(M) = M(x) x^32 mod P (zero init); a block's CRC is
// assembled from per-chunk raw CRCs:  raw(A||B) = raw(A) x^(8|B|) + raw(B).
constexpr u32 POLY = 0x04C11DB7u;

__host__ __device__ __forceinline__ u32 gf_mul(u32 a, u32 b)   // a*b mod P, bit 31 = x^31
{
    u32 r = 0;
#pragma unroll 8
    for (int i = 31; i >= 0; i--) {
        r = (r << 1) ^ ((r & 0x80000000u) ? POLY : 0u);
        if ((b >> i) & 1u) r ^= a;
    }
    return r;
}

struct CrcTables {
    u32 byte_tab[256];     // CRC of one byte
    u32 lane_pow[32];      // x^(8*32*j)
    u32 pow2[48];          // x^(8*2^k)
};
__constant__ CrcTables c_crc;

// x^(8e) mod P, computed by the whole warp: lane k contributes bit k of e
__device__ __forceinline__ u32 warp_xpow8(u64 e)
{
    const u32 lane = lane_id();
    u32 v = 1u;     // the polynomial "1"
    // lanes 0..31 cover bits 0..31; bits 32..47 folded in by lanes 0..15 afterwards
    if ((e >> lane) & 1ull) v = c_crc.pow2[lane];
    if (lane < 16 && ((e >> (32 + lane)) & 1ull)) v = gf_mul(v, c_crc.pow2[32 + lane]);
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) {
        u32 o = __shfl_xor_sync(0xffffffffu, v, d);
        v = gf_mul(v, o);
    }
    return v;      // same value in every lane
}

__global__ void __launch_bounds__(WPB * 32) crc_chunk_kernel(const u8 *__restrict__ in, u64 N, u64 c_begin, u64 n_chunks,
                                                            const RleBlock *__restrict__ blocks, u32 n_blocks,
                                                            u32 *__restrict__ acc)
{
    __shared__ u32 tab[256];
    for (int i = threadIdx.x; i < 256; i += WPB * 32) tab[i] = c_crc.byte_tab[i];
    __syncthreads();
    const u64 c = c_begin + (u64)blockIdx.x * WPB + warp_id();
    if (c >= n_chunks) return;
    const u32 lane = lane_id();
    const u64 xc = c * CH;
    const u64 xe = min(xc + (u64)CH, N);

    u32 lo = 0, hi = n_blocks;
    while (hi - lo > 1) {
        u32 mid = (lo + hi) >> 1;
        if (blocks[mid].s <= xc) lo = mid; else hi = mid;
    }
    u32 k = lo;
    RleBlock bk = blocks[k];

    if (xe - xc == CH && xe <= bk.c && xc >= bk.s) {
        // fast path: full chunk inside one block
        const uint4 *p = reinterpret_cast<const uint4 *>(in + xc) + lane * 2;
        uint4 a = __ldg(p), b = __ldg(p + 1);
        u32 w[8] = { a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w };
        u32 crc = 0;
#pragma unroll
        for (int j = 0; j < 32; j++) {
            u32 byte = (w[j >> 2] >> ((j & 3) * 8)) & 0xffu;
            crc = (crc << 8) ^ tab[(crc >> 24) ^ byte];
        }
        crc = gf_mul(crc, c_crc.lane_pow[31 - lane]);
        crc = __reduce_xor_sync(0xffffffffu, crc);
        u32 sh = warp_xpow8(bk.c - xe);
        if (lane == 0) atomicXor(&acc[k], gf_mul(crc, sh));
    } else {
        // slow path: chunk cut by a block boundary or by the end of the input
        u64 i = max(xc, bk.s);               // bytes before the first local block belong elsewhere
        while (i < xe) {
            while (i >= bk.c && k + 1 < n_blocks) { k++; bk = blocks[k]; }
            if (i >= bk.c) break;            // past the last local block
            u64 stop = min(xe, bk.c);
            u32 crc = 0;
            if (lane == 0) {
                for (u64 q = i; q < stop; q++) crc = (crc << 8) ^ tab[(crc >> 24) ^ in[q]];
            }
            crc = __shfl_sync(0xffffffffu, crc, 0);
            u32 sh = warp_xpow8(bk.c - stop);
            if (lane == 0) atomicXor(&acc[k], gf_mul(crc, sh));
            i = stop;
        }
    }
}

// acc -> CRC-32/BZIP2: fold in init and xorout.  One warp per block.
__global__ void __launch_bounds__(WPB * 32) crc_finalize_kernel(const u32 *__restrict__ acc,
                                                               const RleBlock *__restrict__ blocks, u32 n_blocks,
                                                               u32 *__restrict__ crc)
{
    const u32 b = blockIdx.x * WPB + warp_id();
    if (b >= n_blocks) return;
    const u64 len = blocks[b].c - blocks[b].s;
    const u32 p = warp_xpow8(len);
    if (lane_id() == 0) crc[b] = acc[b] ^ gf_mul(0xFFFFFFFFu, p) ^ 0xFFFFFFFFu;
}

}  // namespace rle

// ---------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------

static rle::CrcTables h_crc;
static bool h_crc_ready = false;

static void crc_tables_init()
{
    if (h_crc_ready) return;
    for (u32 b = 0; b < 256; b++) {
        u32 c = b << 24;
        for (int k = 0; k < 8; k++) c = (c & 0x80000000u) ? (c << 1) ^ rle::POLY : (c << 1);
        h_crc.byte_tab[b] = c;
    }
    // x^8 = 0x100 as a polynomial
    u32 x8 = 0x100u;
    h_crc.pow2[0] = x8;
    for (int k = 1; k < 48; k++) h_crc.pow2[k] = rle::gf_mul(h_crc.pow2[k - 1], h_crc.pow2[k - 1]);
    u32 x256 = h_crc.pow2[5];            // x^(8*32)
    h_crc.lane_pow[0] = 1u;
    for (int j = 1; j < 32; j++) h_crc.lane_pow[j] = rle::gf_mul(h_crc.lane_pow[j - 1], x256);
    h_crc_ready = true;
}

cudaError_t crc_upload_tables()
{
    crc_tables_init();
    return cudaMemcpyToSymbol(rle::c_crc, &h_crc, sizeof h_crc);
}

uint32_t crc_finalize(uint32_t acc, uint64_t len)
{
    crc_tables_init();
    // crc = raw(M) ^ (0xFFFFFFFF * x^(8 len)) ^ 0xFFFFFFFF   (init and xorout of CRC-32/BZIP2)
    u32 p = 1u;
    for (int k = 0; k < 48; k++)
        if ((len >> k) & 1ull) p = rle::gf_mul(p, h_crc.pow2[k]);
    return acc ^ rle::gf_mul(0xFFFFFFFFu, p) ^ 0xFFFFFFFFu;
}

// The chunk tables in three steps, so that several devices can each build the tables of their own
// chunk range [c_base, ...) and exchange two scalars in between (encode.cu, encode_sharded):
//   heads: chunk summaries + per-tile maximum of the last run head      -> tile_head[]
//   oin  : run offset at every chunk start + per-tile cost sums          -> o_in[], tile_sum[]
//          (carry_head = last run head, 1-based global position, before chunk c_base; 0 = none)
//   p    : cost prefix                                                   -> P[], P[c_last]
//          (carry_sum = P at chunk c_base)
// Chunks [c_first, c_last) are processed; (c_first - c_base) must be a multiple of the scan tile and
// the tables / tile aggregates of [c_base, c_first) must already be there (input arriving in pieces).
// All table pointers are indexed by GLOBAL chunk number (the caller rebases them).
static inline unsigned tiles_upto(uint64_t c_base, uint64_t c) { return (unsigned)((c - c_base + rle::STILE - 1) / rle::STILE); }

cudaError_t rle_tables_heads_launch(const uint8_t *d_in, uint64_t N, uint64_t c_base, uint64_t c_first, uint64_t c_last,
                                    uint64_t *d_lasthead, uint32_t *d_meta, uint32_t *d_restsum, uint64_t *d_tile_head,
                                    cudaStream_t st)
{
    if (c_last <= c_first) return cudaSuccess;
    if ((c_first - c_base) % rle::STILE) return cudaErrorInvalidValue;
    unsigned grid = (unsigned)((c_last - c_first + rle::WPB - 1) / rle::WPB);
    rle::rle_summary_kernel<<<grid, rle::WPB * 32, 0, st>>>(d_in, N, c_first, c_last, d_lasthead, d_meta, d_restsum);
    const unsigned tile0 = tiles_upto(c_base, c_first), tile1 = tiles_upto(c_base, c_last);
    rle::rle_scan_heads_kernel<<<tile1 - tile0, rle::ST, 0, st>>>(d_lasthead, c_last, c_base, tile0, d_tile_head);
    return cudaGetLastError();
}

cudaError_t rle_tables_oin_launch(uint64_t c_base, uint64_t c_first, uint64_t c_last, uint64_t carry_head,
                                  const uint64_t *d_lasthead, const uint32_t *d_meta, const uint32_t *d_restsum,
                                  const uint64_t *d_tile_head, uint64_t *d_oin, uint64_t *d_tile_sum, cudaStream_t st)
{
    if (c_last <= c_first) return cudaSuccess;
    const unsigned tile0 = tiles_upto(c_base, c_first), tile1 = tiles_upto(c_base, c_last);
    rle::rle_scan_oin_kernel<<<tile1 - tile0, rle::ST, 0, st>>>(d_lasthead, d_meta, d_restsum, c_last, c_base, tile0, carry_head,
                                                                 d_tile_head, d_oin, d_tile_sum);
    return cudaGetLastError();
}

cudaError_t rle_tables_p_launch(uint64_t c_base, uint64_t c_first, uint64_t c_last, uint64_t carry_sum,
                                const uint32_t *d_meta, const uint32_t *d_restsum, const uint64_t *d_oin,
                                const uint64_t *d_tile_sum, uint64_t *d_P, cudaStream_t st)
{
    if (c_last <= c_first) return cudaSuccess;
    const unsigned tile0 = tiles_upto(c_base, c_first), tile1 = tiles_upto(c_base, c_last);
    rle::rle_scan_p_kernel<<<tile1 - tile0, rle::ST, 0, st>>>(d_meta, d_restsum, d_oin, c_last, c_base, tile0, tile1, carry_sum,
                                                               d_tile_sum, d_P);
    return cudaGetLastError();
}

// all three steps on one device (c_base = 0, no carries)
cudaError_t rle_summary_range_launch(const uint8_t *d_in, uint64_t N, uint64_t n_chunks_total, uint64_t c_first,
                                     uint64_t c_last, uint64_t *d_lasthead, uint32_t *d_meta, uint32_t *d_restsum,
                                     uint64_t *d_oin, uint64_t *d_P, uint64_t *d_tiles, cudaStream_t st)
{
    if (c_last <= c_first) return cudaSuccess;
    const unsigned n_tiles_total = (unsigned)rle_scan_tiles(n_chunks_total);
    uint64_t *tile_head = d_tiles, *tile_sum = d_tiles + n_tiles_total;
    cudaError_t e = rle_tables_heads_launch(d_in, N, 0, c_first, c_last, d_lasthead, d_meta, d_restsum, tile_head, st);
    if (e != cudaSuccess) return e;
    e = rle_tables_oin_launch(0, c_first, c_last, 0, d_lasthead, d_meta, d_restsum, tile_head, d_oin, tile_sum, st);
    if (e != cudaSuccess) return e;
    return rle_tables_p_launch(0, c_first, c_last, 0, d_meta, d_restsum, d_oin, tile_sum, d_P, st);
}

cudaError_t rle_summary_launch(const uint8_t *d_in, uint64_t N, uint64_t n_chunks, uint64_t *d_lasthead,
                               uint32_t *d_meta, uint32_t *d_restsum, uint64_t *d_oin, uint64_t *d_P,
                               uint64_t *d_tiles, cudaStream_t st)
{
    return rle_summary_range_launch(d_in, N, n_chunks, 0, n_chunks, d_lasthead, d_meta, d_restsum, d_oin, d_P, d_tiles, st);
}

uint64_t rle_scan_tile_chunks() { return rle::STILE; }

size_t rle_scan_tiles(uint64_t n_chunks) { return (size_t)((n_chunks + rle::STILE - 1) / rle::STILE); }

cudaError_t rle_emit_launch(const uint8_t *d_in, uint64_t N, uint64_t c_begin, uint64_t c_end,
                            const uint64_t *d_oin, const uint64_t *d_P, const RleBlock *d_blocks,
                            uint32_t n_blocks, uint8_t *d_out, cudaStream_t st)
{
    if (c_end <= c_begin) return cudaSuccess;
    unsigned grid = (unsigned)((c_end - c_begin + rle::WPB - 1) / rle::WPB);
    rle::rle_emit_kernel<<<grid, rle::WPB * 32, 0, st>>>(d_in, N, c_begin, c_end, d_oin, d_P, d_blocks, n_blocks, d_out);
    return cudaGetLastError();
}

// block CRCs (acc must be zeroed): chunk CRCs + combine, then init/xorout; crc[b] is final
cudaError_t crc_launch(const uint8_t *d_in, uint64_t N, uint64_t c_begin, uint64_t c_end, const RleBlock *d_blocks,
                       uint32_t n_blocks, uint32_t *d_crc_acc, uint32_t *d_crc, cudaStream_t st)
{
    if (c_end <= c_begin || n_blocks == 0) return cudaSuccess;
    unsigned grid = (unsigned)((c_end - c_begin + rle::WPB - 1) / rle::WPB);
    rle::crc_chunk_kernel<<<grid, rle::WPB * 32, 0, st>>>(d_in, N, c_begin, c_end, d_blocks, n_blocks, d_crc_acc);
    rle::crc_finalize_kernel<<<(n_blocks + rle::WPB - 1) / rle::WPB, rle::WPB * 32, 0, st>>>(d_crc_acc, d_blocks, n_blocks, d_crc);
    return cudaGetLastError();
}

// The sequential cut chain (the part of lib/lib.rs:101-126 + lib/rle.rs:121-240 that couples
// consecutive blocks).  `in` is the host copy of the input, P/o_in the chunk tables.
// `final` == false: more input follows after in[N-1] (streaming batches); a block whose cut is
// only the end of the available data is then incomplete and is left to the next batch.
// *consumed = input bytes covered by the returned blocks.
int rle_walk_cuts(const uint8_t *in, uint64_t N, int level, const uint64_t *P, const uint64_t *o_in,
                  uint64_t n_chunks, std::vector<RleBlock> &blocks, bool final, uint64_t *consumed)
{
    using rle::need_of;
    const uint64_t M = (uint64_t)100000 * level - 1;
    const uint64_t CHB = RLE_CHUNK;
    const uint64_t lmax = 255 * (M / 5) + ((M % 5 == 4) ? 3 : (M % 5));
    blocks.clear();
    uint64_t s = 0, rle_off = 0;

    // eight bytes at once when no byte equals its predecessor (the common case outside runs):
    // each of them starts a run, costs 1, and leaves the run offset at 1
    auto no_adjacent_equal8 = [&](uint64_t i) -> bool {       // requires 1 <= i and i + 8 <= N
        uint64_t a, b;
        memcpy(&a, in + i, 8);
        memcpy(&b, in + i - 1, 8);
        uint64_t x = a ^ b;                                    // zero byte <=> in[i+k] == in[i+k-1]
        return (((x - 0x0101010101010101ull) & ~x & 0x8080808080808080ull) == 0);
    };

    // P and run-offset (mod 255) at an arbitrary position x, scanning from its chunk start
    auto p_at = [&](uint64_t x, uint32_t &r_out) -> uint64_t {
        uint64_t c = x / CHB;
        if (c >= n_chunks) { r_out = 0; return P[n_chunks]; }
        uint64_t i = c * CHB;
        uint64_t p = P[c];
        uint32_t r = (uint32_t)(o_in[c] % 255);
        while (i < x) {
            if (i >= 1 && i + 8 <= x && no_adjacent_equal8(i)) {
                p += 8;
                r = 1;
                i += 8;
                continue;
            }
            if (i == 0 || in[i] != in[i - 1]) r = 0;
            p += need_of(r);
            r = (r == 254) ? 0 : r + 1;
            i++;
        }
        r_out = r;
        return p;
    };

    while (s < N) {
        const uint8_t b = in[s];
        // end of the run that contains s
        uint64_t e0 = s + 1;
        {
            uint64_t lim = std::min(N, (s / CHB + 1) * CHB);
            while (e0 < lim && in[e0] == b) e0++;
            if (e0 == lim && lim < N && in[lim] == b) {
                // the run crosses chunk boundaries: gallop over chunks using o_in
                uint64_t c_lo = lim / CHB;                       // run known to reach x_{c_lo}
                uint64_t step = 1, c_hi = c_lo;
                auto reaches = [&](uint64_t c) { return c < n_chunks && o_in[c] >= c * CHB - s; };
                while (reaches(c_lo + step)) { c_lo += step; step *= 2; }
                c_hi = std::min(c_lo + step, n_chunks);          // run does not reach x_{c_hi} (or c_hi == n_chunks)
                while (c_hi - c_lo > 1) {
                    uint64_t mid = (c_lo + c_hi) / 2;
                    if (reaches(mid)) c_lo = mid; else c_hi = mid;
                }
                e0 = c_lo * CHB;
                uint64_t lim2 = std::min(N, (c_lo + 1) * CHB);
                while (e0 < lim2 && in[e0] == b) e0++;
            }
        }
        RleBlock bk;
        bk.s = s;
        bk.rle_off = rle_off;
        const uint64_t L0 = e0 - s;
        if (L0 > lmax) {
            // the block fills up inside its first run
            bk.c = s + lmax;
            bk.e0 = bk.c;
            bk.u0 = (uint32_t)(5 * (lmax / 255) + rle::f_of((uint32_t)(lmax % 255)));
            bk.P_e0 = 0;
            bk.n = bk.u0;
        } else {
            const uint64_t u0 = 5 * (L0 / 255) + rle::f_of((uint32_t)(L0 % 255));
            const uint64_t B = M - u0;
            uint32_t r = 0;
            const uint64_t Pe0 = p_at(e0, r);
            const uint64_t target = Pe0 + B;
            // last chunk start with P <= target.  P grows by about one per input byte on ordinary
            // data, so guess the chunk and gallop from there (a plain binary search over the whole
            // table costs ~20 cache misses per block, which dominates at level 1)
            uint64_t c_lo = e0 / CHB, c_hi = n_chunks + 1;       // P[c_lo] <= target (P[c_lo] <= Pe0)
            {
                uint64_t g = c_lo + (target - P[c_lo]) / CHB;
                if (g <= c_lo) g = c_lo + 1;
                if (g > n_chunks) g = n_chunks;
                if (g > c_lo && P[g] <= target) {
                    c_lo = g;
                    uint64_t step = 1;
                    while (c_lo + step <= n_chunks && P[c_lo + step] <= target) { c_lo += step; step *= 2; }
                    c_hi = std::min(c_lo + step, n_chunks + 1);
                } else if (g > c_lo) {
                    c_hi = g;
                    uint64_t step = 1;
                    while (c_hi > c_lo + step && P[c_hi - step] > target) { c_hi -= step; step *= 2; }
                    if (c_hi > c_lo + step) c_lo = c_hi - step;
                }
            }
            while (c_hi - c_lo > 1) {
                uint64_t mid = (c_lo + c_hi) / 2;
                if (P[mid] <= target) c_lo = mid; else c_hi = mid;
            }
            uint64_t cut, pc;
            if (c_lo >= n_chunks) {
                cut = N;
                pc = P[n_chunks];
            } else {
                uint64_t i, p;
                if (c_lo == e0 / CHB) { i = e0; p = Pe0; }
                else { i = c_lo * CHB; p = P[c_lo]; r = (uint32_t)(o_in[c_lo] % 255); }
                while (i < N) {
                    if (i >= 1 && i + 8 <= N && p + 8 <= target && no_adjacent_equal8(i)) {
                        p += 8;
                        r = 1;
                        i += 8;
                        continue;
                    }
                    if (i == 0 || in[i] != in[i - 1]) r = 0;
                    uint32_t nd = need_of(r);
                    if (p + nd > target) break;
                    p += nd;
                    r = (r == 254) ? 0 : r + 1;
                    i++;
                }
                cut = i;
                pc = p;
            }
            bk.c = cut;
            bk.e0 = e0;
            bk.u0 = (uint32_t)u0;
            bk.P_e0 = Pe0;
            bk.n = (uint32_t)(u0 + (pc - Pe0));
        }
        if (bk.c <= bk.s || bk.n == 0 || bk.n > M) return -1;
        if (!final && bk.c >= N) break;          // ran out of data, not out of capacity
        blocks.push_back(bk);
        rle_off += (bk.n + 127) & ~127ull;       // every block's image starts on its own 128-byte line
        s = bk.c;
    }
    if (consumed) *consumed = blocks.empty() ? 0 : blocks.back().c;
    return 0;
}

}  // namespace bnz
