// bwt_cluster.cu — K3/K4, cluster-cooperative variant: one thread-block CLUSTER (C CTAs on C
// SMs) per bzip2 block.  Same algorithm and record format as bwt_sort.cu (cyclic prefix
// doubling, 64-bit records [key:40 | idx:20], LSD radix passes, active-set shrinking, tie rule),
// but a block is spread over C SMs, so only ~148/C blocks are in flight at a time and their
// randomly accessed arrays (rank[] 4 B/byte, S and BWT 1 B/byte) stay resident in the 126 MB
// L2 instead of thrashing HBM with 32-byte sectors for 4-byte accesses.
//
// Cooperation (all through global memory + cluster barriers, no per-tile cluster traffic):
//   * every record array is cut into C contiguous chunks; CTA c streams chunk c tile by tile
//     (TMA bulk copies, running bucket cursors in shared memory) exactly like the 1-CTA kernel;
//   * a stable LSD pass needs, per digit, the number of records in EARLIER chunks: the per-chunk
//     digit histogram of pass p+1's input is accumulated while pass p scatters (the destination
//     position, hence the destination chunk, is known then), pre-aggregated in shared memory and
//     flushed into one of three rotating [C][256] tables in global memory;
//   * the key-build step writes CTA c's still-active records into its own segment, which is the
//     input chunk of pass 0 (so no cross-CTA compaction is needed);
//   * re-ranking needs the positions of the latest group/key heads before a chunk: each CTA
//     finds its chunk's last heads by scanning backwards from the chunk end (one tile in the
//     common case), the cluster exchanges them, then every CTA re-ranks its chunk;
//   * one cluster barrier separates consecutive phases (7 + 2 per round).
#include <cooperative_groups.h>

#include "bwt_common.cuh"
#include "kernels.h"

namespace cg = cooperative_groups;

namespace bnz {
namespace bwtc {

using namespace bwtk;                          // record layout, digit_of, match_digit (bwt_common.cuh)
constexpr int CMAX = BWT_CLUSTER_MAX;
constexpr int MAX_ROUNDS = 40;

struct Ctl {                                   // per cluster, global memory
    u32 chist[3][CMAX][BINS];                  // rotating per-chunk digit histograms
    u32 seg_cnt[2][CMAX];                      // records each CTA produced in the key build
    u32 heads[CMAX][2];                        // last key head / group head of each chunk (1-based)
    u32 totals[2][4];                          // [parity]{active, splits}
    u32 blk;                                   // claimed block id
    u32 pad[7];
};
static_assert(sizeof(Ctl) == BWT_CTL_BYTES, "Ctl size");

template <int T>
struct __align__(128) Smem {
    static constexpr int NW = T / 32;
    u64 inbuf[TILE];
    u64 stage[TILE];
    u32 whist[NW][BINS];                       // per-warp digit counts, then exclusive offsets
    u32 nhist[CMAX][BINS];                     // next pass's (destination chunk, digit) counts
    u32 cursor[BINS];
    u32 binoff[BINS];
    u32 gbase[BINS];
    u64 scratch64[40];
    u64 mbar;
    u32 scratch[40];
    u32 s_count;
    u32 s_flag;
    Period per;                                // periodic-run test of the claimed block (CTA 0)
    KeyCode kc;                                // how the round-0 key packs the block's alphabet (bwt_common.cuh)
    u8 present[256];
};

struct Chunking {
    u32 len;          // records per chunk (multiple of TILE)
    u64 magic;        // ceil(2^44 / len)
    __device__ __forceinline__ u32 chunk_of(u32 q) const { return (u32)(((u64)q * magic) >> 44); }
};
__device__ __forceinline__ Chunking make_chunking(u32 count, u32 C)
{
    Chunking ch;
    u32 tiles = (count + TILE - 1) / TILE;
    u32 per = (tiles + C - 1) / C;
    if (per == 0) per = 1;
    ch.len = per * TILE;
    ch.magic = ((1ull << 44) + ch.len - 1) / ch.len;
    return ch;
}

template <int T>
__device__ __forceinline__ void nhist_clear(Smem<T> &sm)
{
    for (int i = threadIdx.x; i < CMAX * BINS; i += T) (&sm.nhist[0][0])[i] = 0;
}

// add the shared-memory (chunk, digit) counts to table `dst` and clear them
template <int T>
__device__ __forceinline__ void nhist_flush(Smem<T> &sm, u32 (*dst)[BINS], u32 C)
{
    __syncthreads();
    for (u32 i = threadIdx.x; i < C * BINS; i += T) {
        u32 v = (&sm.nhist[0][0])[i];
        if (v) {
            atomicAdd(&dst[0][0] + i, v);
            (&sm.nhist[0][0])[i] = 0;
        }
    }
}

// ---------------------------------------------------------------------------------------
// key build: CTA c scans its index chunk and writes its records to seg (its own segment)
// ---------------------------------------------------------------------------------------
template <int T>
__device__ u32 build_initial(Smem<T> &sm, const u8 *__restrict__ S, u32 n, u32 lo, u32 hi, u64 *seg, u32 c)
{
    constexpr int K = TILE / T;
    for (u32 base = lo; base < hi; base += TILE) {
#pragma unroll
        for (int k = 0; k < K; k++) {
            u32 i = base + k * T + threadIdx.x;
            if (i < hi) {
                const u64 key = (sm.kc.k == 5) ? raw_key5(S, n, i) : packed_key(S, n, i, sm.kc);
                u64 rec = (key << IDX_BITS) | i;
                st_stream(seg + (i - lo), rec);
                atomicAdd(&sm.nhist[c][digit_of(rec, 0)], 1u);
            }
        }
    }
    return hi > lo ? hi - lo : 0;
}

template <int T>
__device__ u32 build_round(Smem<T> &sm, const u32 *rank, u32 n, u32 h, u32 lo, u32 hi, u64 *seg, u32 c)
{
    constexpr int K = TILE / T;
    if (threadIdx.x == 0) sm.s_count = 0;
    __syncthreads();
    const u32 hm = h % n;
    for (u32 base = lo; base < hi; base += TILE) {
        u32 r[K], r2[K];
#pragma unroll
        for (int k = 0; k < K; k++) {
            u32 i = base + k * T + threadIdx.x;
            r[k] = (i < hi) ? ld_keep_cg(rank + i) : DONE;
        }
#pragma unroll
        for (int k = 0; k < K; k++) {
            u32 i = base + k * T + threadIdx.x;
            r2[k] = 0;
            if (!(r[k] & DONE)) {
                u32 j = i + hm;
                if (j >= n) j -= n;
                r2[k] = ld_keep_cg(rank + j);
            }
        }
#pragma unroll
        for (int k = 0; k < K; k++) {
            u32 i = base + k * T + threadIdx.x;
            bool act = !(r[k] & DONE);
            u64 rec = ((u64)r[k] << (IDX_BITS + 20)) | ((u64)(r2[k] & RANK_MASK) << IDX_BITS) | i;
            u32 m = __ballot_sync(0xffffffffu, act);
            if (m) {
                u32 wbase = 0;
                if (lane_id() == 0) wbase = atomicAdd(&sm.s_count, (u32)__popc(m));
                wbase = __shfl_sync(0xffffffffu, wbase, 0);
                if (act) {
                    st_stream(seg + wbase + __popc(m & lanemask_lt()), rec);
                    atomicAdd(&sm.nhist[c][digit_of(rec, 0)], 1u);
                }
            }
        }
    }
    __syncthreads();
    return sm.s_count;
}

// ---------------------------------------------------------------------------------------
// one LSD pass: CTA c sorts its input chunk [in, in + in_cnt) into dst using cursors derived
// from table `tab` (per-chunk histograms of this pass); while scattering it accumulates the
// next pass's per-destination-chunk histogram into sm.nhist.
// ---------------------------------------------------------------------------------------
template <int T>
__device__ void radix_pass(Smem<T> &sm, const u64 *in, u32 in_cnt, u64 *dst, int pass, u32 (*tab)[BINS],
                           u32 C, u32 c, const Chunking &och, bool count_next, u32 &phase)
{
    constexpr int K = TILE / T;
    constexpr int NW = T / 32;
    const u32 tid = threadIdx.x, lane = lane_id(), w = warp_id();

    fence_proxy_async();
    __syncthreads();
    if (tid == 0 && in_cnt > 0) {
        const u32 bytes = (min((u32)TILE, in_cnt) * 8u + 15u) & ~15u;
        mbar_expect_tx(&sm.mbar, bytes);
        tma_load_1d_stream(sm.inbuf, in, bytes, &sm.mbar);
    }

    // cursors: bucket start (all chunks) + records of the same digit in earlier chunks
    if (tid < BINS) {
        u32 col = 0, pre = 0;
        for (u32 cc = 0; cc < C; cc++) {
            u32 v = __ldcg(&tab[cc][tid]);
            if (cc < c) pre += v;
            col += v;
        }
        const u32 inc = warp_incl_sum(col);
        if (lane == 31) sm.scratch[w] = inc;
        asm volatile("bar.sync 1, %0;" ::"n"(BINS) : "memory");
        u32 woff = 0;
        for (u32 q = 0; q < w; q++) woff += sm.scratch[q];
        sm.cursor[tid] = woff + inc - col + pre;
    }
    __syncthreads();

    for (u32 base = 0; base < in_cnt; base += TILE) {
        const u32 tile_n = min((u32)TILE, in_cnt - base);
        u64 rec[K];
        u32 rk[K];
        for (int b = lane; b < BINS; b += 32) sm.whist[w][b] = 0;
        mbar_wait(&sm.mbar, phase);
        phase ^= 1u;
        const u32 wl = w * (K * 32) + lane;
#pragma unroll
        for (int k = 0; k < K; k++) {
            u32 j = wl + k * 32;
            rec[k] = (j < tile_n) ? sm.inbuf[j] : ~0ull;
        }
        __syncwarp();
        // stable in-warp ranking: peers by a ballot loop, then the leader lane of every digit
        // group bumps the warp counter with ONE shared atomic whose return value is the group's
        // base.  Atomics of successive rows to the same counter execute in program order, so
        // the K rows need no warp barrier between them and their latencies overlap.
#pragma unroll
        for (int k = 0; k < K; k++) {
            const u32 d = digit_of(rec[k], pass);
            const u32 peers = match_digit(d);
            const u32 leader = 31 - __clz(peers);
            u32 bcount = 0;
            if (lane == leader) bcount = atomicAdd(&sm.whist[w][d], (u32)__popc(peers));
            bcount = __shfl_sync(0xffffffffu, bcount, leader);
            rk[k] = bcount + __popc(peers & lanemask_lt());
        }
        __syncthreads();                                        // B1

        if (tid == 0 && base + TILE < in_cnt) {
            const u32 nb = (min((u32)TILE, in_cnt - base - TILE) * 8u + 15u) & ~15u;
            // (no proxy fence: see bwt_sort.cu — the reads of inbuf completed before B1)
            mbar_expect_tx(&sm.mbar, nb);
            tma_load_1d_stream(sm.inbuf, in + base + TILE, nb, &sm.mbar);
        }

        if (tid < BINS) {
            u32 run = 0;
#pragma unroll
            for (int ww = 0; ww < NW; ww++) {
                u32 v = sm.whist[ww][tid];
                sm.whist[ww][tid] = run;
                run += v;
            }
            const u32 inc = warp_incl_sum(run);
            if (lane == 31) sm.scratch[w] = inc;
            asm volatile("bar.sync 1, %0;" ::"n"(BINS) : "memory");
            u32 woff = 0;
            for (u32 q = 0; q < w; q++) woff += sm.scratch[q];
            const u32 ex = woff + inc - run;
            const u32 cur = sm.cursor[tid];
            sm.binoff[tid] = ex;
            sm.gbase[tid] = cur - ex;
            sm.cursor[tid] = cur + run;
        }
        __syncthreads();                                        // B2

#pragma unroll
        for (int k = 0; k < K; k++) {
            const u32 d = digit_of(rec[k], pass);
            const u32 pos = sm.binoff[d] + sm.whist[w][d] + rk[k];
            sm.stage[pos] = rec[k];
        }
        __syncthreads();                                        // B3

#pragma unroll
        for (int k = 0; k < K; k++) {
            const u32 j = k * T + tid;
            if (j < tile_n) {
                const u64 r = sm.stage[j];
                const u32 d = digit_of(r, pass);
                const u32 q = sm.gbase[d] + j;
                st_stream(dst + q, r);
                if (count_next) atomicAdd(&sm.nhist[och.chunk_of(q)][digit_of(r, pass + 1)], 1u);
            }
        }
    }
    __syncthreads();
}

// ---------------------------------------------------------------------------------------
// re-rank
// ---------------------------------------------------------------------------------------

// phase A: 1-based positions of the last key head and the last group head inside [lo, hi)
template <int T>
__device__ void last_heads(Smem<T> &sm, const u64 *src, u32 lo, u32 hi, bool initial, u32 &lk_out, u32 &lg_out)
{
    constexpr int K = TILE / T;
    const u64 grp_mask = initial ? 0ull : ((u64)RANK_MASK << 20);
    u32 lk = 0, lg = 0;
    if (hi > lo) {
        u32 te = hi;
        for (;;) {
            const u32 tb = (te - lo > TILE) ? lo + ((te - lo - 1) / TILE) * TILE : lo;
            const u32 j0 = tb + threadIdx.x * K;
            u64 prev = (j0 > 0 && j0 - 1 < te) ? (__ldcg(src + j0 - 1) >> IDX_BITS) : ~0ull;
            u32 mk = 0, mg = 0;
#pragma unroll
            for (int k = 0; k < K; k++) {
                const u32 j = j0 + k;
                if (j < te) {
                    const u64 key = __ldcg(src + j) >> IDX_BITS;
                    if (j == 0 || key != prev) mk = j + 1;
                    if (j == 0 || ((key ^ prev) & grp_mask) != 0) mg = j + 1;
                    prev = key;
                }
            }
            u64 tot2;
            block_excl_max2<T>(((u64)mg << 32) | mk, sm.scratch64, &tot2);
            lk = max(lk, (u32)tot2);
            lg = max(lg, (u32)(tot2 >> 32));
            if (lg != 0 || tb == lo) break;
            te = tb;
        }
    }
    lk_out = lk;
    lg_out = lg;
}

struct RerankOut {
    u32 active, splits;
};

template <int T>
__device__ RerankOut rerank(Smem<T> &sm, const u64 *src, u32 lo, u32 hi, u32 count, bool initial,
                            u32 carry_key, u32 carry_grp, const u8 *__restrict__ S, u32 n, u32 *rank,
                            u8 *__restrict__ bwt_out, u32 *ptr_out)
{
    constexpr int K = TILE / T;
    const u32 tid = threadIdx.x;
    u32 n_active = 0, n_split = 0;
    const u64 grp_mask = initial ? 0ull : ((u64)RANK_MASK << 20);

    for (u32 base = lo; base < hi; base += TILE) {
        const u32 j0 = base + tid * K;
        u64 key[K + 2];
        u32 idx[K];
        key[0] = (j0 > 0 && j0 - 1 < hi) ? (__ldcg(src + j0 - 1) >> IDX_BITS) : ~0ull;
#pragma unroll
        for (int k = 0; k < K; k++) {
            u32 j = j0 + k;
            u64 r = (j < hi) ? __ldcg(src + j) : ~0ull;
            key[k + 1] = r >> IDX_BITS;
            idx[k] = (u32)r & IDX_MASK;
        }
        key[K + 1] = (j0 + K < count) ? (__ldcg(src + j0 + K) >> IDX_BITS) : ~0ull;

        u32 pk[K], pg[K];
        u32 mk = 0, mg = 0;
#pragma unroll
        for (int k = 0; k < K; k++) {
            u32 j = j0 + k;
            bool valid = j < hi;
            bool hk = valid && (j == 0 || key[k + 1] != key[k]);
            bool hg = valid && (j == 0 || ((key[k + 1] ^ key[k]) & grp_mask) != 0);
            if (hk) mk = j + 1;
            if (hg) mg = j + 1;
            pk[k] = mk;
            pg[k] = mg;
        }
        u64 tot2;
        const u64 ex2 = block_excl_max2<T>(((u64)mg << 32) | mk, sm.scratch64, &tot2);
        const u32 ex_k = max((u32)ex2, carry_key);
        const u32 ex_g = max((u32)(ex2 >> 32), carry_grp);

        u32 nrv[K], flg[K], sb[K];
#pragma unroll
        for (int k = 0; k < K; k++) {
            u32 j = j0 + k;
            flg[k] = 0;
            nrv[k] = 0;
            if (j < hi) {
                u32 p_key = max(pk[k], ex_k);
                u32 p_grp = max(pg[k], ex_g);
                bool hk = (pk[k] == j + 1);
                bool hg = (pg[k] == j + 1);
                u32 r1 = initial ? 0u : (u32)(key[k + 1] >> 20) & RANK_MASK;
                nrv[k] = r1 + (p_key - p_grp);
                bool single = hk && (j + 1 == count || key[k + 2] != key[k + 1]);
                flg[k] = 1u | (single ? 2u : 0u);
                // a record that stays in the first subgroup of its old group keeps its rank: its
                // rank[] entry is already correct, skip the (random, 32-byte-sector) store
                if (!single && !initial && nrv[k] == r1) flg[k] |= 4u;
                if (!single) n_active++;
                if (hk && !hg) n_split++;
            }
        }
#pragma unroll
        for (int k = 0; k < K; k++) {
            sb[k] = 0;
            if (flg[k] & 2u) sb[k] = S[idx[k] == 0 ? n - 1 : idx[k] - 1];
        }
#pragma unroll
        for (int k = 0; k < K; k++) {
            if (flg[k] & 1u) {
                const u32 id = idx[k];
                if (flg[k] & 2u) {
                    st_keep(rank + id, nrv[k] | DONE);
                    bwt_out[nrv[k]] = (u8)sb[k];
                    if (id == 0) *ptr_out = nrv[k];
                } else if (!(flg[k] & 4u)) {
                    st_keep(rank + id, nrv[k]);
                }
            }
        }
        carry_key = max(carry_key, (u32)tot2);
        carry_grp = max(carry_grp, (u32)(tot2 >> 32));
    }
    RerankOut o;
    o.active = block_sum<T>(n_active, sm.scratch);
    o.splits = block_sum<T>(n_split, sm.scratch);
    return o;
}

// identical rotations left: one CTA walks the whole sorted array (rare path)
template <int T>
__device__ void finalize_ties(Smem<T> &sm, const u64 *src, u32 count, const u8 *__restrict__ S, u32 n,
                              const u32 *rank, u8 *__restrict__ bwt_out, u32 *ptr_out)
{
    constexpr int K = TILE / T;
    const u32 tid = threadIdx.x;
    u32 carry_grp = 0;
    const u32 base0 = __ldcg(rank);
    const bool zero_tied = !(base0 & DONE);
    u32 size0 = 0;
    for (u32 base = 0; base < count; base += TILE) {
        const u32 j0 = base + tid * K;
        u32 r1[K + 1], idx[K];
        r1[0] = (j0 > 0 && j0 - 1 < count) ? (u32)(__ldcg(src + j0 - 1) >> (IDX_BITS + 20)) : 0xffffffffu;
        u32 pg[K], mg = 0;
#pragma unroll
        for (int k = 0; k < K; k++) {
            u32 j = j0 + k;
            u64 r = (j < count) ? __ldcg(src + j) : ~0ull;
            r1[k + 1] = (u32)(r >> (IDX_BITS + 20));
            idx[k] = (u32)r & IDX_MASK;
            if (j < count && (j == 0 || r1[k + 1] != r1[k])) mg = j + 1;
            pg[k] = mg;
        }
        u32 tot_g;
        u32 ex_g = block_excl_max<T>(mg, sm.scratch, &tot_g);
        ex_g = max(ex_g, carry_grp);
#pragma unroll
        for (int k = 0; k < K; k++) {
            u32 j = j0 + k;
            if (j < count) {
                u32 p_grp = max(pg[k], ex_g);
                u32 pos = r1[k + 1] + (j + 1 - p_grp);
                u32 id = idx[k];
                bwt_out[pos] = S[id == 0 ? n - 1 : id - 1];
                if (zero_tied && r1[k + 1] == base0) size0++;
            }
        }
        carry_grp = max(carry_grp, tot_g);
    }
    u32 s0 = block_sum<T>(size0, sm.scratch);
    if (tid == 0 && zero_tied) *ptr_out = base0 + s0 - 1;
}

// ---------------------------------------------------------------------------------------
// kernel
// ---------------------------------------------------------------------------------------
template <int T>
__global__ void __launch_bounds__(T, (T == 512 ? 2 : 1)) bwt_cluster_kernel(BwtArgs a)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    Smem<T> &sm = *reinterpret_cast<Smem<T> *>(smem_raw);
    cg::cluster_group cluster = cg::this_cluster();
    const u32 C = cluster.num_blocks();
    const u32 c = cluster.block_rank();
    const u32 cl = blockIdx.x / C;
    const u32 tid = threadIdx.x;

    if (tid == 0) {
        mbar_init(&sm.mbar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    nhist_clear(sm);
    __syncthreads();
    u32 phase = 0;

    u64 *bufA = a.ws_rec + (size_t)cl * 2 * a.ws_stride;
    u64 *bufB = bufA + a.ws_stride;
    u32 *rank = a.ws_rank + (size_t)cl * a.ws_stride;
    Ctl *ctl = reinterpret_cast<Ctl *>(a.ws_ctl) + cl;

    for (;;) {
        // claim a block; reset the cluster's control state
        if (c == 0 && tid == 0) ctl->blk = atomicAdd(a.next_block, 1u);
        for (u32 i = tid; i < 3 * BINS; i += T) ctl->chist[i / BINS][c][i % BINS] = 0;
        if (c == 0 && tid < 8) (&ctl->totals[0][0])[tid] = 0;
        for (int i = tid; i < 256; i += T) sm.present[i] = 0;
        cluster.sync();
        const u32 qpos = __ldcg(&ctl->blk);
        if (qpos >= a.n_blocks) break;
        const u32 blk = qpos;

        const u8 *S = a.rle + a.blk_off[blk];
        u8 *bwt_out = a.bwt + a.blk_off[blk];
        const u32 n = a.blk_len[blk];
        u32 *ptr_out = a.ptr + blk;

        if (a.defer_list) {
            // A block with a long periodic run needs ~log2(n) full rounds here; the one-CTA kernel knows the
            // order of such rotations in closed form (bwt_common.cuh: Period) and finishes it in two rounds.
            // Leave it to the follow-up launch of that kernel (stages.cu).
            if (c == 0) {
                detect_period<T>(S, n, sm.scratch, &sm.per);
                if (tid == 0) {
                    ctl->pad[0] = sm.per.p;
                    if (sm.per.p != 0) a.defer_list[atomicAdd(a.defer_count, 1u)] = blk;
                }
            }
            cluster.sync();
            if (__ldcg(&ctl->pad[0]) != 0) continue;
        }

        const Chunking ich = make_chunking(n, C);           // index chunks (key build)
        const u32 ilo = min(n, c * ich.len), ihi = min(n, (c + 1) * ich.len);
        u64 *seg = bufA + (size_t)c * ich.len;              // CTA c's segment of the build output

        // every CTA derives the block's alphabet itself (S comes from L2; no exchange needed)
        build_alphabet<T>(S, n, sm.present, &sm.kc, sm.scratch);

        u32 rounds = 0, tseq = 0, rpar = 0;
        u64 sum_active = 0, sum_active_passes = 0;
        long long cyc_build = 0, cyc_radix = 0, cyc_rerank = 0, t0, t1;
        u32 h = sm.kc.k;
        bool initial = true, tied = false;

        while (rounds < MAX_ROUNDS) {
            // ---- key build into my segment; its digit-0 histogram is row c of the next table
            t0 = clock64();
            u32 my_cnt = initial ? build_initial<T>(sm, S, n, ilo, ihi, seg, c)
                                 : build_round<T>(sm, rank, n, h, ilo, ihi, seg, c);
            nhist_flush(sm, ctl->chist[(tseq + 1) % 3], C);
            for (u32 i = tid; i < BINS; i += T) ctl->chist[(tseq + 2) % 3][c][i] = 0;
            if (tid == 0) ctl->seg_cnt[rpar][c] = my_cnt;
            if (c == 0 && tid < 4) ctl->totals[rpar ^ 1][tid] = 0;
            tseq++;
            fence_proxy_async();
            cluster.sync();
            u32 count = 0;
            for (u32 cc = 0; cc < C; cc++) count += __ldcg(&ctl->seg_cnt[rpar][cc]);
            t1 = clock64();
            cyc_build += t1 - t0;
            if (count == 0) break;
            const Chunking och = make_chunking(count, C);   // chunks of the compact arrays
            const u32 lo = min(count, c * och.len), hi = min(count, (c + 1) * och.len);

            // ---- LSD passes: segments(A) -> B -> A -> B -> A -> B
            u64 *src = bufA, *dst = bufB;
            for (int p = 0; p < PASSES; p++) {
                const u64 *in = (p == 0) ? seg : src + lo;
                const u32 in_cnt = (p == 0) ? my_cnt : hi - lo;
                radix_pass<T>(sm, in, in_cnt, dst, p, ctl->chist[tseq % 3], C, c, och, p + 1 < PASSES, phase);
                nhist_flush(sm, ctl->chist[(tseq + 1) % 3], C);
                for (u32 i = tid; i < BINS; i += T) ctl->chist[(tseq + 2) % 3][c][i] = 0;
                tseq++;
                fence_proxy_async();
                cluster.sync();
                u64 *t = src; src = dst; dst = t;
            }
            sum_active += count;
            sum_active_passes += (u64)count * PASSES;
            rounds++;
            t0 = clock64();
            cyc_radix += t0 - t1;

            // ---- re-rank: exchange chunk-boundary heads, then every CTA handles its chunk
            u32 lk, lg;
            last_heads<T>(sm, src, lo, hi, initial, lk, lg);
            if (tid == 0) {
                ctl->heads[c][0] = lk;
                ctl->heads[c][1] = lg;
            }
            cluster.sync();
            u32 carry_key = 0, carry_grp = 0;
            for (u32 cc = 0; cc < c; cc++) {
                carry_key = max(carry_key, __ldcg(&ctl->heads[cc][0]));
                carry_grp = max(carry_grp, __ldcg(&ctl->heads[cc][1]));
            }
            RerankOut ro = rerank<T>(sm, src, lo, hi, count, initial, carry_key, carry_grp, S, n, rank, bwt_out, ptr_out);
            if (tid == 0) {
                if (ro.active) atomicAdd(&ctl->totals[rpar][0], ro.active);
                if (ro.splits) atomicAdd(&ctl->totals[rpar][1], ro.splits);
            }
            cluster.sync();
            const u32 tot_active = __ldcg(&ctl->totals[rpar][0]);
            const u32 tot_splits = __ldcg(&ctl->totals[rpar][1]);
            cyc_rerank += clock64() - t0;
            if (tot_active > 0 && tot_splits == 0 && !initial) {
                if (c == 0) finalize_ties<T>(sm, src, count, S, n, rank, bwt_out, ptr_out);
                tied = true;
                break;
            }
            if (!initial) h *= 2;
            initial = false;
            rpar ^= 1;
            if (tot_active == 0) break;
        }

        if (a.marks) {
            // rows of the rotations 0, 4096, ... for the self-verification (every CTA's rank stores are
            // visible after the cluster barrier that ended the last round)
            cluster.sync();
            if (c == 0)
                for (u32 i = tid * VERIFY_SPACING; i < n; i += T * VERIFY_SPACING)
                    a.marks[(size_t)blk * VERIFY_MARKS + i / VERIFY_SPACING] = ld_keep_cg(rank + i) & RANK_MASK;
        }
        for (int i = tid; i < 256; i += T)
            if (sm.present[i]) a.has_byte[(size_t)blk * 256 + i] = 1;
        if (c == 0 && tid == 0 && a.stats) {
            BwtStats st;
            st.n = n;
            st.rounds = rounds;
            st.tied = tied ? 1u : 0u;
            st.period = 0;
            st.sum_active = sum_active;
            st.sum_active_passes = sum_active_passes;
            st.sum_tile = 0;
            st.cyc_tile = 0;
            st.cyc_final = 0;
            st.cyc_build = (u64)cyc_build;
            st.cyc_radix = (u64)cyc_radix;
            st.cyc_rerank = (u64)cyc_rerank;
            a.stats[blk] = st;
        }
        if (a.done) {
            // every CTA's stores of this block precede the cluster barrier; publish them, then raise the
            // block's flag (host-mapped memory: the host queues the MTF of finished blocks beside the sort)
            cluster.sync();
            if (c == 0 && tid == 0) {
                __threadfence_system();
                asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(a.done + blk), "r"(1u) : "memory");
            }
        }
        __syncthreads();
    }
}

}  // namespace bwtc

size_t bwtc_smem_bytes(int threads) { return threads == 1024 ? sizeof(bwtc::Smem<1024>) : sizeof(bwtc::Smem<512>); }

// max co-resident clusters of size C for the given CTA size (0 if the shape cannot launch)
cudaError_t bwtc_max_clusters(int threads, int C, int *n_clusters)
{
    *n_clusters = 0;
    const void *fn = threads == 1024 ? (const void *)bwtc::bwt_cluster_kernel<1024> : (const void *)bwtc::bwt_cluster_kernel<512>;
    size_t smem = bwtc_smem_bytes(threads);
    cudaError_t e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    if (C > 8) {
        e = cudaFuncSetAttribute(fn, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
        if (e != cudaSuccess) return e;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(C, 1, 1);
    cfg.blockDim = dim3(threads, 1, 1);
    cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = C;
    at[0].val.clusterDim.y = 1;
    at[0].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    return cudaOccupancyMaxActiveClusters(n_clusters, fn, &cfg);
}

cudaError_t bwtc_launch(const BwtArgs &a, int threads, int C, int n_clusters, cudaStream_t stream)
{
    const void *fn = threads == 1024 ? (const void *)bwtc::bwt_cluster_kernel<1024> : (const void *)bwtc::bwt_cluster_kernel<512>;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(n_clusters * C), 1, 1);
    cfg.blockDim = dim3(threads, 1, 1);
    cfg.dynamicSmemBytes = bwtc_smem_bytes(threads);
    cfg.stream = stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = C;
    at[0].val.clusterDim.y = 1;
    at[0].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    BwtArgs args = a;
    void *params[] = { &args };
    return cudaLaunchKernelExC(&cfg, fn, params);
}

}  // namespace bnz
