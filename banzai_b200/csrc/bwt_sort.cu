// bwt_sort.cu — K3/K4: Burrows-Wheeler transform of many independent bzip2 blocks by a
// cyclic prefix-doubling rotation sort (hand-written LSD radix sort, sm_100a).
//
// Replaces the SA-IS pass of the reference (lib/bwt.rs:526-756) and reproduces its
// contract exactly: bwt[k] = byte preceding the k-th smallest rotation; rotations are
// compared cyclically, EQUAL rotations are ordered by descending start index, so
// origPtr = #{rot < rot0} + #{rot == rot0} - 1 (lib/bwt.rs:564-567, 733-749).
//
// Execution model (B200-first): one persistent CTA per bzip2 block slot.  A CTA claims a
// block from an atomic queue and runs the WHOLE doubling loop for it on its own: no
// inter-CTA synchronisation, no host round trip for the early exit, and blocks that need
// 3 rounds do not wait for blocks that need 20.  With >= 2 x 148 blocks in flight the
// chip is covered; per-CTA state (2 x 8 B records + 4 B rank per byte) streams through HBM.
//
// Per block of n bytes S:
//   records are 64-bit  [ key:40 | idx:20 ]   (n <= 900 000 < 2^20)
//   round 0   : key = S[i..i+5) (cyclic)                           -> sort -> rank_5
//   round h   : key = (rank_h[i] : 20, rank_h[(i+h) mod n] : 20)   -> sort -> rank_2h
//   only rotations whose rank is still shared ("active") are re-sorted; a rotation whose
//   key became unique gets its final rank, its BWT byte is written at once
//   (bwt[rank] = S[i-1]) and it drops out of later rounds.
//   A round that splits no group proves the remaining groups are identical rotations
//   (period | n); their positions inside the group are arbitrary for the BWT bytes and
//   origPtr = group base + group size - 1.
//
// Radix pass (per tile of TILE records, sequential over tiles inside the CTA so that the
// running bucket cursors live in shared memory): warp-striped coalesced load, per-warp
// stable ranking with match.any, cross-warp scan, shared-memory reorder, coalesced
// bucket-run stores.  The per-digit histograms of all passes are taken while the records
// are generated, so each pass costs one read + one write of the records.
#include "common.cuh"
#include "kernels.h"

namespace bnz {
namespace bwt {

constexpr int T = 512;               // threads per CTA
constexpr int NW = T / 32;           // warps per CTA
constexpr int K = 8;                 // records per thread per tile
constexpr int TILE = T * K;          // 4096 records = 32 KB
constexpr int KEY_BITS = 40;
constexpr int IDX_BITS = 20;
constexpr u32 IDX_MASK = (1u << IDX_BITS) - 1u;
constexpr u32 RANK_MASK = (1u << 20) - 1u;
constexpr u32 DONE = 0x80000000u;
constexpr int MAX_ROUNDS = 40;

template <int BITS>
struct Cfg {
    static constexpr int BINS = 1 << BITS;
    static constexpr int PASSES = (KEY_BITS + BITS - 1) / BITS;
    static constexpr int BPT = (BINS + T - 1) / T;      // bins per thread in the bin scans
};

template <int BITS>
struct __align__(16) Smem {
    u64 stage[TILE];                                      // tile reorder buffer
    u16 whist[NW][Cfg<BITS>::BINS];                       // per-warp digit counts / offsets
    u32 cursor[Cfg<BITS>::BINS];                          // running bucket cursors (global)
    u32 binoff[Cfg<BITS>::BINS];                          // exclusive bin offsets inside the tile
    u32 gbase[Cfg<BITS>::BINS];                           // cursor - binoff
    u32 hist[Cfg<BITS>::PASSES][Cfg<BITS>::BINS];         // per-pass digit histograms
    u32 scratch[40];
    u32 s_count;                                          // records appended by build_*
    u32 s_block;                                          // claimed block id
    u32 s_flags[8];
    u8 present[256];                                      // has_byte
};

template <int BITS>
__device__ __forceinline__ u32 digit_of(u64 rec, int pass)
{
    return (u32)(rec >> (IDX_BITS + pass * BITS)) & (u32)(Cfg<BITS>::BINS - 1);
}

template <int BITS>
__device__ __forceinline__ void hist_clear(Smem<BITS> &sm)
{
    for (int i = threadIdx.x; i < Cfg<BITS>::PASSES * Cfg<BITS>::BINS; i += T) (&sm.hist[0][0])[i] = 0;
    if (threadIdx.x == 0) sm.s_count = 0;
    __syncthreads();
}

template <int BITS>
__device__ __forceinline__ void hist_add(Smem<BITS> &sm, u64 rec)
{
#pragma unroll
    for (int p = 0; p < Cfg<BITS>::PASSES; p++) atomicAdd(&sm.hist[p][digit_of<BITS>(rec, p)], 1u);
}

// Round 0: key = the five bytes S[i..i+5) (cyclic), big-endian, so h = 5 afterwards.
template <int BITS>
__device__ void build_initial(Smem<BITS> &sm, const u8 *__restrict__ S, u32 n, u64 *dst)
{
    hist_clear(sm);
    for (int i = threadIdx.x; i < 256; i += T) sm.present[i] = 0;
    __syncthreads();
    for (u32 base = 0; base < n; base += TILE) {
#pragma unroll
        for (int k = 0; k < K; k++) {
            u32 i = base + k * T + threadIdx.x;
            if (i < n) {
                u64 key = 0;
                if (i + 5 <= n) {
#pragma unroll
                    for (int j = 0; j < 5; j++) key = (key << 8) | S[i + j];
                } else {
                    u32 q = i;
                    for (int j = 0; j < 5; j++) {
                        key = (key << 8) | S[q];
                        q = (q + 1 == n) ? 0 : q + 1;
                    }
                }
                sm.present[(u32)(key >> 32)] = 1;        // first byte = S[i]
                u64 rec = (key << IDX_BITS) | i;
                dst[i] = rec;
                hist_add(sm, rec);
            }
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) sm.s_count = n;
    __syncthreads();
}

// Round h: append a record for every still-active rotation, scanning rank[] in index order
// (coalesced reads of rank[i] and rank[i+h]).
template <int BITS>
__device__ void build_round(Smem<BITS> &sm, const u32 *rank, u32 n, u32 h,
                            u64 *dst)
{
    hist_clear(sm);
    const u32 hm = h % n;
    for (u32 base = 0; base < n; base += TILE) {
#pragma unroll
        for (int k = 0; k < K; k++) {
            u32 i = base + k * T + threadIdx.x;
            u32 r = (i < n) ? rank[i] : DONE;
            bool act = !(r & DONE);
            u64 rec = 0;
            if (act) {
                u32 j = i + hm;
                if (j >= n) j -= n;
                u32 r2 = rank[j] & RANK_MASK;
                rec = ((u64)r << (IDX_BITS + 20)) | ((u64)r2 << IDX_BITS) | i;
            }
            u32 m = __ballot_sync(0xffffffffu, act);
            if (m) {
                u32 wbase = 0;
                if (lane_id() == 0) wbase = atomicAdd(&sm.s_count, (u32)__popc(m));
                wbase = __shfl_sync(0xffffffffu, wbase, 0);
                if (act) {
                    dst[wbase + __popc(m & lanemask_lt())] = rec;
                    hist_add(sm, rec);
                }
            }
        }
    }
    __syncthreads();
}

// One LSD pass over `count` records: src -> dst by digit `pass`.
template <int BITS>
__device__ void radix_pass(Smem<BITS> &sm, const u64 *src, u64 *dst,
                           u32 count, int pass)
{
    constexpr int BINS = Cfg<BITS>::BINS;
    constexpr int BPT = Cfg<BITS>::BPT;
    const u32 tid = threadIdx.x, lane = lane_id(), w = warp_id();

    // cursor = exclusive scan of this pass's histogram
    {
        u32 c[BPT], s = 0;
#pragma unroll
        for (int j = 0; j < BPT; j++) {
            u32 b = tid * BPT + j;
            c[j] = (b < BINS) ? sm.hist[pass][b] : 0;
            s += c[j];
        }
        u32 tot;
        u32 ex = block_excl_sum<T>(s, sm.scratch, &tot);
#pragma unroll
        for (int j = 0; j < BPT; j++) {
            u32 b = tid * BPT + j;
            if (b < BINS) sm.cursor[b] = ex;
            ex += c[j];
        }
    }
    __syncthreads();

    for (u32 base = 0; base < count; base += TILE) {
        const u32 tile_n = min((u32)TILE, count - base);
        u64 rec[K];
        u32 rk[K];
        const u32 wbase = base + w * (K * 32) + lane;
#pragma unroll
        for (int k = 0; k < K; k++) {
            u32 j = wbase + k * 32;
            rec[k] = (j < count) ? src[j] : ~0ull;
        }
        for (int b = lane; b < BINS; b += 32) sm.whist[w][b] = 0;
        __syncwarp();
#pragma unroll
        for (int k = 0; k < K; k++) {
            u32 d = digit_of<BITS>(rec[k], pass);
            u32 peers = __match_any_sync(0xffffffffu, d);
            u32 leader = 31 - __clz(peers);
            u32 bcount = 0;
            if (lane == leader) {
                bcount = sm.whist[w][d];
                sm.whist[w][d] = (u16)(bcount + __popc(peers));
            }
            bcount = __shfl_sync(0xffffffffu, bcount, leader);
            rk[k] = bcount + __popc(peers & lanemask_lt());
            __syncwarp();
        }
        __syncthreads();

        // cross-warp exclusive scan per bin, then exclusive scan over bins
        {
            u32 c[BPT], s = 0;
#pragma unroll
            for (int j = 0; j < BPT; j++) {
                u32 b = tid * BPT + j;
                u32 run = 0;
                if (b < BINS) {
#pragma unroll
                    for (int ww = 0; ww < NW; ww++) {
                        u32 v = sm.whist[ww][b];
                        sm.whist[ww][b] = (u16)run;
                        run += v;
                    }
                }
                c[j] = run;
                s += run;
            }
            u32 tot;
            u32 ex = block_excl_sum<T>(s, sm.scratch, &tot);
#pragma unroll
            for (int j = 0; j < BPT; j++) {
                u32 b = tid * BPT + j;
                if (b < BINS) {
                    u32 cur = sm.cursor[b];
                    sm.binoff[b] = ex;
                    sm.gbase[b] = cur - ex;
                    sm.cursor[b] = cur + c[j];
                }
                ex += c[j];
            }
        }
        __syncthreads();

#pragma unroll
        for (int k = 0; k < K; k++) {
            u32 d = digit_of<BITS>(rec[k], pass);
            u32 pos = sm.binoff[d] + sm.whist[w][d] + rk[k];
            sm.stage[pos] = rec[k];
        }
        __syncthreads();

#pragma unroll
        for (int k = 0; k < K; k++) {
            u32 j = k * T + tid;
            if (j < tile_n) {
                u64 r = sm.stage[j];
                u32 d = digit_of<BITS>(r, pass);
                dst[sm.gbase[d] + j] = r;
            }
        }
        __syncthreads();
    }
}

struct RerankOut {
    u32 active;      // records that still share their key after this round
    u32 splits;      // key heads that are not group heads (new groups created)
};

// Walk the sorted records, assign new ranks, retire singletons (writing their BWT byte).
// `initial`: all records belong to one group with base rank 0 (round 0).
template <int BITS>
__device__ RerankOut rerank(Smem<BITS> &sm, const u64 *src, u32 count, bool initial,
                            const u8 *__restrict__ S, u32 n, u32 *rank,
                            u8 *__restrict__ bwt_out, u32 *ptr_out)
{
    const u32 tid = threadIdx.x;
    u32 carry_grp = 0, carry_key = 0;       // 1-based positions of the latest heads so far
    u32 n_active = 0, n_split = 0;
    const u64 grp_mask = initial ? 0ull : ((u64)RANK_MASK << 20);   // bits of r1 inside key40

    for (u32 base = 0; base < count; base += TILE) {
        const u32 j0 = base + tid * K;
        u64 key[K + 2];                     // key[0] = left neighbour, key[K+1] = right neighbour
        u32 idx[K];
        key[0] = (j0 > 0 && j0 - 1 < count) ? (src[j0 - 1] >> IDX_BITS) : ~0ull;
#pragma unroll
        for (int k = 0; k < K; k++) {
            u32 j = j0 + k;
            u64 r = (j < count) ? src[j] : ~0ull;
            key[k + 1] = r >> IDX_BITS;
            idx[k] = (u32)r & IDX_MASK;
        }
        key[K + 1] = (j0 + K < count) ? (src[j0 + K] >> IDX_BITS) : ~0ull;

        u32 pk[K], pg[K];
        u32 mk = 0, mg = 0;
#pragma unroll
        for (int k = 0; k < K; k++) {
            u32 j = j0 + k;
            bool valid = j < count;
            bool hk = valid && (j == 0 || key[k + 1] != key[k]);
            bool hg = valid && (j == 0 || ((key[k + 1] ^ key[k]) & grp_mask) != 0);
            if (hk) mk = j + 1;
            if (hg) mg = j + 1;
            pk[k] = mk;
            pg[k] = mg;
        }
        u32 tot_k, tot_g;
        u32 ex_k = block_excl_max<T>(mk, sm.scratch, &tot_k);
        u32 ex_g = block_excl_max<T>(mg, sm.scratch, &tot_g);
        ex_k = max(ex_k, carry_key);
        ex_g = max(ex_g, carry_grp);

#pragma unroll
        for (int k = 0; k < K; k++) {
            u32 j = j0 + k;
            if (j < count) {
                u32 p_key = max(pk[k], ex_k);        // 1-based
                u32 p_grp = max(pg[k], ex_g);
                bool hk = (pk[k] == j + 1);
                bool hg = (pg[k] == j + 1);
                u32 r1 = initial ? 0u : (u32)(key[k + 1] >> 20) & RANK_MASK;
                u32 nr = r1 + (p_key - p_grp);
                bool single = hk && (j + 1 == count || key[k + 2] != key[k + 1]);
                u32 id = idx[k];
                if (single) {
                    rank[id] = nr | DONE;
                    bwt_out[nr] = S[id == 0 ? n - 1 : id - 1];
                    if (id == 0) *ptr_out = nr;
                } else {
                    rank[id] = nr;
                    n_active++;
                }
                if (hk && !hg) n_split++;
            }
        }
        carry_key = max(carry_key, tot_k);
        carry_grp = max(carry_grp, tot_g);
        __syncthreads();
    }
    RerankOut o;
    o.active = block_sum<T>(n_active, sm.scratch);
    o.splits = block_sum<T>(n_split, sm.scratch);
    return o;
}

// Remaining groups are sets of identical rotations: give them distinct positions inside
// their group (any order yields the same BWT bytes) and derive origPtr by the
// descending-index rule.
template <int BITS>
__device__ void finalize_ties(Smem<BITS> &sm, const u64 *src, u32 count,
                              const u8 *__restrict__ S, u32 n, const u32 *rank,
                              u8 *__restrict__ bwt_out, u32 *ptr_out)
{
    const u32 tid = threadIdx.x;
    u32 carry_grp = 0;
    const u32 base0 = rank[0];
    const bool zero_tied = !(base0 & DONE);
    u32 size0 = 0;
    for (u32 base = 0; base < count; base += TILE) {
        const u32 j0 = base + tid * K;
        u32 r1[K + 1], idx[K];
        r1[0] = (j0 > 0 && j0 - 1 < count) ? (u32)(src[j0 - 1] >> (IDX_BITS + 20)) : 0xffffffffu;
        u32 pg[K], mg = 0;
#pragma unroll
        for (int k = 0; k < K; k++) {
            u32 j = j0 + k;
            u64 r = (j < count) ? src[j] : ~0ull;
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
        __syncthreads();
    }
    u32 s0 = block_sum<T>(size0, sm.scratch);
    if (tid == 0 && zero_tied) *ptr_out = base0 + s0 - 1;
}

template <int BITS>
__global__ void __launch_bounds__(T, 2) bwt_sort_kernel(BwtArgs a)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    Smem<BITS> &sm = *reinterpret_cast<Smem<BITS> *>(smem_raw);
    constexpr int PASSES = Cfg<BITS>::PASSES;
    const u32 tid = threadIdx.x;

    u64 *bufA = a.ws_rec + (size_t)blockIdx.x * 2 * a.ws_stride;
    u64 *bufB = bufA + a.ws_stride;
    u32 *rank = a.ws_rank + (size_t)blockIdx.x * a.ws_stride;

    for (;;) {
        if (tid == 0) sm.s_block = atomicAdd(a.next_block, 1u);
        __syncthreads();
        const u32 blk = sm.s_block;
        __syncthreads();
        if (blk >= a.n_blocks) break;

        const u8 *S = a.rle + a.blk_off[blk];
        u8 *bwt_out = a.bwt + a.blk_off[blk];
        const u32 n = a.blk_len[blk];
        u32 *ptr_out = a.ptr + blk;

        u32 rounds = 0;
        u64 sum_active = 0, sum_active_passes = 0;
        u32 h = 5;
        u32 count = n;
        bool initial = true;
        bool tied = false;

        while (count > 0 && rounds < MAX_ROUNDS) {
            if (initial) build_initial<BITS>(sm, S, n, bufA);
            else build_round<BITS>(sm, rank, n, h, bufA);
            count = sm.s_count;

            u64 *src = bufA, *dst = bufB;
            u32 passes_run = 0;
            for (int p = 0; p < PASSES; p++) {
                // skip a pass whose digit is the same for every record
                if (tid == 0) sm.s_flags[0] = 0;
                __syncthreads();
                for (int b = tid; b < Cfg<BITS>::BINS; b += T)
                    if (sm.hist[p][b] == count) sm.s_flags[0] = 1;
                __syncthreads();
                const bool skip = sm.s_flags[0] != 0;
                __syncthreads();
                if (skip) continue;
                radix_pass<BITS>(sm, src, dst, count, p);
                u64 *t = src; src = dst; dst = t;
                passes_run++;
            }
            sum_active += count;
            sum_active_passes += (u64)count * passes_run;
            rounds++;

            RerankOut ro = rerank<BITS>(sm, src, count, initial, S, n, rank, bwt_out, ptr_out);
            __syncthreads();
            if (ro.active > 0 && ro.splits == 0 && !initial) {
                finalize_ties<BITS>(sm, src, count, S, n, rank, bwt_out, ptr_out);
                tied = true;
                break;
            }
            if (!initial) h *= 2;
            initial = false;
            count = ro.active;
            // make this round's rank[] stores visible to every thread of the CTA
            __threadfence_block();
            __syncthreads();
        }

        if (tid < 256) a.has_byte[(size_t)blk * 256 + tid] = sm.present[tid];
        if (tid == 0 && a.stats) {
            BwtStats st;
            st.n = n;
            st.rounds = rounds;
            st.tied = tied ? 1u : 0u;
            st.pad = 0;
            st.sum_active = sum_active;
            st.sum_active_passes = sum_active_passes;
            a.stats[blk] = st;
        }
        __syncthreads();
    }
}

}  // namespace bwt

size_t bwt_smem_bytes(int bits)
{
    return bits == 8 ? sizeof(bwt::Smem<8>) : sizeof(bwt::Smem<10>);
}

int bwt_passes(int bits) { return bits == 8 ? bwt::Cfg<8>::PASSES : bwt::Cfg<10>::PASSES; }

cudaError_t bwt_max_ctas(int bits, int *ctas_per_sm)
{
    cudaError_t e;
    size_t smem = bwt_smem_bytes(bits);
    if (bits == 8) {
        e = cudaFuncSetAttribute(bwt::bwt_sort_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        return cudaOccupancyMaxActiveBlocksPerMultiprocessor(ctas_per_sm, bwt::bwt_sort_kernel<8>, bwt::T, smem);
    }
    e = cudaFuncSetAttribute(bwt::bwt_sort_kernel<10>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    return cudaOccupancyMaxActiveBlocksPerMultiprocessor(ctas_per_sm, bwt::bwt_sort_kernel<10>, bwt::T, smem);
}

cudaError_t bwt_launch(const BwtArgs &a, int bits, int grid, cudaStream_t stream)
{
    size_t smem = bwt_smem_bytes(bits);
    if (bits == 8) bwt::bwt_sort_kernel<8><<<grid, bwt::T, smem, stream>>>(a);
    else bwt::bwt_sort_kernel<10><<<grid, bwt::T, smem, stream>>>(a);
    return cudaGetLastError();
}

}  // namespace bnz
