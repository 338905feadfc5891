// bwt_sort.cu — K3/K4: Burrows-Wheeler transform of many independent bzip2 blocks by a
// cyclic prefix-doubling rotation sort with in-shared-memory group refinement (sm_100a).
//
// Replaces the SA-IS pass of the reference (lib/bwt.rs:526-756) and reproduces its
// contract exactly: bwt[k] = byte preceding the k-th smallest rotation; rotations are
// compared cyclically, EQUAL rotations are ordered by descending start index, so
// origPtr = #{rot < rot0} + #{rot == rot0} - 1 (lib/bwt.rs:564-567, 733-749).
//
// Execution model (B200-first): one persistent CTA per bzip2 block slot.  A CTA claims a
// block from an atomic queue and runs the WHOLE doubling loop for it on its own: no
// inter-CTA synchronisation, no host round trip for the early exit, and blocks that need
// 3 rounds do not wait for blocks that need 20.
//
// Per block of n bytes S (n <= 900 000 < 2^20):
//   round 0 : 64-bit records [ key:40 | idx:20 ], key = the first k symbols of the rotation: 5 raw
//             bytes, or 6-8 symbols of a smaller alphabet packed base-sigma plus a coarse next symbol
//             (bwt_common.cuh: KeyCode), h = k afterwards; five LSD radix
//             passes through HBM (TMA-streamed tiles, see radix_pass); the re-rank step gives
//             every rotation the position of its group's first record as rank[i] and leaves the
//             ACTIVE LIST: the records [ rank:20 | idx:20 ] of all rotations whose key is
//             shared, in sorted order, so that every group is a contiguous run of the list.
//   round h : (h = k, 2k, 4k, ...) a group only has to be sorted by rank[idx + h] WITHIN
//             itself.  The list is walked in tiles of whole groups (<= 4096 records): one
//             coalesced read of the list, one gather of rank[idx + h], a stable LSD radix sort of
//             (group number in tile : 12, rank[idx+h] : 20) entirely in shared memory, new ranks
//             from head-flag bitmaps, one scatter of the changed ranks, and the records that
//             still share their key go back to the list in place (order-preserving compaction).
//             A round therefore moves a record through HBM once instead of five times.  Groups
//             larger than a tile (long repeats, periodic data) take the global path: keys
//             [ rank : rank[idx+h] : idx ] are written out and sorted by the 20 bits of
//             rank[idx+h] with the radix passes of round 0 (the passes whose digit is the same in
//             every record are skipped: three instead of five).
//   Rank stores: round 0 writes all n ranks; scattered as 4-byte stores they cost a 32-byte DRAM
//   sector read + write each, and with 296 CTAs doing so at once the kernel was bound by exactly
//   that.  Its ranks therefore go through apply_ranks_bucketed (one more radix pass by idx >> 13,
//   then rank[] is written chunk by chunk with coalesced stores).  Later rounds store ranks at once:
//   all ranks of a tile's groups are replaced between two CTA barriers and nothing is gathered
//   meanwhile, so a reader sees a group either completely refined or not at all — both are
//   consistent with the final order, the refined one merely carries more information, which
//   saves record-rounds (measured: 3.03 against 3.17 per byte on the mixed corpus).
//   A round that splits no group proves the remaining groups are identical rotations
//   (period | n); their positions inside the group are arbitrary for the BWT bytes and
//   origPtr = group base + group size - 1.
//   Last, one pass in index order writes bwt[rank[i]] = S[i-1] (coalesced reads, a one-byte
//   scatter that covers the block's output at once and merges in L2).
// DESIGN.md §4 has the measurements behind these choices.
#include "bwt_common.cuh"
#include "kernels.h"

namespace bnz {
namespace bwt {

using namespace bwtk;                // record layout, digit_of, match_digit (bwt_common.cuh)
constexpr int T = 512;               // threads per CTA
constexpr int NW = T / 32;           // warps per CTA
constexpr int K = 8;                 // records per thread per tile
constexpr int WORDS = BINS / 2;      // packed counter words per warp row (two 16-bit bins per word)
constexpr int MAX_ROUNDS = 48;
constexpr int BMW = TILE / 32;       // words of a per-tile bitmap
constexpr int UPD_SHIFT = 13;        // round-0 rank updates are bucketed by idx >> 13 ...
constexpr u32 UPD_CHUNK = 1u << UPD_SHIFT;   // ... so that a bucket covers 8192 ranks = 32 KB of rank[]
static_assert(TILE == T * K && WORDS <= T && BMW == 128 && NW == 16, "scan layouts below assume 512 threads, 4096-record tiles");

struct __align__(128) Smem {
    u64 buf0[TILE];                  // TMA landing zone of the global passes | ping buffer of the in-tile sort
    u64 buf1[TILE];                  // reorder buffer of the global passes (digit histograms while keys are built) | pong
    u32 whist[NW][WORDS];            // per-warp digit counts / offsets, two 16-bit bins per word
    u32 cursor[BINS];                // running bucket cursors of a global pass
    u32 gbase[BINS];                 // cursor - binoff
    u16 binoff[BINS];                // exclusive bin offsets inside the tile
    u32 r1tab[TILE];                 // tile path: rank (= position of the first record) of each group of the tile
    u32 bm_k[BMW], bm_g[BMW], bm_a[BMW];      // per position: key head | group head | stays active
    u32 pre_k[BMW], pre_g[BMW], pre_a[BMW];   // per bitmap word: last head before it (1-based) | active records before it
    u64 scratch64[40];
    u64 mbar;                        // mbarrier of the TMA tile pipeline
    u32 scratch[40];
    u32 wcnt[NW], wlast[NW];         // tile path: group heads per warp, last head position per warp
    u32 s_count;                     // records appended by build_*
    u32 s_block;                     // claimed block id
    u32 s_tot;                       // tile path: records that stay active in this tile
    u32 s_la;                        // tile path: the record after the tile starts a new group
    u32 s_min;                       // group-end search
    u32 s_flags[8];                  // per pass: every record has the same digit
    u64 acc[8];                      // per block statistics, kept by thread 0 (see ACC_*)
    Period per;                      // the block's periodic run (per.p == 0: none), bwt_common.cuh
    u8 present[256];                 // has_byte
    KeyCode kc;                      // build_initial: how the round-0 key packs the block's alphabet
};
enum { ACC_ACTIVE = 0, ACC_PASSES, ACC_TILE, ACC_CYC_BUILD, ACC_CYC_RADIX, ACC_CYC_RERANK, ACC_CYC_TILE, ACC_CYC_FINAL };

// per-pass digit histograms: built in the (then idle) reorder buffer, parked in global memory
__device__ __forceinline__ u32 *hist_of(Smem &sm) { return reinterpret_cast<u32 *>(sm.buf1); }

// update record of the deferred rank scatter: [ done:1 rank:20 | idx >> 13 : 8 | 0:7 | idx & 8191 : 13 ]
__device__ __forceinline__ u64 upd_record(u32 nr, bool done, u32 id)
{
    const u64 val = (u64)nr | (done ? (1ull << 20) : 0ull);
    return (val << 28) | ((u64)(id >> UPD_SHIFT) << IDX_BITS) | (id & (UPD_CHUNK - 1));
}
__device__ __forceinline__ u32 upd_rank_word(u64 e)
{
    const u32 val = (u32)(e >> 28);
    return (val & RANK_MASK) | ((val >> 20) ? DONE : 0u);
}

__device__ __forceinline__ u32 lanemask_le()
{
    u32 m;
    asm("mov.u32 %0, %%lanemask_le;" : "=r"(m));
    return m;
}

__device__ __forceinline__ void hist_clear(Smem &sm)
{
    u32 *hist = hist_of(sm);
    for (int i = threadIdx.x; i < PASSES * BINS; i += T) hist[i] = 0;
    if (threadIdx.x == 0) sm.s_count = 0;
    __syncthreads();
}

__device__ __forceinline__ void hist_add(Smem &sm, u64 rec)
{
    u32 *hist = hist_of(sm);
#pragma unroll
    for (int p = 0; p < PASSES; p++) atomicAdd(&hist[p * BINS + digit_of(rec, p)], 1u);
}

// Round 0: key = the first k symbols of the rotation (bwt_common.cuh: KeyCode), so h = k afterwards.
// Returns k.
__device__ __noinline__ u32 build_initial(Smem &sm, const u8 *__restrict__ S, u32 n, u64 *dst)
{
    const u32 tid = threadIdx.x;
    hist_clear(sm);
    for (int i = tid; i < 256; i += T) sm.present[i] = 0;
    __syncthreads();
    build_alphabet<T>(S, n, sm.present, &sm.kc, sm.scratch);
    const u32 k = sm.kc.k;
    if (k == 5) {
        for (u32 base = 0; base < n; base += TILE) {
#pragma unroll
            for (int kk = 0; kk < K; kk++) {
                const u32 i = base + kk * T + tid;
                if (i < n) {
                    const u64 rec = (raw_key5(S, n, i) << IDX_BITS) | i;
                    st_stream(dst + i, rec);
                    hist_add(sm, rec);
                }
            }
        }
    } else {
        for (u32 base = 0; base < n; base += TILE) {
#pragma unroll 2
            for (int kk = 0; kk < K; kk++) {
                const u32 i = base + kk * T + tid;
                if (i < n) {
                    const u64 rec = (packed_key(S, n, i, sm.kc) << IDX_BITS) | i;
                    st_stream(dst + i, rec);
                    hist_add(sm, rec);
                }
            }
        }
    }
    __syncthreads();
    if (tid == 0) sm.s_count = n;
    __syncthreads();
    return k;
}

// Stable rank of this warp's records among the warp's records with the same digit.  The warp owns
// the tile positions [w*256, w*256+256), row k = positions w*256 + 32k + lane.  One shared atomic
// with return value per digit group and row; atomics of successive rows to the same counter execute
// in program order, so rows need no barrier.  (A warp holds at most 256 records of a tile, so the
// 16-bit halves of a counter word never carry.)  Leaves the per-warp digit counts in whist[w].
__device__ __forceinline__ void rank_rows(Smem &sm, const u64 (&rec)[K], u32 nrows, int pass, u32 (&rk)[K])
{
    const u32 lane = lane_id(), w = warp_id();
    {   // zero this warp's counter row with 16-byte stores
        uint4 *row = reinterpret_cast<uint4 *>(sm.whist[w]);
        for (int b = lane; b < WORDS / 4; b += 32) row[b] = make_uint4(0, 0, 0, 0);
    }
    __syncwarp();
#pragma unroll
    for (int k = 0; k < K; k++) {
        rk[k] = 0;
        if ((u32)k < nrows) {                                   // (warp-uniform)
            const u32 d = digit_of(rec[k], pass);
            const u32 peers = match_digit(d);
            const u32 leader = 31 - __clz(peers);
            const u32 sh = (d & 1u) * 16u;
            u32 bcount = 0;
            if (lane == leader) bcount = atomicAdd(&sm.whist[w][d >> 1], (u32)__popc(peers) << sh);
            bcount = __shfl_sync(0xffffffffu, bcount, leader);
            rk[k] = ((bcount >> sh) & 0xffffu) + __popc(peers & lanemask_lt());
        }
    }
}

// Key of a group that is too large for a tile: [ rank : rank[idx+h] : idx ] for the records
// alist[0..cnt) (one group of the active list), with the digit histograms of all passes.
__device__ void build_group(Smem &sm, const u32 *rank, u32 n, u32 hm, const u64 *alist, u32 cnt, u64 *dst)
{
    hist_clear(sm);
    for (u32 base = 0; base < cnt; base += TILE) {
        u64 e[K];
        u32 r2[K];
#pragma unroll
        for (int k = 0; k < K; k++) {
            u32 j = base + k * T + threadIdx.x;
            e[k] = (j < cnt) ? alist[j] : 0;
        }
#pragma unroll
        for (int k = 0; k < K; k++) {
            u32 j = base + k * T + threadIdx.x;
            r2[k] = 0;
            if (j < cnt) {
                u32 q = ((u32)e[k] & IDX_MASK) + hm;
                if (q >= n) q -= n;
                r2[k] = ld_keep(rank + q) & RANK_MASK;
            }
        }
#pragma unroll
        for (int k = 0; k < K; k++) {
            u32 j = base + k * T + threadIdx.x;
            if (j < cnt) {
                const u32 idx = (u32)e[k] & IDX_MASK;
                const u64 r1 = (e[k] >> IDX_BITS) & RANK_MASK;
                const u64 rec = (r1 << (IDX_BITS + 20)) | ((u64)r2[k] << IDX_BITS) | idx;
                st_stream(dst + j, rec);
                hist_add(sm, rec);
            }
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) sm.s_count = cnt;
    __syncthreads();
}

// After the keys are built: park the per-pass histograms in global memory (the reorder buffer they
// were built in is needed by the passes) and note the passes whose digit is the same for every record.
__device__ void hist_park(Smem &sm, u32 *ghist, u32 count)
{
    const u32 *hist = hist_of(sm);
    if (threadIdx.x < 8) sm.s_flags[threadIdx.x] = 0;
    __syncthreads();
    for (int i = threadIdx.x; i < PASSES * BINS; i += T) {
        const u32 v = hist[i];
        ghist[i] = v;
        if (v == count) sm.s_flags[i / BINS] = 1;
    }
    __syncthreads();
}

// One LSD pass over `count` records through HBM: src -> dst by digit `pass`.
// Tiles are streamed through shared memory with TMA bulk copies (cp.async.bulk + mbarrier):
// the copy of tile t+1 is in flight while tile t is ranked, reordered and stored.
__device__ void radix_pass(Smem &sm, const u64 *src, u64 *dst, u32 count, int pass, const u32 *ghist, u32 &phase)
{
    constexpr int BPT = (BINS + T - 1) / T;
    const u32 tid = threadIdx.x, lane = lane_id(), w = warp_id();

    // records were written with generic-proxy stores; order them before the async-proxy reads
    fence_proxy_async();
    __syncthreads();
    if (tid == 0) {
        const u32 bytes = (min((u32)TILE, count) * 8u + 15u) & ~15u;
        mbar_expect_tx(&sm.mbar, bytes);
        tma_load_1d_stream(sm.buf0, src, bytes, &sm.mbar);
    }

    // cursor = exclusive scan of this pass's histogram
    {
        u32 c[BPT], sum = 0;
#pragma unroll
        for (int j = 0; j < BPT; j++) {
            u32 b = tid * BPT + j;
            c[j] = (b < BINS) ? ghist[pass * BINS + b] : 0;
            sum += c[j];
        }
        u32 tot;
        u32 ex = block_excl_sum<T>(sum, sm.scratch, &tot);
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
        mbar_wait(&sm.mbar, phase);
        phase ^= 1u;
        const u32 wl = w * (K * 32) + lane;
        if (tile_n == TILE) {
#pragma unroll
            for (int k = 0; k < K; k++) rec[k] = sm.buf0[wl + k * 32];
        } else {
#pragma unroll
            for (int k = 0; k < K; k++) {
                u32 j = wl + k * 32;
                rec[k] = (j < tile_n) ? sm.buf0[j] : ~0ull;
            }
        }
        rank_rows(sm, rec, K, pass, rk);
        __syncthreads();                                        // B1: buf0 consumed, whist complete

        if (tid == 0 && base + TILE < count) {                  // prefetch the next tile
            const u32 nb = (min((u32)TILE, count - base - TILE) * 8u + 15u) & ~15u;
            // Write-after-read across proxies: every thread's generic reads of buf0 have returned (their
            // values feed the ranking above) and B1 orders them before this copy is issued.  This is the
            // consumer-release / producer-acquire handshake of every TMA pipeline (no proxy fence on the
            // release side); a fence.proxy.async here compiles to a GPU-scope MEMBAR per tile.
            mbar_expect_tx(&sm.mbar, nb);
            tma_load_1d_stream(sm.buf0, src + base + TILE, nb, &sm.mbar);
        }

        // cross-warp exclusive scan per bin (two bins per packed word; a tile holds 4096 records, so
        // the halves never carry), then exclusive scan over bins
        if (tid < WORDS) {
            u32 run = 0;
#pragma unroll
            for (int ww = 0; ww < NW; ww++) {
                const u32 v = sm.whist[ww][tid];
                sm.whist[ww][tid] = run;
                run += v;
            }
            const u32 lo = run & 0xffffu, hi = run >> 16, sum = lo + hi;
            const u32 inc = warp_incl_sum(sum);
            if (lane == 31) sm.scratch[w] = inc;
            asm volatile("bar.sync 1, %0;" ::"n"(WORDS) : "memory");
            u32 woff = 0;
            for (u32 q = 0; q < w; q++) woff += sm.scratch[q];
            const u32 ex = woff + inc - sum;
            const u32 b0 = 2 * tid;
            const u32 cur0 = sm.cursor[b0], cur1 = sm.cursor[b0 + 1];
            sm.binoff[b0] = (u16)ex;
            sm.binoff[b0 + 1] = (u16)(ex + lo);
            sm.gbase[b0] = cur0 - ex;
            sm.gbase[b0 + 1] = cur1 - (ex + lo);
            sm.cursor[b0] = cur0 + lo;
            sm.cursor[b0 + 1] = cur1 + hi;
        }
        __syncthreads();                                        // B2

#pragma unroll
        for (int k = 0; k < K; k++) {
            const u32 d = digit_of(rec[k], pass);
            const u32 pos = sm.binoff[d] + ((sm.whist[w][d >> 1] >> ((d & 1u) * 16u)) & 0xffffu) + rk[k];
            sm.buf1[pos] = rec[k];
        }
        __syncthreads();                                        // B3

        if (tile_n == TILE) {
#pragma unroll
            for (int k = 0; k < K; k++) {
                const u32 j = k * T + tid;
                const u64 r = sm.buf1[j];
                st_stream(dst + sm.gbase[digit_of(r, pass)] + j, r);
            }
        } else {
#pragma unroll
            for (int k = 0; k < K; k++) {
                const u32 j = k * T + tid;
                if (j < tile_n) {
                    const u64 r = sm.buf1[j];
                    st_stream(dst + sm.gbase[digit_of(r, pass)] + j, r);
                }
            }
        }
        // no barrier here: the next tile's B1 orders these reads before buf1/whist are reused
    }
    __syncthreads();
}

// Scatter one tile of records (K per thread, any arrangement: stability is not needed) into the
// buckets of digit `pass`, whose running cursors live in sm.cursor: the tile body of radix_pass
// without the TMA load.  Records of threads beyond the valid ones must be ~0 (they sort last and are
// not stored); `tile_n` = valid records of the tile.
__device__ __forceinline__ void partition_tile(Smem &sm, const u64 (&rec)[K], u32 tile_n, int pass, u64 *dst)
{
    const u32 tid = threadIdx.x, lane = lane_id(), w = warp_id();
    u32 rk[K];
    rank_rows(sm, rec, K, pass, rk);
    __syncthreads();
    if (tid < WORDS) {
        u32 run = 0;
#pragma unroll
        for (int ww = 0; ww < NW; ww++) {
            const u32 v = sm.whist[ww][tid];
            sm.whist[ww][tid] = run;
            run += v;
        }
        const u32 lo = run & 0xffffu, hi = run >> 16, sum = lo + hi;
        const u32 inc = warp_incl_sum(sum);
        if (lane == 31) sm.scratch[w] = inc;
        asm volatile("bar.sync 1, %0;" ::"n"(WORDS) : "memory");
        u32 woff = 0;
        for (u32 q = 0; q < w; q++) woff += sm.scratch[q];
        const u32 ex = woff + inc - sum;
        const u32 b0 = 2 * tid;
        const u32 cur0 = sm.cursor[b0], cur1 = sm.cursor[b0 + 1];
        sm.binoff[b0] = (u16)ex;
        sm.binoff[b0 + 1] = (u16)(ex + lo);
        sm.gbase[b0] = cur0 - ex;
        sm.gbase[b0 + 1] = cur1 - (ex + lo);
        sm.cursor[b0] = cur0 + lo;
        sm.cursor[b0 + 1] = cur1 + hi;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < K; k++) {
        const u32 d = digit_of(rec[k], pass);
        const u32 pos = sm.binoff[d] + ((sm.whist[w][d >> 1] >> ((d & 1u) * 16u)) & 0xffffu) + rk[k];
        sm.buf1[pos] = rec[k];
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < K; k++) {
        const u32 j = k * T + tid;
        if (j < tile_n) {
            const u64 r = sm.buf1[j];
            st_stream(dst + sm.gbase[digit_of(r, pass)] + j, r);
        }
    }
    __syncthreads();
}

struct RerankOut {
    u32 active;      // records that still share their key after this step
    u32 splits;      // key heads that are not group heads (new groups created)
};

// Walk records sorted by their 40-bit key (round 0: all of the block, `initial`; later: one group
// that went through the global passes), assign new ranks, retire singletons, and append the
// records that stay active to the active list at list[out_pos ...] IN SORTED ORDER.
// `upd` != nullptr (round 0): instead of scattering the ranks, one update record per rotation goes to
// upd[], bucketed by idx >> 13, for apply_ranks_bucketed.
// (m_val, m_cnt): the group has m_cnt further members whose rank[idx+h] is m_val; they are not among the
// records (refine_majority) but take their places in the numbering.
__device__ RerankOut rerank(Smem &sm, const u64 *src, u32 count, bool initial, u32 *rank, u64 *list, u32 out_pos,
                            u64 *upd = nullptr, u32 m_val = 0xffffffffu, u32 m_cnt = 0)
{
    const u32 tid = threadIdx.x;
    u32 carry_grp = 0, carry_key = 0;       // 1-based positions of the latest heads so far
    u32 n_active = 0, n_split = 0;
    const u64 grp_mask = initial ? 0ull : ((u64)RANK_MASK << 20);   // bits of r1 inside key40
    if (upd) {
        // the update records are scattered at once into the buckets of apply_ranks_bucketed: bucket b
        // (the ranks of the positions [8192 b, 8192 b + 8192)) lives at upd[8192 b ...] and receives
        // exactly its positions, so its cursor starts there and no histogram is needed
        for (u32 b = tid; b < (u32)BINS; b += T) sm.cursor[b] = min(b << UPD_SHIFT, count);
        __syncthreads();
    }

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
        u64 tot2;
        const u64 ex2 = block_excl_max2<T>(((u64)mg << 32) | mk, sm.scratch64, &tot2);
        const u32 ex_k = max((u32)ex2, carry_key);
        const u32 ex_g = max((u32)(ex2 >> 32), carry_grp);
        const u32 tot_k = (u32)tot2, tot_g = (u32)(tot2 >> 32);

        u32 nrv[K];
        u32 flg[K];                          // bit0 valid, bit1 singleton, bit2 rank unchanged
        u32 mine = 0;                        // records of this thread that stay active
#pragma unroll
        for (int k = 0; k < K; k++) {
            u32 j = j0 + k;
            flg[k] = 0;
            nrv[k] = 0;
            if (j < count) {
                u32 p_key = max(pk[k], ex_k);        // 1-based
                u32 p_grp = max(pg[k], ex_g);
                bool hk = (pk[k] == j + 1);
                bool hg = (pg[k] == j + 1);
                u32 r1 = initial ? 0u : (u32)(key[k + 1] >> 20) & RANK_MASK;
                nrv[k] = r1 + (p_key - p_grp) + ((((u32)key[k + 1] & RANK_MASK) > m_val) ? m_cnt : 0u);
                bool single = hk && (j + 1 == count || key[k + 2] != key[k + 1]);
                flg[k] = 1u | (single ? 2u : 0u);
                // a record that stays in the first subgroup of its old group keeps its rank: its
                // rank[] entry is already correct, skip the (random, 32-byte-sector) store
                if (!single && !initial && nrv[k] == r1) flg[k] |= 4u;
                if (!single) mine++;
                if (hk && !hg) n_split++;
            }
        }
        // order-preserving compaction: thread t holds K consecutive records
        u32 tot_a;
        u32 at = out_pos + n_active + block_excl_sum<T>(mine, sm.scratch, &tot_a);
#pragma unroll
        for (int k = 0; k < K; k++) {
            if ((flg[k] & 3u) == 1u) st_stream(list + at++, ((u64)nrv[k] << IDX_BITS) | idx[k]);
        }
        n_active += tot_a;
        if (upd) {
            u64 urec[K];
#pragma unroll
            for (int k = 0; k < K; k++) urec[k] = (flg[k] & 1u) ? upd_record(nrv[k], (flg[k] & 2u) != 0, idx[k]) : ~0ull;
            partition_tile(sm, urec, min((u32)TILE, count - base), 0, upd);
        } else {
#pragma unroll
            for (int k = 0; k < K; k++) {
                if (flg[k] & 1u) {
                    const u32 id = idx[k];
                    if (flg[k] & 2u) {
                        st_keep(rank + id, nrv[k] | DONE);
                    } else if (!(flg[k] & 4u)) {
                        st_keep(rank + id, nrv[k]);
                    }
                }
            }
        }
        carry_key = max(carry_key, tot_k);
        carry_grp = max(carry_grp, tot_g);
    }
    RerankOut o;
    o.active = n_active;                     // (block-uniform: sum of the tile totals)
    o.splits = block_sum<T>(n_split, sm.scratch);
    return o;
}

// The ranks of round 0 reach rank[] without random DRAM accesses.  rerank scattered one update
// record per rotation into buckets by idx >> 13 (a bucket = the 8192 positions of one 32 KB chunk
// of rank[], at upd[8192 b ...]; the scatter is the tile body of a radix pass, fused into the
// re-rank walk); here every chunk is assembled in shared memory and written out with coalesced
// stores.  20 bytes of streaming traffic per rotation instead of a 32-byte sector read + write:
// with all CTAs scattering 4-byte ranks at once the kernel was bound by exactly those sectors
// (profiles/README.md; doing the same in the later rounds costs their in-place refinement and
// measured slower on the mixed corpus).
__device__ void apply_ranks_bucketed(Smem &sm, const u64 *upd, u32 n, u32 *rank)
{
    const u32 tid = threadIdx.x;
    u32 *chunk = reinterpret_cast<u32 *>(sm.buf0);              // 8192 ranks
    for (u32 lo = 0; lo < n; lo += UPD_CHUNK) {
        const u32 len = min(UPD_CHUNK, n - lo);
#pragma unroll
        for (int k = 0; k < 2 * K; k++) {
            const u32 j = k * T + tid;
            if (j < len) {
                const u64 e = ld_stream64(upd + lo + j);
                chunk[(u32)e & (UPD_CHUNK - 1)] = upd_rank_word(e);
            }
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < 2 * K; k++) {
            const u32 j = k * T + tid;
            if (j < len) st_keep(rank + lo + j, chunk[j]);
        }
        __syncthreads();
    }
}

__device__ __forceinline__ void tile_finish(Smem &sm, u64 (&rec)[K], u32 nrows, u32 tile_n, u32 G, u64 *list, u32 *rank,
                                            u32 &out_pos, u32 &n_active, u32 &n_split, u32 m_val, u32 m_cnt);

// One tile of the active list: the whole groups among list[p .. p+TILE), sorted by rank[idx + h]
// inside shared memory.  Returns the number of list records consumed; 0 = the group that starts at
// list[p] does not fit a tile (nothing was done).  The records that stay active are written back to
// list[out_pos ...] (out_pos <= p: the compaction is in place), out_pos advances.
// PERIODIC: the block has a periodic run (sm.per, bwt_common.cuh); a second instantiation, so that the
// code of ordinary blocks is exactly what it is without that case (in one function the extra branch
// cost them 17 % more cycles per tile in spills).
template <bool PERIODIC>
__device__ u32 refine_tile(Smem &sm, u64 *list, u32 p, u32 count, u32 hm, u32 n, u32 *rank, u32 &out_pos,
                           u32 &n_active, u32 &n_split)
{
    const u32 tid = threadIdx.x, lane = lane_id(), w = warp_id();
    const u32 q0 = w * (K * 32) + lane;
    const u32 avail = min((u32)TILE, count - p);
    const bool more = p + TILE < count;                 // records follow this tile

    // ---- records and group heads
    u64 rec[K];
    u32 hb[K];
#pragma unroll
    for (int k = 0; k < K; k++) {
        const u32 q = q0 + k * 32;
        rec[k] = (q < avail) ? list[p + q] : ~0ull;
    }
    u32 left0 = 0xffffffffu;                             // r1 of the record before this warp's first one
    if (lane == 0 && w > 0 && w * (K * 32) - 1 < avail) left0 = (u32)(list[p + w * (K * 32) - 1] >> IDX_BITS) & RANK_MASK;
    if (tid == T - 1) {
        // does the record after the tile start a new group?
        u32 la = 1;
        if (more) la = (((u32)(list[p + TILE] >> IDX_BITS) & RANK_MASK) != ((u32)(rec[K - 1] >> IDX_BITS) & RANK_MASK)) ? 1u : 0u;
        sm.s_la = la;
    }
    u32 cnt = 0, last = 0;                               // heads of this warp; last head position + 1 (heads at q >= 1 only)
    {
        u32 prev_row_last = left0;
#pragma unroll
        for (int k = 0; k < K; k++) {
            const u32 q = q0 + k * 32;
            const u32 r1 = (u32)(rec[k] >> IDX_BITS) & RANK_MASK;
            u32 lf = __shfl_up_sync(0xffffffffu, r1, 1);
            if (lane == 0) lf = prev_row_last;
            const bool head = q < avail && (q == 0 || r1 != lf);
            hb[k] = __ballot_sync(0xffffffffu, head);
            prev_row_last = __shfl_sync(0xffffffffu, r1, 31);
            cnt += __popc(hb[k]);
            u32 hq = hb[k];
            if (w == 0 && k == 0) hq &= ~1u;             // position 0 is a head by construction
            if (hq) last = w * (K * 32) + k * 32 + (31 - __clz(hq)) + 1;
        }
    }
    if (lane == 0) {
        sm.wcnt[w] = cnt;
        sm.wlast[w] = last;
    }
    __syncthreads();
    u32 woff = 0, total = 0, lh = 0;
#pragma unroll
    for (int v = 0; v < NW; v++) {
        const u32 c = sm.wcnt[v];
        if ((u32)v < w) woff += c;
        total += c;
        lh = max(lh, sm.wlast[v]);
    }
    u32 tile_n = avail, G = total;
    if (more && !sm.s_la) {
        if (lh == 0) return 0;                           // one group fills the tile and goes on (block-uniform)
        tile_n = lh - 1;                                 // cut at the last head: the groups before it are complete
        G = total - 1;
    }
    const u32 nrows = (tile_n > w * (K * 32)) ? min((u32)K, (tile_n - w * (K * 32) + 31) / 32) : 0u;

    // ---- rank[idx + h], group numbers, sort items [ group:12 | rank[idx+h]:20 | idx:20 ]
    if (PERIODIC) {
        // The block has a periodic run (bwt_common.cuh): a group whose members all lie inside the run and
        // in one class mod p is ordered by index, so its sort key is the index (or its complement) and no
        // rank is gathered; every member becomes a singleton.  r1tab[g] = [ bad:1 | class:11 | rank:20 ].
        const u32 ps = sm.per.s, pe = sm.per.e, asc = sm.per.asc;
        u32 rowbase = woff;
#pragma unroll
        for (int k = 0; k < K; k++) {
            const u32 q = q0 + k * 32;
            if (q < tile_n) {
                const u32 g = rowbase + __popc(hb[k] & lanemask_le()) - 1;
                if ((hb[k] >> lane) & 1u)
                    sm.r1tab[g] = ((u32)(rec[k] >> IDX_BITS) & RANK_MASK) | (sm.per.cls((u32)rec[k] & IDX_MASK) << 20);
                rec[k] = ((u64)g << (IDX_BITS + 20)) | ((u32)rec[k] & IDX_MASK);
            } else {
                rec[k] = ~0ull;
            }
            rowbase += __popc(hb[k]);
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < K; k++) {
            if (rec[k] != ~0ull) {
                const u32 id = (u32)rec[k] & IDX_MASK, g = (u32)(rec[k] >> (IDX_BITS + 20));
                if (id < ps || id >= pe || sm.per.cls(id) != ((sm.r1tab[g] >> 20) & 0x7ffu)) atomicOr(&sm.r1tab[g], 0x80000000u);
            }
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < K; k++) {
            if (rec[k] != ~0ull) {
                const u32 id = (u32)rec[k] & IDX_MASK, g = (u32)(rec[k] >> (IDX_BITS + 20));
                u32 r2;
                if (sm.r1tab[g] >> 31) {
                    u32 j = id + hm;
                    if (j >= n) j -= n;
                    r2 = ld_keep(rank + j) & RANK_MASK;
                } else {
                    r2 = asc ? id : (IDX_MASK - id);
                }
                rec[k] |= (u64)r2 << IDX_BITS;
            }
        }
    } else {
        u32 r2[K];
#pragma unroll
        for (int k = 0; k < K; k++) {
            const u32 q = q0 + k * 32;
            r2[k] = 0;
            if (q < tile_n) {
                u32 j = ((u32)rec[k] & IDX_MASK) + hm;
                if (j >= n) j -= n;
                r2[k] = ld_keep(rank + j) & RANK_MASK;
            }
        }
        u32 rowbase = woff;
#pragma unroll
        for (int k = 0; k < K; k++) {
            const u32 q = q0 + k * 32;
            if (q < tile_n) {
                const u32 g = rowbase + __popc(hb[k] & lanemask_le()) - 1;
                if ((hb[k] >> lane) & 1u) sm.r1tab[g] = (u32)(rec[k] >> IDX_BITS) & RANK_MASK;
                rec[k] = ((u64)g << (IDX_BITS + 20)) | ((u64)r2[k] << IDX_BITS) | ((u32)rec[k] & IDX_MASK);
            } else {
                rec[k] = ~0ull;
            }
            rowbase += __popc(hb[k]);
        }
    }

    tile_finish(sm, rec, nrows, tile_n, G, list, rank, out_pos, n_active, n_split, 0xffffffffu, 0u);
    return tile_n;
}

// Second half of a tile: `rec` holds the sort items [ group:12 | rank[idx+h]:20 | idx:20 ] of the
// tile positions w*256 + 32k + lane (~0 beyond tile_n), sm.r1tab the rank of every group.
// Sorts them in shared memory, derives the new ranks, stores the changed ones, and writes the
// records that stay active to list[out_pos ...].  (m_val, m_cnt): the group also has m_cnt members
// whose rank[idx+h] is m_val; they are not in the tile (refine_majority) but take their places in
// the numbering, between the items below and the items above m_val.
__device__ __forceinline__ void tile_finish(Smem &sm, u64 (&rec)[K], u32 nrows, u32 tile_n, u32 G, u64 *list, u32 *rank,
                                            u32 &out_pos, u32 &n_active, u32 &n_split, u32 m_val, u32 m_cnt)
{
    const u32 tid = threadIdx.x, lane = lane_id(), w = warp_id();
    const u32 q0 = w * (K * 32) + lane;
    // ---- stable LSD radix sort of the tile in shared memory (20 + log2(G) key bits)
    const int npass = (G <= 16) ? 3 : 4;
    u64 *fin = nullptr;
    for (int pass = 0; pass < npass; pass++) {
        u64 *dstb = (pass & 1) ? sm.buf0 : sm.buf1;
        if (pass > 0) {
            const u64 *srcb = (pass & 1) ? sm.buf1 : sm.buf0;
#pragma unroll
            for (int k = 0; k < K; k++) rec[k] = ((u32)k < nrows) ? srcb[q0 + k * 32] : ~0ull;
        }
        u32 rk[K];
        rank_rows(sm, rec, nrows, pass, rk);
        __syncthreads();
        if (tid < WORDS) {
            u32 run = 0;
#pragma unroll
            for (int ww = 0; ww < NW; ww++) {
                const u32 v = sm.whist[ww][tid];
                sm.whist[ww][tid] = run;
                run += v;
            }
            const u32 lo = run & 0xffffu, hi = run >> 16, sum = lo + hi;
            const u32 inc = warp_incl_sum(sum);
            if (lane == 31) sm.scratch[w] = inc;
            asm volatile("bar.sync 1, %0;" ::"n"(WORDS) : "memory");
            u32 wo = 0;
            for (u32 q = 0; q < w; q++) wo += sm.scratch[q];
            const u32 ex = wo + inc - sum;
            sm.binoff[2 * tid] = (u16)ex;
            sm.binoff[2 * tid + 1] = (u16)(ex + lo);
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < K; k++) {
            if ((u32)k < nrows) {
                const u32 d = digit_of(rec[k], pass);
                const u32 pos = sm.binoff[d] + ((sm.whist[w][d >> 1] >> ((d & 1u) * 16u)) & 0xffffu) + rk[k];
                dstb[pos] = rec[k];
            }
        }
        __syncthreads();
        fin = dstb;
    }

    // ---- head flags of the sorted tile (kept as bitmaps in shared memory: one word per warp row)
    u64 *spare = (fin == sm.buf0) ? sm.buf1 : sm.buf0;
    {
#pragma unroll
        for (int k = 0; k < K; k++) {
            const u32 q = q0 + k * 32;
            const u32 keyhi = ((u32)k < nrows) ? (u32)(fin[q] >> IDX_BITS) : 0xffffffffu;   // [ group:12 | r2:20 ]
            u32 lf = __shfl_up_sync(0xffffffffu, keyhi, 1);
            if (lane == 0) lf = (q > 0 && (u32)k < nrows) ? (u32)(fin[q - 1] >> IDX_BITS) : 0xffffffffu;
            const bool valid = q < tile_n;
            const bool ghead = valid && (q == 0 || (keyhi >> 20) != (lf >> 20));
            const bool khead = valid && (q == 0 || keyhi != lf);
            const u32 kbw = __ballot_sync(0xffffffffu, khead);
            const u32 gbw = __ballot_sync(0xffffffffu, ghead);
            if (lane == 0) {
                sm.bm_k[w * K + k] = kbw;
                sm.bm_g[w * K + k] = gbw;
            }
        }
    }
    __syncthreads();
    if (w < 2) {
        // per bitmap word: 1-based position of the last head before the word (exclusive running maximum)
        const u32 *bm = w ? sm.bm_g : sm.bm_k;
        u32 *pre = w ? sm.pre_g : sm.pre_k;
        u32 loc[4], run = 0;
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const u32 wd = lane * 4 + j;
            loc[j] = run;
            const u32 m = bm[wd];
            if (m) run = wd * 32 + (31 - __clz(m)) + 1;
        }
        const u32 inc = warp_incl_max(run);
        u32 ex = __shfl_up_sync(0xffffffffu, inc, 1);
        if (lane == 0) ex = 0;
#pragma unroll
        for (int j = 0; j < 4; j++) pre[lane * 4 + j] = max(loc[j], ex);
    }
    __syncthreads();

    // ---- new ranks, singletons, rank[] scatter; the new list records wait in the spare buffer
#pragma unroll
    for (int k = 0; k < K; k++) {
        if ((u32)k < nrows) {                                     // (warp-uniform)
            const u32 q = q0 + k * 32;
            const u32 wd = w * K + k;
            const bool valid = q < tile_n;
            const u32 kbw = sm.bm_k[wd], gbw = sm.bm_g[wd];
            const u32 mk = kbw & lanemask_le(), mg = gbw & lanemask_le();
            const u32 pk = mk ? wd * 32 + (31 - __clz(mk)) + 1 : sm.pre_k[wd];
            const u32 pg = mg ? wd * 32 + (31 - __clz(mg)) + 1 : sm.pre_g[wd];
            const u64 it = fin[q];
            const u32 g = (u32)(it >> (IDX_BITS + 20)) & 0xfffu;
            const u32 id = (u32)it & IDX_MASK;
            const u32 r1 = valid ? (sm.r1tab[g] & RANK_MASK) : 0u;      // (the upper bits: refine_tile on periodic blocks)
            const u32 nr = r1 + (pk - pg) + ((((u32)(it >> IDX_BITS) & RANK_MASK) > m_val) ? m_cnt : 0u);
            const bool khead = (kbw >> lane) & 1u;
            u32 nxt;                                              // is position q + 1 a key head (or the end)?
            if (lane < 31) nxt = (kbw >> (lane + 1)) & 1u;
            else nxt = (wd + 1 < BMW) ? (sm.bm_k[(wd + 1) & (BMW - 1)] & 1u) : 1u;
            const bool single = khead && (q + 1 >= tile_n || nxt);
            const bool active = valid && !single;
            const u32 abw = __ballot_sync(0xffffffffu, active);
            if (lane == 0) {
                sm.bm_a[wd] = abw;
                n_active += __popc(abw);
                n_split += __popc(kbw & ~gbw);
            }
            // rank stores: singletons are final, the first subgroup of a group keeps its rank
            if (valid && (single || nr != r1)) st_keep(rank + id, nr | (single ? DONE : 0u));
            if (active) spare[q] = ((u64)nr << IDX_BITS) | id;
        } else if (lane == 0) {
            sm.bm_a[w * K + k] = 0;
        }
    }
    __syncthreads();
    if (w == 0) {
        u32 loc[4], run = 0;
#pragma unroll
        for (int j = 0; j < 4; j++) {
            loc[j] = run;
            run += __popc(sm.bm_a[lane * 4 + j]);
        }
        const u32 inc = warp_incl_sum(run);
        const u32 ex = inc - run;
#pragma unroll
        for (int j = 0; j < 4; j++) sm.pre_a[lane * 4 + j] = ex + loc[j];
        if (lane == 31) sm.s_tot = inc;
    }
    __syncthreads();
    // ---- the records that stay active go back to the list, in sorted order
#pragma unroll
    for (int k = 0; k < K; k++) {
        if ((u32)k < nrows) {
            const u32 wd = w * K + k;
            const u32 abw = sm.bm_a[wd];
            if ((abw >> lane) & 1u) st_stream(list + out_pos + sm.pre_a[wd] + __popc(abw & lanemask_lt()), spare[q0 + k * 32]);
        }
    }
    out_pos += sm.s_tot;
    __syncthreads();
}

// first position > p + TILE of the active list whose rank differs from r1p (the list is sorted
// by rank, the records up to p + TILE are known to carry r1p), or count
__device__ u32 group_end(Smem &sm, const u64 *list, u32 p, u32 count, u32 r1p)
{
    const u32 tid = threadIdx.x;
    u32 a = p + TILE + 1, b = count;          // answer in [a, b]; every position < a matches, b mismatches (or is the end)
    if (a > b) a = b;
    while (a < b) {
        const u32 span = b - a;
        const u32 step = (span + T - 1) / T;
        if (tid == 0) sm.s_min = T;
        __syncthreads();
        const u32 pos = a + tid * step;
        const bool mism = pos >= b || (((u32)(list[pos] >> IDX_BITS) & RANK_MASK) != r1p);
        if (mism) atomicMin(&sm.s_min, tid);
        __syncthreads();
        const u32 f = sm.s_min;
        __syncthreads();
        if (f == T) {
            a = a + (T - 1) * step + 1;
            if (a > b) a = b;
        } else {
            const u32 nb = min(b, a + f * step);
            if (f > 0) a = a + (f - 1) * step + 1;
            b = nb;
            if (a > b) a = b;
        }
    }
    return a;
}

// A group larger than a tile in which (nearly) every member has the same rank[idx + h] = M: the
// signature of periodic data, where a group of ~n/period rotations loses only the few members
// within h of the point where the period breaks per round (the reference's SA-IS has no such
// worst case, README.md:7; plain doubling pays 17 rounds of full-group sorts for it).  Sorting is
// then a three-way split: the members equal to M keep their order and stay one group, and the
// <= 4096 others are sorted as one tile.  Two streaming passes over the group (one gather each
// way) instead of key build + three radix passes + re-rank.  Returns 0 (nothing done) if the group
// does not have that shape, 1 if it is done, 2 if the members equal to M are done and the others
// (more than a tile, at most half of the group) wait in outl[0 .. *n_others) as key records for the
// global passes; *m_val = M, *m_cnt = members equal to M.
__device__ int refine_majority(Smem &sm, u64 *list, u32 p, u32 m, u32 hm, u32 n, u32 *rank, u64 *outl, u32 *bitmap,
                               u32 &out_pos, u32 &n_active, u32 &n_split, u32 *m_val, u32 *m_cnt, u32 *n_others)
{
    const u32 tid = threadIdx.x, lane = lane_id(), w = warp_id();
    const u32 r1 = (u32)(list[p] >> IDX_BITS) & RANK_MASK;
    // ---- M from 32 samples; give up unless most of them agree
    if (w == 0) {
        const u32 j = p + (u32)(((u64)m * (2 * lane + 1)) >> 6);
        u32 q = ((u32)list[j] & IDX_MASK) + hm;
        if (q >= n) q -= n;
        const u32 r2 = ld_keep(rank + q) & RANK_MASK;
        const u32 mv = __shfl_sync(0xffffffffu, r2, 16);
        const u32 agree = __popc(__ballot_sync(0xffffffffu, r2 == mv));
        if (lane == 0) {
            sm.wcnt[0] = mv;
            sm.wcnt[1] = agree;
            sm.s_tot = 0;                                    // members that differ from M, appended to outl
        }
    }
    __syncthreads();
    const u32 M = sm.wcnt[0];
    if (sm.wcnt[1] < 28) return 0;                           // (block-uniform)

    // ---- pass A: classify; the members that differ from M go to outl[], one bit per member to bitmap[]
    u32 n_lt = 0;
    for (u32 base = 0; base < m; base += TILE) {
        u64 rec[K];
        u32 r2[K];
#pragma unroll
        for (int k = 0; k < K; k++) {
            const u32 j = base + w * (K * 32) + k * 32 + lane;
            rec[k] = (j < m) ? list[p + j] : 0ull;
        }
#pragma unroll
        for (int k = 0; k < K; k++) {
            const u32 j = base + w * (K * 32) + k * 32 + lane;
            r2[k] = M;
            if (j < m) {
                u32 q = ((u32)rec[k] & IDX_MASK) + hm;
                if (q >= n) q -= n;
                r2[k] = ld_keep(rank + q) & RANK_MASK;
            }
        }
#pragma unroll
        for (int k = 0; k < K; k++) {
            const u32 j = base + w * (K * 32) + k * 32 + lane;
            const bool odd = r2[k] != M;                     // (never for j >= m)
            const u32 ob = __ballot_sync(0xffffffffu, odd);
            if (base + w * (K * 32) + k * 32 < m) {          // (warp-uniform)
                if (lane == 0) bitmap[(base + w * (K * 32) + k * 32) >> 5] = ob;
                if (ob) {
                    u32 at = 0;
                    if (lane == 0) at = atomicAdd(&sm.s_tot, (u32)__popc(ob));
                    at = __shfl_sync(0xffffffffu, at, 0) + __popc(ob & lanemask_lt());
                    if (odd) outl[at] = ((u64)r1 << (IDX_BITS + 20)) | ((u64)r2[k] << IDX_BITS) | ((u32)rec[k] & IDX_MASK);
                    if (odd && r2[k] < M) n_lt++;
                }
            }
            (void)j;
        }
    }
    n_lt = block_sum<T>(n_lt, sm.scratch);                   // (ends with barriers: s_tot, outl and bitmap are complete)
    const u32 n_out = sm.s_tot;
    __syncthreads();
    if (n_out > (u32)TILE && (u64)n_out * 2 > m) return 0;   // (nothing but scratch was written so far)
    const u32 n_eq = m - n_out;
    *m_val = M;
    *m_cnt = n_eq;
    *n_others = n_out;
    const u32 nr_eq = r1 + n_lt;
    const bool eq_single = n_eq == 1;

    // ---- pass C: the members equal to M, in their old order, to the front of the group's place in the
    // list (in place: a tile is read completely before anything of it is written, and nothing moves right)
    if (n_eq > 0) {
        u32 done_eq = 0;
        for (u32 base = 0; base < m; base += TILE) {
            u64 rec[K];
            u32 eb[K];
            u32 cnt = 0;
#pragma unroll
            for (int k = 0; k < K; k++) {
                const u32 row = base + w * (K * 32) + k * 32;
                const u32 j = row + lane;
                rec[k] = (j < m) ? list[p + j] : 0ull;
                u32 ob = (row < m) ? bitmap[row >> 5] : 0xffffffffu;
                if (row + 32 > m && row < m) ob |= ~((1u << (m - row)) - 1u);        // beyond the group: not a member
                eb[k] = ~ob;
                cnt += __popc(eb[k]);
            }
            if (lane == 0) sm.wcnt[w] = cnt;
            __syncthreads();
            u32 woff = 0, total = 0;
#pragma unroll
            for (int v = 0; v < NW; v++) {
                const u32 c = sm.wcnt[v];
                if ((u32)v < w) woff += c;
                total += c;
            }
            u32 rowbase = out_pos + done_eq + woff;
#pragma unroll
            for (int k = 0; k < K; k++) {
                if ((eb[k] >> lane) & 1u) {
                    const u32 id = (u32)rec[k] & IDX_MASK;
                    if (!eq_single) st_stream(list + rowbase + __popc(eb[k] & lanemask_lt()), ((u64)nr_eq << IDX_BITS) | id);
                    if (eq_single || nr_eq != r1) st_keep(rank + id, nr_eq | (eq_single ? DONE : 0u));
                }
                rowbase += __popc(eb[k]);
            }
            done_eq += total;
            __syncthreads();
        }
        if (!eq_single) {
            out_pos += n_eq;
            if (tid == 0) n_active += n_eq;
        }
    }

    if (n_out > (u32)TILE) {
        if (tid == 0 && n_eq > 0) n_split++;                 // the members equal to M split from the others
        return 2;
    }
    // ---- the others: one tile, sorted in shared memory, numbered around the members equal to M
    if (n_out > 0) {
        u64 rec[K];
        const u32 nrows = (n_out > w * (K * 32)) ? min((u32)K, (n_out - w * (K * 32) + 31) / 32) : 0u;
#pragma unroll
        for (int k = 0; k < K; k++) {
            const u32 q = w * (K * 32) + k * 32 + lane;
            rec[k] = (q < n_out) ? (outl[q] & ((1ull << (IDX_BITS + 20)) - 1ull)) : ~0ull;      // group number 0
        }
        if (tid == 0) {
            sm.r1tab[0] = r1;
            if (n_eq > 0) n_split++;                         // the members equal to M split from the others
        }
        __syncthreads();
        tile_finish(sm, rec, nrows, n_out, 1, list, rank, out_pos, n_active, n_split, M, n_eq);
    }
    return 1;
}

// A group larger than a tile (list[p .. p+m)) of a block with a periodic run (bwt_common.cuh): if its
// members are exactly the arithmetic progression lo, lo + p, ..., hi inside the run — what a run of
// period p leaves in one group: "abab..." two groups of n/2 rotations — their final order is their
// index order (ascending or descending by per.asc), so every member gets its final rank at once: two
// streaming passes over the group, no gather, no further rounds.  Returns false (nothing done) if the
// group does not have that shape.
__device__ bool refine_periodic_big(Smem &sm, const u64 *list, u32 p, u32 m, u32 *rank, u32 &n_split)
{
    const u32 tid = threadIdx.x;
    const u32 pp = sm.per.p, ps = sm.per.s, pe = sm.per.e;
    const u32 r1 = (u32)(list[p] >> IDX_BITS) & RANK_MASK;
    const u32 c0 = sm.per.cls((u32)list[p] & IDX_MASK);
    if (tid == 0) {
        sm.s_min = 0xffffffffu;
        sm.s_tot = 0;
    }
    __syncthreads();
    u32 mn = 0xffffffffu, mx = 0;
    int bad = 0;
    for (u32 j = tid; j < m; j += T) {
        const u32 id = (u32)list[p + j] & IDX_MASK;
        if (id < ps || id >= pe || sm.per.cls(id) != c0) bad = 1;
        mn = min(mn, id);
        mx = max(mx, id);
    }
    mn = __reduce_min_sync(0xffffffffu, mn);
    mx = __reduce_max_sync(0xffffffffu, mx);
    if (lane_id() == 0) {
        atomicMin(&sm.s_min, mn);
        atomicMax(&sm.s_tot, mx);
    }
    const int anybad = __syncthreads_or(bad);
    const u32 lo = sm.s_min, hi = sm.s_tot;
    __syncthreads();
    // m distinct indices of one class in [lo, hi] with (hi - lo) / p + 1 == m: every class member in between
    if (anybad || (u64)(hi - lo) != (u64)(m - 1) * pp) return false;
    const u32 asc = sm.per.asc;
    for (u32 j = tid; j < m; j += T) {
        const u32 id = (u32)list[p + j] & IDX_MASK;
        const u32 pos = asc ? sm.per.div(id - lo) : sm.per.div(hi - id);
        st_keep(rank + id, (r1 + pos) | DONE);
    }
    if (tid == 0) n_split += m - 1;
    __syncthreads();
    return true;
}

// digit histograms of key records that are already in place (the others of refine_majority)
__device__ void hist_records(Smem &sm, const u64 *recs, u32 cnt)
{
    hist_clear(sm);
    for (u32 base = 0; base < cnt; base += TILE) {
#pragma unroll
        for (int k = 0; k < K; k++) {
            const u32 j = base + k * T + threadIdx.x;
            if (j < cnt) hist_add(sm, recs[j]);
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) sm.s_count = cnt;
    __syncthreads();
}

// Remaining groups are sets of identical rotations: give them distinct positions inside
// their group (any order yields the same BWT bytes) and derive origPtr by the
// descending-index rule.  `list` = the active list [ rank:20 | idx:20 ].
__device__ void finalize_ties(Smem &sm, const u64 *list, u32 count, const u8 *__restrict__ S, u32 n, const u32 *rank,
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
        r1[0] = (j0 > 0 && j0 - 1 < count) ? (u32)(list[j0 - 1] >> IDX_BITS) & RANK_MASK : 0xffffffffu;
        u32 pg[K], mg = 0;
#pragma unroll
        for (int k = 0; k < K; k++) {
            u32 j = j0 + k;
            u64 r = (j < count) ? list[j] : ~0ull;
            r1[k + 1] = (u32)(r >> IDX_BITS) & RANK_MASK;
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

// (a call, not inlined: the test must not disturb the register allocation of the kernel of ordinary blocks)
__device__ __noinline__ void detect_period_call(const u8 *S, u32 n, u32 *sh, Period *out) { detect_period<T>(S, n, sh, out); }

// PERIODIC = false: the kernel of ordinary blocks.  With a.defer_list it tests every block it claims for a
// long periodic run and leaves those to the follow-up launch of the PERIODIC = true instantiation, which
// takes its blocks from a.blk_list and knows the order of such rotations in closed form (bwt_common.cuh:
// Period).  Two instantiations, because the extra branches cost ordinary blocks 4-5 % (register spills)
// when both lived in one kernel.
template <bool PERIODIC>
__global__ void __launch_bounds__(T, 2) bwt_sort_kernel(BwtArgs a)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    Smem &sm = *reinterpret_cast<Smem *>(smem_raw);
    const u32 tid = threadIdx.x;

    if (tid == 0) {
        mbar_init(&sm.mbar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    u32 phase = 0;                       // parity of the next TMA completion

    u64 *list = a.ws_rec + (size_t)blockIdx.x * 3 * a.ws_stride;     // the active list
    u64 *bufC = list + a.ws_stride;                                   // key records of the global passes
    u64 *bufD = bufC + a.ws_stride;
    u32 *rank = a.ws_rank + (size_t)blockIdx.x * a.ws_stride;
    u32 *ghist = a.ws_hist + (size_t)blockIdx.x * BWT_HIST_WORDS;

    for (;;) {
        if (tid == 0) {
            // (the list of a follow-up launch: the blocks the cluster kernel left to this one)
            const u32 pos = atomicAdd(a.next_block, 1u);
            const u32 nb = a.n_blocks_dev ? __ldcg(a.n_blocks_dev) : a.n_blocks;
            sm.s_block = pos >= nb ? 0xffffffffu : (a.blk_list ? __ldcg(a.blk_list + pos) : pos);
        }
        __syncthreads();
        const u32 blk = sm.s_block;
        __syncthreads();
        if (blk == 0xffffffffu) break;

        const u8 *S = a.rle + a.blk_off[blk];
        u8 *bwt_out = a.bwt + a.blk_off[blk];
        const u32 n = a.blk_len[blk];
        u32 *ptr_out = a.ptr + blk;

        if (PERIODIC) {
            detect_period<T>(S, n, sm.scratch, &sm.per);
        } else {
            if (a.defer_list) {
                detect_period_call(S, n, sm.scratch, &sm.per);
                if (sm.per.p != 0) {                             // (block-uniform)
                    if (tid == 0) a.defer_list[atomicAdd(a.defer_count, 1u)] = blk;
                    __syncthreads();
                    continue;
                }
            }
        }

        u32 rounds = 1;
        bool tied = false;
        if (tid == 0) {
            for (int i = 0; i < 8; i++) sm.acc[i] = 0;
            sm.acc[ACC_ACTIVE] = n;
        }
        auto acc = [&](int which, u64 v) { if (tid == 0) sm.acc[which] += v; };

        // global passes over `cnt` key records in bufC; returns the buffer that holds the sorted records
        auto sort_keys = [&](u32 cnt) -> const u64 * {
            const long long c0 = clock64();
            hist_park(sm, ghist, cnt);
            u64 *src = bufC, *dst = bufD;
            u32 passes_run = 0;
            for (int p = 0; p < PASSES; p++) {
                if (sm.s_flags[p]) continue;               // every record has the same digit: nothing moves
                radix_pass(sm, src, dst, cnt, p, ghist, phase);
                u64 *t = src; src = dst; dst = t;
                passes_run++;
            }
            acc(ACC_PASSES, (u64)cnt * passes_run);
            acc(ACC_CYC_RADIX, (u64)(clock64() - c0));
            return src;
        };

        // ---- round 0: all rotations by their first h0 symbols (5 bytes, or more symbols of a small alphabet)
        u32 count, h0;
        {
            long long c0 = clock64();
            h0 = build_initial(sm, S, n, bufC);
            acc(ACC_CYC_BUILD, (u64)(clock64() - c0));
            const u64 *sorted = sort_keys(n);
            c0 = clock64();
            u64 *upd = (sorted == bufC) ? bufD : bufC;             // the pass buffer that is free
            RerankOut ro = rerank(sm, sorted, n, true, rank, list, 0, upd);
            count = ro.active;
            __syncthreads();
            apply_ranks_bucketed(sm, upd, n, rank);
            acc(ACC_CYC_RERANK, (u64)(clock64() - c0));
        }

        // ---- rounds h = 5, 10, 20, ...: refine the groups of the active list
        u32 h = h0;
        while (count > 0 && rounds < MAX_ROUNDS) {
            const u32 hm = h % n;
            u32 p = 0, out_pos = 0, n_act = 0, n_spl = 0, big_spl = 0;
            acc(ACC_ACTIVE, count);
            rounds++;
            while (p < count) {
                long long c0 = clock64();
                const u32 used = (PERIODIC && sm.per.p != 0) ? refine_tile<true>(sm, list, p, count, hm, n, rank, out_pos, n_act, n_spl)
                                                             : refine_tile<false>(sm, list, p, count, hm, n, rank, out_pos, n_act, n_spl);
                if (used) {
                    acc(ACC_TILE, used);
                    acc(ACC_CYC_TILE, (u64)(clock64() - c0));
                    p += used;
                    continue;
                }
                // a group larger than a tile: sort it by rank[idx + h] through HBM
                const u32 ge = group_end(sm, list, p, count, (u32)(list[p] >> IDX_BITS) & RANK_MASK);
                const u32 m = ge - p;
                if (PERIODIC && sm.per.p != 0 && refine_periodic_big(sm, list, p, m, rank, n_spl)) {
                    acc(ACC_TILE, m);
                    acc(ACC_CYC_TILE, (u64)(clock64() - c0));
                    p = ge;
                    continue;
                }
                u32 m_val = 0xffffffffu, m_cnt = 0, n_keys = m;
                const int how = refine_majority(sm, list, p, m, hm, n, rank, bufC, reinterpret_cast<u32 *>(bufD), out_pos, n_act,
                                                n_spl, &m_val, &m_cnt, &n_keys);
                if (how == 1) {
                    acc(ACC_TILE, m);
                    acc(ACC_CYC_TILE, (u64)(clock64() - c0));
                    p = ge;
                    continue;
                }
                if (how == 2) {
                    acc(ACC_TILE, m_cnt);
                    hist_records(sm, bufC, n_keys);          // (the other members are in bufC already)
                } else {
                    build_group(sm, rank, n, hm, list + p, m, bufC);
                }
                acc(ACC_CYC_BUILD, (u64)(clock64() - c0));
                const u64 *srt = sort_keys(n_keys);
                c0 = clock64();
                RerankOut rg = rerank(sm, srt, n_keys, false, rank, list, out_pos, nullptr, m_val, m_cnt);
                out_pos += rg.active;
                big_spl += rg.splits;
                __syncthreads();
                acc(ACC_CYC_RERANK, (u64)(clock64() - c0));
                p = ge;
            }
            const u32 splits = block_sum<T>(n_spl, sm.scratch) + big_spl;
            count = out_pos;
            if (count > 0 && splits == 0) {
                finalize_ties(sm, list, count, S, n, rank, bwt_out, ptr_out);
                tied = true;
                break;
            }
            if (h < (1u << 30)) h *= 2;
            (void)n_act;
        }

        // BWT bytes of every rotation that was ranked uniquely: bwt[rank[i]] = S[i-1].  rank[] and S
        // are read in index order (coalesced); the one-byte scatter covers the block's whole output
        // within this one short pass, so the sectors fill up in L2 instead of costing a 32-byte
        // DRAM gather per rotation inside the re-rank steps.
        __syncthreads();
        const long long cf0 = clock64();
        for (u32 base = 0; base < n; base += 4 * T) {
            u32 r[4], c[4];
#pragma unroll
            for (int k = 0; k < 4; k++) {
                const u32 i = base + k * T + tid;
                r[k] = (i < n) ? rank[i] : 0u;
            }
#pragma unroll
            for (int k = 0; k < 4; k++) {
                const u32 i = base + k * T + tid;
                c[k] = (i < n) ? S[i == 0 ? n - 1 : i - 1] : 0u;
            }
#pragma unroll
            for (int k = 0; k < 4; k++) {
                if (r[k] & DONE) bwt_out[r[k] & RANK_MASK] = (u8)c[k];
                const u32 i = base + k * T + tid;
                if (a.marks && i < n && (i % VERIFY_SPACING) == 0) a.marks[(size_t)blk * VERIFY_MARKS + i / VERIFY_SPACING] = r[k] & RANK_MASK;
            }
        }
        acc(ACC_CYC_FINAL, (u64)(clock64() - cf0));
        if (tid == 0 && (rank[0] & DONE)) *ptr_out = rank[0] & RANK_MASK;
        if (tid < 256) a.has_byte[(size_t)blk * 256 + tid] = sm.present[tid];
        if (tid == 0 && a.stats) {
            BwtStats st;
            st.n = n;
            st.rounds = rounds;
            st.tied = tied ? 1u : 0u;
            st.period = PERIODIC ? sm.per.p : 0u;
            st.sum_active = sm.acc[ACC_ACTIVE];
            st.sum_active_passes = sm.acc[ACC_PASSES];
            st.sum_tile = sm.acc[ACC_TILE];
            st.cyc_build = sm.acc[ACC_CYC_BUILD];
            st.cyc_radix = sm.acc[ACC_CYC_RADIX];
            st.cyc_rerank = sm.acc[ACC_CYC_RERANK];
            st.cyc_tile = sm.acc[ACC_CYC_TILE];
            st.cyc_final = sm.acc[ACC_CYC_FINAL];
            a.stats[blk] = st;
        }
        __syncthreads();
        if (a.done && tid == 0) {
            // every thread's stores of this block precede the barrier; publish them, then raise the
            // block's flag (host-mapped memory: the host launches the follow-up work of finished blocks)
            __threadfence_system();
            asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(a.done + blk), "r"(1u) : "memory");
        }
    }
}

}  // namespace bwt

size_t bwt_smem_bytes() { return sizeof(bwt::Smem); }

cudaError_t bwt_max_ctas(int *ctas_per_sm)
{
    size_t smem = bwt_smem_bytes();
    cudaError_t e = cudaFuncSetAttribute(bwt::bwt_sort_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(bwt::bwt_sort_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    int p0 = 0, p1 = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&p0, bwt::bwt_sort_kernel<false>, bwt::T, smem);
    if (e != cudaSuccess) return e;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&p1, bwt::bwt_sort_kernel<true>, bwt::T, smem);
    *ctas_per_sm = p0 < p1 ? p0 : p1;
    return e;
}

cudaError_t bwt_launch(const BwtArgs &a, int grid, cudaStream_t stream, bool periodic)
{
    if (periodic) bwt::bwt_sort_kernel<true><<<grid, bwt::T, bwt_smem_bytes(), stream>>>(a);
    else bwt::bwt_sort_kernel<false><<<grid, bwt::T, bwt_smem_bytes(), stream>>>(a);
    return cudaGetLastError();
}

}  // namespace bnz
