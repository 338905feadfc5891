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
//   key became unique gets its final rank and drops out of later rounds.
//   A round that splits no group proves the remaining groups are identical rotations
//   (period | n); their positions inside the group are arbitrary for the BWT bytes and
//   origPtr = group base + group size - 1.
//   Last, one pass in index order writes bwt[rank[i]] = S[i-1] (coalesced reads, a one-byte
//   scatter that covers the block's output at once and merges in L2).
//
// Radix pass (per tile of TILE records, sequential over tiles inside the CTA so that the
// running bucket cursors live in shared memory): TMA bulk copy of the next tile while this one
// is processed, per-warp stable ranking (one ballot per digit bit finds the lanes with the
// same digit, one shared atomic with return value per digit group; counters are 16-bit pairs),
// cross-warp scan, shared-memory reorder, coalesced bucket-run stores with an L2 evict_first
// policy.  The per-digit histograms of all passes are taken while the records are generated
// (in the then idle reorder buffer, parked in global memory during the passes), so each pass
// costs one read + one write of the records.
// DESIGN.md §4 has the measurements behind these choices (what bounds the kernel, what was
// tried and dropped).
#include "common.cuh"
#include "kernels.h"

namespace bnz {
namespace bwt {

// Phase timing of the radix pass for tools/bwt_phase_prof.py (`make prof` builds a second library
// with -DBWT_PHASE_PROF=<thread id>): cycles of thread BWT_PHASE_PROF of every CTA between the
// marks, summed over the launch.  Compiled out of the product library.
#ifdef BWT_PHASE_PROF
__device__ unsigned long long g_prof[8];
#define PROF_DECL long long prof_t[8] = { 0, 0, 0, 0, 0, 0, 0, 0 }, prof_prev = clock64(), prof_now
#define PROF_MARK(i) (prof_now = clock64(), prof_t[i] += prof_now - prof_prev, prof_prev = prof_now)
#define PROF_FLUSH()                                                                               \
    do {                                                                                           \
        if (threadIdx.x == BWT_PHASE_PROF)                                                         \
            for (int i_ = 0; i_ < 8; i_++) atomicAdd(&g_prof[i_], (unsigned long long)prof_t[i_]); \
    } while (0)
#else
#define PROF_DECL
#define PROF_MARK(i)
#define PROF_FLUSH()
#endif

#ifndef BWT_T
#define BWT_T 512
#endif
constexpr int T = BWT_T;             // threads per CTA
constexpr int NW = T / 32;           // warps per CTA
#ifndef BWT_K
#define BWT_K 8
#endif
#ifndef BWT_MINCTA
#define BWT_MINCTA (1024 / BWT_T)
#endif
constexpr int K = BWT_K;             // records per thread per tile
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
struct __align__(128) Smem {
    u64 inbuf[TILE];                                      // TMA landing zone for the next tile
    u64 stage[TILE];                                      // tile reorder buffer (digit histograms while keys are built)
    u32 whist[NW][Cfg<BITS>::BINS / 2];                   // per-warp digit counts / offsets, two 16-bit bins per word
    u32 cursor[Cfg<BITS>::BINS];                          // running bucket cursors (global)
    u32 gbase[Cfg<BITS>::BINS];                           // cursor - binoff
    u16 binoff[Cfg<BITS>::BINS];                          // exclusive bin offsets inside the tile
    u64 scratch64[40];
    u64 mbar;                                             // mbarrier of the TMA tile pipeline
    u32 scratch[40];
    u32 s_count;                                          // records appended by build_*
    u32 s_list;                                           // active-list length appended by rerank
    u32 s_block;                                          // claimed block id
    u32 s_flags[8];                                       // per pass: every record has the same digit
    u8 present[256];                                      // has_byte
};
static_assert(Cfg<10>::PASSES * Cfg<10>::BINS * 4 <= TILE * 8, "histograms must fit the reorder buffer");

// per-pass digit histograms: built in the (then idle) reorder buffer, parked in global memory
template <int BITS>
__device__ __forceinline__ u32 *hist_of(Smem<BITS> &sm) { return reinterpret_cast<u32 *>(sm.stage); }

template <int BITS>
__device__ __forceinline__ u32 digit_of(u64 rec, int pass)
{
    return (u32)(rec >> (IDX_BITS + pass * BITS)) & (u32)(Cfg<BITS>::BINS - 1);
}

template <int BITS>
__device__ __forceinline__ void hist_clear(Smem<BITS> &sm)
{
    u32 *hist = hist_of(sm);
    for (int i = threadIdx.x; i < Cfg<BITS>::PASSES * Cfg<BITS>::BINS; i += T) hist[i] = 0;
    if (threadIdx.x == 0) sm.s_count = 0;
    __syncthreads();
}

template <int BITS>
__device__ __forceinline__ void hist_add(Smem<BITS> &sm, u64 rec)
{
    u32 *hist = hist_of(sm);
#pragma unroll
    for (int p = 0; p < Cfg<BITS>::PASSES; p++) atomicAdd(&hist[p * Cfg<BITS>::BINS + digit_of<BITS>(rec, p)], 1u);
}

// Round 0: key = the five bytes S[i..i+5) (cyclic), big-endian, so h = 5 afterwards.
// S is 16-byte aligned and padded to 16 bytes, so two aligned 32-bit loads cover any 5-byte window.
template <int BITS>
__device__ void build_initial(Smem<BITS> &sm, const u8 *__restrict__ S, u32 n, u64 *dst)
{
    hist_clear(sm);
    for (int i = threadIdx.x; i < 256; i += T) sm.present[i] = 0;
    __syncthreads();
    const u32 *S32 = reinterpret_cast<const u32 *>(S);
    for (u32 base = 0; base < n; base += TILE) {
#pragma unroll
        for (int k = 0; k < K; k++) {
            u32 i = base + k * T + threadIdx.x;
            if (i < n) {
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
                sm.present[(u32)(key >> 32)] = 1;        // first byte = S[i]
                u64 rec = (key << IDX_BITS) | i;
                st_stream(dst + i, rec);
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
        // all loads first (rank[i], then the rank[i+h] gathers) so their latencies overlap
        u32 r[K], r2[K];
#pragma unroll
        for (int k = 0; k < K; k++) {
            u32 i = base + k * T + threadIdx.x;
            r[k] = (i < n) ? ld_keep(rank + i) : DONE;
        }
#pragma unroll
        for (int k = 0; k < K; k++) {
            u32 i = base + k * T + threadIdx.x;
            r2[k] = 0;
            if (!(r[k] & DONE)) {
                u32 j = i + hm;
                if (j >= n) j -= n;
                r2[k] = ld_keep(rank + j);
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
                    st_stream(dst + wbase + __popc(m & lanemask_lt()), rec);
                    hist_add(sm, rec);
                }
            }
        }
    }
    __syncthreads();
}

// warp-wide "which lanes hold my digit": BITS ballots (cheaper than match.any here)
template <int BITS>
__device__ __forceinline__ u32 match_digit(u32 d)
{
    // peers = lanes whose digit equals mine: AND over the digit's bits of XNOR(ballot(bit), my bit).
    // Inline PTX keeps it at and/setp + vote + predicated not + and per bit.
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

// Round h when few rotations are still active: the previous re-rank left the list of active
// (new rank, idx) pairs, so only those are touched instead of scanning all n ranks.
template <int BITS>
__device__ void build_round_list(Smem<BITS> &sm, const u32 *rank, u32 n, u32 h, const u64 *alist, u32 cnt,
                                 u64 *dst)
{
    hist_clear(sm);
    const u32 hm = h % n;
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
                const u64 r1 = e[k] >> IDX_BITS;
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

// After the keys of a round are built: park the per-pass histograms in global memory (the
// reorder buffer they were built in is needed by the passes) and note the passes whose digit is
// the same for every record.
template <int BITS>
__device__ void hist_park(Smem<BITS> &sm, u32 *ghist, u32 count)
{
    constexpr int BINS = Cfg<BITS>::BINS, PASSES = Cfg<BITS>::PASSES;
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

// One LSD pass over `count` records: src -> dst by digit `pass`.
// Tiles are streamed through shared memory with TMA bulk copies (cp.async.bulk + mbarrier):
// the copy of tile t+1 is in flight while tile t is ranked, reordered and stored.
template <int BITS>
__device__ void radix_pass(Smem<BITS> &sm, const u64 *src, u64 *dst, u32 count, int pass, const u32 *ghist,
                           u32 &phase)
{
    constexpr int BINS = Cfg<BITS>::BINS;
    constexpr int BPT = Cfg<BITS>::BPT;
    constexpr int WORDS = BINS / 2;                             // packed counter words per warp row
    constexpr int NSCAN = WORDS < T ? WORDS : T;                // threads that own counter words in the scan
    static_assert(WORDS <= T, "one counter word per scan thread");
    const u32 tid = threadIdx.x, lane = lane_id(), w = warp_id();

    // records were written with generic-proxy stores; order them before the async-proxy reads
    fence_proxy_async();
    __syncthreads();
    if (tid == 0) {
        const u32 bytes = (min((u32)TILE, count) * 8u + 15u) & ~15u;
        mbar_expect_tx(&sm.mbar, bytes);
        tma_load_1d_stream(sm.inbuf, src, bytes, &sm.mbar);
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

    PROF_DECL;
    for (u32 base = 0; base < count; base += TILE) {
        const u32 tile_n = min((u32)TILE, count - base);
        u64 rec[K];
        u32 rk[K];
        {   // zero this warp's counter row with 16-byte stores
            uint4 *row = reinterpret_cast<uint4 *>(sm.whist[w]);
            for (int b = lane; b < WORDS / 4; b += 32) row[b] = make_uint4(0, 0, 0, 0);
        }
        PROF_MARK(0);                                           // loop overhead + counter zeroing
        mbar_wait(&sm.mbar, phase);
        phase ^= 1u;
        PROF_MARK(1);                                           // wait for the tile
        const u32 wl = w * (K * 32) + lane;
        if (tile_n == TILE) {
#pragma unroll
            for (int k = 0; k < K; k++) rec[k] = sm.inbuf[wl + k * 32];
        } else {
#pragma unroll
            for (int k = 0; k < K; k++) {
                u32 j = wl + k * 32;
                rec[k] = (j < tile_n) ? sm.inbuf[j] : ~0ull;
            }
        }
        __syncwarp();
#pragma unroll
        for (int k = 0; k < K; k++) {
            const u32 d = digit_of<BITS>(rec[k], pass);
            const u32 peers = match_digit<BITS>(d);
            const u32 leader = 31 - __clz(peers);
            const u32 sh = (d & 1u) * 16u;
            // one shared atomic per digit group; its return value is the group's base.  Atomics of
            // successive rows to the same counter execute in program order, so rows need no barrier.
            // (a warp holds at most 256 records of a tile, so the 16-bit halves never carry)
            u32 bcount = 0;
            if (lane == leader) bcount = atomicAdd(&sm.whist[w][d >> 1], (u32)__popc(peers) << sh);
            bcount = __shfl_sync(0xffffffffu, bcount, leader);
            rk[k] = ((bcount >> sh) & 0xffffu) + __popc(peers & lanemask_lt());
        }
        PROF_MARK(2);                                           // ranking
        __syncthreads();                                        // B1: inbuf consumed, whist complete
        PROF_MARK(3);                                           // wait at B1

        if (tid == 0 && base + TILE < count) {                  // prefetch the next tile
            const u32 nb = (min((u32)TILE, count - base - TILE) * 8u + 15u) & ~15u;
            // no proxy fence here: every thread's reads of inbuf have returned (their values feed the
            // ranking above) and B1 orders them before this copy; a fence.proxy.async in this spot
            // costs a GPU-scope MEMBAR per tile that stalls the whole CTA behind warp 0
            mbar_expect_tx(&sm.mbar, nb);
            tma_load_1d_stream(sm.inbuf, src + base + TILE, nb, &sm.mbar);
        }

        // cross-warp exclusive scan per bin (two bins per packed word; a tile holds 4096 records, so
        // the halves never carry), then exclusive scan over bins
        if (tid < NSCAN) {
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
            asm volatile("bar.sync 1, %0;" ::"n"(NSCAN) : "memory");
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
        PROF_MARK(4);                                           // scan (B1 -> B2)

#pragma unroll
        for (int k = 0; k < K; k++) {
            const u32 d = digit_of<BITS>(rec[k], pass);
            const u32 pos = sm.binoff[d] + ((sm.whist[w][d >> 1] >> ((d & 1u) * 16u)) & 0xffffu) + rk[k];
            sm.stage[pos] = rec[k];
        }
        __syncthreads();                                        // B3
        PROF_MARK(5);                                           // reorder (B2 -> B3)

        if (tile_n == TILE) {
#pragma unroll
            for (int k = 0; k < K; k++) {
                const u32 j = k * T + tid;
                const u64 r = sm.stage[j];
                st_stream(dst + sm.gbase[digit_of<BITS>(r, pass)] + j, r);
            }
        } else {
#pragma unroll
            for (int k = 0; k < K; k++) {
                const u32 j = k * T + tid;
                if (j < tile_n) {
                    const u64 r = sm.stage[j];
                    st_stream(dst + sm.gbase[digit_of<BITS>(r, pass)] + j, r);
                }
            }
        }
        // no barrier here: the next tile's B1 orders these reads before stage/whist are reused
        PROF_MARK(6);                                           // bucket stores
    }
    PROF_FLUSH();
    __syncthreads();
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
                            u8 *__restrict__ bwt_out, u32 *ptr_out, u64 *alist)
{
    const u32 tid = threadIdx.x;
    if (tid == 0) sm.s_list = 0;
    __syncthreads();
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
        u64 tot2;
        const u64 ex2 = block_excl_max2<T>(((u64)mg << 32) | mk, sm.scratch64, &tot2);
        const u32 ex_k = max((u32)ex2, carry_key);
        const u32 ex_g = max((u32)(ex2 >> 32), carry_grp);
        const u32 tot_k = (u32)tot2, tot_g = (u32)(tot2 >> 32);

        u32 nrv[K];
        u32 flg[K];                          // bit0 valid, bit1 singleton
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
        if (alist) {
            // compact (new rank, idx) of the rotations that stay active (order is irrelevant)
#pragma unroll
            for (int k = 0; k < K; k++) {
                const bool act = (flg[k] & 3u) == 1u;
                const u32 m = __ballot_sync(0xffffffffu, act);
                if (m) {
                    u32 wbase = 0;
                    if (lane_id() == 0) wbase = atomicAdd(&sm.s_list, (u32)__popc(m));
                    wbase = __shfl_sync(0xffffffffu, wbase, 0);
                    if (act) alist[wbase + __popc(m & lanemask_lt())] = ((u64)nrv[k] << IDX_BITS) | idx[k];
                }
            }
        }
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
        carry_key = max(carry_key, tot_k);
        carry_grp = max(carry_grp, tot_g);
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
__global__ void __launch_bounds__(T, BWT_MINCTA) bwt_sort_kernel(BwtArgs a)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    Smem<BITS> &sm = *reinterpret_cast<Smem<BITS> *>(smem_raw);
    constexpr int PASSES = Cfg<BITS>::PASSES;
    const u32 tid = threadIdx.x;

    if (tid == 0) {
        mbar_init(&sm.mbar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    u32 phase = 0;                       // parity of the next TMA completion

    u64 *bufA = a.ws_rec + (size_t)blockIdx.x * 2 * a.ws_stride;
    u64 *bufB = bufA + a.ws_stride;
    u32 *rank = a.ws_rank + (size_t)blockIdx.x * a.ws_stride;

    for (;;) {
        if (tid == 0) sm.s_block = atomicAdd(a.next_block, 1u);
        __syncthreads();
        const u32 qpos = sm.s_block;
        __syncthreads();
        if (qpos >= a.n_blocks) break;
        const u32 blk = a.order ? a.order[qpos] : qpos;

        const u8 *S = a.rle + a.blk_off[blk];
        u8 *bwt_out = a.bwt + a.blk_off[blk];
        const u32 n = a.blk_len[blk];
        u32 *ptr_out = a.ptr + blk;

        u32 rounds = 0;
        u64 sum_active = 0, sum_active_passes = 0;
        long long cyc_build = 0, cyc_radix = 0, cyc_rerank = 0, t0, t1;
        u32 h = 5;
        u32 count = n;
        bool initial = true;
        bool tied = false;
        const u64 *alist = nullptr;           // active list left by the previous re-rank (or null)
        u32 alist_cnt = 0;

        while (count > 0 && rounds < MAX_ROUNDS) {
            t0 = clock64();
            if (initial) build_initial<BITS>(sm, S, n, bufA);
            else if (alist) build_round_list<BITS>(sm, rank, n, h, alist, alist_cnt, bufA);
            else build_round<BITS>(sm, rank, n, h, bufA);
            count = sm.s_count;
            t1 = clock64();
            cyc_build += t1 - t0;

            u32 *ghist = a.ws_hist + (size_t)blockIdx.x * (PASSES * Cfg<BITS>::BINS);
            hist_park<BITS>(sm, ghist, count);
            u64 *src = bufA, *dst = bufB;
            u32 passes_run = 0;
            for (int p = 0; p < PASSES; p++) {
                if (sm.s_flags[p]) continue;               // every record has the same digit: nothing moves
                radix_pass<BITS>(sm, src, dst, count, p, ghist, phase);
                u64 *t = src; src = dst; dst = t;
                passes_run++;
            }
            sum_active += count;
            sum_active_passes += (u64)count * passes_run;
            rounds++;

            t0 = clock64();
            cyc_radix += t0 - t1;
            // few active rotations: let the re-rank leave their list in the tail of the free buffer
            // (the next key build writes < n/8 records at the front of bufA, the list sits at the end)
            u64 *next_list = (!initial && (u64)count * 8 < n) ? dst + (a.ws_stride - count) : nullptr;
            RerankOut ro = rerank<BITS>(sm, src, count, initial, S, n, rank, bwt_out, ptr_out, next_list);
            __syncthreads();
            alist = next_list;
            alist_cnt = sm.s_list;
            cyc_rerank += clock64() - t0;
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

        // BWT bytes of every rotation that was ranked uniquely: bwt[rank[i]] = S[i-1].  rank[] and S
        // are read in index order (coalesced); the one-byte scatter covers the block's whole output
        // within this one short pass, so the sectors fill up in L2 instead of costing a 32-byte
        // DRAM gather per rotation inside the re-rank steps (measured: -5..7 % sort time).
        __threadfence_block();
        __syncthreads();
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
            for (int k = 0; k < 4; k++)
                if (r[k] & DONE) bwt_out[r[k] & RANK_MASK] = (u8)c[k];
        }
        if (tid == 0 && (rank[0] & DONE)) *ptr_out = rank[0] & RANK_MASK;
        if (tid < 256) a.has_byte[(size_t)blk * 256 + tid] = sm.present[tid];
        if (tid == 0 && a.stats) {
            BwtStats st;
            st.n = n;
            st.rounds = rounds;
            st.tied = tied ? 1u : 0u;
            st.pad = 0;
            st.sum_active = sum_active;
            st.sum_active_passes = sum_active_passes;
            st.cyc_build = (u64)cyc_build;
            st.cyc_radix = (u64)cyc_radix;
            st.cyc_rerank = (u64)cyc_rerank;
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

// ---------------------------------------------------------------------------------------
// cost predictor: blocks with many long repeats need many doubling rounds.  Windows are sampled
// by CONTENT (a 4-byte hash decides), so both copies of a repeat are sampled, and a Bloom filter
// in shared memory tells whether the 24-byte window was seen before.
// ---------------------------------------------------------------------------------------
constexpr int PT = 512;
constexpr u32 BLOOM_BITS = 1u << 19;                 // 64 KB
__global__ void __launch_bounds__(PT) bwt_predict_kernel(const u8 *__restrict__ rle, const u64 *__restrict__ blk_off,
                                                        const u32 *__restrict__ blk_len, u32 *__restrict__ score)
{
    extern __shared__ u32 bloom[];
    __shared__ u32 scratch[40];
    const u32 b = blockIdx.x, tid = threadIdx.x;
    const u8 *S = rle + blk_off[b];
    const u32 n = blk_len[b];
    for (u32 i = tid; i < BLOOM_BITS / 32; i += PT) bloom[i] = 0;
    __syncthreads();
    u32 hits = 0;
    // lanes take consecutive positions (coalesced); 4 positions per thread and iteration
    for (u32 base = 0; base + 24 <= n; base += PT * 4) {
        const u32 i0 = base + tid * 4;
        if (i0 + 28 > n) continue;
        // 8 bytes starting at i0 cover the four 4-byte windows i0..i0+3
        u32 lo = 0, hi = 0;
#pragma unroll
        for (int j = 0; j < 4; j++) {
            lo |= (u32)S[i0 + j] << (8 * j);
            hi |= (u32)S[i0 + 4 + j] << (8 * j);
        }
#pragma unroll
        for (int q = 0; q < 4; q++) {
            const u32 w0 = q ? __funnelshift_r(lo, hi, 8 * q) : lo;
            if (((w0 * 2654435761u) >> 28) != 0) continue;      // content-defined 1/16 sampling
            const u32 i = i0 + q;
            u32 h1 = 2166136261u, h2 = 0x9747b28cu;
#pragma unroll
            for (int j = 0; j < 24; j++) {
                u32 c = S[i + j];
                h1 = (h1 ^ c) * 16777619u;
                h2 = (h2 + c) * 0xcc9e2d51u;
                h2 = (h2 << 13) | (h2 >> 19);
            }
            h1 &= BLOOM_BITS - 1;
            h2 &= BLOOM_BITS - 1;
            u32 o1 = atomicOr(&bloom[h1 >> 5], 1u << (h1 & 31));
            u32 o2 = atomicOr(&bloom[h2 >> 5], 1u << (h2 & 31));
            if (((o1 >> (h1 & 31)) & 1u) && ((o2 >> (h2 & 31)) & 1u)) hits++;
        }
    }
    u32 tot = block_sum<PT>(hits, scratch);
    if (tid == 0) score[b] = tot;
}

}  // namespace bwt

#ifdef BWT_PHASE_PROF
extern "C" __attribute__((visibility("default"))) void bnz_prof_read(unsigned long long *out)
{
    unsigned long long z[8] = { 0 };
    cudaMemcpyFromSymbol(out, bwt::g_prof, sizeof z);
    cudaMemcpyToSymbol(bwt::g_prof, z, sizeof z);
}
#endif

cudaError_t bwt_predict_launch(const uint8_t *d_rle, const uint64_t *d_blk_off, const uint32_t *d_blk_len,
                               uint32_t n_blocks, uint32_t *d_score, cudaStream_t stream)
{
    if (n_blocks == 0) return cudaSuccess;
    size_t smem = bwt::BLOOM_BITS / 8;
    cudaError_t e = cudaFuncSetAttribute(bwt::bwt_predict_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    bwt::bwt_predict_kernel<<<n_blocks, bwt::PT, smem, stream>>>(d_rle, d_blk_off, d_blk_len, d_score);
    return cudaGetLastError();
}

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
