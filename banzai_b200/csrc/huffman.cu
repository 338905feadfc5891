// huffman.cu — K6 (table modelling), K7 (code-length construction), K8 (serialisation and
// bit packing).  Replaces huffman::encode (reference lib/huffman.rs:313-575), the framing
// helpers write_block_header / write_sym_map (lib/lib.rs:24-64) and the bit writer
// (lib/out.rs) for a batch of blocks.
//
// Reference behaviour reproduced exactly:
//   Q8   table count from the ALPHABET size: num_syms <= 199 -> 2, else 3 (4..6 unreachable but
//        implemented)                                              huffman.rs:319-326
//   Q9   initial tables: contiguous symbol ranges of ~equal mass, length 15 inside the range and
//        0 outside (sic), odd interior tables give back one symbol  huffman.rs:333-376
//   Q10  iteration 0 assigns every 50-symbol group to the cheapest table (strict <, lowest
//        index wins) and adds the group histogram to table_freqs[best]; iterations 1..3 first
//        zero the code-length tables (huffman.rs:403-409), so every cost is 0, every group goes
//        to table 0 and table_freqs[0] grows by the whole-block histogram each time;
//        table_freqs is never cleared.  Hence after the 4 iterations
//            table_freqs[0] = A_0 + 3 G,  table_freqs[t>0] = A_t,  all selectors = 0,
//        which is what huff_build_kernel folds in (the tables built after iterations 0..2 are
//        dead: they are zeroed before use).  tests/test_oracle.py proves this closed form
//        against the literal loop of the oracle.  With bnz_ctx_set("huff_literal", 1) the device
//        runs the reference's loop literally instead (huff_launch_literal: four rounds of
//        per-group cost / strict-< argmin over the T tables, table_freqs accumulation, table
//        rebuild, selectors recorded in the last round, symbols coded with selectors[i / 50]);
//        tests/test_mtf_huff_gpu.py checks that both give the oracle's bits.
//   Q11  code lengths: the reference's own binary heap with its tie behaviour, priorities
//        (weight, depth), scaling retry until max length <= 17      huffman.rs:161-298
//   Q12/13 serialisation order and MSB-first packing               huffman.rs:462-575, out.rs
#include "common.cuh"
#include "kernels.h"

namespace bnz {
namespace huff {

constexpr int MAXS = HUFF_MAX_SYMS;      // 258
constexpr int MAXT = HUFF_MAX_TABLES;    // 6
constexpr int GROUP = 50;                // huffman.rs:310 SEGMENT_WIDTH
constexpr int MAXLEN = 17;               // huffman.rs:13

__device__ __forceinline__ u32 table_count(u32 num_syms)     // huffman.rs:319-326
{
    if (num_syms <= 199) return 2;
    if (num_syms <= 599) return 3;
    if (num_syms <= 1199) return 4;
    if (num_syms <= 2399) return 5;
    return 6;
}

// ------------------------------------------------------------------ H1: initial tables
__global__ void huff_init_kernel(HuffArgs a)
{
    const u32 b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= a.n_blocks) return;
    const u32 num_syms = a.num_names[b] + 2;
    const u32 m = a.sym_len[b];
    const u32 *freqs = a.freqs + (size_t)b * MAXS;
    const u32 T = table_count(num_syms);
    u8 *lens = a.lens + (size_t)b * MAXT * MAXS;
    u32 freq_remaining = m, sym_left = 0;
    for (u32 cur = 0; cur < T; cur++) {
        u32 freq_target = freq_remaining / (T - cur);
        u32 acc = 0, right = sym_left;
        for (;;) {
            acc += (right < MAXS) ? freqs[right] : 0;     // index >= 258 is the reference's latent panic (A-Q14)
            if (acc >= freq_target || right + 1 == num_syms || right >= MAXS) break;
            right++;
        }
        if (right > sym_left && cur != 0 && cur != T - 1 && (cur & 1)) {
            acc -= freqs[right];
            right--;
        }
        for (u32 s = 0; s < MAXS; s++) lens[cur * MAXS + s] = (s < num_syms && s >= sym_left && s <= right) ? 15 : 0;
        sym_left = right + 1;
        freq_remaining -= acc;
    }
    a.num_tables[b] = T;
    a.num_sel[b] = (m + GROUP - 1) / GROUP;
}

// ------------------------------------------------------------------ H2: group assignment (iteration 0)
constexpr int GT = 256;                  // threads per CTA
constexpr int GPC = 256;                 // groups per CTA
__global__ void __launch_bounds__(GT) huff_assign_kernel(HuffArgs a)
{
    __shared__ u8 lens[MAXT * MAXS];
    __shared__ u32 hist[MAXT * MAXS];
    const u32 tid = threadIdx.x, lane = lane_id(), w = warp_id();
    // which block / span
    u32 lo = 0, hi = a.n_blocks;
    while (hi - lo > 1) {
        u32 mid = (lo + hi) >> 1;
        if (a.span_base[mid] <= blockIdx.x) lo = mid; else hi = mid;
    }
    const u32 b = lo;
    const u32 span = blockIdx.x - a.span_base[b];
    const u32 T = a.num_tables[b];
    const u32 m = a.sym_len[b];
    const u16 *syms = a.syms + a.sym_off[b];
    for (u32 i = tid; i < T * MAXS; i += GT) {
        lens[i] = a.lens[(size_t)b * MAXT * MAXS + i];
        hist[i] = 0;
    }
    __syncthreads();
    const u32 ngroups = (m + GROUP - 1) / GROUP;
    const u32 g0 = span * GPC, g1 = min(g0 + GPC, ngroups);
    for (u32 g = g0 + w; g < g1; g += GT / 32) {
        const u32 base = g * GROUP;
        const u32 glen = min((u32)GROUP, m - base);
        u32 s1 = (lane < glen) ? syms[base + lane] : 0xffffu;
        u32 s2 = (lane + 32 < glen) ? syms[base + 32 + lane] : 0xffffu;
        u32 best = 0, best_cost = 0xffffffffu;
        for (u32 t = 0; t < T; t++) {
            u32 c = (s1 != 0xffffu ? lens[t * MAXS + s1] : 0) + (s2 != 0xffffu ? lens[t * MAXS + s2] : 0);
            c = __reduce_add_sync(0xffffffffu, c);
            if (c < best_cost) { best = t; best_cost = c; }
        }
        if (s1 != 0xffffu) atomicAdd(&hist[best * MAXS + s1], 1u);
        if (s2 != 0xffffu) atomicAdd(&hist[best * MAXS + s2], 1u);
        if (a.sel_out && lane == 0) a.sel_out[(size_t)b * a.sel_stride + g] = (u8)best;     // huffman.rs:446-448
    }
    __syncthreads();
    u32 *tf = a.tf + (size_t)b * MAXT * MAXS;
    for (u32 i = tid; i < T * MAXS; i += GT)
        if (hist[i]) atomicAdd(&tf[i], hist[i]);
}

// ------------------------------------------------------------------ H3/H4: table_freqs closed form + code lengths + codes
// one warp per (block, table); lane 0 replays the reference's heap literally.
struct HeapMem {
    u64 prio[MAXS + 2];        // (weight << 8) | depth : lexicographic order == integer order
    u16 id[MAXS + 2];
    u16 parent[2 * MAXS + 2];
    u8 depth[2 * MAXS + 2];
};

__device__ __forceinline__ void heap_insert(HeapMem &h, u32 &len, u16 id, u64 pr)    // huffman.rs:196-222
{
    u32 init_idx = len + 1;
    h.prio[len] = pr;
    h.id[len] = id;
    len++;
    if (init_idx == 1) return;
    u32 idx = init_idx;
    for (;;) {
        u32 above = idx >> 1;
        u64 ap = h.prio[above - 1];
        if (pr < ap) {
            h.prio[idx - 1] = ap;
            h.id[idx - 1] = h.id[above - 1];
            idx = above;
            if (idx == 1) break;
        } else break;
    }
    if (idx != init_idx) {
        h.prio[idx - 1] = pr;
        h.id[idx - 1] = id;
    }
}

__device__ __forceinline__ void heap_extract(HeapMem &h, u32 &len, u16 &id_out, u64 &pr_out)   // huffman.rs:225-267
{
    u64 lp = h.prio[len - 1];
    u16 lid = h.id[len - 1];
    len--;
    if (len == 0) { id_out = lid; pr_out = lp; return; }
    id_out = h.id[0];
    pr_out = h.prio[0];
    u32 idx = 1;
    for (;;) {
        u32 left = idx << 1;
        if (left > len) break;
        u32 right = left + 1;
        u32 below = left;
        u64 bp = h.prio[left - 1];
        if (right <= len) {
            u64 rp = h.prio[right - 1];
            if (rp < bp) { below = right; bp = rp; }
        }
        if (lp < bp) break;
        h.prio[idx - 1] = bp;
        h.id[idx - 1] = h.id[below - 1];
        idx = below;
    }
    h.prio[idx - 1] = lp;
    h.id[idx - 1] = lid;
}

// The reference retries the whole build with the weights halved until no code is longer than 17
// bits (huffman.rs:271-298); large blocks need several tries.  The tries are independent, so
// NTRY lanes replay the heap for scaling 1, 2, 4, ... at the same time (each in its own heap) and
// the lowest scaling that fits wins — the same table the sequential retry loop ends with.
constexpr int BW = 1;            // warps (jobs) per CTA
constexpr int NTRY = 8;          // scalings tried at once
__global__ void __launch_bounds__(BW * 32) huff_build_kernel(HuffArgs a)
{
    __shared__ HeapMem mem[BW][NTRY];
    __shared__ u32 fr[BW][MAXS];
    const u32 w = warp_id(), lane = lane_id();
    const u32 job = blockIdx.x * BW + w;
    if (job >= a.n_blocks * MAXT) return;
    const u32 b = job / MAXT, t = job % MAXT;
    if (t >= a.num_tables[b]) return;
    const u32 num_syms = a.num_names[b] + 2;
    u32 *tf = a.tf + ((size_t)b * MAXT + t) * MAXS;
    const u32 *G = a.freqs + (size_t)b * MAXS;
    // iterations 1..3 (zeroed tables): every group adds its histogram to table 0
    for (u32 s = lane; s < MAXS; s += 32) {
        u32 v = tf[s];
        if (t == 0 && !a.literal) v += (HUFF_REFINEMENTS - 1) * G[s];
        tf[s] = v;
        fr[w][s] = v;
    }
    __syncwarp();
    u8 *lens = a.lens + ((size_t)b * MAXT + t) * MAXS;
    u32 *codes = a.codes + ((size_t)b * MAXT + t) * MAXS;

    u32 winner = 0;
    for (u32 round = 0;; round++) {
        bool fits = false;
        if (lane < NTRY) {
            HeapMem &h = mem[w][lane];
            const u32 shift = round * NTRY + lane;           // scaling = 2^shift
            u32 len = 0;
            for (u32 s = 0; s < num_syms; s++)
                heap_insert(h, len, (u16)(s + 1), ((u64)((shift < 32 ? fr[w][s] >> shift : 0u) + 1)) << 8);
            u32 nodes = num_syms + 1;                    // root (0) + leaves (1..n)
            for (;;) {
                u16 i1, i2;
                u64 p1, p2;
                heap_extract(h, len, i1, p1);
                heap_extract(h, len, i2, p2);
                u32 par;
                if (nodes == num_syms * 2 - 1) par = 0;  // Tree::tie, huffman.rs:60-74
                else par = nodes++;
                h.parent[i1] = (u16)par;
                h.parent[i2] = (u16)par;
                if (par == 0) break;
                u64 d1 = p1 & 0xff, d2 = p2 & 0xff;
                u64 np = (((p1 >> 8) + (p2 >> 8)) << 8) | ((d1 > d2 ? d1 : d2) + 1);
                heap_insert(h, len, (u16)par, np);
            }
            // depths: parents are created after their children, the root last
            h.depth[0] = 0;
            u32 maxd = 0;
            for (u32 idn = num_syms * 2 - 2; idn >= 1; idn--) {
                u32 d = h.depth[h.parent[idn]] + 1;
                h.depth[idn] = (u8)d;
                if (idn <= num_syms && d > maxd) maxd = d;
            }
            fits = maxd <= MAXLEN;
        }
        const u32 ok = __ballot_sync(0xffffffffu, fits);
        if (ok) {
            winner = __ffs(ok) - 1;                          // the smallest scaling that fits
            break;
        }
    }
    __syncwarp();
    const HeapMem &h = mem[w][winner];
    for (u32 s = lane; s < MAXS; s += 32) lens[s] = (s < num_syms) ? h.depth[s + 1] : (u8)0;
    if (lane == 0) {
        // canonical codes (huffman.rs:550-561): per length, symbols in ascending order get consecutive
        // words; the word doubles from one length to the next
        u32 cnt[MAXLEN + 2], first[MAXLEN + 2];
        for (u32 l = 0; l <= MAXLEN + 1; l++) cnt[l] = 0;
        u32 minl = 255, maxl = 0;
        for (u32 s = 0; s < num_syms; s++) {
            const u32 l = h.depth[s + 1];
            cnt[l]++;
            minl = min(minl, l);
            maxl = max(maxl, l);
        }
        u32 word = 0;
        for (u32 l = minl; l <= maxl; l++) {
            first[l] = word;
            word = (word + cnt[l]) << 1;
            cnt[l] = 0;
        }
        for (u32 s = 0; s < num_syms; s++) {
            const u32 l = h.depth[s + 1];
            codes[s] = (l << 24) | (first[l] + cnt[l]++);
        }
    }
}

// ------------------------------------------------------------------ H5: per-block header bits
struct BitW {
    u32 *words;
    u64 acc;
    u32 nacc, wpos;
    u64 total;
    __device__ __forceinline__ void put(u32 val, u32 nbits)
    {
        acc = (acc << nbits) | val;
        nacc += nbits;
        total += nbits;
        if (nacc >= 32) {
            words[wpos++] = (u32)(acc >> (nacc - 32));
            nacc -= 32;
        }
    }
    __device__ __forceinline__ void zeros(u32 nbits)
    {
        while (nbits >= 32) { put(0, 16); put(0, 16); nbits -= 32; }
        if (nbits > 16) { put(0, 16); nbits -= 16; }
        if (nbits) put(0, nbits);
    }
    __device__ __forceinline__ void flush()
    {
        if (nacc) words[wpos++] = (u32)(acc << (32 - nacc));
    }
};

__global__ void huff_header_kernel(HuffArgs a)
{
    const u32 b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= a.n_blocks) return;
    const u32 num_syms = a.num_names[b] + 2;
    const u32 T = a.num_tables[b];
    const u32 S = a.num_sel[b];
    BitW w;
    w.words = a.hdr + (size_t)b * a.hdr_stride;
    w.acc = 0;
    w.nacc = 0;
    w.wpos = 0;
    w.total = 0;

    if (a.with_block_header) {
        // write_block_header, lib.rs:24-36
        w.put(0x314159u, 24);
        w.put(0x265359u, 24);
        const u32 crc = a.crc[b];
        w.put(crc >> 16, 16);
        w.put(crc & 0xffffu, 16);
        w.put(0, 1);
        w.put(a.ptr[b] & 0xffffffu, 24);
        // write_sym_map, lib.rs:39-64
        const u8 *has = a.has_byte + (size_t)b * 256;
        u32 sector_map = 0, sectors[16], ns = 0;
        for (u32 s = 0; s < 16; s++) {
            u32 sec = 0;
            for (u32 k = 0; k < 16; k++) sec = (sec << 1) | (has[s * 16 + k] ? 1u : 0u);
            sector_map <<= 1;
            if (sec) { sector_map |= 1; sectors[ns++] = sec; }
        }
        w.put(sector_map, 16);
        for (u32 k = 0; k < ns; k++) w.put(sectors[k], 16);
    }

    // huffman.rs:465-469
    w.put(T, 3);
    w.put(S, 15);
    // selectors, MTF + unary (huffman.rs:472-503)
    const u8 *sel = a.selectors ? a.selectors + (size_t)b * a.sel_stride : nullptr;
    if (!sel) {
        w.zeros(S);                                     // every selector is 0 (A-Q10): one '0' bit each
    } else {
        u32 mtfl[MAXT];
        for (u32 i = 0; i < MAXT; i++) mtfl[i] = i;
        for (u32 k = 0; k < S; k++) {
            u32 sv = sel[k];
            u32 bump = mtfl[0];
            if (bump == sv) w.put(0, 1);
            else {
                u32 idx = 1;
                for (;;) {
                    u32 st = mtfl[idx];
                    mtfl[idx] = bump;
                    if (st == sv) { w.put((1u << (idx + 1)) - 2, idx + 1); break; }
                    bump = st;
                    idx++;
                }
                mtfl[0] = sv;
            }
        }
    }
    // delta-coded tables (huffman.rs:509-545)
    for (u32 t = 0; t < T; t++) {
        const u8 *table = a.lens + ((size_t)b * MAXT + t) * MAXS;
        u32 acc = table[0];
        w.put(acc, 5);
        for (u32 s = 0; s < num_syms; s++) {
            u32 l = table[s];
            while (l != acc) {
                if (l > acc) { w.put(2, 2); acc++; }
                else { w.put(3, 2); acc--; }
            }
            w.put(0, 1);
        }
    }
    w.flush();
    a.hdr_bits[b] = (u32)w.total;

    // symbol bits: all groups use table 0 in the final iteration (A-Q10)
    u64 sym_bits = 0;
    if (!sel) {
        const u32 *G = a.freqs + (size_t)b * MAXS;
        const u8 *t0 = a.lens + (size_t)b * MAXT * MAXS;
        for (u32 s = 0; s < num_syms; s++) sym_bits += (u64)G[s] * t0[s];
    }
    a.blk_bits[b] = w.total + sym_bits;
}

// symbol bits of every block when groups use different tables: sum over the groups of the code
// lengths of their symbols in the table selectors[g] names (huffman.rs:565-572); added to blk_bits
__global__ void __launch_bounds__(GT) huff_symbits_kernel(HuffArgs a)
{
    __shared__ u8 lens[MAXT * MAXS];
    __shared__ unsigned long long total;
    const u32 tid = threadIdx.x, lane = lane_id(), w = warp_id();
    u32 lo = 0, hi = a.n_blocks;
    while (hi - lo > 1) {
        u32 mid = (lo + hi) >> 1;
        if (a.span_base[mid] <= blockIdx.x) lo = mid; else hi = mid;
    }
    const u32 b = lo;
    const u32 span = blockIdx.x - a.span_base[b];
    const u32 T = a.num_tables[b];
    const u32 m = a.sym_len[b];
    const u16 *syms = a.syms + a.sym_off[b];
    const u8 *sel = a.selectors + (size_t)b * a.sel_stride;
    for (u32 i = tid; i < T * MAXS; i += GT) lens[i] = a.lens[(size_t)b * MAXT * MAXS + i];
    if (tid == 0) total = 0;
    __syncthreads();
    const u32 ngroups = (m + GROUP - 1) / GROUP;
    const u32 g0 = span * GPC, g1 = min(g0 + GPC, ngroups);
    u32 mine = 0;
    for (u32 g = g0 + w; g < g1; g += GT / 32) {
        const u32 base = g * GROUP;
        const u32 glen = min((u32)GROUP, m - base);
        const u32 t = sel[g];
        if (lane < glen) mine += lens[t * MAXS + syms[base + lane]];
        if (lane + 32 < glen) mine += lens[t * MAXS + syms[base + 32 + lane]];
    }
    mine = __reduce_add_sync(0xffffffffu, mine);
    if (lane == 0 && mine) atomicAdd(&total, (unsigned long long)mine);
    __syncthreads();
    if (tid == 0 && total) atomicAdd(reinterpret_cast<unsigned long long *>(a.blk_bits + b), total);
}

// exclusive scan of blk_bits -> blk_bitoff (+ base); single CTA
__global__ void __launch_bounds__(1024) huff_scan_kernel(HuffArgs a)
{
    __shared__ u64 sh[40];
    u64 carry = a.bit_base;
    const u32 lane = lane_id(), w = warp_id();
    for (u32 base = 0; base < a.n_blocks; base += 1024) {
        u32 b = base + threadIdx.x;
        u64 v = (b < a.n_blocks) ? a.blk_bits[b] : 0;
        u64 inc = warp_incl_sum64(v);
        if (lane == 31) sh[w] = inc;
        __syncthreads();
        if (w == 0) {
            u64 x = sh[lane];
            u64 xi = warp_incl_sum64(x);
            sh[lane] = xi - x;
            if (lane == 31) sh[32] = xi;
        }
        __syncthreads();
        if (b < a.n_blocks)
            a.blk_bitoff[b] = a.fixed_stride_bits ? (u64)b * a.fixed_stride_bits : carry + sh[w] + inc - v;
        carry += sh[32];
        __syncthreads();
    }
    if (threadIdx.x == 0) *a.total_bits = carry;
}

// ------------------------------------------------------------------ H6: bit packing
constexpr int PT = 512;
constexpr int PK = 8;
constexpr int PTILE = PT * PK;                         // 4096 symbols
constexpr int PWORDS = PTILE * MAXLEN / 32 + 4;        // tile bit buffer

__device__ __forceinline__ u32 bswap32(u32 x) { return __byte_perm(x, 0, 0x0123); }

__global__ void __launch_bounds__(PT) huff_pack_kernel(HuffArgs a)
{
    __shared__ u32 codes[MAXT * MAXS];
    __shared__ u32 buf[PWORDS];
    __shared__ u32 scratch[40];
    const u32 tid = threadIdx.x;
    const u32 b = blockIdx.x;
    const u32 T = a.num_tables[b];
    const u32 m = a.sym_len[b];
    const u16 *syms = a.syms + a.sym_off[b];
    const u8 *sel = a.selectors ? a.selectors + (size_t)b * a.sel_stride : nullptr;
    u32 *out = a.out_words;
    for (u32 i = tid; i < T * MAXS; i += PT) codes[i] = a.codes[(size_t)b * MAXT * MAXS + i];

    // global bit cursor; `first_word` = first global word this block touches
    u64 q = a.blk_bitoff[b];
    const u64 first_word = q >> 5;
    // ---- header: copy hdr bits to the stream at bit q
    {
        const u32 hb = a.hdr_bits[b];
        const u32 *hw = a.hdr + (size_t)b * a.hdr_stride;
        const u32 nw = (hb + 31) / 32;
        const u32 sh = (u32)(q & 31);
        // output word j (relative to first_word) = hw[j-1] << (32-sh) | hw[j] >> sh
        const u32 now = (u32)((sh + hb + 31) / 32);
        for (u32 j = tid; j < now; j += PT) {
            u32 hi = (j >= 1 && j - 1 < nw) ? hw[j - 1] : 0;
            u32 lo = (j < nw) ? hw[j] : 0;
            u32 v = sh ? ((hi << (32 - sh)) | (lo >> sh)) : lo;
            // mask bits beyond the header end (hdr words are zero padded already)
            const bool shared_word = (j == 0 && sh != 0) || (j == now - 1 && ((sh + hb) & 31) != 0);
            if (v) {
                if (shared_word) atomicOr(&out[first_word + j], bswap32(v));
                else out[first_word + j] = bswap32(v);
            } else if (!shared_word) out[first_word + j] = 0;
        }
        q += hb;
    }
    __syncthreads();

    // ---- symbols
    for (u32 base = 0; base < m; base += PTILE) {
        for (u32 i = tid; i < PWORDS; i += PT) buf[i] = 0;
        const u32 j0 = base + tid * PK;
        u32 cw[PK];
        u32 bits = 0;
#pragma unroll
        for (int k = 0; k < PK; k++) {
            u32 j = j0 + k;
            cw[k] = 0;
            if (j < m) {
                u32 t = sel ? sel[j / GROUP] : 0;
                cw[k] = codes[t * MAXS + syms[j]];
                bits += cw[k] >> 24;
            }
        }
        u32 tot;
        u32 ex = block_excl_sum<PT>(bits, scratch, &tot);     // barriers inside: buf is zeroed
        const u32 sh = (u32)(q & 31);
        u32 pos = sh + ex;                                    // bit position inside buf
#pragma unroll
        for (int k = 0; k < PK; k++) {
            u32 len = cw[k] >> 24;
            if (len) {
                u32 code = cw[k] & 0xffffffu;
                u32 wi = pos >> 5, bo = pos & 31;
                // place `len` bits so that the first bit lands at bit (31 - bo) of word wi
                u64 v = (u64)code << (64 - len - bo);
                atomicOr(&buf[wi], (u32)(v >> 32));
                if (bo + len > 32) atomicOr(&buf[wi + 1], (u32)v);
                pos += len;
            }
        }
        __syncthreads();
        const u64 w0 = q >> 5;
        const u32 endbit = sh + tot;
        const u32 nwords = (endbit + 31) / 32;
        for (u32 j = tid; j < nwords; j += PT) {
            const bool partial_last = (j == nwords - 1) && (endbit & 31);
            const bool partial_first = (j == 0) && sh != 0;
            u32 v = buf[j];
            if (partial_first || partial_last) { if (v) atomicOr(&out[w0 + j], bswap32(v)); }
            else out[w0 + j] = bswap32(v);
        }
        q += tot;
        __syncthreads();
    }
}

}  // namespace huff

cudaError_t huff_launch(const HuffArgs &a, uint32_t total_spans, cudaStream_t st, uint32_t *launches)
{
    if (a.n_blocks == 0) return cudaSuccess;
    huff::huff_init_kernel<<<(a.n_blocks + 127) / 128, 128, 0, st>>>(a);
    huff::huff_assign_kernel<<<total_spans, huff::GT, 0, st>>>(a);
    unsigned jobs = a.n_blocks * huff::MAXT;
    huff::huff_build_kernel<<<(jobs + huff::BW - 1) / huff::BW, huff::BW * 32, 0, st>>>(a);
    huff::huff_header_kernel<<<(a.n_blocks + 31) / 32, 32, 0, st>>>(a);
    huff::huff_scan_kernel<<<1, 1024, 0, st>>>(a);
    if (launches) *launches += 5;
    return cudaGetLastError();
}

// huffman::encode's modelling loop run literally (huffman.rs:399-460): HUFF_REFINEMENTS rounds of
// { zero the tables unless it is the first round; assign every group to its cheapest table and add
// its histogram to that table's frequencies; rebuild all tables }, selectors recorded in the last
// round.  `a.selectors` / `a.sel_out` must point to [n_blocks][a.sel_stride] bytes.
cudaError_t huff_launch_literal(HuffArgs a, uint32_t total_spans, cudaStream_t st, uint32_t *launches)
{
    if (a.n_blocks == 0) return cudaSuccess;
    uint8_t *sel = a.sel_out;
    const unsigned jobs = a.n_blocks * huff::MAXT;
    a.literal = 1;
    a.selectors = nullptr;
    huff::huff_init_kernel<<<(a.n_blocks + 127) / 128, 128, 0, st>>>(a);
    for (int it = 0; it < HUFF_REFINEMENTS; it++) {
        if (it != 0) {
            cudaError_t e = cudaMemsetAsync(a.lens, 0, (size_t)a.n_blocks * huff::MAXT * huff::MAXS, st);
            if (e != cudaSuccess) return e;
        }
        a.sel_out = (it == HUFF_REFINEMENTS - 1) ? sel : nullptr;
        huff::huff_assign_kernel<<<total_spans, huff::GT, 0, st>>>(a);
        huff::huff_build_kernel<<<(jobs + huff::BW - 1) / huff::BW, huff::BW * 32, 0, st>>>(a);
    }
    a.sel_out = nullptr;
    a.selectors = sel;
    huff::huff_header_kernel<<<(a.n_blocks + 31) / 32, 32, 0, st>>>(a);
    huff::huff_symbits_kernel<<<total_spans, huff::GT, 0, st>>>(a);
    huff::huff_scan_kernel<<<1, 1024, 0, st>>>(a);
    if (launches) *launches += 4 + 2 * HUFF_REFINEMENTS;
    return cudaGetLastError();
}

cudaError_t huff_pack_launch(const HuffArgs &a, cudaStream_t st, uint32_t *launches)
{
    if (a.n_blocks == 0) return cudaSuccess;
    huff::huff_pack_kernel<<<a.n_blocks, huff::PT, 0, st>>>(a);
    if (launches) *launches += 1;
    return cudaGetLastError();
}

cudaError_t huff_rescan_launch(const HuffArgs &a, cudaStream_t st, uint32_t *launches)
{
    if (a.n_blocks == 0) return cudaSuccess;
    huff::huff_scan_kernel<<<1, 1024, 0, st>>>(a);
    if (launches) *launches += 1;
    return cudaGetLastError();
}

uint32_t huff_groups_per_span() { return huff::GPC; }

}  // namespace bnz
