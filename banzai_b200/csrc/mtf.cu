// mtf.cu — K5: move-to-front + RUNA/RUNB zero-run coding + symbol histogram for a batch of
// BWT blocks.  Replaces mtf::mtf_and_rle (reference lib/mtf.rs:14-121).
//
// MTF is sequential through its 256-entry recency list, so each block is cut into SEG-byte
// segments and the list is made composable:
//   M1  summary : per segment, the distinct bytes in order of LAST occurrence (newest first),
//                 found by one backward walk with a 256-bit seen-set.
//   M2  compose : per block, one warp folds the summaries left to right:
//                 list' = summary ++ (list \ summary); the list each segment starts from is
//                 stored (256 B per segment).  The initial list is the present bytes in
//                 ascending order (lib/mtf.rs:17-24,40-43: names are order preserving, so
//                 ranking raw bytes instead of names gives identical indices).
//   M3  apply   : per segment, one thread derives the indices the list shuffle of lib/mtf.rs:86-100
//                 would give, without shuffling a list: index = number of bytes whose last
//                 occurrence is more recent than that of the current byte (a popcount over a
//                 position mask); one MTF index byte per position.
//   M4  rle2    : per block, a CTA turns index bytes into symbols: index r >= 1 -> r + 1;
//                 a maximal zero run of length z -> the bits of z + 1 below its top bit,
//                 LSB first, as RUNA(0)/RUNB(1) (lib/mtf.rs:46-65); EOB = names + 1 closes
//                 the block (:112-113).  Offsets come from block-wide scans; the symbol
//                 histogram freqs[258] is accumulated on the way.
#include "common.cuh"
#include "kernels.h"

namespace bnz {
namespace mtf {

constexpr int SEG = MTF_SEG;          // bytes per segment
constexpr int NT1 = 128;              // threads per CTA in M1 / M3

__device__ __forceinline__ u32 find_block(const u32 *__restrict__ seg_base, u32 n_blocks, u32 seg)
{
    u32 lo = 0, hi = n_blocks;        // seg_base[lo] <= seg < seg_base[hi]
    while (hi - lo > 1) {
        u32 mid = (lo + hi) >> 1;
        if (seg_base[mid] <= seg) lo = mid; else hi = mid;
    }
    return lo;
}

// ------------------------------------------------------------------ M1: segment summaries
__global__ void __launch_bounds__(NT1) mtf_summary_kernel(MtfArgs a)
{
    __shared__ u32 seen[8][NT1];      // seen[k][tid]: conflict-free per-thread 256-bit set
    const u32 tid = threadIdx.x;
    const u32 segc = blockIdx.x * NT1 + tid;             // position in this launch's segment list
    if (segc >= a.total_segs) return;
    const u32 bc = find_block(a.cseg_base, a.n_blocks, segc);
    const u32 s = segc - a.cseg_base[bc];
    const u32 b = a.ids[bc];
    const u32 seg = a.seg_base[b] + s;
    const u32 n = a.blk_len[b];
    const u8 *src = a.bwt + a.blk_off[b];
    const u32 start = s * SEG, end = min(start + SEG, n);
#pragma unroll
    for (int k = 0; k < 8; k++) seen[k][tid] = 0;
    u8 *list = a.seg_list + (size_t)seg * 256;
    u32 cnt = 0;
    u32 p = end;
    // unaligned tail first
    while (p > start && (p & 15u)) {
        p--;
        u32 c = src[p];
        u32 wd = seen[c >> 5][tid], bit = 1u << (c & 31);
        if (!(wd & bit)) { seen[c >> 5][tid] = wd | bit; list[cnt++] = (u8)c; }
    }
    while (p > start) {
        p -= 16;
        uint4 v = *reinterpret_cast<const uint4 *>(src + p);
        u32 w[4] = { v.x, v.y, v.z, v.w };
#pragma unroll
        for (int j = 15; j >= 0; j--) {
            u32 c = (w[j >> 2] >> ((j & 3) * 8)) & 0xffu;
            u32 wd = seen[c >> 5][tid], bit = 1u << (c & 31);
            if (!(wd & bit)) { seen[c >> 5][tid] = wd | bit; list[cnt++] = (u8)c; }
        }
    }
    a.seg_cnt[seg] = cnt;
}

// ------------------------------------------------------------------ M2: compose per block
constexpr int W2 = 4;                 // warps (blocks) per CTA
__global__ void __launch_bounds__(W2 * 32) mtf_compose_kernel(MtfArgs a)
{
    __shared__ __align__(8) u8 L[W2][2][256];
    __shared__ u32 mask[W2][8];
    const u32 w = warp_id(), lane = lane_id();
    const u32 bc = blockIdx.x * W2 + w;
    if (bc >= a.n_blocks) return;
    const u32 b = a.ids[bc];
    const u8 *has = a.has_byte + (size_t)b * 256;

    // initial list: present bytes ascending
    u32 nn;
    {
        u32 flags = 0;
#pragma unroll
        for (int k = 0; k < 8; k++) if (has[lane * 8 + k]) flags |= 1u << k;
        u32 c = __popc(flags);
        u32 inc = warp_incl_sum(c);
        u32 pos = inc - c;
#pragma unroll
        for (int k = 0; k < 8; k++) if (flags & (1u << k)) L[w][0][pos++] = (u8)(lane * 8 + k);
        nn = __shfl_sync(0xffffffffu, inc, 31);
        if (lane == 0) a.num_names[b] = nn;
    }
    __syncwarp();
    int cur = 0;
    const u32 nseg = a.seg_base[b + 1] - a.seg_base[b];
    const u32 seg_first = a.seg_base[b];
    // the summary of segment s is fetched one step ahead, so that a step costs shared-memory work
    // only (the dependent global load was most of its latency)
    u32 cnt_n = 0;
    uint2 sv_n = make_uint2(0, 0);
    if (nseg > 1) {
        cnt_n = a.seg_cnt[seg_first];
        sv_n = *reinterpret_cast<const uint2 *>(a.seg_list + (size_t)seg_first * 256 + lane * 8);
    }
    for (u32 s = 0; s < nseg; s++) {
        const u32 seg = seg_first + s;
        // the list this segment starts from
        *reinterpret_cast<uint2 *>(a.seg_state + (size_t)seg * 256 + lane * 8) =
            *reinterpret_cast<const uint2 *>(&L[w][cur][lane * 8]);
        if (s + 1 == nseg) break;
        const u32 cnt = cnt_n;
        const uint2 sv = sv_n;
        if (s + 2 < nseg) {
            cnt_n = a.seg_cnt[seg + 1];
            sv_n = *reinterpret_cast<const uint2 *>(a.seg_list + (size_t)(seg + 1) * 256 + lane * 8);
        }
        if (lane < 8) mask[w][lane] = 0;
        __syncwarp();
        u32 sw[2] = { sv.x, sv.y };
#pragma unroll
        for (int k = 0; k < 8; k++) {
            u32 pos = lane * 8 + k;
            if (pos < cnt) {
                u32 c = (sw[k >> 2] >> ((k & 3) * 8)) & 0xffu;
                atomicOr(&mask[w][c >> 5], 1u << (c & 31));
                L[w][cur ^ 1][pos] = (u8)c;
            }
        }
        __syncwarp();
        u32 keep = 0, vals[8];
#pragma unroll
        for (int k = 0; k < 8; k++) {
            u32 pos = lane * 8 + k;
            u32 c = L[w][cur][pos];
            vals[k] = c;
            if (pos < nn && !((mask[w][c >> 5] >> (c & 31)) & 1u)) keep |= 1u << k;
        }
        u32 kc = __popc(keep);
        u32 kin = warp_incl_sum(kc);
        u32 pos = cnt + kin - kc;
#pragma unroll
        for (int k = 0; k < 8; k++) if (keep & (1u << k)) L[w][cur ^ 1][pos++] = (u8)vals[k];
        __syncwarp();
        cur ^= 1;
    }
}

// ------------------------------------------------------------------ M3: apply per segment
// One thread per segment, but without the recency list: the MTF index of byte c at position p is
// the number of DISTINCT bytes seen since c's previous occurrence, i.e. the number of bytes whose
// LAST occurrence lies after last[c].  The thread keeps last[256] (position of the last occurrence
// of every byte) and a bit mask `alive` over positions (bit q set <=> q is the last occurrence of
// its byte so far); the starting list of the segment is a virtual prefix of 256 positions (list[0]
// the most recent).  Index = popcount(alive over (last[c], p)) — one or two mask words for the
// recent bytes of text, six for the ~180 positions random data looks back, where the list shuffle
// it replaces walked and shifted up to 32 eight-entry words per byte (17 ms per GiB, 10 of 32 lanes
// active; profiles/README.md).  The most recent byte is kept in registers (not in the tables)
// until another byte displaces it, so runs cost two instructions per byte.
// Tables are transposed in shared memory ([entry][thread]): conflict-free whatever the lanes index.
constexpr int NT3 = 160;              // threads (segments) per CTA: 672 B of tables each, two CTAs = 10 warps per SM (128: 8 warps)
constexpr int MW = (256 + SEG) / 32;  // mask words per segment

__global__ void __launch_bounds__(NT3) mtf_apply_kernel(MtfArgs a)
{
    extern __shared__ __align__(16) u32 sm3[];
    u32 *mask = sm3;                                            // mask[w * NT3 + tid]
    u16 *last = reinterpret_cast<u16 *>(sm3 + MW * NT3);        // last[c * NT3 + tid]
    const u32 tid = threadIdx.x;
    const u32 segc = blockIdx.x * NT3 + tid;                     // position in this launch's segment list
    if (segc >= a.total_segs) return;
    const u32 bc = find_block(a.cseg_base, a.n_blocks, segc);
    const u32 s = segc - a.cseg_base[bc];
    const u32 b = a.ids[bc];
    const u32 seg = a.seg_base[b] + s;
    const u32 n = a.blk_len[b];
    const u8 *src = a.bwt + a.blk_off[b];
    u8 *dst = a.idx + a.blk_off[b];
    const u32 start = s * SEG, end = min(start + SEG, n);

    // ---- the starting list as a virtual prefix: list[i] sits at position 255 - i.  Entries behind the
    // block's nn names are padding: written first, so that a real entry of the same value wins.
    {
        const uint4 *row = reinterpret_cast<const uint4 *>(a.seg_state + (size_t)seg * 256);
        for (int i = 15; i >= 0; i--) {
            const uint4 v = row[i];
            const u32 wv[4] = { v.x, v.y, v.z, v.w };
#pragma unroll
            for (int j = 15; j >= 0; j--) {
                const u32 c = (wv[j >> 2] >> ((j & 3) * 8)) & 0xffu;
                last[c * NT3 + tid] = (u16)(255 - (i * 16 + j));
            }
        }
        const u32 nn = a.num_names[b];                           // (compose wrote it before this launch)
#pragma unroll
        for (int w = 0; w < 8; w++) {
            // bits 256 - nn .. 255 alive
            const int lo = 256 - (int)nn - w * 32;               // first alive bit of this word (may be <= 0 or >= 32)
            mask[w * NT3 + tid] = lo >= 32 ? 0u : (lo <= 0 ? 0xffffffffu : (0xffffffffu << lo));
        }
        for (int w = 8; w < MW; w++) mask[w * NT3 + tid] = 0u;
    }

    // the most recent byte and its position are kept out of the tables
    u32 rc = a.seg_state[(size_t)seg * 256];                      // list[0]
    u32 rpos = 255;
    {   // take it out of the tables
        mask[7 * NT3 + tid] &= 0x7fffffffu;
    }

    for (u32 p = start; p < end; p += 16) {
        u32 w4[4];
        if (p + 16 <= end) {
            uint4 v = *reinterpret_cast<const uint4 *>(src + p);
            w4[0] = v.x; w4[1] = v.y; w4[2] = v.z; w4[3] = v.w;
        } else {
            w4[0] = w4[1] = w4[2] = w4[3] = 0;
            for (u32 j = 0; p + j < end; j++) w4[j >> 2] |= (u32)src[p + j] << ((j & 3) * 8);
        }
        u32 o4[4] = { 0, 0, 0, 0 };
        const u32 base = 256 + (p - start);                      // position of byte 0 of this vector
#pragma unroll
        for (int j = 0; j < 16; j++) {
            // a whole word equal to the current byte is four zero indices: nothing moves
            if ((j & 3) == 0 && p + j + 4 <= end && w4[j >> 2] == rc * 0x01010101u) {
                rpos = base + j + 3;
                j += 3;
                continue;
            }
            if (p + j < end) {
                const u32 c = (w4[j >> 2] >> ((j & 3) * 8)) & 0xffu;
                const u32 cur = base + j;
                u32 k = 0;
                if (c != rc) {
                    // commit the displaced byte: it is the last occurrence of rc
                    mask[(rpos >> 5) * NT3 + tid] |= 1u << (rpos & 31);
                    last[rc * NT3 + tid] = (u16)rpos;
                    // alive positions in (lp, cur): every one is a distinct byte more recent than c
                    const u32 lp = last[c * NT3 + tid];
                    const u32 lo = lp + 1, hi = cur - 1;          // inclusive range; never empty (rpos = cur - 1 > lp)
                    const u32 wl = lo >> 5, wr = hi >> 5;
                    const u32 ml = 0xffffffffu << (lo & 31), mr = 0xffffffffu >> (31 - (hi & 31));
                    if (wl == wr) {
                        k = __popc(mask[wl * NT3 + tid] & ml & mr);
                    } else {
                        k = __popc(mask[wl * NT3 + tid] & ml) + __popc(mask[wr * NT3 + tid] & mr);
                        for (u32 w = wl + 1; w < wr; w++) k += __popc(mask[w * NT3 + tid]);
                    }
                    mask[(lp >> 5) * NT3 + tid] &= ~(1u << (lp & 31));
                    rc = c;
                }
                rpos = cur;
                o4[j >> 2] |= k << ((j & 3) * 8);
            }
        }
        if (p + 16 <= end) {
            *reinterpret_cast<uint4 *>(dst + p) = make_uint4(o4[0], o4[1], o4[2], o4[3]);
        } else {
            for (u32 j = 0; p + j < end; j++) dst[p + j] = (u8)(o4[j >> 2] >> ((j & 3) * 8));
        }
    }
}

// ------------------------------------------------------------------ M4: RUNA/RUNB + compaction
constexpr int T4 = 512;
constexpr int K4 = 8;
constexpr int TILE4 = T4 * K4;

__global__ void __launch_bounds__(T4) mtf_rle2_kernel(MtfArgs a)
{
    __shared__ u32 hist[260];
    __shared__ u32 scratch[40];
    const u32 tid = threadIdx.x;
    const u32 b = a.ids[blockIdx.x];
    const u32 n = a.blk_len[b];
    const u8 *idx = a.idx + a.blk_off[b];
    u16 *out = a.syms + a.sym_off[b];
    for (int i = tid; i < 260; i += T4) hist[i] = 0;
    __syncthreads();

    u32 carry_nz = 0;          // 1-based position of the last nonzero index so far
    u32 carry_out = 0;         // symbols emitted so far
    for (u32 base = 0; base < n; base += TILE4) {
        const u32 j0 = base + tid * K4;
        u32 v[K4 + 1];
        {
            u32 lo = 0, hi = 0;
            if (j0 + K4 <= n) {
                uint2 t = *reinterpret_cast<const uint2 *>(idx + j0);
                lo = t.x; hi = t.y;
            } else {
                for (u32 k = 0; k < K4 && j0 + k < n; k++) {
                    u32 x = idx[j0 + k];
                    if (k < 4) lo |= x << (k * 8); else hi |= x << ((k - 4) * 8);
                }
            }
#pragma unroll
            for (int k = 0; k < 4; k++) { v[k] = (lo >> (k * 8)) & 0xffu; v[k + 4] = (hi >> (k * 8)) & 0xffu; }
        }
        // first index of the next thread (run-end detection); 1 = "nonzero" sentinel at block end
        u32 nxt = __shfl_down_sync(0xffffffffu, v[0], 1);
        if (lane_id() == 31) nxt = (j0 + K4 < n) ? idx[j0 + K4] : 1u;
        v[K4] = nxt;

        u32 lastnz = 0;
#pragma unroll
        for (int k = 0; k < K4; k++) if (j0 + k < n && v[k] != 0) lastnz = j0 + k + 1;
        u32 tot_nz;
        u32 ex_nz = block_excl_max<T4>(lastnz, scratch, &tot_nz);
        ex_nz = max(ex_nz, carry_nz);

        // count
        u32 cnt = 0;
        {
            u32 nz = ex_nz;
#pragma unroll
            for (int k = 0; k < K4; k++) {
                u32 j = j0 + k;
                if (j < n) {
                    if (v[k] != 0) { cnt++; nz = j + 1; }
                    else {
                        bool endrun = (j + 1 == n) || (v[k + 1] != 0);
                        if (endrun) cnt += 31 - __clz(j + 1 - nz + 1);
                    }
                }
            }
        }
        u32 tot_out;
        u32 ex_out = block_excl_sum<T4>(cnt, scratch, &tot_out);
        u32 o = carry_out + ex_out;

        // emit
        u32 runa = 0, runb = 0;
        {
            u32 nz = ex_nz;
#pragma unroll
            for (int k = 0; k < K4; k++) {
                u32 j = j0 + k;
                u32 sym = 0xffffffffu;
                if (j < n) {
                    if (v[k] != 0) {
                        sym = v[k] + 1;
                        out[o++] = (u16)sym;
                        nz = j + 1;
                    } else {
                        bool endrun = (j + 1 == n) || (v[k + 1] != 0);
                        if (endrun) {
                            u32 code = j + 1 - nz + 1;          // zero_count + 1
                            int nd = 31 - __clz(code);
                            runb += __popc(code & ((1u << nd) - 1u));
                            runa += nd - __popc(code & ((1u << nd) - 1u));
                            for (int d = 0; d < nd; d++) out[o++] = (u16)((code >> d) & 1u);
                        }
                    }
                }
                // histogram of the index symbols (shared atomics: ~1-3 SM-cycles per warp-op on this
                // part, whereas match.any costs ~1000 cycles per call — tools/micro/match_bench.cu)
                if (sym != 0xffffffffu) atomicAdd(&hist[sym], 1u);
            }
        }
        runa = __reduce_add_sync(0xffffffffu, runa);
        runb = __reduce_add_sync(0xffffffffu, runb);
        if (lane_id() == 0) {
            if (runa) atomicAdd(&hist[0], runa);
            if (runb) atomicAdd(&hist[1], runb);
        }
        carry_out += tot_out;
        carry_nz = max(carry_nz, tot_nz);
        __syncthreads();
    }
    const u32 nn = a.num_names[b];
    if (tid == 0) {
        out[carry_out] = (u16)(nn + 1);                         // EOB
        hist[nn + 1] = 1;
        a.sym_len[b] = carry_out + 1;
    }
    __syncthreads();
    for (int i = tid; i < 258; i += T4) a.freqs[(size_t)b * 258 + i] = hist[i];
}

}  // namespace mtf

cudaError_t mtf_launch(const MtfArgs &a, cudaStream_t st, uint32_t *launches)
{
    if (a.n_blocks == 0) return cudaSuccess;
    unsigned g1 = (a.total_segs + mtf::NT1 - 1) / mtf::NT1;
    mtf::mtf_summary_kernel<<<g1, mtf::NT1, 0, st>>>(a);
    mtf::mtf_compose_kernel<<<(a.n_blocks + mtf::W2 - 1) / mtf::W2, mtf::W2 * 32, 0, st>>>(a);
    const size_t smem3 = (size_t)mtf::NT3 * (mtf::MW * 4 + 256 * 2);
    {   // (per device; cheap)
        cudaError_t e = cudaFuncSetAttribute(mtf::mtf_apply_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem3);
        if (e != cudaSuccess) return e;
    }
    mtf::mtf_apply_kernel<<<(a.total_segs + mtf::NT3 - 1) / mtf::NT3, mtf::NT3, smem3, st>>>(a);
    mtf::mtf_rle2_kernel<<<a.n_blocks, mtf::T4, 0, st>>>(a);
    if (launches) *launches += 4;
    return cudaGetLastError();
}

}  // namespace bnz
