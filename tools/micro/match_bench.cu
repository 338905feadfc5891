// micro-benchmark: warp match.any vs ballot-loop vs shared atomics (cycles per warp-op at full occupancy)
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t match_ballot(uint32_t d, int bits) {
    uint32_t peers = 0xffffffffu;
    for (int b = 0; b < bits; b++) { uint32_t bit = (d >> b) & 1u; uint32_t v = __ballot_sync(0xffffffffu, bit); peers &= bit ? v : ~v; }
    return peers;
}
template <int MODE>
__global__ void k(uint32_t *out, int iters, long long *cyc) {
    __shared__ uint32_t cnt[1024];
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) cnt[i] = 0;
    __syncthreads();
    uint32_t x = threadIdx.x * 2654435761u + blockIdx.x * 40503u, acc = 0;
    long long t0 = clock64();
    for (int i = 0; i < iters; i++) {
        x = x * 1664525u + 1013904223u;
        uint32_t d = (x >> 13) & 0xff;
        if (MODE == 0) acc += __match_any_sync(0xffffffffu, d);
        else if (MODE == 1) acc += match_ballot(d, 8);
        else if (MODE == 2) acc += atomicAdd(&cnt[d + (threadIdx.x >> 5 & 3) * 256], 1u);
        else if (MODE == 3) acc += __match_any_sync(0xffffffffu, (x >> 13) & 0x3ff);
        else if (MODE == 4) { uint32_t p = __match_any_sync(0xffffffffu, d); acc += __popc(p) + (31 - __clz(p)); }
    }
    long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
int main() {
    uint32_t *out; long long *cyc; cudaMalloc(&out, 148 * 4 * 512 * 4); cudaMallocManaged(&cyc, 8);
    const int iters = 4096;
    const char *names[] = {"match.any 8b", "ballot x8", "smem atomicAdd ret", "match.any 10b", "match+popc+flo"};
    for (int warps_per_sm : {16, 32, 64}) {
        int ctas = warps_per_sm / 16;   // 512-thread CTAs
        for (int m = 0; m < 5; m++) {
            for (int rep = 0; rep < 2; rep++) {
                if (m == 0) k<0><<<148 * ctas, 512>>>(out, iters, cyc);
                if (m == 1) k<1><<<148 * ctas, 512>>>(out, iters, cyc);
                if (m == 2) k<2><<<148 * ctas, 512>>>(out, iters, cyc);
                if (m == 3) k<3><<<148 * ctas, 512>>>(out, iters, cyc);
                if (m == 4) k<4><<<148 * ctas, 512>>>(out, iters, cyc);
                cudaDeviceSynchronize();
            }
            double per_sm_cyc_per_warpop = (double)*cyc / iters / warps_per_sm;
            printf("warps/SM=%2d %-20s: %8.1f cycles per iteration per warp, %6.2f SM-cycles per warp-op\n", warps_per_sm, names[m], (double)*cyc / iters, per_sm_cyc_per_warpop);
        }
    }
    return 0;
}
