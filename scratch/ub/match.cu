#include <cstdio>
#include <cuda_runtime.h>
#include <stdint.h>
__device__ __forceinline__ uint32_t peers_ballot(uint32_t d) {
    uint32_t peers = 0xffffffffu;
#pragma unroll
    for (int b = 0; b < 8; ++b) {
        const bool bit = (d >> b) & 1;
        const uint32_t m = __ballot_sync(0xffffffffu, bit);
        peers &= bit ? m : ~m;
    }
    return peers;
}
template <int MODE>
__global__ void k(uint32_t *out, int iters, uint32_t seed) {
    uint32_t x = (threadIdx.x * 2654435761u + blockIdx.x * 40503u + seed);
    uint32_t acc = 0;
    for (int i = 0; i < iters; ++i) {
        x = x * 1664525u + 1013904223u;
        const uint32_t d = (x >> 13) & 255u;
        uint32_t p;
        if (MODE == 0) p = __match_any_sync(0xffffffffu, d);
        else if (MODE == 1) p = peers_ballot(d);
        else p = d;
        acc += __popc(p);
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}
int main() {
    uint32_t *out; cudaMalloc(&out, 148 * 8 * 256 * 4);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 4096;
    for (int mode = 0; mode < 3; ++mode) {
        for (int wpb : {32, 256, 1024}) for (int bps : {1, 2, 8}) {
            if (wpb * bps > 2048) continue;
            int grid = 148 * bps;
            for (int rep = 0; rep < 2; ++rep) {
                cudaEventRecord(e0);
                if (mode == 0) k<0><<<grid, wpb>>>(out, iters, rep);
                else if (mode == 1) k<1><<<grid, wpb>>>(out, iters, rep);
                else k<2><<<grid, wpb>>>(out, iters, rep);
                cudaEventRecord(e1); cudaEventSynchronize(e1);
            }
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            double warps_per_sm = (double)wpb / 32 * bps;
            double cyc = ms * 1e-3 * 1.965e9;
            printf("mode %d threads/blk %4d blk/SM %d: %.3f ms  -> %.1f cyc per warp-iter per SM (%.1f warps/SM)\n", mode, wpb, bps, ms,
                   cyc / (iters * warps_per_sm), warps_per_sm);
        }
    }
    return 0;
}
