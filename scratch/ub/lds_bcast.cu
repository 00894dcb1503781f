// Shared-memory load throughput when the lanes of a warp read 1 / 2 / 4 / 8 distinct 16-byte (or 4-byte) addresses:
// decides whether half-warp (4x4 pixel) lists in the compositing kernels double the shared-memory wavefronts.
#include <cstdio>
#include <cuda_runtime.h>
#include <stdint.h>
template <int GROUPS, int VEC>
__global__ void k(float *out, int iters) {
    __shared__ __align__(16) float s[4096];
    for (int i = threadIdx.x; i < 4096; i += blockDim.x) s[i] = (float)i;
    __syncthreads();
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t grp = lane / (32 / GROUPS);
    uint32_t base = (uint32_t)__cvta_generic_to_shared(s);
    // each group reads its own 48-byte record; records of different groups are 13 records apart (odd bank offsets)
    uint32_t addr = base + grp * 13 * 48;
    float acc = 0.f;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            if (VEC == 4) {
                float4 v;
                asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr + u * 48));
                acc += v.x + v.w;
            } else {
                float v;
                asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr + u * 48));
                acc += v;
            }
        }
        addr = base + ((addr - base + 400) & 8191);
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}
template <int GROUPS, int VEC>
void run(float *out, const char *name) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 2048;
    float ms = 0;
    for (int rep = 0; rep < 3; ++rep) {
        cudaEventRecord(e0);
        k<GROUPS, VEC><<<148 * 4, 256>>>(out, iters);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1);
    }
    double loads_per_sm = (double)iters * 8 * 8 * 4;  // warp-level load instructions per SM
    double cyc = ms * 1e-3 * 1.965e9;
    printf("%-28s %.3f ms  -> %.2f cycles per warp load instruction per SM\n", name, ms, cyc / loads_per_sm);
}
int main() {
    float *out; cudaMalloc(&out, 148 * 4 * 256 * 4);
    run<1, 4>(out, "LDS.128 1 address/warp");
    run<2, 4>(out, "LDS.128 2 addresses/warp");
    run<4, 4>(out, "LDS.128 4 addresses/warp");
    run<8, 4>(out, "LDS.128 8 addresses/warp");
    run<1, 1>(out, "LDS.32  1 address/warp");
    run<2, 1>(out, "LDS.32  2 addresses/warp");
    run<4, 1>(out, "LDS.32  4 addresses/warp");
    return 0;
}
