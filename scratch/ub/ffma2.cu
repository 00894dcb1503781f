#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
__device__ __forceinline__ u64 pack(float a, float b) { u64 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ void unpack(u64 r, float &a, float &b) { asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(r)); }
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) { u64 d; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
// MODE 0: 8 independent scalar FFMA chains; MODE 1: 4 independent FFMA2 chains (same flops); MODE 2: 8 FFMA + 8 IADD; MODE 3: 4 FFMA2 + 8 IADD
template <int MODE>
__global__ void k(float *out, int iters, float s) {
    float a[8]; u64 p[4]; int q[8];
    for (int i = 0; i < 8; ++i) { a[i] = threadIdx.x * 0.001f + i; q[i] = threadIdx.x + i; }
    for (int i = 0; i < 4; ++i) p[i] = pack(a[2 * i], a[2 * i + 1]);
    const u64 ss = pack(s, s), tt = pack(0.5f, 0.25f);
    for (int it = 0; it < iters; ++it) {
        if (MODE == 0 || MODE == 2) {
#pragma unroll
            for (int i = 0; i < 8; ++i) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a[i]) : "f"(s), "f"(0.5f));
        } else {
#pragma unroll
            for (int i = 0; i < 4; ++i) p[i] = fma2(p[i], ss, tt);
        }
        if (MODE >= 2) {
#pragma unroll
            for (int i = 0; i < 8; ++i) asm volatile("add.s32 %0, %0, %1;" : "+r"(q[i]) : "r"(it));
        }
    }
    float r = 0; int qq = 0;
    for (int i = 0; i < 8; ++i) { r += a[i]; qq += q[i]; }
    for (int i = 0; i < 4; ++i) { float x, y; unpack(p[i], x, y); r += x + y; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = r + qq;
}
int main() {
    float *out; cudaMalloc(&out, 148 * 2048 * 4);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 20000;
    for (int mode = 0; mode < 4; ++mode) {
        for (int rep = 0; rep < 2; ++rep) {
            cudaEventRecord(e0);
            if (mode == 0) k<0><<<148 * 2, 1024>>>(out, iters, 0.999f);
            if (mode == 1) k<1><<<148 * 2, 1024>>>(out, iters, 0.999f);
            if (mode == 2) k<2><<<148 * 2, 1024>>>(out, iters, 0.999f);
            if (mode == 3) k<3><<<148 * 2, 1024>>>(out, iters, 0.999f);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
        }
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        double cyc = ms * 1e-3 * 1.965e9;
        // per SM: 64 warps, each iter: 8 scalar fma-equivalents (+8 iadd)
        printf("mode %d: %.3f ms -> %.2f SM-cycles per warp-iteration (8 fma lanes-worth%s)\n", mode, ms, cyc / (iters * 64.0), mode >= 2 ? " + 8 IADD" : "");
    }
}
