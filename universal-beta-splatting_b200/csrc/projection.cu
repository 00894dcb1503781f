// K5/K6: world -> screen projection of already-conditioned 3-D primitives (means [N,3], covars [N,6]).
// Drop-in for the reference's fully_fused_projection_{fwd,bwd} (covars path, perspective camera).
#include "common.cuh"
#include "proj_math.cuh"

namespace ubs {

constexpr int kProjThreads = 256;

__global__ void __launch_bounds__(kProjThreads)
projection_fwd_kernel(int C, int64_t N, const float *__restrict__ means, const float *__restrict__ covars,
                      const float *__restrict__ viewmats, const float *__restrict__ Ks, uint32_t width, uint32_t height,
                      float eps2d, float near_plane, float far_plane, float radius_clip, int32_t *__restrict__ radii,
                      float *__restrict__ means2d, float *__restrict__ depths, float *__restrict__ conics,
                      float *__restrict__ compensations) {
    // grid.y = camera; grid.x covers primitives. One thread per (camera, primitive).
    const int cid = blockIdx.y;
    const int64_t gid = (int64_t)blockIdx.x * kProjThreads + threadIdx.x;
    if (gid >= N) return;
    const Cam cam = load_cam(viewmats + cid * 16, Ks + cid * 9);

    float p[3], s6[6];
    p[0] = __ldg(means + gid * 3 + 0);
    p[1] = __ldg(means + gid * 3 + 1);
    p[2] = __ldg(means + gid * 3 + 2);
    const float2 *c2 = reinterpret_cast<const float2 *>(covars + gid * 6);  // 24-byte rows are 8-byte aligned
    const float2 a = __ldg(c2), b = __ldg(c2 + 1), c = __ldg(c2 + 2);
    s6[0] = a.x, s6[1] = a.y, s6[2] = b.x, s6[3] = b.y, s6[4] = c.x, s6[5] = c.y;

    const Splat2D o = project_splat(cam, p, s6, width, height, eps2d, near_plane, far_plane, radius_clip);
    const int64_t idx = (int64_t)cid * N + gid;
    radii[idx] = o.radius;
    reinterpret_cast<float2 *>(means2d)[idx] = make_float2(o.mean2d[0], o.mean2d[1]);
    depths[idx] = o.depth;
    conics[idx * 3 + 0] = o.conic[0];
    conics[idx * 3 + 1] = o.conic[1];
    conics[idx * 3 + 2] = o.conic[2];
    if (compensations != nullptr) compensations[idx] = o.compensation;
}

__global__ void __launch_bounds__(kProjThreads)
projection_bwd_kernel(int C, int64_t N, const float *__restrict__ means, const float *__restrict__ covars,
                      const float *__restrict__ viewmats, const float *__restrict__ Ks, uint32_t width, uint32_t height,
                      float eps2d, const int32_t *__restrict__ radii, const float *__restrict__ conics,
                      const float *__restrict__ compensations, const float *__restrict__ v_means2d,
                      const float *__restrict__ v_depths, const float *__restrict__ v_conics,
                      const float *__restrict__ v_compensations, float *__restrict__ v_means,
                      float *__restrict__ v_covars, float *__restrict__ v_viewmats) {
    // One thread per primitive, looping over cameras: the per-primitive sums need no atomics and are
    // deterministic (the reference uses label-partitioned warp sums + atomics, _bwd.cu:176-201).
    const int64_t gid = (int64_t)blockIdx.x * kProjThreads + threadIdx.x;
    const bool active = gid < N;
    float p[3] = {0.f, 0.f, 0.f}, s6[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    if (active) {
#pragma unroll
        for (int k = 0; k < 3; ++k) p[k] = __ldg(means + gid * 3 + k);
#pragma unroll
        for (int k = 0; k < 6; ++k) s6[k] = __ldg(covars + gid * 6 + k);
    }
    float vp[3] = {0.f, 0.f, 0.f}, vs[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    for (int cid = 0; cid < C; ++cid) {
        float vR[9] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f}, vt[3] = {0.f, 0.f, 0.f};
        const int64_t idx = (int64_t)cid * N + gid;
        if (active && radii[idx] > 0) {
            const Cam cam = load_cam(viewmats + cid * 16, Ks + cid * 9);
            const float conic[3] = {conics[idx * 3], conics[idx * 3 + 1], conics[idx * 3 + 2]};
            const float vm[2] = {v_means2d[idx * 2], v_means2d[idx * 2 + 1]};
            const float vc[3] = {v_conics[idx * 3], v_conics[idx * 3 + 1], v_conics[idx * 3 + 2]};
            const bool comp = v_compensations != nullptr;
            project_splat_vjp(cam, p, s6, width, height, eps2d, conic, comp ? compensations + idx : nullptr, vm,
                              v_depths[idx], vc, comp ? v_compensations + idx : nullptr, vp, vs,
                              v_viewmats ? vR : nullptr, v_viewmats ? vt : nullptr);
        }
        if (v_viewmats != nullptr) {
            // warp reduce, then one atomic per warp per entry
#pragma unroll
            for (int k = 0; k < 12; ++k) {
                float v = k < 9 ? vR[k] : vt[k - 9];
#pragma unroll
                for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
                if ((threadIdx.x & 31) == 0 && v != 0.f) {
                    const int i = k < 9 ? k / 3 : k - 9, j = k < 9 ? k % 3 : 3;
                    atomicAdd(v_viewmats + cid * 16 + i * 4 + j, v);
                }
            }
        }
    }
    if (active) {
#pragma unroll
        for (int k = 0; k < 3; ++k) v_means[gid * 3 + k] = vp[k];
#pragma unroll
        for (int k = 0; k < 6; ++k) v_covars[gid * 6 + k] = vs[k];
    }
}

}  // namespace ubs

extern "C" int ubs_projection_fwd(int C, int64_t N, const float *means, const float *covars, const float *viewmats,
                                  const float *Ks, int width, int height, float eps2d, float near_plane,
                                  float far_plane, float radius_clip, int32_t *radii, float *means2d, float *depths,
                                  float *conics, float *compensations, void *stream) {
    using namespace ubs;
    UBS_CHECK_ARG(C >= 0 && N >= 0 && width > 0 && height > 0, "projection_fwd: bad sizes C=%d N=%lld %dx%d", C,
                  (long long)N, width, height);
    if (C == 0 || N == 0) return UBS_OK;
    UBS_CHECK_ARG(means && covars && viewmats && Ks && radii && means2d && depths && conics,
                  "projection_fwd: null pointer");
    UBS_CHECK_ARG(C <= 65535, "projection_fwd: C=%d exceeds 65535", C);
    dim3 grid((unsigned)ceil_div(N, kProjThreads), (unsigned)C);
    projection_fwd_kernel<<<grid, kProjThreads, 0, (cudaStream_t)stream>>>(
        C, N, means, covars, viewmats, Ks, (uint32_t)width, (uint32_t)height, eps2d, near_plane, far_plane,
        radius_clip, radii, means2d, depths, conics, compensations);
    UBS_LAUNCH_CHECK("projection_fwd_kernel");
    return UBS_OK;
}

extern "C" int ubs_projection_bwd(int C, int64_t N, const float *means, const float *covars, const float *viewmats,
                                  const float *Ks, int width, int height, float eps2d, const int32_t *radii,
                                  const float *conics, const float *compensations, const float *v_means2d,
                                  const float *v_depths, const float *v_conics, const float *v_compensations,
                                  float *v_means, float *v_covars, float *v_viewmats, void *stream) {
    using namespace ubs;
    UBS_CHECK_ARG(C >= 0 && N >= 0 && width > 0 && height > 0, "projection_bwd: bad sizes");
    if (N == 0) return UBS_OK;
    UBS_CHECK_ARG(means && covars && viewmats && Ks && radii && conics && v_means2d && v_depths && v_conics &&
                      v_means && v_covars,
                  "projection_bwd: null pointer");
    UBS_CHECK_ARG((v_compensations == nullptr) || (compensations != nullptr),
                  "projection_bwd: v_compensations given without compensations");
    cudaStream_t s = (cudaStream_t)stream;
    if (v_viewmats != nullptr) UBS_CUDA_TRY(cudaMemsetAsync(v_viewmats, 0, sizeof(float) * 16 * C, s));
    projection_bwd_kernel<<<(unsigned)ceil_div(N, kProjThreads), kProjThreads, 0, s>>>(
        C, N, means, covars, viewmats, Ks, (uint32_t)width, (uint32_t)height, eps2d, radii, conics, compensations,
        v_means2d, v_depths, v_conics, v_compensations, v_means, v_covars, v_viewmats);
    UBS_LAUNCH_CHECK("projection_bwd_kernel");
    return UBS_OK;
}
