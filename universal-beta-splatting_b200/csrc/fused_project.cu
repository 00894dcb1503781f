// North-star kernel (1): fused activation + covariance build + conditioning + projection + tile count, straight
// from the PACKED primitive records.  No single reference counterpart: it replaces the chain
//   softplus/sigmoid/exp/cat/contiguous (scene/beta_model.py:103-121) -> K1 -> K2 -> view-dir glue (:675-690)
//   -> K3 -> 3x3->6 gather (rendering.py:55-56) -> K5 -> first pass of K7
// which in the reference is ~25 launches and ~1.3 KB of HBM traffic per primitive, with one pass that reads the
// 144 B (D=6) / 176 B (D=7) record once and writes 52 B per (camera, primitive).
//
// Record staging: each CTA owns kFusedThreads consecutive records = one contiguous, 16-byte aligned span of
// global memory, fetched with ONE TMA bulk copy (cp.async.bulk ... mbarrier::complete_tx) into shared memory;
// threads then read their own record with conflict-free 128-bit shared loads (row strides of 36 / 44 floats map
// each quarter-warp onto 32 distinct banks).
#include "common.cuh"
#include "cond_math.cuh"
#include "isect.cuh"
#include "proj_math.cuh"

namespace ubs {
namespace {

constexpr int kFusedThreads = 128;

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_bulk_g2s(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t phase) {
    uint32_t done = 0;
    while (!done) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(smem_u32(bar)), "r"(phase)
            : "memory");
    }
}

// Everything about one primitive that does not depend on the camera.
template <int D>
struct PrimState {
    static constexpr int C = D - 3;
    float xyz[3];
    float mu2[C];
    float rgb[3];
    float opacity;       // sigmoid
    float beta0;         // spatial beta
    float beta_c[C];     // conditional betas
    CondPrep<C> prep;
};

template <int D>
__device__ __forceinline__ void decode_record(const float *rec, PrimState<D> &ps) {
    constexpr int C = D - 3, M = NdDims<D>::M;
    // layout (ubs_b200.h): xyz | mean | rgb | opacity | beta(D-2) | scale(D) | l_triangle(M)
#pragma unroll
    for (int k = 0; k < 3; ++k) ps.xyz[k] = rec[k];
#pragma unroll
    for (int k = 0; k < C; ++k) ps.mu2[k] = rec[3 + k];
#pragma unroll
    for (int k = 0; k < 3; ++k) ps.rgb[k] = rec[D + k];
    ps.opacity = sigmoid_f(rec[D + 3]);
    ps.beta0 = beta_act_f(rec[D + 4]);
#pragma unroll
    for (int k = 0; k < C; ++k) ps.beta_c[k] = beta_act_f(rec[D + 5 + k]);
    float s[D], lt[M];
#pragma unroll
    for (int k = 0; k < D; ++k) s[k] = softplus_f(rec[2 * D + 2 + k]);
#pragma unroll
    for (int k = 0; k < M; ++k) lt[k] = rec[3 * D + 2 + k];
    const float R[9] = {1.f, lt[0], lt[1], -lt[0], 1.f, lt[2], -lt[1], -lt[2], 1.f};  // K1: I + skew
    float L[D * D], S[D * D];
    build_L<D>(R, s, lt, L);
    covar_from_L<D>(L, s, S, D);
    float V11[9], V12[3 * C], V21[C * 3], V22[C * C];
#pragma unroll
    for (int r = 0; r < 3; ++r) {
#pragma unroll
        for (int c = 0; c < 3; ++c) V11[r * 3 + c] = S[r * D + c];
#pragma unroll
        for (int c = 0; c < C; ++c) {
            V12[r * C + c] = S[r * D + 3 + c];
            V21[c * 3 + r] = S[(3 + c) * D + r];
        }
    }
#pragma unroll
    for (int r = 0; r < C; ++r)
#pragma unroll
        for (int c = 0; c < C; ++c) V22[r * C + c] = S[(3 + r) * D + 3 + c];
    cond_prepare<C>(V11, V12, V21, V22, ps.beta_c, ps.prep);
}

template <int D>
__global__ void __launch_bounds__(kFusedThreads)
fused_project_fwd_kernel(int C, int64_t N, const float *__restrict__ records, const float *__restrict__ viewmats,
                         const float *__restrict__ Ks, const float *__restrict__ cam_pos,
                         const float *__restrict__ timestamps, const uint8_t *__restrict__ prim_mask, uint32_t width,
                         uint32_t height, float eps2d, float near_plane, float far_plane, float radius_clip,
                         int calc_comp, uint32_t tile_size, uint32_t tile_width, uint32_t tile_height,
                         int32_t *__restrict__ radii, float *__restrict__ means2d, float *__restrict__ depths,
                         float *__restrict__ conics, float *__restrict__ opacities, float *__restrict__ betas,
                         float *__restrict__ colors, int32_t *__restrict__ tiles_per_gauss) {
    constexpr int Cd = D - 3;
    constexpr int STRIDE = UBS_RECORD_STRIDE(D);
    __shared__ __align__(128) float s_rec[kFusedThreads * STRIDE];
    __shared__ __align__(8) uint64_t s_bar;

    const int64_t base = (int64_t)blockIdx.x * kFusedThreads;
    const int n_here = (int)min((int64_t)kFusedThreads, N - base);
    if (threadIdx.x == 0) mbar_init(&s_bar, 1);
    __syncthreads();
    if (threadIdx.x == 0) {
        const uint32_t bytes = (uint32_t)n_here * STRIDE * sizeof(float);
        mbar_expect_tx(&s_bar, bytes);
        tma_bulk_g2s(s_rec, records + base * STRIDE, bytes, &s_bar);
    }
    mbar_wait(&s_bar, 0);

    const int64_t gid = base + threadIdx.x;
    const bool active = threadIdx.x < n_here && (prim_mask == nullptr || prim_mask[gid]);
    PrimState<D> ps;
    if (active) {
        float rec[STRIDE];
        const float4 *src = reinterpret_cast<const float4 *>(s_rec + threadIdx.x * STRIDE);
#pragma unroll
        for (int k = 0; k < STRIDE / 4; ++k) {
            const float4 v = src[k];
            rec[4 * k + 0] = v.x, rec[4 * k + 1] = v.y, rec[4 * k + 2] = v.z, rec[4 * k + 3] = v.w;
        }
        decode_record<D>(rec, ps);
    }

    for (int cid = blockIdx.y; cid < C; cid += gridDim.y) {
        if (threadIdx.x >= n_here) continue;
        const int64_t idx = (int64_t)cid * N + gid;
        Splat2D o;
        o.radius = 0;
        o.mean2d[0] = o.mean2d[1] = o.depth = o.conic[0] = o.conic[1] = o.conic[2] = o.compensation = 0.f;
        float opac = 0.f;
        if (active) {
            const Cam cam = load_cam(viewmats + cid * 16, Ks + cid * 9);
            // query: unit view direction (+ timestamp)   (scene/beta_model.py:675-690)
            float x[Cd];
            {
                const float dx = ps.xyz[0] - cam_pos[cid * 3 + 0], dy = ps.xyz[1] - cam_pos[cid * 3 + 1],
                            dz = ps.xyz[2] - cam_pos[cid * 3 + 2];
                const float nrm = __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz)));
                x[0] = __fdiv_rn(dx, nrm) - ps.mu2[0];
                x[1] = __fdiv_rn(dy, nrm) - ps.mu2[1];
                x[2] = __fdiv_rn(dz, nrm) - ps.mu2[2];
                if constexpr (Cd > 3) {
                    x[3] = timestamps[cid] - ps.mu2[3];
#pragma unroll
                    for (int k = 4; k < Cd; ++k) x[k] = -ps.mu2[k];
                }
            }
            float mean[3];
            cond_apply<Cd>(ps.prep, ps.xyz, x, ps.opacity, ps.beta_c, mean, opac);
            // upper triangle of the (unsymmetrised) conditional covariance, as rendering.py:55-56 gathers it
            const float s6[6] = {ps.prep.cov[0], ps.prep.cov[1], ps.prep.cov[2], ps.prep.cov[4], ps.prep.cov[5], ps.prep.cov[8]};
            o = project_splat(cam, mean, s6, width, height, eps2d, near_plane, far_plane, radius_clip);
            if (calc_comp) opac *= o.compensation;
        }
        int32_t cnt = 0;
        if (o.radius > 0) {
            const TileRect t = tile_rect(o.mean2d[0], o.mean2d[1], o.radius, tile_size, tile_width, tile_height);
            cnt = (int32_t)((t.y1 - t.y0) * (t.x1 - t.x0));
        }
        radii[idx] = o.radius;
        reinterpret_cast<float2 *>(means2d)[idx] = make_float2(o.mean2d[0], o.mean2d[1]);
        depths[idx] = o.depth;
        conics[idx * 3 + 0] = o.conic[0];
        conics[idx * 3 + 1] = o.conic[1];
        conics[idx * 3 + 2] = o.conic[2];
        opacities[idx] = o.radius > 0 ? opac : 0.f;
        betas[idx] = active ? ps.beta0 : 0.f;
        if (colors != nullptr) {
            colors[idx * 3 + 0] = active ? ps.rgb[0] : 0.f;
            colors[idx * 3 + 1] = active ? ps.rgb[1] : 0.f;
            colors[idx * 3 + 2] = active ? ps.rgb[2] : 0.f;
        }
        tiles_per_gauss[idx] = cnt;
    }
}

}  // namespace
}  // namespace ubs

extern "C" int ubs_fused_project_fwd(int C, int64_t N, int D, const float *records, const float *viewmats,
                                     const float *Ks, const float *cam_pos, const float *timestamps,
                                     const uint8_t *prim_mask, int width, int height, float eps2d, float near_plane,
                                     float far_plane, float radius_clip, int calc_compensations, int tile_size,
                                     int tile_width, int tile_height, int32_t *radii, float *means2d, float *depths,
                                     float *conics, float *opacities, float *betas, float *colors,
                                     int32_t *tiles_per_gauss, int64_t *n_isects, void *workspace,
                                     size_t workspace_bytes, void *stream) {
    using namespace ubs;
    UBS_CHECK_ARG(C >= 0 && N >= 0 && width > 0 && height > 0 && tile_size > 0, "fused_project_fwd: bad sizes");
    UBS_CHECK_ARG(D == 6 || D == 7, "fused_project_fwd: D must be 6 or 7 (got %d)", D);
    UBS_CHECK_ARG(n_isects != nullptr, "fused_project_fwd: n_isects is null");
    cudaStream_t s = (cudaStream_t)stream;
    const int64_t CN = (int64_t)C * N;
    if (CN == 0) {
        UBS_CUDA_TRY(cudaMemsetAsync(n_isects, 0, sizeof(int64_t), s));
        return UBS_OK;
    }
    UBS_CHECK_ARG(records && viewmats && Ks && cam_pos && radii && means2d && depths && conics && opacities && betas &&
                      tiles_per_gauss && workspace,
                  "fused_project_fwd: null pointer");
    UBS_CHECK_ARG(D != 7 || timestamps != nullptr, "fused_project_fwd: D=7 needs timestamps");
    UBS_CHECK_ARG(((uintptr_t)records & 15) == 0, "fused_project_fwd: records must be 16-byte aligned");
    UBS_CHECK_ARG(CN < ((int64_t)1 << 31), "fused_project_fwd: C*N must fit int32 flatten ids");
    if (workspace_bytes < ubs_isect_workspace_bytes(CN, 0)) {
        set_error("fused_project_fwd: workspace %zu < %zu", workspace_bytes, ubs_isect_workspace_bytes(CN, 0));
        return UBS_ENOSPC;
    }
    const unsigned gx = (unsigned)ceil_div(N, kFusedThreads);
    // enough CTAs to fill the machine: split the camera loop over grid.y only when there are few primitives
    int sm = ubs_device_sm_count();
    if (sm <= 0) sm = 148;
    unsigned gy = 1;
    while ((int64_t)gx * gy < (int64_t)sm * 8 && (int)gy < C) gy *= 2;
    if ((int)gy > C) gy = (unsigned)C;
    dim3 grid(gx, gy);
#define UBS_FUSED_LAUNCH(DD)                                                                                           \
    fused_project_fwd_kernel<DD><<<grid, kFusedThreads, 0, s>>>(                                                       \
        C, N, records, viewmats, Ks, cam_pos, timestamps, prim_mask, (uint32_t)width, (uint32_t)height, eps2d,         \
        near_plane, far_plane, radius_clip, calc_compensations, (uint32_t)tile_size, (uint32_t)tile_width,             \
        (uint32_t)tile_height, radii, means2d, depths, conics, opacities, betas, colors, tiles_per_gauss)
    if (D == 6) UBS_FUSED_LAUNCH(6);
    else UBS_FUSED_LAUNCH(7);
#undef UBS_FUSED_LAUNCH
    UBS_LAUNCH_CHECK("fused_project_fwd_kernel");
    return isect_blocksums_from_counts(CN, tiles_per_gauss, workspace, n_isects, s);
}
