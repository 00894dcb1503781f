// K1-K4 as stand-alone operators (drop-ins for the reference's l_triangle_to_rotmat, rot_scale_l_triangle_to_covar
// and cond_mean_convariance_opacity, forward and backward).  One thread per primitive, D dispatched to a
// compile-time template so all small matrices live in registers.
#include "common.cuh"
#include "cond_math.cuh"

namespace ubs {
namespace {

constexpr int kCondThreads = 128;

__global__ void __launch_bounds__(256)
l_triangle_to_rotmat_fwd_kernel(int64_t N, const float *__restrict__ lt, float *__restrict__ R) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    const float a0 = lt[i * 3 + 0], a1 = lt[i * 3 + 1], a2 = lt[i * 3 + 2];
    float *r = R + i * 9;
    r[0] = 1.f, r[1] = a0, r[2] = a1;
    r[3] = -a0, r[4] = 1.f, r[5] = a2;
    r[6] = -a1, r[7] = -a2, r[8] = 1.f;
}

__global__ void __launch_bounds__(256)
l_triangle_to_rotmat_bwd_kernel(int64_t N, const float *__restrict__ vR, float *__restrict__ vlt) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    const float *g = vR + i * 9;
    vlt[i * 3 + 0] = g[1] - g[3];
    vlt[i * 3 + 1] = g[2] - g[6];
    vlt[i * 3 + 2] = g[5] - g[7];
}

template <int D>
__global__ void __launch_bounds__(kCondThreads)
covar_fwd_kernel(int64_t N, int spatial, const float *__restrict__ rot, const float *__restrict__ scale,
                 const float *__restrict__ ltri, float *__restrict__ covar) {
    constexpr int M = NdDims<D>::M;
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    float R[9], s[D], lt[M], L[D * D], S[D * D];
#pragma unroll
    for (int k = 0; k < 9; ++k) R[k] = rot[i * 9 + k];
#pragma unroll
    for (int k = 0; k < D; ++k) s[k] = scale[i * D + k];
#pragma unroll
    for (int k = 0; k < M; ++k) lt[k] = ltri[i * M + k];
    build_L<D>(R, s, lt, L);
    if (spatial) {
        covar_from_L<D>(L, s, S, 3);
#pragma unroll
        for (int r = 0; r < 3; ++r)
#pragma unroll
            for (int c = 0; c < 3; ++c) covar[i * 9 + r * 3 + c] = S[r * D + c];
    } else {
        covar_from_L<D>(L, s, S, D);
#pragma unroll
        for (int k = 0; k < D * D; ++k) covar[i * D * D + k] = S[k];
    }
}

template <int D>
__global__ void __launch_bounds__(kCondThreads)
covar_bwd_kernel(int64_t N, int spatial, const float *__restrict__ rot, const float *__restrict__ scale,
                 const float *__restrict__ ltri, const float *__restrict__ v_covar, float *__restrict__ v_rot,
                 float *__restrict__ v_scale, float *__restrict__ v_ltri) {
    constexpr int M = NdDims<D>::M;
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    float R[9], s[D], lt[M], L[D * D];
#pragma unroll
    for (int k = 0; k < 9; ++k) R[k] = rot[i * 9 + k];
#pragma unroll
    for (int k = 0; k < D; ++k) s[k] = scale[i * D + k];
#pragma unroll
    for (int k = 0; k < M; ++k) lt[k] = ltri[i * M + k];
    build_L<D>(R, s, lt, L);
    float vR[9], vs[D], vlt[M];
    if (spatial) {
        float G[9];
#pragma unroll
        for (int k = 0; k < 9; ++k) G[k] = v_covar[i * 9 + k];
        covar_vjp<D>(R, s, L, G, 3, vR, vs, vlt);
    } else {
        float G[D * D];
#pragma unroll
        for (int k = 0; k < D * D; ++k) G[k] = v_covar[i * D * D + k];
        covar_vjp<D>(R, s, L, G, D, vR, vs, vlt);
    }
#pragma unroll
    for (int k = 0; k < 9; ++k) v_rot[i * 9 + k] = vR[k];
#pragma unroll
    for (int k = 0; k < D; ++k) v_scale[i * D + k] = vs[k];
#pragma unroll
    for (int k = 0; k < M; ++k) v_ltri[i * M + k] = vlt[k];
}

template <int D>
__device__ __forceinline__ void load_cond_inputs(int64_t i, const float *__restrict__ means,
                                                 const float *__restrict__ covars, const float *__restrict__ betas,
                                                 const float *__restrict__ query, float mu1[3], float x[D - 3],
                                                 float V11[9], float V12[3 * (D - 3)], float V21[(D - 3) * 3],
                                                 float V22[(D - 3) * (D - 3)], float beta[D - 3]) {
    constexpr int C = D - 3;
    const float *m = means + i * D;
    const float *v = covars + i * D * D;
#pragma unroll
    for (int k = 0; k < 3; ++k) mu1[k] = m[k];
#pragma unroll
    for (int j = 0; j < C; ++j) {
        x[j] = query[i * C + j] - m[3 + j];
        beta[j] = betas[i * C + j];
    }
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int c = 0; c < 3; ++c) V11[r * 3 + c] = v[r * D + c];
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int c = 0; c < C; ++c) V12[r * C + c] = v[r * D + 3 + c];
#pragma unroll
    for (int r = 0; r < C; ++r)
#pragma unroll
        for (int c = 0; c < 3; ++c) V21[r * 3 + c] = v[(3 + r) * D + c];
#pragma unroll
    for (int r = 0; r < C; ++r)
#pragma unroll
        for (int c = 0; c < C; ++c) V22[r * C + c] = v[(3 + r) * D + 3 + c];
}

template <int D>
__global__ void __launch_bounds__(kCondThreads)
cond_fwd_kernel(int64_t N, const float *__restrict__ means, const float *__restrict__ covars,
                const float *__restrict__ opacities, const float *__restrict__ betas, const float *__restrict__ query,
                float *__restrict__ out_means, float *__restrict__ out_covars, float *__restrict__ out_opac) {
    constexpr int C = D - 3;
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    float mu1[3], x[C], V11[9], V12[3 * C], V21[C * 3], V22[C * C], beta[C];
    load_cond_inputs<D>(i, means, covars, betas, query, mu1, x, V11, V12, V21, V22, beta);
    const CondOut<C> o = cond_forward<C>(mu1, x, V11, V12, V21, V22, opacities[i], beta);
#pragma unroll
    for (int k = 0; k < 3; ++k) out_means[i * 3 + k] = o.mean[k];
#pragma unroll
    for (int k = 0; k < 9; ++k) out_covars[i * 9 + k] = o.cov[k];
    out_opac[i] = o.opacity;
}

template <int D>
__global__ void __launch_bounds__(kCondThreads)
cond_bwd_kernel(int64_t N, const float *__restrict__ means, const float *__restrict__ covars,
                const float *__restrict__ opacities, const float *__restrict__ betas, const float *__restrict__ query,
                const float *__restrict__ v_om, const float *__restrict__ v_oc, const float *__restrict__ v_oo,
                float *__restrict__ v_means, float *__restrict__ v_covars, float *__restrict__ v_opac,
                float *__restrict__ v_betas) {
    constexpr int C = D - 3;
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    float mu1[3], x[C], V11[9], V12[3 * C], V21[C * 3], V22[C * C], beta[C];
    load_cond_inputs<D>(i, means, covars, betas, query, mu1, x, V11, V12, V21, V22, beta);
    float gM[3], gV[9];
#pragma unroll
    for (int k = 0; k < 3; ++k) gM[k] = v_om[i * 3 + k];
#pragma unroll
    for (int k = 0; k < 9; ++k) gV[k] = v_oc[i * 9 + k];
    float g_mu1[3], g_mu2[C], g11[9], g12[3 * C], g21[C * 3], g22[C * C], g_o, g_beta[C];
    cond_backward<C>(x, V12, V21, V22, opacities[i], beta, gM, gV, v_oo[i], g_mu1, g_mu2, g11, g12, g21, g22, g_o,
                     g_beta);
    float *gm = v_means + i * D;
    float *gv = v_covars + i * D * D;
#pragma unroll
    for (int k = 0; k < 3; ++k) gm[k] = g_mu1[k];
#pragma unroll
    for (int k = 0; k < C; ++k) gm[3 + k] = g_mu2[k];
#pragma unroll
    for (int r = 0; r < 3; ++r) {
#pragma unroll
        for (int c = 0; c < 3; ++c) gv[r * D + c] = g11[r * 3 + c];
#pragma unroll
        for (int c = 0; c < C; ++c) gv[r * D + 3 + c] = g12[r * C + c];
    }
#pragma unroll
    for (int r = 0; r < C; ++r) {
#pragma unroll
        for (int c = 0; c < 3; ++c) gv[(3 + r) * D + c] = g21[r * 3 + c];
#pragma unroll
        for (int c = 0; c < C; ++c) gv[(3 + r) * D + 3 + c] = g22[r * C + c];
    }
    v_opac[i] = g_o;
#pragma unroll
    for (int k = 0; k < C; ++k) v_betas[i * C + k] = g_beta[k];
}

}  // namespace
}  // namespace ubs

#define UBS_DISPATCH_D(D, ...)                                                                                         \
    switch (D) {                                                                                                       \
        case 4: { constexpr int kD = 4; __VA_ARGS__; break; }                                                          \
        case 5: { constexpr int kD = 5; __VA_ARGS__; break; }                                                          \
        case 6: { constexpr int kD = 6; __VA_ARGS__; break; }                                                          \
        case 7: { constexpr int kD = 7; __VA_ARGS__; break; }                                                          \
        case 8: { constexpr int kD = 8; __VA_ARGS__; break; }                                                          \
        default:                                                                                                       \
            ubs::set_error("unsupported D=%d (supported: 4..8)", D);                                                   \
            return UBS_EUNSUPPORTED;                                                                                   \
    }

extern "C" int ubs_l_triangle_to_rotmat_fwd(int64_t N, const float *l_triangle, float *rot, void *stream) {
    using namespace ubs;
    UBS_CHECK_ARG(N >= 0, "l_triangle_to_rotmat_fwd: N < 0");
    if (N == 0) return UBS_OK;
    UBS_CHECK_ARG(l_triangle && rot, "l_triangle_to_rotmat_fwd: null pointer");
    l_triangle_to_rotmat_fwd_kernel<<<(unsigned)ceil_div(N, 256), 256, 0, (cudaStream_t)stream>>>(N, l_triangle, rot);
    UBS_LAUNCH_CHECK("l_triangle_to_rotmat_fwd_kernel");
    return UBS_OK;
}

extern "C" int ubs_l_triangle_to_rotmat_bwd(int64_t N, const float *v_rot, float *v_l_triangle, void *stream) {
    using namespace ubs;
    UBS_CHECK_ARG(N >= 0, "l_triangle_to_rotmat_bwd: N < 0");
    if (N == 0) return UBS_OK;
    UBS_CHECK_ARG(v_rot && v_l_triangle, "l_triangle_to_rotmat_bwd: null pointer");
    l_triangle_to_rotmat_bwd_kernel<<<(unsigned)ceil_div(N, 256), 256, 0, (cudaStream_t)stream>>>(N, v_rot,
                                                                                                 v_l_triangle);
    UBS_LAUNCH_CHECK("l_triangle_to_rotmat_bwd_kernel");
    return UBS_OK;
}

extern "C" int ubs_rot_scale_l_triangle_to_covar_fwd(int64_t N, int D, int spatial_block, const float *rot,
                                                     const float *scale, const float *l_triangle, float *covar,
                                                     void *stream) {
    using namespace ubs;
    UBS_CHECK_ARG(N >= 0, "covar_fwd: N < 0");
    if (N == 0) return UBS_OK;
    UBS_CHECK_ARG(rot && scale && l_triangle && covar, "covar_fwd: null pointer");
    const unsigned grid = (unsigned)ceil_div(N, kCondThreads);
    UBS_DISPATCH_D(D, covar_fwd_kernel<kD><<<grid, kCondThreads, 0, (cudaStream_t)stream>>>(N, spatial_block, rot,
                                                                                              scale, l_triangle, covar));
    UBS_LAUNCH_CHECK("covar_fwd_kernel");
    return UBS_OK;
}

extern "C" int ubs_rot_scale_l_triangle_to_covar_bwd(int64_t N, int D, int spatial_block, const float *rot,
                                                     const float *scale, const float *l_triangle, const float *v_covar,
                                                     float *v_rot, float *v_scale, float *v_l_triangle, void *stream) {
    using namespace ubs;
    UBS_CHECK_ARG(N >= 0, "covar_bwd: N < 0");
    if (N == 0) return UBS_OK;
    UBS_CHECK_ARG(rot && scale && l_triangle && v_covar && v_rot && v_scale && v_l_triangle, "covar_bwd: null pointer");
    const unsigned grid = (unsigned)ceil_div(N, kCondThreads);
    UBS_DISPATCH_D(D, covar_bwd_kernel<kD><<<grid, kCondThreads, 0, (cudaStream_t)stream>>>(
                          N, spatial_block, rot, scale, l_triangle, v_covar, v_rot, v_scale, v_l_triangle));
    UBS_LAUNCH_CHECK("covar_bwd_kernel");
    return UBS_OK;
}

extern "C" int ubs_cond_mean_covar_opacity_fwd(int64_t N, int D, const float *means, const float *covars,
                                               const float *opacities, const float *betas, const float *query,
                                               float *out_means, float *out_covars, float *out_opacities,
                                               void *stream) {
    using namespace ubs;
    UBS_CHECK_ARG(N >= 0, "cond_fwd: N < 0");
    if (N == 0) return UBS_OK;
    UBS_CHECK_ARG(means && covars && opacities && betas && query && out_means && out_covars && out_opacities,
                  "cond_fwd: null pointer");
    const unsigned grid = (unsigned)ceil_div(N, kCondThreads);
    UBS_DISPATCH_D(D, cond_fwd_kernel<kD><<<grid, kCondThreads, 0, (cudaStream_t)stream>>>(
                          N, means, covars, opacities, betas, query, out_means, out_covars, out_opacities));
    UBS_LAUNCH_CHECK("cond_fwd_kernel");
    return UBS_OK;
}

extern "C" int ubs_cond_mean_covar_opacity_bwd(int64_t N, int D, const float *means, const float *covars,
                                               const float *opacities, const float *betas, const float *query,
                                               const float *v_out_means, const float *v_out_covars,
                                               const float *v_out_opacities, float *v_means, float *v_covars,
                                               float *v_opacities, float *v_betas, void *stream) {
    using namespace ubs;
    UBS_CHECK_ARG(N >= 0, "cond_bwd: N < 0");
    if (N == 0) return UBS_OK;
    UBS_CHECK_ARG(means && covars && opacities && betas && query && v_out_means && v_out_covars && v_out_opacities &&
                      v_means && v_covars && v_opacities && v_betas,
                  "cond_bwd: null pointer");
    const unsigned grid = (unsigned)ceil_div(N, kCondThreads);
    UBS_DISPATCH_D(D, cond_bwd_kernel<kD><<<grid, kCondThreads, 0, (cudaStream_t)stream>>>(
                          N, means, covars, opacities, betas, query, v_out_means, v_out_covars, v_out_opacities,
                          v_means, v_covars, v_opacities, v_betas));
    UBS_LAUNCH_CHECK("cond_bwd_kernel");
    return UBS_OK;
}
