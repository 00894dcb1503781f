// Device math for the N-D part of a Beta primitive: activations, covariance build (K1+K2), conditioning on the
// query (K3) and their VJPs (K4, K2-bwd, K1-bwd).  D (6 or 7 on the UBS path; 4..8 supported) is a template
// parameter, so every matrix is a fully unrolled register array -- no per-thread local-memory tables
// (the reference keeps dynamically indexed MAX_C=8 / int[64] arrays in local memory: 560 B - 3.7 KB per thread).
//
// Accumulation orders follow the reference kernels (rot_scale_l_triangle_to_covar_fwd.cu:28-187,
// cond_mean_convariance_opacity_fwd.cu:135-289) so the conditional mean -- which decides depth bits -- rounds
// the same way under the same compiler flags.
#pragma once
#include <cuda_runtime.h>
#include <float.h>
#include <stdint.h>

namespace ubs {

template <int D>
struct NdDims {
    static constexpr int C = D - 3;            // conditioning dims
    static constexpr int M = D * (D - 1) / 2;  // strictly-lower entries
};

// index of strictly-lower entry (r, k), k < r, in torch.tril_indices(D, D, -1) order
__device__ __host__ constexpr int tril_idx(int r, int k) { return r * (r - 1) / 2 + k; }

// ---- activations (scene/beta_model.py:36-52) ----------------------------------------------------------------
// The reference applies them with torch (F.softplus, torch.sigmoid, 4 * torch.exp), whose CUDA kernels are built
// WITHOUT fast-math: libdevice's full-precision expf / log1pf and IEEE division.  This library is a --use_fast_math
// build (the reference's own kernels are, and the tile lists must match bit for bit), under which `expf` would be
// ex2.approx -- enough to move a radius across an integer boundary for a few primitives in a million.  So the precise
// libdevice entry points are named explicitly: activated values are bit-identical to torch's, and with them every
// integer output of the fused path (radii, tile counts, sorted pair lists) equals the reference chain's.
extern "C" __device__ float __nv_expf(float);
extern "C" __device__ float __nv_log1pf(float);
__device__ __forceinline__ float softplus_f(float x) { return x > 20.f ? x : __nv_log1pf(__nv_expf(x)); }  // F.softplus
__device__ __forceinline__ float sigmoid_f(float x) { return __fdiv_rn(1.f, __fadd_rn(1.f, __nv_expf(-x))); }
__device__ __forceinline__ float beta_act_f(float x) { return __fmul_rn(4.f, __nv_expf(x)); }
// The backward pass recomputes the activated values only to evaluate derivatives (FP32 tolerance, no integer decided
// by them): there the fast-math forms are used -- 35 us of a 0.39 ms kernel at 3M primitives.
__device__ __forceinline__ float softplus_fast(float x) { return x > 20.f ? x : log1pf(expf(x)); }
__device__ __forceinline__ float sigmoid_fast(float x) { return 1.f / (1.f + expf(-x)); }
__device__ __forceinline__ float beta_act_fast(float x) { return 4.f * expf(x); }

// ---- K1 + K2: L and Sigma = L L^T -----------------------------------------------------------------------------
// L is the full D x D lower-triangular factor: L[:3,:3] = R diag(s0..s2) (R = any 3x3, row-major),
// L[r][k] = l_triangle[tril_idx(r,k)] for r >= 3, k < r, L[r][r] = s_r.
template <int D>
__device__ __forceinline__ void build_L(const float R[9], const float s[D], const float lt[NdDims<D>::M], float L[D * D]) {
#pragma unroll
    for (int i = 0; i < D * D; ++i) L[i] = 0.f;
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) L[i * D + j] = R[i * 3 + j] * s[j];
#pragma unroll
    for (int r = 3; r < D; ++r) {
#pragma unroll
        for (int k = 0; k < r; ++k) L[r * D + k] = lt[tril_idx(r, k)];
        L[r * D + r] = s[r];
    }
}

// Sigma (full symmetric D x D, row-major) from L.  Only the first `dim` rows/cols are produced (dim = 3 for the
// spatial block).
template <int D>
__device__ __forceinline__ void covar_from_L(const float L[D * D], const float s[D], float S[D * D], int dim = D) {
    // spatial block
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = i; j < 3; ++j) {
            const float c = L[i * D + 0] * L[j * D + 0] + L[i * D + 1] * L[j * D + 1] + L[i * D + 2] * L[j * D + 2];
            S[i * D + j] = c;
            S[j * D + i] = c;
        }
    if (dim <= 3) return;
    // cross block
#pragma unroll
    for (int c = 3; c < D; ++c)
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            float acc = 0.f;
#pragma unroll
            for (int k = 0; k < 3; ++k) acc += L[i * D + k] * L[c * D + k];
            S[i * D + c] = acc;
            S[c * D + i] = acc;
        }
    // diagonal, rows >= 3
#pragma unroll
    for (int r = 3; r < D; ++r) {
        // the reference rounds s_r^2 before the run-time accumulation loop; keep the compiler from contracting
        // it into the first FMA of the unrolled chain
        float acc = __fmul_rn(s[r], s[r]);
#pragma unroll
        for (int k = 0; k < r; ++k) acc = __fmaf_rn(L[r * D + k], L[r * D + k], acc);
        S[r * D + r] = acc;
    }
    // off-diagonal, rows/cols >= 3
#pragma unroll
    for (int r = 4; r < D; ++r)
#pragma unroll
        for (int c = 3; c < r; ++c) {
            float acc = 0.f;
#pragma unroll
            for (int k = 0; k < c; ++k) acc += L[r * D + k] * L[c * D + k];
            acc += s[c] * L[r * D + c];
            S[r * D + c] = acc;
            S[c * D + r] = acc;
        }
}

// VJP of Sigma = L L^T w.r.t. (R, s, l_triangle): P = (G + G^T) L, masked to the structure of L
// (rot_scale_l_triangle_to_covar_bwd.cu:68-235).  G is dim x dim (dim = 3 or D), row stride `dim`.
template <int D>
__device__ __forceinline__ void covar_vjp(const float R[9], const float s[D], const float L[D * D], const float *G,
                                          int dim, float vR[9], float vs[D], float vlt[NdDims<D>::M]) {
#pragma unroll
    for (int i = 0; i < D; ++i) vs[i] = 0.f;
#pragma unroll
    for (int i = 0; i < NdDims<D>::M; ++i) vlt[i] = 0.f;
    float H[9];
    if (dim <= 3) {
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                float acc = 0.f;
#pragma unroll
                for (int j = 0; j < 3; ++j) acc += (G[i * 3 + j] + G[j * 3 + i]) * L[j * D + k];
                H[i * 3 + k] = acc;
            }
    } else {
        float Ssym[D * D];
#pragma unroll
        for (int i = 0; i < D; ++i)
#pragma unroll
            for (int j = 0; j < D; ++j) Ssym[i * D + j] = G[i * D + j] + G[j * D + i];
        // P = Ssym L ; only lower-triangular entries of P are needed (k <= r)
#pragma unroll
        for (int r = 0; r < D; ++r)
#pragma unroll
            for (int k = 0; k <= (r < 3 ? 2 : r); ++k) {
                float acc = 0.f;
#pragma unroll
                for (int j = (k < 3 ? 0 : k); j < D; ++j) acc += Ssym[r * D + j] * L[j * D + k];  // L[j][k]=0 for j<k (k>=3)
                if (r < 3) H[r * 3 + k] = acc;
                else if (k == r) vs[r] = acc;
                else vlt[tril_idx(r, k)] = acc;
            }
    }
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        float acc = 0.f;
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            vR[i * 3 + j] = H[i * 3 + j] * s[j];
            acc += H[i * 3 + j] * R[i * 3 + j];
        }
        vs[j] = acc;
    }
}

// ---- small dense helpers -----------------------------------------------------------------------------------------
// closed-form 3x3 inverse (cond_mean_convariance_opacity_fwd.cu:76-104)
__device__ __forceinline__ void invert_3x3(const float a[9], float inv[9]) {
    const float c00 = a[4] * a[8] - a[5] * a[7];
    const float c01 = -(a[3] * a[8] - a[5] * a[6]);
    const float c02 = a[3] * a[7] - a[4] * a[6];
    const float c10 = -(a[1] * a[8] - a[2] * a[7]);
    const float c11 = a[0] * a[8] - a[2] * a[6];
    const float c12 = -(a[0] * a[7] - a[1] * a[6]);
    const float c20 = a[1] * a[5] - a[2] * a[4];
    const float c21 = -(a[0] * a[5] - a[2] * a[3]);
    const float c22 = a[0] * a[4] - a[1] * a[3];
    float det = a[0] * c00 + a[1] * c01 + a[2] * c02;
    if (det == 0.f) det = 1e-20f;
    const float invdet = 1.f / det;
    inv[0] = c00 * invdet, inv[1] = c10 * invdet, inv[2] = c20 * invdet;
    inv[3] = c01 * invdet, inv[4] = c11 * invdet, inv[5] = c21 * invdet;
    inv[6] = c02 * invdet, inv[7] = c12 * invdet, inv[8] = c22 * invdet;
}

// Gauss-Jordan with partial pivoting, n = C compile-time (cond_mean_convariance_opacity_fwd.cu:26-74).
// Row swaps are predicated on the (run-time) pivot row so that the augmented matrix stays in registers.
template <int C>
__device__ __forceinline__ void invert_gauss_jordan(const float A[C * C], float inv[C * C]) {
    float aug[C][2 * C];
#pragma unroll
    for (int r = 0; r < C; ++r)
#pragma unroll
        for (int c = 0; c < C; ++c) {
            aug[r][c] = A[r * C + c];
            aug[r][C + c] = (r == c) ? 1.f : 0.f;
        }
#pragma unroll
    for (int col = 0; col < C; ++col) {
        int piv = col;
        float maxabs = fabsf(aug[col][col]);
#pragma unroll
        for (int r = col + 1; r < C; ++r) {
            const float v = fabsf(aug[r][col]);
            if (v > maxabs) {
                maxabs = v;
                piv = r;
            }
        }
#pragma unroll
        for (int r = col + 1; r < C; ++r) {
            if (piv == r) {
#pragma unroll
                for (int c = 0; c < 2 * C; ++c) {
                    const float t = aug[col][c];
                    aug[col][c] = aug[r][c];
                    aug[r][c] = t;
                }
            }
        }
        float diag = aug[col][col];
        if (diag == 0.f) diag = 1e-20f;
        const float invdiag = 1.f / diag;
#pragma unroll
        for (int c = 0; c < 2 * C; ++c) aug[col][c] *= invdiag;
#pragma unroll
        for (int r = 0; r < C; ++r) {
            if (r != col) {
                const float f = aug[r][col];
                if (f != 0.f) {
#pragma unroll
                    for (int c = 0; c < 2 * C; ++c) aug[r][c] -= f * aug[col][c];
                }
            }
        }
    }
#pragma unroll
    for (int r = 0; r < C; ++r)
#pragma unroll
        for (int c = 0; c < C; ++c) inv[r * C + c] = aug[r][C + c];
}

template <int C>
__device__ __forceinline__ void invert_small(const float A[C * C], float inv[C * C]) {
    if constexpr (C == 3) invert_3x3(A, inv);
    else invert_gauss_jordan<C>(A, inv);
}

// lower Cholesky with the reference's guards (cond_mean_convariance_opacity_fwd.cu:250-271)
template <int C>
__device__ __forceinline__ void cholesky_guarded(const float A[C * C], float Lc[C * C]) {
#pragma unroll
    for (int i = 0; i < C * C; ++i) Lc[i] = 0.f;
#pragma unroll
    for (int i = 0; i < C; ++i)
#pragma unroll
        for (int j = 0; j <= i; ++j) {
            float sum = A[i * C + j];
#pragma unroll
            for (int k = 0; k < j; ++k) sum -= Lc[i * C + k] * Lc[j * C + k];
            if (i == j) {
                if (sum <= 0.f) sum = 1e-20f;
                Lc[i * C + j] = sqrtf(sum);
            } else {
                float denom = Lc[j * C + j];
                if (denom == 0.f) denom = 1e-20f;
                Lc[i * C + j] = sum / denom;
            }
        }
}

__device__ __forceinline__ float guard_denom(float d) { return d == 0.f ? 1e-20f : d; }

// ---- K3: conditioning -------------------------------------------------------------------------------------------
template <int C>
struct CondOut {
    float mean[3];
    float cov[9];  // full 3x3, not symmetrised (as the reference)
    float opacity;
};

// Camera-independent part of the conditioning: everything except x = q - mu2.  With many cameras per launch the
// inverse, the regression matrix, the conditional covariance and the Cholesky factor are computed once.
template <int C>
struct CondPrep {
    float rb[3 * C];   // V12 V22^-1 diag(beta_adj)
    float cov[9];      // V11 - rb V21 (full 3x3, not symmetrised, as the reference)
    float Lc[C * C];   // guarded Cholesky factor of V22
};

template <int C>
__device__ __forceinline__ void cond_prepare(const float V11[9], const float V12[3 * C], const float V21[C * 3],
                                             const float V22[C * C], const float beta[C], CondPrep<C> &p) {
    float beta_adj[C];
#pragma unroll
    for (int j = 0; j < C; ++j) {
        const float t = beta[j] * 0.25f;
        beta_adj[j] = t < 1.f ? t : 1.f;
    }
    float i22[C * C];
    invert_small<C>(V22, i22);
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int c = 0; c < C; ++c) {
            float acc = 0.f;
#pragma unroll
            for (int k = 0; k < C; ++k) acc += V12[r * C + k] * i22[k * C + c];
            p.rb[r * C + c] = acc * beta_adj[c];
        }
    float vc[9] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int k = 0; k < C; ++k)
#pragma unroll
        for (int r = 0; r < 3; ++r)
#pragma unroll
            for (int c = 0; c < 3; ++c) vc[r * 3 + c] += p.rb[r * C + k] * V21[k * 3 + c];
#pragma unroll
    for (int i = 0; i < 9; ++i) p.cov[i] = V11[i] - vc[i];
    cholesky_guarded<C>(V22, p.Lc);
}

// Camera-dependent part: conditional mean and opacity for x = q - mu2.
template <int C>
__device__ __forceinline__ void cond_apply(const CondPrep<C> &p, const float mu1[3], const float x[C], float o_in,
                                           const float beta[C], float mean[3], float &opacity) {
    float mc[3] = {0.f, 0.f, 0.f};
#pragma unroll
    for (int c = 0; c < C; ++c) {
        mc[0] += p.rb[0 * C + c] * x[c];
        mc[1] += p.rb[1 * C + c] * x[c];
        mc[2] += p.rb[2 * C + c] * x[c];
    }
#pragma unroll
    for (int r = 0; r < 3; ++r) mean[r] = mu1[r] + mc[r];
    float y[C];
#pragma unroll
    for (int i = 0; i < C; ++i) {
        float sum = x[i];
#pragma unroll
        for (int k = 0; k < i; ++k) sum -= p.Lc[i * C + k] * y[k];
        y[i] = sum / guard_denom(p.Lc[i * C + i]);
    }
    float o_change = 1.f;
    const float upper = 1.f - FLT_EPSILON;
#pragma unroll
    for (int i = 0; i < C; ++i) {
        float d = tanhf(y[i] * y[i]);
        if (d < 0.f) d = 0.f;
        if (d > upper) d = upper;
        o_change *= powf(1.f - d, beta[i]);
    }
    opacity = o_in * o_change;
}

// mu1[3], x = q - mu2 [C], V11[9], V12[3*C], V21[C*3], V22[C*C], o, beta[C]
template <int C>
__device__ __forceinline__ CondOut<C> cond_forward(const float mu1[3], const float x[C], const float V11[9],
                                                   const float V12[3 * C], const float V21[C * 3],
                                                   const float V22[C * C], float o_in, const float beta[C]) {
    CondOut<C> out;
    CondPrep<C> p;
    cond_prepare<C>(V11, V12, V21, V22, beta, p);
    cond_apply<C>(p, mu1, x, o_in, beta, out.mean, out.opacity);
#pragma unroll
    for (int i = 0; i < 9; ++i) out.cov[i] = p.cov[i];
    return out;
}

// ---- K4: VJP of the conditioning (cond_mean_convariance_opacity_bwd.cu:227-562) --------------------------------
// Outputs: g_mu[3+C], g_V (blocks: g11[9], g12[3C], g21[C3], g22[CC]), g_o, g_beta[C].  No gradient for q.
template <int C>
__device__ __forceinline__ void cond_backward(const float x[C], const float V12[3 * C], const float V21[C * 3],
                                              const float V22[C * C], float o_in, const float beta[C],
                                              const float gM[3], const float gV[9], float gO, float g_mu1[3],
                                              float g_mu2[C], float g11[9], float g12[3 * C], float g21[C * 3],
                                              float g22[C * C], float &g_o, float g_beta[C]) {
    float beta_adj[C];
#pragma unroll
    for (int j = 0; j < C; ++j) {
        const float t = beta[j] * 0.25f;
        beta_adj[j] = t < 1.f ? t : 1.f;
    }
    float i22[C * C];
    invert_small<C>(V22, i22);
    float rg[3 * C], rb[3 * C];
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int c = 0; c < C; ++c) {
            float acc = 0.f;
#pragma unroll
            for (int k = 0; k < C; ++k) acc += V12[r * C + k] * i22[k * C + c];
            rg[r * C + c] = acc;
            rb[r * C + c] = acc * beta_adj[c];
        }
#pragma unroll
    for (int r = 0; r < 3; ++r) g_mu1[r] = gM[r];
#pragma unroll
    for (int i = 0; i < 9; ++i) g11[i] = gV[i];

    // Gr = gM (x) x - gV V21^T ;  gx = rb^T gM
    float Gr[3 * C], gx[C];
#pragma unroll
    for (int c = 0; c < C; ++c) {
#pragma unroll
        for (int r = 0; r < 3; ++r)
            Gr[r * C + c] = gM[r] * x[c] - (gV[r * 3 + 0] * V21[c * 3 + 0] + gV[r * 3 + 1] * V21[c * 3 + 1] +
                                            gV[r * 3 + 2] * V21[c * 3 + 2]);
        gx[c] = rb[0 * C + c] * gM[0] + rb[1 * C + c] * gM[1] + rb[2 * C + c] * gM[2];
    }
    // g21 = -rb^T gV
#pragma unroll
    for (int rr = 0; rr < C; ++rr)
#pragma unroll
        for (int c = 0; c < 3; ++c)
            g21[rr * 3 + c] = -(rb[0 * C + rr] * gV[0 * 3 + c] + rb[1 * C + rr] * gV[1 * 3 + c] + rb[2 * C + rr] * gV[2 * 3 + c]);
    // beta_adj and r paths
    float G_r[3 * C];
#pragma unroll
    for (int c = 0; c < C; ++c) {
        const float dL_dba = Gr[0 * C + c] * rg[0 * C + c] + Gr[1 * C + c] * rg[1 * C + c] + Gr[2 * C + c] * rg[2 * C + c];
        g_beta[c] = dL_dba * (beta[c] < 4.f ? 0.25f : 0.f);
#pragma unroll
        for (int r = 0; r < 3; ++r) G_r[r * C + c] = Gr[r * C + c] * beta_adj[c];
    }
    // g12 = G_r i22^T
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int c = 0; c < C; ++c) {
            float acc = 0.f;
#pragma unroll
            for (int k = 0; k < C; ++k) acc += G_r[r * C + k] * i22[c * C + k];
            g12[r * C + c] = acc;
        }
    // Gi22 = V12^T G_r ; g22 = -i22^T Gi22 i22^T
    float Gi[C * C], tmp[C * C];
#pragma unroll
    for (int r = 0; r < C; ++r)
#pragma unroll
        for (int c = 0; c < C; ++c) {
            float acc = 0.f;
#pragma unroll
            for (int k = 0; k < 3; ++k) acc += V12[k * C + r] * G_r[k * C + c];
            Gi[r * C + c] = acc;
        }
#pragma unroll
    for (int r = 0; r < C; ++r)
#pragma unroll
        for (int c = 0; c < C; ++c) {
            float acc = 0.f;
#pragma unroll
            for (int k = 0; k < C; ++k) acc += Gi[r * C + k] * i22[c * C + k];
            tmp[r * C + c] = acc;
        }
#pragma unroll
    for (int r = 0; r < C; ++r)
#pragma unroll
        for (int c = 0; c < C; ++c) {
            float acc = 0.f;
#pragma unroll
            for (int k = 0; k < C; ++k) acc += i22[k * C + r] * tmp[k * C + c];
            g22[r * C + c] = -acc;
        }

    // ---- opacity path -------------------------------------------------------------------------------------------
    float Lc[C * C];
    cholesky_guarded<C>(V22, Lc);
    float y[C];
#pragma unroll
    for (int i = 0; i < C; ++i) {
        float sum = x[i];
#pragma unroll
        for (int k = 0; k < i; ++k) sum -= Lc[i * C + k] * y[k];
        y[i] = sum / guard_denom(Lc[i * C + i]);
    }
    const float upper = 1.f - FLT_EPSILON;
    float d_i[C], base[C], o_change = 1.f;
#pragma unroll
    for (int i = 0; i < C; ++i) {
        float d = tanhf(y[i] * y[i]);
        if (d < 0.f) d = 0.f;
        if (d > upper) d = upper;
        d_i[i] = d;
        base[i] = 1.f - d;
        o_change *= powf(base[i], beta[i]);
    }
    g_o = gO * o_change;
    const float g_oc = gO * o_in;
    float g_y[C];
#pragma unroll
    for (int i = 0; i < C; ++i) {
        float g_d = 0.f;
        if (base[i] > 0.f) {
            g_beta[i] += g_oc * o_change * logf(base[i]);
            if (d_i[i] > 0.f && d_i[i] < upper) g_d = g_oc * (-o_change * beta[i] / base[i]);
        }
        g_y[i] = g_d * (2.f * y[i] * (1.f - d_i[i] * d_i[i]));
    }
    // a = Lc^{-T} g_y
    float a[C];
#pragma unroll
    for (int i = C - 1; i >= 0; --i) {
        float sum = g_y[i];
#pragma unroll
        for (int k = i + 1; k < C; ++k) sum -= Lc[k * C + i] * a[k];
        a[i] = sum / guard_denom(Lc[i * C + i]);
    }
#pragma unroll
    for (int i = 0; i < C; ++i) g_mu2[i] = -(gx[i] + a[i]);
    // gL = -tril(a (x) y); Cholesky backward: S = tril(L^T gL), diag/2; G = L^{-T} S L^{-1}; g22 += (G+G^T)/2
    float Sm[C * C];
#pragma unroll
    for (int r = 0; r < C; ++r)
#pragma unroll
        for (int c = 0; c < C; ++c) {
            float acc = 0.f;
            if (c <= r) {
#pragma unroll
                for (int k = r; k < C; ++k)  // L^T(r,k) = L(k,r) (k >= r); gL(k,c) = -a[k] y[c] for c <= k
                    acc += Lc[k * C + r] * (-(a[k] * y[c]));
                if (c == r) acc *= 0.5f;
            }
            Sm[r * C + c] = acc;
        }
    float U[C * C];  // U = L^{-T} S
#pragma unroll
    for (int col = 0; col < C; ++col)
#pragma unroll
        for (int i = C - 1; i >= 0; --i) {
            float sum = Sm[i * C + col];
#pragma unroll
            for (int k = i + 1; k < C; ++k) sum -= Lc[k * C + i] * U[k * C + col];
            U[i * C + col] = sum / guard_denom(Lc[i * C + i]);
        }
    float Gm[C * C];  // Gm(:,col) = L^{-1} U(col,:)^T  => G = Gm^T ... symmetrised below, so orientation is moot
#pragma unroll
    for (int col = 0; col < C; ++col)
#pragma unroll
        for (int i = 0; i < C; ++i) {
            float sum = U[col * C + i];
#pragma unroll
            for (int k = 0; k < i; ++k) sum -= Lc[i * C + k] * Gm[k * C + col];
            Gm[i * C + col] = sum / guard_denom(Lc[i * C + i]);
        }
#pragma unroll
    for (int r = 0; r < C; ++r)
#pragma unroll
        for (int c = 0; c < C; ++c) g22[r * C + c] += 0.5f * (Gm[r * C + c] + Gm[c * C + r]);
}

}  // namespace ubs
