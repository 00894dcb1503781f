// Adam update shared by the stand-alone pass (optim.cu) and the projection-backward epilogue (fused_project.cu).
#pragma once
#include <math.h>

#include "common.cuh"

namespace ubs {

constexpr int kAdamMaxStride = 64;

struct AdamParams {
    float step_size[kAdamMaxStride];  // lr[col] / (1 - beta1^t); 0 for padding columns
    float w1, beta2, w2;              // 1 - beta1, beta2, 1 - beta2
    float inv_bc2_sqrt, eps;          // 1 / sqrt(1 - beta2^t) rounded to FP32, eps
    float reg_opacity, reg_scale;     // regulariser coefficients already divided by their mean's element count
    int col_opacity, col_scale, D;
};

// torch/optim/adam.py (_single_tensor_adam): python-double scalars, rounded to FP32 where they meet a tensor.
static inline AdamParams make_adam_params(int64_t N, int D, const double *h_lr, double beta1, double beta2, double eps,
                                          int64_t step, double opacity_reg, double scale_reg) {
    AdamParams a;
    const int stride = UBS_RECORD_STRIDE(D), n_floats = UBS_RECORD_FLOATS(D);
    const double bc1 = 1.0 - pow(beta1, (double)step), bc2 = 1.0 - pow(beta2, (double)step);
    for (int c = 0; c < kAdamMaxStride; ++c) a.step_size[c] = c < n_floats ? (float)(h_lr[c] / bc1) : 0.f;
    (void)stride;
    a.w1 = (float)(1.0 - beta1);
    a.beta2 = (float)beta2;
    a.w2 = (float)(1.0 - beta2);
    a.inv_bc2_sqrt = 1.0f / (float)sqrt(bc2);
    a.eps = (float)eps;
    a.reg_opacity = (float)(opacity_reg / (double)N);                    // mean over [N,1]
    a.reg_scale = (float)(scale_reg / (double)((N < 3 ? N : 3) * D));    // mean over get_scale[:3] = [3,D]
    a.col_opacity = D + 3;
    a.col_scale = 2 * D + 2;
    a.D = D;
    return a;
}

// the library is compiled with --use_fast_math (expf -> ex2.approx); the regulariser touches one column per row, so
// its sigmoid is evaluated in FP64 and rounded once
static __device__ __noinline__ float precise_sigmoid(float x) { return (float)(1.0 / (1.0 + exp(-(double)x))); }

// gradient of the regularisers of train.py:122-124 w.r.t. the raw parameter p in column c of row `row`
__device__ __forceinline__ float adam_reg_grad(const AdamParams &a, int64_t row, int c, float p) {
    float g = 0.f;
    if (a.reg_opacity != 0.f && c == a.col_opacity) {
        // d/d raw of reg * mean(|sigmoid(raw)|): sigmoid > 0, so the |.| passes the derivative through
        const float sg = precise_sigmoid(p);
        g += a.reg_opacity * sg * (1.f - sg);
    }
    if (a.reg_scale != 0.f && row < 3 && c >= a.col_scale && c < a.col_scale + a.D) {
        // train.py:124 regularises get_scale[:3] -- the first three PRIMITIVES, all D scales (reproduced);
        // d softplus(raw) / d raw = sigmoid(raw)
        g += a.reg_scale * precise_sigmoid(p);
    }
    return g;
}

// one element, in the operation order of torch's CUDA Adam (torch/optim/adam.py _single_tensor_adam + the
// elementwise CUDA kernels it dispatches to), split in two so that a caller holding the gradient in registers can
// run the cheap first half unrolled and the expensive second half as a rolled loop
__device__ __forceinline__ void adam_moments(const AdamParams &a, float g, float &m, float &v) {
    m = fmaf(a.w1, g - m, m);            // exp_avg.lerp_(grad, 1 - beta1)
    v = fmaf(a.w2 * g, g, v * a.beta2);  // exp_avg_sq.mul_(beta2).addcmul_(grad, grad, value=1 - beta2)
}
// sqrt and division to within 1 ulp without the IEEE instructions' range-check branches: a zero gradient history
// (m = v = 0: every primitive no camera has seen yet, and the padding columns) sends div.rn / sqrt.rn down their
// slow paths, which a warp then executes for all of its lanes -- measured 100 instructions per element.
__device__ __forceinline__ float sqrt_1ulp(float v) {  // v >= 0
    float rs;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(rs) : "f"(v));
    const float s = v * rs;
    const float s2 = fmaf(fmaf(-s, s, v), 0.5f * rs, s);  // one Newton step
    return v > 0.f ? s2 : 0.f;
}
__device__ __forceinline__ float div_1ulp(float a, float b) {  // b > 0, well inside the normal range
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(b));
    const float q = a * r;
    return fmaf(fmaf(-b, q, a), r, q);  // one Newton step on the quotient
}
__device__ __forceinline__ float adam_apply(const AdamParams &a, float step_size, float p, float m, float v) {
    // (exp_avg_sq.sqrt() / bias_correction2_sqrt).add_(eps): a tensor divided by a python scalar is a multiplication
    // by its FP32 reciprocal in torch's CUDA kernel
    const float denom = __fadd_rn(__fmul_rn(sqrt_1ulp(v), a.inv_bc2_sqrt), a.eps);
    return fmaf(-step_size, div_1ulp(m, denom), p);  // param.addcdiv_(exp_avg, denom, value=-step_size)
}
__device__ __forceinline__ void adam_update(const AdamParams &a, float step_size, float &p, float g, float &m,
                                            float &v) {
    adam_moments(a, g, m, v);
    p = adam_apply(a, step_size, p, m, v);
}

}  // namespace ubs
