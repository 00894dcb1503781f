// Photometric loss of the training step, forward and backward in two launches:
//     loss = (1 - lambda) * mean|img - gt| + lambda * (1 - mean(SSIM(img, gt)))
// (train.py:118-121; utils/loss_utils.py:18-19,45-85; the reference calls the third-party fused_ssim CUDA package
// -- rahul-goel/fused-ssim@1272e21, absent from the reference tree -- whose published algorithm is the same SSIM:
// 11x11 Gaussian window, sigma 1.5, zero "same" padding, C1 = 0.01^2, C2 = 0.03^2, mean over all elements).
//
// Both kernels fetch the next channel's input tiles into registers while the current channel is convolved (the
// staging loads were the exposed latency: long-scoreboard 42 % / 63 % of the stall samples before, 279 -> 169 us).
// Kernel 1 (ssim_fwd): per 32x32 tile and channel, the five windowed moments E[x], E[x^2], E[y], E[y^2], E[xy] by a
// separable convolution out of shared memory with register sliding windows (each thread produces a strip of
// outputs from one strip of loads), then the SSIM value and its three partial derivatives w.r.t. the moments of x
// (the maps the backward needs); block-reduced partial sums of |x - y| and SSIM go to two doubles.
// Kernel 2 (ssim_bwd): convolves the three derivative maps with the same (symmetric) window -- the adjoint of a
// zero-padded "same" correlation -- and combines them with the L1 sign term into d loss / d img, written in the
// layout of the rendered image so that it is the compositing backward's `v_render_colors` without a copy.
#include <math.h>

#include "common.cuh"

namespace ubs {
namespace {

__device__ __forceinline__ float rcp_1ulp(float b) {  // b > 0, normal
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(b));
    return fmaf(fmaf(-b, r, 1.f), r, r);
}


constexpr int kR = 5;              // window radius (11 taps)
constexpr int kTX = 32, kTY = 32;  // outputs per CTA
constexpr int kIN = kTX + 2 * kR;  // 42 staged rows / columns
constexpr int kThreads = 256;

struct Window {
    float g[2 * kR + 1];
};

struct ImageView {  // element (n, c, y, x) at base[n * sn + c * sc + y * sy + x * sx]
    const float *base;
    int64_t sn, sc, sy, sx;
};

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// stage a kIN x kIN tile (zero outside the image) of one (camera, channel) plane: the global loads go to registers
// first (issued a whole phase before they are needed) and to shared memory later
constexpr int kStageRegs = (kIN * kIN + kThreads - 1) / kThreads;  // 7

__device__ __forceinline__ void tile_load(float (&r)[kStageRegs], const float *plane, int64_t sy, int64_t sx, int H,
                                          int W, int y0, int x0) {
#pragma unroll
    for (int j = 0; j < kStageRegs; ++j) {
        const int idx = threadIdx.x + j * kThreads;
        const int row = idx / kIN, c = idx - row * kIN;
        const int gy = y0 - kR + row, gx = x0 - kR + c;
        r[j] = (idx < kIN * kIN && gy >= 0 && gy < H && gx >= 0 && gx < W) ? plane[gy * sy + gx * sx] : 0.f;
    }
}
__device__ __forceinline__ void tile_store(float (*dst)[kIN + 1], const float (&r)[kStageRegs]) {
#pragma unroll
    for (int j = 0; j < kStageRegs; ++j) {
        const int idx = threadIdx.x + j * kThreads;
        const int row = idx / kIN, c = idx - row * kIN;
        if (idx < kIN * kIN) dst[row][c] = r[j];
    }
}

// two CTAs per SM: the prefetch registers do not fit the 80-register budget of three (148 B of spills, 175 us against
// 169 us for loss + gradient at 1080p)
__global__ void __launch_bounds__(kThreads, 2)
ssim_fwd_kernel(int CH, int H, int W, ImageView img, ImageView gt, Window win, float *__restrict__ maps,
                double *__restrict__ sums) {
    __shared__ float s_x[kIN][kIN + 1], s_y[kIN][kIN + 1];
    __shared__ float s_h[5][kIN][kTX + 1];
    __shared__ float s_red[2][kThreads / 32];
    const int cam = blockIdx.z, x0 = blockIdx.x * kTX, y0 = blockIdx.y * kTY;
    const int tid = threadIdx.x;
    const size_t plane_elems = (size_t)H * W;
    float l1_acc = 0.f, ssim_acc = 0.f;
    constexpr float C1 = 0.01f * 0.01f, C2 = 0.03f * 0.03f;

    float rx[kStageRegs], ry[kStageRegs];
    tile_load(rx, img.base + cam * img.sn, img.sy, img.sx, H, W, y0, x0);
    tile_load(ry, gt.base + cam * gt.sn, gt.sy, gt.sx, H, W, y0, x0);
    for (int ch = 0; ch < CH; ++ch) {
        __syncthreads();  // previous channel's vertical pass has finished reading s_h / s_x / s_y
        tile_store(s_x, rx);
        tile_store(s_y, ry);
        __syncthreads();
        if (ch + 1 < CH) {  // next channel's tile: in flight during both passes of this one
            tile_load(rx, img.base + cam * img.sn + (ch + 1) * img.sc, img.sy, img.sx, H, W, y0, x0);
            tile_load(ry, gt.base + cam * gt.sn + (ch + 1) * gt.sc, gt.sy, gt.sx, H, W, y0, x0);
        }
        // horizontal pass: item = (row, strip of 8 output columns); 18 loads per image feed 8 outputs x 11 taps
        if (tid < kIN * (kTX / 8)) {
            const int row = tid % kIN, c0 = (tid / kIN) * 8;
            float xs[8 + 2 * kR], ys[8 + 2 * kR];
#pragma unroll
            for (int k = 0; k < 8 + 2 * kR; ++k) {
                xs[k] = s_x[row][c0 + k];
                ys[k] = s_y[row][c0 + k];
            }
#pragma unroll
            for (int o = 0; o < 8; ++o) {
                float m1 = 0.f, m11 = 0.f, m2 = 0.f, m22 = 0.f, m12 = 0.f;
#pragma unroll
                for (int k = 0; k < 2 * kR + 1; ++k) {
                    const float w = win.g[k], wx = w * xs[o + k], wy = w * ys[o + k];
                    m1 += wx;
                    m2 += wy;
                    m11 = fmaf(wx, xs[o + k], m11);
                    m22 = fmaf(wy, ys[o + k], m22);
                    m12 = fmaf(wx, ys[o + k], m12);
                }
                s_h[0][row][c0 + o] = m1;
                s_h[1][row][c0 + o] = m11;
                s_h[2][row][c0 + o] = m2;
                s_h[3][row][c0 + o] = m22;
                s_h[4][row][c0 + o] = m12;
            }
        }
        __syncthreads();
        // vertical pass: item = (column, strip of 4 output rows); lanes = consecutive columns
        {
            const int col = tid & 31, r0 = (tid >> 5) * 4;
            float acc[5][4];
#pragma unroll
            for (int q = 0; q < 5; ++q) {
                float v[4 + 2 * kR];
#pragma unroll
                for (int k = 0; k < 4 + 2 * kR; ++k) v[k] = s_h[q][r0 + k][col];
#pragma unroll
                for (int o = 0; o < 4; ++o) {
                    float a = 0.f;
#pragma unroll
                    for (int k = 0; k < 2 * kR + 1; ++k) a = fmaf(win.g[k], v[o + k], a);
                    acc[q][o] = a;
                }
            }
            const int gx = x0 + col;
#pragma unroll
            for (int o = 0; o < 4; ++o) {
                const int gy = y0 + r0 + o;
                if (gx < W && gy < H) {
                    const float mu1 = acc[0][o], mu2 = acc[2][o];
                    const float mu1_sq = mu1 * mu1, mu2_sq = mu2 * mu2, mu12 = mu1 * mu2;
                    const float sig1 = acc[1][o] - mu1_sq, sig2 = acc[3][o] - mu2_sq, sig12 = acc[4][o] - mu12;
                    const float A1 = 2.f * mu12 + C1, A2 = 2.f * sig12 + C2;
                    const float B1 = mu1_sq + mu2_sq + C1, B2 = sig1 + sig2 + C2;
                    // B1 >= C1 and B2 >= C2 - rounding: well inside the normal range, so MUFU.RCP + one Newton step
                    // (<= 1 ulp, 4 instructions) replaces the IEEE division and its slow-path checks
                    const float inv_b1 = rcp_1ulp(B1), inv_b2 = rcp_1ulp(B2);
                    const float S = (A1 * A2) * (inv_b1 * inv_b2);
                    ssim_acc += S;
                    l1_acc += fabsf(s_x[r0 + o + kR][col + kR] - s_y[r0 + o + kR][col + kR]);
                    if (maps != nullptr) {
                        // dS/dmu1 (with sigma1^2 = E[x^2] - mu1^2 and sigma12 = E[xy] - mu1 mu2), dS/dE[x^2], dS/dE[xy]
                        const float d_mu1 = 2.f * mu2 * (A2 - A1) * (inv_b1 * inv_b2) - 2.f * mu1 * S * (inv_b1 - inv_b2);
                        const float d_x2 = -S * inv_b2;
                        const float d_xy = 2.f * A1 * (inv_b1 * inv_b2);
                        const size_t o_idx = ((size_t)(cam * CH + ch) * 3) * plane_elems + (size_t)gy * W + gx;
                        maps[o_idx] = d_mu1;
                        maps[o_idx + plane_elems] = d_x2;
                        maps[o_idx + 2 * plane_elems] = d_xy;
                    }
                }
            }
        }
    }
    l1_acc = warp_sum(l1_acc);
    ssim_acc = warp_sum(ssim_acc);
    if ((tid & 31) == 0) {
        s_red[0][tid >> 5] = l1_acc;
        s_red[1][tid >> 5] = ssim_acc;
    }
    __syncthreads();
    if (tid < 2) {
        double t = 0.0;
        for (int w = 0; w < kThreads / 32; ++w) t += (double)s_red[tid][w];
        atomicAdd(sums + tid, t);
    }
}

__global__ void __launch_bounds__(kThreads, 4)
ssim_bwd_kernel(int CH, int H, int W, ImageView img, ImageView gt, Window win, const float *__restrict__ maps,
                float l1_coeff, float ssim_coeff, float *__restrict__ v_img) {
    __shared__ float s_m[3][kIN][kIN + 1];
    __shared__ float s_h[3][kIN][kTX + 1];
    const int cam = blockIdx.z, x0 = blockIdx.x * kTX, y0 = blockIdx.y * kTY;
    const int tid = threadIdx.x;
    const size_t plane_elems = (size_t)H * W;
    const int col = tid & 31, r0 = (tid >> 5) * 4;
    const int gx = x0 + col;

    float rm[3][kStageRegs];
#pragma unroll
    for (int q = 0; q < 3; ++q)
        tile_load(rm[q], maps + ((size_t)(cam * CH) * 3 + q) * plane_elems, (int64_t)W, 1, H, W, y0, x0);
    for (int ch = 0; ch < CH; ++ch) {
        __syncthreads();  // previous channel's passes have finished reading s_m / s_h
#pragma unroll
        for (int q = 0; q < 3; ++q) tile_store(s_m[q], rm[q]);
        __syncthreads();
        if (ch + 1 < CH) {  // the next channel's three maps: in flight during this channel's passes
#pragma unroll
            for (int q = 0; q < 3; ++q)
                tile_load(rm[q], maps + ((size_t)(cam * CH + ch + 1) * 3 + q) * plane_elems, (int64_t)W, 1, H, W, y0, x0);
        }
        // the pixel's own values for the epilogue, also early
        float xv[4], yv[4];
        {
            const float *xp = img.base + cam * img.sn + ch * img.sc;
            const float *yp = gt.base + cam * gt.sn + ch * gt.sc;
#pragma unroll
            for (int o = 0; o < 4; ++o) {
                const int gy = y0 + r0 + o;
                const bool in = gx < W && gy < H;
                xv[o] = in ? xp[gy * img.sy + gx * img.sx] : 0.f;
                yv[o] = in ? yp[gy * gt.sy + gx * gt.sx] : 0.f;
            }
        }
        // horizontal pass over the three maps: item = (map, row, strip of 8 columns)
        for (int item = tid; item < 3 * kIN * (kTX / 8); item += kThreads) {
            const int q = item / (kIN * (kTX / 8)), rem = item - q * (kIN * (kTX / 8));
            const int row = rem % kIN, c0 = (rem / kIN) * 8;
            float v[8 + 2 * kR];
#pragma unroll
            for (int k = 0; k < 8 + 2 * kR; ++k) v[k] = s_m[q][row][c0 + k];
#pragma unroll
            for (int o = 0; o < 8; ++o) {
                float a = 0.f;
#pragma unroll
                for (int k = 0; k < 2 * kR + 1; ++k) a = fmaf(win.g[k], v[o + k], a);
                s_h[q][row][c0 + o] = a;
            }
        }
        __syncthreads();
        float conv[3][4];
#pragma unroll
        for (int q = 0; q < 3; ++q) {
            float v[4 + 2 * kR];
#pragma unroll
            for (int k = 0; k < 4 + 2 * kR; ++k) v[k] = s_h[q][r0 + k][col];
#pragma unroll
            for (int o = 0; o < 4; ++o) {
                float a = 0.f;
#pragma unroll
                for (int k = 0; k < 2 * kR + 1; ++k) a = fmaf(win.g[k], v[o + k], a);
                conv[q][o] = a;
            }
        }
        if (gx < W) {
#pragma unroll
            for (int o = 0; o < 4; ++o) {
                const int gy = y0 + r0 + o;
                if (gy < H) {
                    const float x = xv[o], y = yv[o];
                    const float d = x - y;
                    const float sgn = d > 0.f ? 1.f : (d < 0.f ? -1.f : 0.f);
                    const float ds = conv[0][o] + 2.f * x * conv[1][o] + y * conv[2][o];
                    v_img[cam * img.sn + ch * img.sc + gy * img.sy + gx * img.sx] = l1_coeff * sgn + ssim_coeff * ds;
                }
            }
        }
    }
}

__global__ void loss_finalize_kernel(const double *__restrict__ sums, double inv_count, float lambda,
                                     float *__restrict__ loss_out) {
    const double l1 = sums[0] * inv_count, ssim = sums[1] * inv_count;
    loss_out[0] = (float)l1;
    loss_out[1] = (float)ssim;
    loss_out[2] = (float)((1.0 - (double)lambda) * l1 + (double)lambda * (1.0 - ssim));
}

}  // namespace
}  // namespace ubs

extern "C" size_t ubs_l1_ssim_workspace_bytes(int C, int channels, int height, int width) {
    return (size_t)3 * (size_t)C * channels * height * width * sizeof(float) + 2 * sizeof(double);
}

extern "C" int ubs_l1_ssim_loss(int C, int channels, int height, int width, const float *img, int64_t img_sn,
                                int64_t img_sc, int64_t img_sy, int64_t img_sx, const float *gt, int64_t gt_sn,
                                int64_t gt_sc, int64_t gt_sy, int64_t gt_sx, float lambda_dssim, float grad_scale,
                                float *loss_out, float *v_img, void *workspace, size_t workspace_bytes, void *stream) {
    using namespace ubs;
    UBS_CHECK_ARG(C >= 0 && channels > 0 && height > 0 && width > 0, "l1_ssim_loss: bad sizes");
    if (C == 0) return UBS_OK;
    UBS_CHECK_ARG(img && gt && workspace, "l1_ssim_loss: null pointer");
    UBS_CHECK_ARG(loss_out || v_img, "l1_ssim_loss: neither loss_out nor v_img requested");
    UBS_CHECK_ARG(C <= 65535, "l1_ssim_loss: C=%d exceeds 65535", C);
    const size_t need = v_img ? ubs_l1_ssim_workspace_bytes(C, channels, height, width) : 2 * sizeof(double);
    if (workspace_bytes < need) {
        set_error("l1_ssim_loss: workspace of %zu bytes, %zu needed", workspace_bytes, need);
        return UBS_ENOSPC;
    }
    cudaStream_t s = (cudaStream_t)stream;
    // workspace: [2 doubles: sum |x - y|, sum SSIM][3 maps of C*channels*H*W floats]
    double *sums = (double *)workspace;
    float *maps = v_img ? (float *)(sums + 2) : nullptr;
    UBS_CUDA_TRY(cudaMemsetAsync(sums, 0, 2 * sizeof(double), s));
    Window win;
    {
        // utils/loss_utils.py:26-33: the taps are FP32 roundings of exp(-(k-5)^2 / (2 sigma^2)), divided by their
        // FP32 sum
        float g[2 * kR + 1], tot = 0.f;
        for (int k = 0; k <= 2 * kR; ++k) {
            g[k] = (float)exp(-(double)((k - kR) * (k - kR)) / (2.0 * 1.5 * 1.5));
            tot += g[k];
        }
        for (int k = 0; k <= 2 * kR; ++k) win.g[k] = g[k] / tot;
    }
    const ImageView iv{img, img_sn, img_sc, img_sy, img_sx}, gv{gt, gt_sn, gt_sc, gt_sy, gt_sx};
    const dim3 grid((unsigned)ceil_div(width, kTX), (unsigned)ceil_div(height, kTY), (unsigned)C);
    ssim_fwd_kernel<<<grid, kThreads, 0, s>>>(channels, height, width, iv, gv, win, maps, sums);
    UBS_LAUNCH_CHECK("ssim_fwd_kernel");
    const double count = (double)C * channels * height * width;
    if (loss_out) {
        loss_finalize_kernel<<<1, 1, 0, s>>>(sums, 1.0 / count, lambda_dssim, loss_out);
        UBS_LAUNCH_CHECK("loss_finalize_kernel");
    }
    if (v_img) {
        const float l1_coeff = (float)((double)grad_scale * (1.0 - (double)lambda_dssim) / count);
        const float ssim_coeff = (float)(-(double)grad_scale * (double)lambda_dssim / count);
        ssim_bwd_kernel<<<grid, kThreads, 0, s>>>(channels, height, width, iv, gv, win, maps, l1_coeff, ssim_coeff,
                                                  v_img);
        UBS_LAUNCH_CHECK("ssim_bwd_kernel");
    }
    return UBS_OK;
}
