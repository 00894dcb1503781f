// The step after the backward pass, on the packed [N, stride] buffers:
//   * ubs_adam_step: torch.optim.Adam(eps=1e-15) of scene/beta_model.py:239-268 -- seven parameter groups, one
//     learning rate each -- as ONE bandwidth-bound pass over params / grads / exp_avg / exp_avg_sq with a learning
//     rate per record column (read 4, write 3 record-sized streams).  The opacity and scale regularisers of
//     train.py:122-124 are added to the gradient inside the same pass.
//   * ubs_mcmc_relocate: the deterministic part of relocate_gs / add_new_gs (scene/beta_model.py:548-657): copy the
//     sampled source rows into the destination rows with the opacity rescaled by the sampling multiplicity
//     (_update_params, :548-565), write the new opacity back to the sources and reset their Adam moments
//     (replace_tensors_to_optimizer(inds=...), :512-546).  The sampling itself (torch.multinomial) stays with the
//     caller: it is defined by torch's RNG stream.
#include <math.h>

#include "common.cuh"
#include "optim.cuh"

namespace ubs {
namespace {

__global__ void __launch_bounds__(256)
adam_kernel(int64_t n_vec, int vec_per_row, int64_t row_begin, float4 *__restrict__ params,
            const float4 *__restrict__ grads, float4 *__restrict__ exp_avg, float4 *__restrict__ exp_avg_sq,
            AdamParams a) {
    __shared__ float s_step[kAdamMaxStride];
    if (threadIdx.x < kAdamMaxStride) s_step[threadIdx.x] = a.step_size[threadIdx.x];
    __syncthreads();
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_vec; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t row_local = i / vec_per_row, row = row_begin + row_local;  // `row` = global primitive index
        const int c0 = (int)(i - row_local * vec_per_row) * 4;
        float4 p4 = params[i], g4 = __ldcs(grads + i), m4 = exp_avg[i], v4 = exp_avg_sq[i];
        float p[4] = {p4.x, p4.y, p4.z, p4.w}, g[4] = {g4.x, g4.y, g4.z, g4.w};
        float m[4] = {m4.x, m4.y, m4.z, m4.w}, v[4] = {v4.x, v4.y, v4.z, v4.w};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            g[k] += adam_reg_grad(a, row, c0 + k, p[k]);
            adam_update(a, s_step[c0 + k], p[k], g[k], m[k], v[k]);
        }
        params[i] = make_float4(p[0], p[1], p[2], p[3]);
        exp_avg[i] = make_float4(m[0], m[1], m[2], m[3]);
        exp_avg_sq[i] = make_float4(v[0], v[1], v[2], v[3]);
    }
}

struct PeerRecords {
    float4 *p[UBS_MAX_RANKS];
};

// Owner side of the sharded step: rows [row_begin, row_begin + n_rows) belong to this rank.  Their gradient is the sum
// of the `world` staging slots (fixed order: every rank would compute the same bits), Adam runs on the shard's
// moments, and the new parameters are stored into every rank's record buffer (peer addresses over NVLink).
__global__ void __launch_bounds__(256)
reduce_adam_gather_kernel(int64_t n_vec, int vec_per_row, int64_t row_begin, int world, int64_t slot_vecs,
                          const float4 *__restrict__ staging, float4 *__restrict__ exp_avg,
                          float4 *__restrict__ exp_avg_sq, PeerRecords peers, float4 *__restrict__ mc_records,
                          int rank, AdamParams a) {
    __shared__ float s_step[kAdamMaxStride];
    if (threadIdx.x < kAdamMaxStride) s_step[threadIdx.x] = a.step_size[threadIdx.x];
    __syncthreads();
    const int64_t rec_off = row_begin * vec_per_row;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_vec; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t row_local = i / vec_per_row, row = row_begin + row_local;
        const int c0 = (int)(i - row_local * vec_per_row) * 4;
        float4 g4 = __ldcs(staging + i);
        for (int j = 1; j < world; ++j) {
            const float4 t = __ldcs(staging + j * slot_vecs + i);
            g4.x += t.x, g4.y += t.y, g4.z += t.z, g4.w += t.w;
        }
        float4 p4 = peers.p[rank][rec_off + i], m4 = exp_avg[i], v4 = exp_avg_sq[i];
        float p[4] = {p4.x, p4.y, p4.z, p4.w}, g[4] = {g4.x, g4.y, g4.z, g4.w};
        float m[4] = {m4.x, m4.y, m4.z, m4.w}, v[4] = {v4.x, v4.y, v4.z, v4.w};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            g[k] += adam_reg_grad(a, row, c0 + k, p[k]);
            adam_update(a, s_step[c0 + k], p[k], g[k], m[k], v[k]);
        }
        exp_avg[i] = make_float4(m[0], m[1], m[2], m[3]);
        exp_avg_sq[i] = make_float4(v[0], v[1], v[2], v[3]);
        const float4 out = make_float4(p[0], p[1], p[2], p[3]);
        if (mc_records != nullptr) {
            // one store to the multicast address: the NVSwitch replicates it into every rank's records (this rank's
            // included), so the owner's link carries its shard once instead of world - 1 times
            asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(mc_records + rec_off + i),
                         "f"(out.x), "f"(out.y), "f"(out.z), "f"(out.w)
                         : "memory");
        } else {
            for (int j = 0; j < world; ++j) peers.p[j][rec_off + i] = out;
        }
    }
}

__global__ void count_sources_kernel(int64_t K, const int64_t *__restrict__ src, int32_t *__restrict__ counts) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < K) atomicAdd(counts + src[i], 1);
}

// one warp per relocated row: copies the source record and rescales the opacity
__global__ void __launch_bounds__(256)
relocate_copy_kernel(int64_t K, int stride, int col_opacity, float *__restrict__ records,
                     const int64_t *__restrict__ dst, const int64_t *__restrict__ src,
                     const int32_t *__restrict__ counts) {
    const int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (i >= K) return;
    const int64_t s = src[i], d = dst[i];
    const float *in = records + s * stride;
    float *out = records + d * stride;
    for (int c = lane; c < stride; c += 32) {
        float val = in[c];
        if (c == col_opacity) {
            // new = 1 - (1 - sigmoid(raw))^(1 / (ratio + 1)), clamped to [0.005, 1 - eps], stored as a logit
            // (scene/beta_model.py:548-558); evaluated in FP64 and rounded once
            const double op = (double)(1.f / (1.f + (float)exp(-(double)val)));
            const float expo = 1.0f / (float)(counts[s] + 1);
            float nw = 1.0f - (float)pow((double)(1.0f - (float)op), (double)expo);
            nw = fminf(fmaxf(nw, 0.005f), 1.0f - 1.1920928955078125e-07f);
            val = (float)log((double)nw / (double)(1.0f - nw));
        }
        out[c] = val;
    }
}

// after every copy has been made: sources take the rescaled opacity of (any of) their copies -- all copies of one
// source carry the same value -- and lose their Adam moments
__global__ void __launch_bounds__(256)
relocate_fixup_kernel(int64_t K, int stride, int col_opacity, float *__restrict__ records,
                      float *__restrict__ exp_avg, float *__restrict__ exp_avg_sq, const int64_t *__restrict__ dst,
                      const int64_t *__restrict__ src, int64_t m_begin, int64_t m_count) {
    const int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (i >= K) return;
    const int64_t s = src[i], d = dst[i];
    if (lane == 0) records[s * stride + col_opacity] = records[d * stride + col_opacity];
    // the moment buffers hold rows [m_begin, m_begin + m_count) (all rows, or this rank's shard of the sharded step)
    if (exp_avg != nullptr && s >= m_begin && s < m_begin + m_count)
        for (int c = lane; c < stride; c += 32) {
            exp_avg[(s - m_begin) * stride + c] = 0.f;
            exp_avg_sq[(s - m_begin) * stride + c] = 0.f;
        }
}

// train.py:156-163 (every densification interval): xyz += Sigma_xyz (noise (1 - opacity)^100 noise_lr xyz_lr) with
// Sigma_xyz = get_xyz_covariance = (R diag(s)) (R diag(s))^T, R = I + skew(l_triangle[:3]), s = softplus(scale[:3])
// (rot_scale_l_triangle_to_covar_fwd.cu:28-49, spatial_block = true).  The N(0,1) draw (torch.randn_like) stays with
// the caller: its values are defined by torch's RNG stream.  One thread per primitive; runs once per 100 iterations.
__global__ void __launch_bounds__(256)
sgld_noise_kernel(int64_t N, int stride, int D, float *__restrict__ records, const float *__restrict__ noise,
                  float noise_lr, float xyz_lr) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    float *rec = records + i * stride;
    // the library is built with --use_fast_math: sigmoid, softplus and the 100th power are evaluated in FP64 and
    // rounded once (torch's CUDA kernels are precise-math FP32)
    const float op = (float)(1.0 / (1.0 + exp(-(double)rec[D + 3])));
    const float w = (float)pow((double)(1.f - op), 100.0);
    float s[3], n[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const float raw = rec[2 * D + 2 + k];
        s[k] = raw > 20.f ? raw : (float)log1p(exp((double)raw));
        n[k] = __fmul_rn(__fmul_rn(__fmul_rn(noise[i * 3 + k], w), noise_lr), xyz_lr);
    }
    const float a01 = rec[3 * D + 2], a02 = rec[3 * D + 3], a12 = rec[3 * D + 4];
    const float R[9] = {1.f, a01, a02, -a01, 1.f, a12, -a02, -a12, 1.f};
    float L[9];
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int c = 0; c < 3; ++c) L[r * 3 + c] = __fmul_rn(R[r * 3 + c], s[c]);
#pragma unroll
    for (int r = 0; r < 3; ++r) {
        float acc = 0.f;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const float cov = __fmaf_rn(L[r * 3 + 2], L[c * 3 + 2], __fmaf_rn(L[r * 3 + 1], L[c * 3 + 1], __fmul_rn(L[r * 3], L[c * 3])));
            acc = __fmaf_rn(cov, n[c], acc);
        }
        rec[r] += acc;
    }
}

}  // namespace
}  // namespace ubs

extern "C" int ubs_adam_step(int64_t N, int D, int64_t row_begin, int64_t row_count, float *records,
                             const float *grads, float *exp_avg, float *exp_avg_sq, const double *h_lr, double beta1, double beta2, double eps, int64_t step,
                             double opacity_reg, double scale_reg, void *stream) {
    using namespace ubs;
    UBS_CHECK_ARG(N >= 0 && D >= 4 && D <= 8, "adam_step: bad sizes (N=%lld, D=%d)", (long long)N, D);
    UBS_CHECK_ARG(row_begin >= 0 && row_count >= 0 && row_begin + row_count <= N, "adam_step: rows [%lld, +%lld) outside N=%lld",
                  (long long)row_begin, (long long)row_count, (long long)N);
    if (N == 0 || row_count == 0) return UBS_OK;
    UBS_CHECK_ARG(records && grads && exp_avg && exp_avg_sq && h_lr, "adam_step: null pointer");
    UBS_CHECK_ARG(step >= 1, "adam_step: step counts from 1 (got %lld)", (long long)step);
    const int stride = UBS_RECORD_STRIDE(D);
    UBS_CHECK_ARG(stride <= kAdamMaxStride, "adam_step: stride %d exceeds %d", stride, kAdamMaxStride);
    const AdamParams a = make_adam_params(N, D, h_lr, beta1, beta2, eps, step, opacity_reg, scale_reg);
    const int64_t n_vec = row_count * (stride / 4);
    const size_t off = (size_t)row_begin * stride;
    int sm = 148;
    {
        int dev = 0;
        UBS_CUDA_TRY(cudaGetDevice(&dev));
        UBS_CUDA_TRY(cudaDeviceGetAttribute(&sm, cudaDevAttrMultiProcessorCount, dev));
    }
    const int64_t blocks = ceil_div(n_vec, 256);
    const unsigned grid = (unsigned)(blocks < (int64_t)sm * 16 ? blocks : (int64_t)sm * 16);
    adam_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(n_vec, stride / 4, row_begin, (float4 *)(records + off),
                                                        (const float4 *)(grads + off), (float4 *)(exp_avg + off),
                                                        (float4 *)(exp_avg_sq + off), a);
    UBS_LAUNCH_CHECK("adam_kernel");
    return UBS_OK;
}

extern "C" int ubs_mcmc_relocate(int64_t N, int D, float *records, float *exp_avg, float *exp_avg_sq,
                                 int64_t moment_row_begin, int64_t moment_row_count, int64_t K,
                                 const int64_t *dst_idx, const int64_t *src_idx, int32_t *counts, void *stream) {
    using namespace ubs;
    UBS_CHECK_ARG(N >= 0 && K >= 0 && D >= 4 && D <= 8, "mcmc_relocate: bad sizes");
    if (K == 0 || N == 0) return UBS_OK;
    UBS_CHECK_ARG(records && dst_idx && src_idx && counts, "mcmc_relocate: null pointer");
    UBS_CHECK_ARG((exp_avg == nullptr) == (exp_avg_sq == nullptr), "mcmc_relocate: give both moment buffers or none");
    cudaStream_t s = (cudaStream_t)stream;
    const int stride = UBS_RECORD_STRIDE(D), col_opacity = D + 3;
    UBS_CUDA_TRY(cudaMemsetAsync(counts, 0, (size_t)N * sizeof(int32_t), s));
    count_sources_kernel<<<(unsigned)ceil_div(K, 256), 256, 0, s>>>(K, src_idx, counts);
    UBS_LAUNCH_CHECK("count_sources_kernel");
    const unsigned grid = (unsigned)ceil_div(K * 32, 256);
    relocate_copy_kernel<<<grid, 256, 0, s>>>(K, stride, col_opacity, records, dst_idx, src_idx, counts);
    UBS_LAUNCH_CHECK("relocate_copy_kernel");
    relocate_fixup_kernel<<<grid, 256, 0, s>>>(K, stride, col_opacity, records, exp_avg, exp_avg_sq, dst_idx, src_idx,
                                               moment_row_begin, moment_row_count);
    UBS_LAUNCH_CHECK("relocate_fixup_kernel");
    return UBS_OK;
}

extern "C" int ubs_reduce_adam_gather(int64_t N, int D, int world, int rank, int64_t shard_rows, const float *staging,
                                      float *exp_avg_shard, float *exp_avg_sq_shard, float *const *h_peer_records,
                                      float *mc_records, const double *h_lr, double beta1, double beta2, double eps, int64_t step,
                                      double opacity_reg, double scale_reg, void *stream) {
    using namespace ubs;
    UBS_CHECK_ARG(N >= 0 && D >= 4 && D <= 8, "reduce_adam_gather: bad sizes (N=%lld, D=%d)", (long long)N, D);
    UBS_CHECK_ARG(world >= 1 && world <= UBS_MAX_RANKS && rank >= 0 && rank < world,
                  "reduce_adam_gather: rank %d of %d (at most %d ranks)", rank, world, UBS_MAX_RANKS);
    UBS_CHECK_ARG(shard_rows > 0 && shard_rows * world >= N, "reduce_adam_gather: shards do not cover N");
    UBS_CHECK_ARG(step >= 1, "reduce_adam_gather: step counts from 1 (got %lld)", (long long)step);
    UBS_CHECK_ARG(staging && exp_avg_shard && exp_avg_sq_shard && h_peer_records && h_lr,
                  "reduce_adam_gather: null pointer");
    const int64_t row_begin = (int64_t)rank * shard_rows;
    const int64_t n_rows = N - row_begin < shard_rows ? N - row_begin : shard_rows;
    if (n_rows <= 0) return UBS_OK;  // this rank's shard is all padding
    const int stride = UBS_RECORD_STRIDE(D);
    PeerRecords peers{};
    for (int g = 0; g < world; ++g) {
        UBS_CHECK_ARG(h_peer_records[g] != nullptr, "reduce_adam_gather: peer_records[%d] is null", g);
        peers.p[g] = (float4 *)h_peer_records[g];
    }
    const AdamParams a = make_adam_params(N, D, h_lr, beta1, beta2, eps, step, opacity_reg, scale_reg);
    const int64_t n_vec = n_rows * (stride / 4);
    int sm = 148;
    {
        int dev = 0;
        UBS_CUDA_TRY(cudaGetDevice(&dev));
        UBS_CUDA_TRY(cudaDeviceGetAttribute(&sm, cudaDevAttrMultiProcessorCount, dev));
    }
    const int64_t blocks = ceil_div(n_vec, 256);
    const unsigned grid = (unsigned)(blocks < (int64_t)sm * 16 ? blocks : (int64_t)sm * 16);
    reduce_adam_gather_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(
        n_vec, stride / 4, row_begin, world, shard_rows * (stride / 4), (const float4 *)staging, (float4 *)exp_avg_shard,
        (float4 *)exp_avg_sq_shard, peers, (float4 *)mc_records, rank, a);
    UBS_LAUNCH_CHECK("reduce_adam_gather_kernel");
    return UBS_OK;
}

extern "C" int ubs_sgld_noise(int64_t N, int D, float *records, const float *noise, double noise_lr, double xyz_lr,
                              void *stream) {
    using namespace ubs;
    UBS_CHECK_ARG(N >= 0 && D >= 4 && D <= 8, "sgld_noise: bad sizes (N=%lld, D=%d)", (long long)N, D);
    if (N == 0) return UBS_OK;
    UBS_CHECK_ARG(records && noise, "sgld_noise: null pointer");
    sgld_noise_kernel<<<(unsigned)ceil_div(N, 256), 256, 0, (cudaStream_t)stream>>>(N, UBS_RECORD_STRIDE(D), D, records, noise,
                                                                            (float)noise_lr, (float)xyz_lr);
    UBS_LAUNCH_CHECK("sgld_noise_kernel");
    return UBS_OK;
}
