// Packing of the reference's separate per-primitive tensors into the [N, stride] record buffer the fused kernels
// read, and the reverse for the gradient records.  Used by the zero-edit drop-in route (ubs_b200/dropin.py): the
// reference caller (scene/beta_model.py:154-159,697-711) hands the operator chain seven separate ACTIVATED tensors; one
// bandwidth-bound pass (read 140 / 176 B, write 144 / 176 B per primitive) turns them into records, and the gradient
// records go back to seven separate arrays the same way.  No reference counterpart (the reference never packs).
//
// One CTA handles 128 consecutive rows: every source tile [128, w] is a contiguous span of global memory and is read
// with fully coalesced loads into a shared [128, stride] tile, which then leaves as coalesced 128-bit stores.
#include "common.cuh"
#include "pack.cuh"

namespace ubs {
namespace {

constexpr int kPackRows = 128;
constexpr int kPackThreads = 256;

template <int D, bool PACK>
__global__ void __launch_bounds__(kPackThreads)
pack_kernel(int64_t N, PackSegs segs, float *__restrict__ records) {
    constexpr int STRIDE = UBS_RECORD_STRIDE(D);
    __shared__ __align__(16) float s_rec[kPackRows * STRIDE];
    const int64_t base = (int64_t)blockIdx.x * kPackRows;
    const int n_here = (int)min((int64_t)kPackRows, N - base);
    float4 *rec4 = reinterpret_cast<float4 *>(records + base * STRIDE);
    float4 *s4 = reinterpret_cast<float4 *>(s_rec);
    if constexpr (PACK) {
        // the padding columns of the record are zero
        for (int i = threadIdx.x; i < n_here; i += kPackThreads)
#pragma unroll
            for (int c = UBS_RECORD_FLOATS(D); c < STRIDE; ++c) s_rec[i * STRIDE + c] = 0.f;
    } else {
        for (int i = threadIdx.x; i < n_here * (STRIDE / 4); i += kPackThreads) s4[i] = rec4[i];
        __syncthreads();
    }
#pragma unroll
    for (int sgm = 0; sgm < 7; ++sgm) {
        const int w = seg_width<D>(sgm), c0 = seg_col<D>(sgm);
        float *g = segs.ptr[sgm];
        if (g == nullptr) continue;  // unpack: the caller does not want this gradient
        g += base * w;
        for (int i = threadIdx.x; i < n_here * w; i += kPackThreads) {
            const int r = i / w, k = i - r * w;
            if constexpr (PACK) s_rec[r * STRIDE + c0 + k] = g[i];
            else g[i] = s_rec[r * STRIDE + c0 + k];
        }
    }
    if constexpr (PACK) {
        __syncthreads();
        for (int i = threadIdx.x; i < n_here * (STRIDE / 4); i += kPackThreads) rec4[i] = s4[i];
    }
}

template <bool PACK>
int launch_pack(int64_t N, int D, const PackSegs &segs, float *records, cudaStream_t s) {
    const unsigned grid = (unsigned)ceil_div(N, kPackRows);
    if (D == 6) pack_kernel<6, PACK><<<grid, kPackThreads, 0, s>>>(N, segs, records);
    else pack_kernel<7, PACK><<<grid, kPackThreads, 0, s>>>(N, segs, records);
    UBS_LAUNCH_CHECK("pack_kernel");
    return UBS_OK;
}

// Separate screen-space gradient arrays -> 48-byte gradient rows in the "direct" form (rows_form = 0 of
// ubs_fused_project_bwd*): one thread per (camera, primitive); a plain streaming pass (40-44 B read, 48 B written).
__global__ void __launch_bounds__(256)
pack_gradient_rows_kernel(int64_t CN, const float *__restrict__ v_means2d, const float *__restrict__ v_depths,
                          const float *__restrict__ v_conics, const float *__restrict__ v_opacities,
                          const float *__restrict__ v_betas, const float *__restrict__ v_colors,
                          float4 *__restrict__ v_rows) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= CN) return;
    float4 q0, q1, q2;
    q0.x = v_colors != nullptr ? v_colors[i * 3 + 0] : 0.f;
    q0.y = v_colors != nullptr ? v_colors[i * 3 + 1] : 0.f;
    q0.z = v_colors != nullptr ? v_colors[i * 3 + 2] : 0.f;
    q0.w = v_conics[i * 3 + 0];
    q1.x = v_conics[i * 3 + 1];
    q1.y = v_conics[i * 3 + 2];
    q1.z = v_means2d[i * 2 + 0];
    q1.w = v_means2d[i * 2 + 1];
    q2.x = v_opacities[i];
    q2.y = v_betas[i];
    q2.z = v_depths != nullptr ? v_depths[i] : 0.f;
    q2.w = 0.f;
    v_rows[i * 3 + 0] = q0;
    v_rows[i * 3 + 1] = q1;
    v_rows[i * 3 + 2] = q2;
}

}  // namespace
}  // namespace ubs

extern "C" int ubs_pack_records(int64_t N, int D, const float *mean, const float *rgb, const float *opacity,
                                const float *beta0, const float *beta_c, const float *scale, const float *l_triangle,
                                float *records, void *stream) {
    using namespace ubs;
    UBS_CHECK_ARG(N >= 0 && (D == 6 || D == 7), "pack_records: N >= 0 and D in {6, 7} (got %lld, %d)", (long long)N, D);
    if (N == 0) return UBS_OK;
    UBS_CHECK_ARG(mean && rgb && opacity && beta0 && beta_c && scale && l_triangle && records,
                  "pack_records: null pointer");
    UBS_CHECK_ARG(((uintptr_t)records & 15) == 0, "pack_records: records must be 16-byte aligned");
    PackSegs segs{{const_cast<float *>(mean), const_cast<float *>(rgb), const_cast<float *>(opacity),
                   const_cast<float *>(beta0), const_cast<float *>(beta_c), const_cast<float *>(scale),
                   const_cast<float *>(l_triangle)}};
    return launch_pack<true>(N, D, segs, records, (cudaStream_t)stream);
}

extern "C" int ubs_unpack_records(int64_t N, int D, const float *records, float *mean, float *rgb, float *opacity,
                                  float *beta0, float *beta_c, float *scale, float *l_triangle, void *stream) {
    using namespace ubs;
    UBS_CHECK_ARG(N >= 0 && (D == 6 || D == 7), "unpack_records: N >= 0 and D in {6, 7} (got %lld, %d)", (long long)N, D);
    if (N == 0) return UBS_OK;
    UBS_CHECK_ARG(records != nullptr, "unpack_records: null pointer");
    UBS_CHECK_ARG(((uintptr_t)records & 15) == 0, "unpack_records: records must be 16-byte aligned");
    PackSegs segs{{mean, rgb, opacity, beta0, beta_c, scale, l_triangle}};
    return launch_pack<false>(N, D, segs, const_cast<float *>(records), (cudaStream_t)stream);
}

extern "C" int ubs_pack_gradient_rows(int64_t CN, const float *v_means2d, const float *v_depths, const float *v_conics,
                                      const float *v_opacities, const float *v_betas, const float *v_colors,
                                      float *v_rows, void *stream) {
    using namespace ubs;
    UBS_CHECK_ARG(CN >= 0, "pack_gradient_rows: bad size");
    if (CN == 0) return UBS_OK;
    UBS_CHECK_ARG(v_means2d && v_conics && v_opacities && v_betas && v_rows, "pack_gradient_rows: null pointer");
    UBS_CHECK_ARG(((uintptr_t)v_rows & 15) == 0, "pack_gradient_rows: v_rows must be 16-byte aligned");
    cudaStream_t s = (cudaStream_t)stream;
    pack_gradient_rows_kernel<<<(unsigned)ceil_div(CN, (int64_t)256), 256, 0, s>>>(
        CN, v_means2d, v_depths, v_conics, v_opacities, v_betas, v_colors, (float4 *)v_rows);
    UBS_LAUNCH_CHECK("pack_gradient_rows_kernel");
    return UBS_OK;
}
