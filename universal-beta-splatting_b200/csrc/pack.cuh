// Segment layout of the packed record (include/ubs_b200.h): the reference's seven per-primitive tensors as column
// ranges of one [N, stride] row.  Shared by the pack / unpack kernels (csrc/pack.cu) and the projection backward's
// unpacked epilogue (csrc/fused_project.cu).
#pragma once
#include "common.cuh"

namespace ubs {

struct PackSegs {
    float *ptr[7];  // mean[N,D] (xyz | conditional mean), rgb[N,3], opacity[N], beta0[N], beta_c[N,D-3], scale[N,D], l_triangle[N,M]
};

// width and first record column of segment `s` (layout of include/ubs_b200.h)
template <int D>
__host__ __device__ constexpr int seg_width(int s) {
    return s == 0 ? D : s == 1 ? 3 : s == 2 ? 1 : s == 3 ? 1 : s == 4 ? D - 3 : s == 5 ? D : D * (D - 1) / 2;
}
template <int D>
__host__ __device__ constexpr int seg_col(int s) {
    return s == 0 ? 0 : s == 1 ? D : s == 2 ? D + 3 : s == 3 ? D + 4 : s == 4 ? D + 5 : s == 5 ? 2 * D + 2 : 3 * D + 2;
}

// [n_here, STRIDE] tile in shared memory -> the seven separate arrays (rows base .. base + n_here), coalesced stores;
// a NULL destination is skipped.  All NT threads of the CTA take part.
template <int D, int NT>
__device__ __forceinline__ void unpack_tile(const float *s_rec, const PackSegs &segs, int64_t base, int n_here) {
    constexpr int STRIDE = UBS_RECORD_STRIDE(D);
#pragma unroll
    for (int sgm = 0; sgm < 7; ++sgm) {
        const int w = seg_width<D>(sgm), c0 = seg_col<D>(sgm);
        float *g = segs.ptr[sgm];
        if (g == nullptr) continue;
        g += base * w;
        for (int i = threadIdx.x; i < n_here * w; i += NT) {
            const int r = i / w, k = i - r * w;
            g[i] = s_rec[r * STRIDE + c0 + k];
        }
    }
}

}  // namespace ubs
