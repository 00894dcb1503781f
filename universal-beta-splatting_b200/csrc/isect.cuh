// Shared device helpers for tile intersection (used by isect.cu and the fused projection kernel).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace ubs {

constexpr int kIsectThreads = 256;

// Truncation report of the capacity-bounded pair list (one thread per call): status[0] = was THIS call's list
// truncated (rewritten by every call, so it always describes the most recent frame), status[1] = number of truncated
// calls so far.  The Adam kernels read status[0] as their skip flag: a truncated frame never updates the parameters.
__device__ __forceinline__ void report_truncation(int32_t *status, bool truncated) {
    if (status == nullptr) return;
    status[0] = truncated ? 1 : 0;
    if (truncated) atomicAdd(status + 1, 1);
}

struct TileRect {
    uint32_t x0, y0, x1, y1;  // min inclusive, max exclusive
};

// Tile AABB of a projected primitive (isect_tiles.cu:56-66).  The float->uint32 casts of negative values
// saturate to 0 (PTX cvt.rzi.u32.f32), which is what makes the reference's `max(0, (uint32_t)...)` work.
__device__ __forceinline__ TileRect tile_rect(float mx, float my, int32_t radius_i, uint32_t tile_size,
                                              uint32_t tile_width, uint32_t tile_height) {
    const float radius = (float)radius_i;
    const float tile_radius = radius / static_cast<float>(tile_size);
    const float tile_x = mx / static_cast<float>(tile_size);
    const float tile_y = my / static_cast<float>(tile_size);
    TileRect t;
    t.x0 = min(max(0u, (uint32_t)floorf(tile_x - tile_radius)), tile_width);
    t.y0 = min(max(0u, (uint32_t)floorf(tile_y - tile_radius)), tile_height);
    t.x1 = min(max(0u, (uint32_t)ceilf(tile_x + tile_radius)), tile_width);
    t.y1 = min(max(0u, (uint32_t)ceilf(tile_y + tile_radius)), tile_height);
    return t;
}

// Tile binning (bin_sort.cu): a primitive covering tiles [x0,x1) x [y0,y1) adds +1/-1 at the four corners of that
// rectangle in the camera's (tile_height+1) x (tile_width+1) delta grid; the 2-D prefix sum of the grid is the
// number of pairs per tile.
// Cells are kDeltaStride ints apart (one per 32-byte sector): packed 4-byte cells put the whole grid into a few
// L2 slices whose atomic units then saturate.
constexpr int kDeltaStride = 8;
__device__ __forceinline__ void add_tile_deltas(int32_t *delta, const TileRect &t, uint32_t tile_width) {
    const uint32_t gw = tile_width + 1;
    atomicAdd(delta + (size_t)(t.y0 * gw + t.x0) * kDeltaStride, 1);
    atomicAdd(delta + (size_t)(t.y0 * gw + t.x1) * kDeltaStride, -1);
    atomicAdd(delta + (size_t)(t.y1 * gw + t.x0) * kDeltaStride, -1);
    atomicAdd(delta + (size_t)(t.y1 * gw + t.x1) * kDeltaStride, 1);
}

// Block-wide (kIsectThreads) int64 sum; result valid in thread 0.
__device__ __forceinline__ int64_t block_reduce_sum_i64(int64_t v) {
    __shared__ int64_t s_red[kIsectThreads / 32];
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) v += __shfl_down_sync(0xffffffffu, v, off);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) s_red[warp] = v;
    __syncthreads();
    int64_t total = 0;
    if (threadIdx.x == 0) {
#pragma unroll
        for (int w = 0; w < kIsectThreads / 32; ++w) total += s_red[w];
    }
    return total;
}

// Block-wide (kIsectThreads) exclusive int64 scan.
__device__ __forceinline__ int64_t block_exclusive_scan_i64(int64_t v) {
    __shared__ int64_t s_scan[kIsectThreads / 32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int64_t incl = v;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
        const int64_t t = __shfl_up_sync(0xffffffffu, incl, off);
        if (lane >= off) incl += t;
    }
    if (lane == 31) s_scan[warp] = incl;
    __syncthreads();
    int64_t base = 0;
#pragma unroll
    for (int w = 0; w < kIsectThreads / 32; ++w)
        if (w < warp) base += s_scan[w];
    return base + incl - v;
}

size_t bin_delta_bytes(int C, int tile_width, int tile_height);  // bytes of the delta grid at the start of the bin workspace

int isect_scan_and_total(int64_t n_blocks, int64_t *block_sums, int64_t *n_isects, cudaStream_t s);
// block sums + scan + total from already computed per-(camera, primitive) tile counts; `workspace` is the buffer
// later handed to ubs_isect_emit_sort.
int isect_blocksums_from_counts(int64_t CN, const int32_t *tiles_per_gauss, void *workspace, int64_t *n_isects,
                                cudaStream_t s);

}  // namespace ubs
