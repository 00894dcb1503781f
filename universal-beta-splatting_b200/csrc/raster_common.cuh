// Constants and helpers shared by the compositing kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace ubs {
constexpr int kTile = 16;                   // tile edge in pixels (the reference caller always uses 16)
constexpr int kTilePixels = kTile * kTile;  // one thread per pixel, one CTA per tile
constexpr int kSubW = 8, kSubH = 4;         // each warp owns an 8x4 pixel sub-tile (2 x 4 sub-tiles per tile)

struct SubTile {
    uint32_t bx, by;  // sub-tile coordinates inside the tile
    uint32_t px, py;  // pixel coordinates inside the tile
};

__device__ __forceinline__ SubTile sub_tile_of(uint32_t thread_rank) {
    const uint32_t warp = thread_rank >> 5, lane = thread_rank & 31;
    SubTile s;
    s.bx = warp & 1;
    s.by = warp >> 1;
    s.px = s.bx * kSubW + (lane & (kSubW - 1));
    s.py = s.by * kSubH + (lane >> 3);
    return s;
}

// Axis-aligned bounding box of the support {sigma < 1} of a 2-D Beta primitive with conic (a, b, c):
// half extents sqrt(c / det), sqrt(a / det) with det = ac - b^2, inflated so that rounding can never cull a
// pixel the exact test would accept.  Degenerate conics get an unbounded box (never culled).
__device__ __forceinline__ float4 support_bbox(float mx, float my, float a, float b, float c) {
    const float det = a * c - b * b;
    float hx = 3.0e38f, hy = 3.0e38f;
    if (det > 0.f && a > 0.f && c > 0.f) {
        const float rdet = 1.f / det;
        hx = sqrtf(c * rdet) * 1.001f + 0.01f;
        hy = sqrtf(a * rdet) * 1.001f + 0.01f;
    }
    return make_float4(mx - hx, mx + hx, my - hy, my + hy);
}
}  // namespace ubs
