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

// Shared-memory loads through an explicit 32-bit shared-window address: the address is formed once outside the
// hot loop (the compiler otherwise re-derives the shared base from SR_CgaCtaId inside it).
__device__ __forceinline__ uint32_t smem_addr(const void *p) {
    uint32_t a = (uint32_t)__cvta_generic_to_shared(p);
    asm volatile("" : "+r"(a));  // opaque to the optimiser: keeps the address in a register (no rematerialisation)
    return a;
}
__device__ __forceinline__ float4 lds_f4(uint32_t addr) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
    return v;
}
// MUFU.RCP without the range fix-up code that `1.f / x` expands to even under --use_fast_math
__device__ __forceinline__ float fast_rcp(float x) {
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ uint4 lds_u4(uint32_t addr) {
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
    return v;
}
// -|x| in one LOP (the compiler otherwise forms it with two FADDs)
__device__ __forceinline__ float set_sign(float x) {
    uint32_t b = __float_as_uint(x);
    asm("or.b32 %0, %0, 0x80000000;" : "+r"(b));
    return __uint_as_float(b);
}
__device__ __forceinline__ uint32_t lds_u16(uint32_t addr) {
    uint32_t v;
    asm volatile("ld.shared.u16 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
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

// Which of the tile's eight 8x4 sub-tiles (bit w = warp w: column w & 1, row w >> 1; see sub_tile_of) can the
// sigma < 1 box `bb` = (xmin, xmax, ymin, ymax) touch.  (tx0, ty0) = centre of the tile's first pixel.  Computed once
// per staged pair by the staging thread; each warp then only tests its bit.
__device__ __forceinline__ uint32_t sub_tile_mask(float4 bb, float tx0, float ty0) {
    uint32_t mx = 0, my = 0;
    if (bb.x <= tx0 + 7.f && bb.y >= tx0) mx |= 0x55u;
    if (bb.x <= tx0 + 15.f && bb.y >= tx0 + 8.f) mx |= 0xAAu;
    if (bb.z <= ty0 + 3.f && bb.w >= ty0) my |= 0x03u;
    if (bb.z <= ty0 + 7.f && bb.w >= ty0 + 4.f) my |= 0x0Cu;
    if (bb.z <= ty0 + 11.f && bb.w >= ty0 + 8.f) my |= 0x30u;
    if (bb.z <= ty0 + 15.f && bb.w >= ty0 + 12.f) my |= 0xC0u;
    return mx & my;
}

// Refinement of sub_tile_mask for the backward pass (where a (warp, pair) evaluation costs ~120 instructions): drop the
// sub-tiles whose rectangle of pixel centres the ellipse {sigma < 1} misses although its bounding box overlaps them
// (corners, thin tilted ellipses).  Exact minimum of the convex quadratic sigma over the rectangle: when the centre is
// outside, the minimiser lies on an edge facing the centre, so at most two clamped 1-D minimisations are needed.
// Conservative by a margin far above rounding: a sub-tile is only dropped when min sigma >= 1.001.
__device__ __forceinline__ uint32_t refine_sub_tile_mask(uint32_t mask, float mx, float my, float a, float b, float c,
                                                         float tx0, float ty0) {
    if (!(a > 0.f && c > 0.f && a * c - b * b > 0.f)) return mask;  // degenerate conic: keep the box test
    const float rc = 1.f / c, ra = 1.f / a;
#pragma unroll
    for (int w = 0; w < 8; ++w) {
        if (!((mask >> w) & 1u)) continue;
        const float x0 = tx0 + (float)((w & 1) * kSubW), x1 = x0 + (float)(kSubW - 1);
        const float y0 = ty0 + (float)((w >> 1) * kSubH), y1 = y0 + (float)(kSubH - 1);
        const bool in_x = mx >= x0 && mx <= x1, in_y = my >= y0 && my <= y1;
        if (in_x && in_y) continue;  // centre inside: sigma = 0 there
        float qmin = 3.0e38f;
        if (!in_x) {  // vertical edge facing the centre
            const float dx = (mx < x0 ? x0 : x1) - mx;
            const float dy = fminf(fmaxf(my - b * dx * rc, y0), y1) - my;
            qmin = a * dx * dx + 2.f * b * dx * dy + c * dy * dy;
        }
        if (!in_y) {  // horizontal edge facing the centre
            const float dy = (my < y0 ? y0 : y1) - my;
            const float dx = fminf(fmaxf(mx - b * dy * ra, x0), x1) - mx;
            qmin = fminf(qmin, a * dx * dx + 2.f * b * dx * dy + c * dy * dy);
        }
        if (qmin >= 1.001f) mask &= ~(1u << w);
    }
    return mask;
}
// Which (16 / COLS) x 4 pixel blocks of the tile can {sigma < 1} of a conic (a, b, c) centred at (mx, my) touch: bit
// COLS r + cx = block column cx of block row r (COLS = 2: the eight 8x4 sub-tiles in sub_tile_of's numbering, COLS = 4:
// sixteen 4x4 blocks).  Exact up to the stated margins: for each of the four block rows, the x-extent of the ellipse
// restricted to the rows' y-slab (the right edge x = (-b v + sqrt(a - det v^2)) / a is concave in v = y - my with its
// maximum at v* = -b / sqrt(det c); the left edge mirrors it), compared with the block columns.  (tx0, ty0) = centre
// of the tile's first pixel.  Degenerate conics are never culled.
template <int COLS>
__device__ __forceinline__ uint32_t slab_mask(float mx, float my, float a, float b, float c, float tx0, float ty0) {
    constexpr int CW = kTile / COLS, RH = 4;
    const float det = a * c - b * b;
    if (!(det > 0.f && a > 0.f && c > 0.f)) return (1u << (4 * COLS)) - 1u;
    const float hy = sqrtf(a / det);
    const float vstar = -b * rsqrtf(det * c);
    const float ra = 1.f / a;
    const float u0 = tx0 - mx;  // first pixel column relative to the centre
    uint32_t mask = 0;
#pragma unroll
    for (int r = 0; r < kTile / RH; ++r) {
        const float d0 = ty0 + (float)(RH * r) - my;
        const float lo = fmaxf(d0 - 0.01f, -hy), hi = fminf(d0 + (float)(RH - 1) + 0.01f, hy);
        if (lo > hi) continue;
        const float vr = fminf(fmaxf(vstar, lo), hi), vl = fminf(fmaxf(-vstar, lo), hi);
        float xr = (sqrtf(fmaxf(0.f, a - det * vr * vr)) - b * vr) * ra;
        float xl = (-sqrtf(fmaxf(0.f, a - det * vl * vl)) - b * vl) * ra;
        xr += 0.01f + 0.001f * fabsf(xr);
        xl -= 0.01f + 0.001f * fabsf(xl);
        uint32_t cols = 0;
#pragma unroll
        for (int cx = 0; cx < COLS; ++cx)
            if (xl <= u0 + (float)(CW * cx + CW - 1) && xr >= u0 + (float)(CW * cx)) cols |= 1u << cx;
        mask |= cols << (COLS * r);
    }
    return mask;
}
}  // namespace ubs
