// K11: backward of the per-tile compositing.  Drop-in for rasterize_to_pixels_bwd (rasterize_to_pixels_bwd.cu:16-276):
// re-walks each tile back to front from T_final / last_ids and produces gradients of means2d, conics, colours,
// opacities and betas (accumulated into caller-zeroed arrays or 48-byte rows).
//
// Three kernels:
//   * rasterize_bwd3_pairlane_kernel (RGB; the fused path's kernel, ubs_rasterize_bwd_rows, and the default of
//     ubs_rasterize_bwd[_splats] at 3 channels): lane = pair, loop = the pixels of a 4x4 block, the sequential walk of a
//     pixel's pairs as one warp scan of affine maps -- see the comment at the kernel;
//   * rasterize_bwd_kernel<CH> (other channel counts: depth modes, N-D colour chunks) and rasterize_bwd3_kernel (the
//     round-1 RGB kernel, UBS_BWD3_VARIANT=0, kept for A/B): lane = pixel.  Their reduction strategy (the reference
//     does 50 SHFL + up to 8x10 global atomics per (tile, pair)):
//       1. each warp owns an 8x4 pixel sub-tile and first compacts the batch to the pairs whose sigma < 1 support can
//          touch it (same cull as the forward pass), so most (warp, pair) combinations cost nothing;
//       2. for 3 channels, three pairs x 10 gradient components are reduced together with one 31-shuffle transposing
//          butterfly, after which lane l holds the warp total of component l;
//       3. 30 lanes add those totals into a shared-memory accumulator [pair][component] in one instruction;
//       4. after the batch, each thread flushes one pair with at most 10 global atomics -- one set per (tile, pair).
#include <stdlib.h>

#include "common.cuh"
#include "raster_common.cuh"

namespace ubs {
namespace {

constexpr int kGrad3 = 10;  // rgb(3) conic(3) xy(2) opacity beta

template <int CH>
__global__ void __launch_bounds__(kTilePixels)
rasterize_bwd_kernel(int C, int64_t N, const int64_t *__restrict__ n_isects_dev, int64_t isect_capacity,
                     const float2 *__restrict__ means2d, const float *__restrict__ conics,
                     const float *__restrict__ colors, const float *__restrict__ opacities,
                     const float *__restrict__ betas, const float *__restrict__ backgrounds,
                     const uint8_t *__restrict__ masks, uint32_t width, uint32_t height, uint32_t tile_width,
                     uint32_t tile_height, const int32_t *__restrict__ tile_offsets,
                     const int32_t *__restrict__ flatten_ids, const float *__restrict__ render_alphas,
                     const int32_t *__restrict__ last_ids, const float *__restrict__ v_render_colors,
                     const float *__restrict__ v_render_alphas, float *__restrict__ v_means2d,
                     float *__restrict__ v_conics, float *__restrict__ v_colors, float *__restrict__ v_opacities,
                     float *__restrict__ v_betas, const float4 *__restrict__ splats, bool splat_colors,
                     float *__restrict__ v_depths) {
    constexpr int NG = 7 + CH;           // gradient components per pair
    constexpr bool kSmemAcc = CH <= 4;   // wide colour vectors go straight to global atomics (shared memory budget)
    const uint32_t cam = blockIdx.z;
    const uint32_t tile_id = blockIdx.y * tile_width + blockIdx.x;
    const uint32_t tr = threadIdx.x, lane = tr & 31, warp = tr >> 5;
    const SubTile st = sub_tile_of(tr);
    const uint32_t i = blockIdx.y * kTile + st.py;
    const uint32_t j = blockIdx.x * kTile + st.px;
    const float px = (float)j + 0.5f, py = (float)i + 0.5f;
    const bool inside = (i < height && j < width);

    tile_offsets += (size_t)cam * tile_height * tile_width;
    if (backgrounds != nullptr) backgrounds += cam * CH;
    if (masks != nullptr && !masks[(size_t)cam * tile_height * tile_width + tile_id]) return;

    const int64_t n_isects = min(*n_isects_dev, isect_capacity);
    const int32_t range_start = tile_offsets[tile_id];
    const int32_t range_end = (cam == (uint32_t)C - 1 && tile_id == tile_width * tile_height - 1)
                                  ? (int32_t)n_isects
                                  : tile_offsets[tile_id + 1];
    const int32_t num_batches = (range_end - range_start + kTilePixels - 1) / kTilePixels;
    if (num_batches <= 0) return;

    __shared__ int32_t s_id[kTilePixels];
    __shared__ float4 s_xyob[kTilePixels];
    __shared__ float4 s_conic[kTilePixels];
    __shared__ float4 s_bbox[kTilePixels];
    __shared__ float s_color[kTilePixels * CH];
    __shared__ float s_acc[kSmemAcc ? kTilePixels * NG : 1];
    __shared__ uint8_t s_list[kTilePixels / 32][kTilePixels];

    const float wx0 = (float)(blockIdx.x * kTile + st.bx * kSubW) + 0.5f, wx1 = wx0 + (float)(kSubW - 1);
    const float wy0 = (float)(blockIdx.y * kTile + st.by * kSubH) + 0.5f, wy1 = wy0 + (float)(kSubH - 1);

    // per-pixel state
    const size_t pix = inside ? ((size_t)cam * height + i) * width + j : 0;
    const float T_final = inside ? 1.f - render_alphas[pix] : 1.f;
    float T = T_final;
    float buffer[CH], v_rc[CH];
    float bg_dot = 0.f;
#pragma unroll
    for (int k = 0; k < CH; ++k) {
        buffer[k] = 0.f;
        v_rc[k] = inside ? v_render_colors[pix * CH + k] : 0.f;
        if (backgrounds != nullptr) bg_dot += backgrounds[k] * v_rc[k];
    }
    const float v_ra = inside ? v_render_alphas[pix] : 0.f;
    const int32_t bin_final = inside ? last_ids[pix] : 0;
    int32_t warp_bin_final = bin_final;
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) warp_bin_final = max(warp_bin_final, __shfl_xor_sync(0xffffffffu, warp_bin_final, off));

    if constexpr (kSmemAcc)
        for (int k = tr; k < kTilePixels * NG; k += kTilePixels) s_acc[k] = 0.f;

    // gradient of one (pixel, pair); writes NG values (zeros when the pair does not contribute to the pixel)
    auto pair_grad = [&](uint32_t p, int32_t batch_end, float *g) {
#pragma unroll
        for (int k = 0; k < NG; ++k) g[k] = 0.f;
        if (!inside || batch_end - (int32_t)p > bin_final) return false;
        const float4 conic = s_conic[p];
        const float4 xyob = s_xyob[p];
        const float opac = xyob.z, beta = xyob.w;
        const float dx = xyob.x - px, dy = xyob.y - py;
        const float sigma = (conic.x * dx * dx + conic.z * dy * dy) + 2.f * conic.y * dx * dy;
        if (sigma < 0.f || sigma >= 1.f) return false;
        const float vis = __powf(1.f - sigma, beta);
        const float alpha = fminf(0.999f, opac * vis);
        const float ra = fast_rcp(1.f - alpha);
        T *= ra;
        const float fac = alpha * T;
        float v_alpha = 0.f;
#pragma unroll
        for (int k = 0; k < CH; ++k) {
            const float c = s_color[p * CH + k];
            g[k] = fac * v_rc[k];
            v_alpha += (c * T - buffer[k] * ra) * v_rc[k];
            buffer[k] += c * fac;
        }
        v_alpha += T_final * ra * v_ra;
        if (backgrounds != nullptr) v_alpha += -T_final * ra * bg_dot;
        if (opac * vis <= 0.999f) {
            const float v_sigma = -v_alpha * opac * beta * __powf(1.f - sigma, beta - 1.f);
            g[CH + 0] = dx * dx * v_sigma;
            g[CH + 1] = 2.f * dx * dy * v_sigma;
            g[CH + 2] = dy * dy * v_sigma;
            g[CH + 3] = 2.f * v_sigma * (conic.x * dx + conic.y * dy);
            g[CH + 4] = 2.f * v_sigma * (conic.y * dx + conic.z * dy);
            g[CH + 5] = vis * v_alpha;
            g[CH + 6] = v_alpha * opac * vis * __logf(1.f - sigma);
        }
        return true;
    };

    for (int32_t b = 0; b < num_batches; ++b) {
        __syncthreads();  // previous batch fully consumed and flushed
        const int32_t batch_end = range_end - 1 - kTilePixels * b;  // pair index held by slot 0 (furthest back)
        const int32_t batch_size = min((int32_t)kTilePixels, batch_end + 1 - range_start);
        const int32_t idx = batch_end - (int32_t)tr;
        if (idx >= range_start) {
            const int32_t g = flatten_ids[idx];
            s_id[tr] = g;
            float4 xyob, cn;
            if (splats != nullptr) {  // 48-byte rows of the fused projection kernel (see rasterize_fwd.cu)
                xyob = splats[(size_t)g * 3];
                cn = splats[(size_t)g * 3 + 1];
                cn.w = 0.f;
            } else {
                const float2 xy = means2d[g];
                xyob = make_float4(xy.x, xy.y, opacities[g], betas[g]);
                cn = make_float4(conics[(size_t)g * 3], conics[(size_t)g * 3 + 1], conics[(size_t)g * 3 + 2], 0.f);
            }
            s_xyob[tr] = xyob;
            s_conic[tr] = cn;
            s_bbox[tr] = support_bbox(xyob.x, xyob.y, cn.x, cn.y, cn.z);
            if ((CH == 1 || CH == 3 || CH == 4) && splats != nullptr && splat_colors) {
                if constexpr (CH == 1) {
                    s_color[tr] = splats[(size_t)g * 3 + 1].w;  // depth as the colour (rasterize_fwd.cu)
                } else {
                    const float4 c4 = splats[(size_t)g * 3 + 2];
                    s_color[tr * CH + 0] = c4.x;
                    if constexpr (CH > 1) s_color[tr * CH + 1] = c4.y;
                    if constexpr (CH > 2) s_color[tr * CH + 2] = c4.z;
                    if constexpr (CH > 3) s_color[tr * CH + 3] = c4.w;
                }
            } else {
#pragma unroll
                for (int k = 0; k < CH; ++k) s_color[tr * CH + k] = colors[(size_t)g * CH + k];
            }
        }
        __syncthreads();

        // pairs behind every pixel's last contributor are skipped outright (rasterize_to_pixels_bwd.cu:157)
        const int32_t t_begin = max(0, batch_end - warp_bin_final);
        uint32_t cnt = 0;
        for (int32_t p0 = t_begin & ~31; p0 < batch_size; p0 += 32) {
            const int32_t p = p0 + (int32_t)lane;
            bool hit = false;
            if (p >= t_begin && p < batch_size) {
                const float4 bb = s_bbox[p];
                hit = (bb.x <= wx1) && (bb.y >= wx0) && (bb.z <= wy1) && (bb.w >= wy0);
            }
            const uint32_t m = __ballot_sync(0xffffffffu, hit);
            if (hit) s_list[warp][cnt + __popc(m & ((1u << lane) - 1u))] = (uint8_t)p;
            cnt += __popc(m);
        }
        __syncwarp();

        if constexpr (CH == 3) {
            for (uint32_t t = 0; t < cnt; t += 3) {
                float v[32];
                v[30] = 0.f, v[31] = 0.f;
                bool any_valid = false;
#pragma unroll
                for (int q = 0; q < 3; ++q) {
                    if (t + q < cnt) {
                        any_valid |= pair_grad(s_list[warp][t + q], batch_end, v + q * kGrad3);
                    } else {
#pragma unroll
                        for (int k = 0; k < kGrad3; ++k) v[q * kGrad3 + k] = 0.f;
                    }
                }
                if (!__any_sync(0xffffffffu, any_valid)) continue;
                // transposing butterfly: afterwards v[0] of lane l is the warp-wide sum of component l
#pragma unroll
                for (int off = 16; off >= 1; off >>= 1) {
                    const bool upper = (lane & off) != 0;
#pragma unroll
                    for (int k = 0; k < off; ++k) {
                        const float send = upper ? v[k] : v[k + off];
                        const float keep = upper ? v[k + off] : v[k];
                        v[k] = keep + __shfl_xor_sync(0xffffffffu, send, off);
                    }
                }
                const uint32_t q = lane / kGrad3, comp = lane - q * kGrad3;
                if (lane < 3 * kGrad3 && t + q < cnt && v[0] != 0.f)
                    atomicAdd(&s_acc[(uint32_t)s_list[warp][t + q] * NG + comp], v[0]);
            }
        } else {
            for (uint32_t t = 0; t < cnt; ++t) {
                float g[NG];
                const uint32_t p = s_list[warp][t];
                const bool valid = pair_grad(p, batch_end, g);
                if (!__any_sync(0xffffffffu, valid)) continue;
#pragma unroll
                for (int k = 0; k < NG; ++k) {
                    float x = g[k];
#pragma unroll
                    for (int off = 16; off > 0; off >>= 1) x += __shfl_xor_sync(0xffffffffu, x, off);
                    if (lane == 0 && x != 0.f) {
                        if constexpr (kSmemAcc) {
                            atomicAdd(&s_acc[p * NG + k], x);
                        } else {
                            const size_t g_id = (size_t)s_id[p];
                            float *dst = k < CH       ? v_colors + g_id * CH + k
                                         : k < CH + 3 ? v_conics + g_id * 3 + (k - CH)
                                         : k < CH + 5 ? v_means2d + g_id * 2 + (k - CH - 3)
                                         : k == CH + 5 ? v_opacities + g_id
                                                       : v_betas + g_id;
                            atomicAdd(dst, x);
                        }
                    }
                }
            }
        }
        __syncthreads();

        // flush: one set of global atomics per (tile, pair)
        if (kSmemAcc && (int32_t)tr < batch_size) {
            const int32_t g = s_id[tr];
            float *acc = s_acc + tr * NG;
            bool nz = false;
#pragma unroll
            for (int k = 0; k < NG; ++k) nz |= (acc[k] != 0.f);
            if (nz) {
                if (v_depths != nullptr) {
                    // colours came out of the splat rows: the RGB part goes to v_colors [C,N,3], the depth channel
                    // (the last one) to v_depths [C,N] -- the two arrays ubs_fused_project_bwd consumes
#pragma unroll
                    for (int k = 0; k < CH - 1; ++k) atomicAdd(v_colors + (size_t)g * 3 + k, acc[k]);
                    atomicAdd(v_depths + g, acc[CH - 1]);
                } else {
#pragma unroll
                    for (int k = 0; k < CH; ++k) atomicAdd(v_colors + (size_t)g * CH + k, acc[k]);
                }
                atomicAdd(v_conics + (size_t)g * 3 + 0, acc[CH + 0]);
                atomicAdd(v_conics + (size_t)g * 3 + 1, acc[CH + 1]);
                atomicAdd(v_conics + (size_t)g * 3 + 2, acc[CH + 2]);
                atomicAdd(v_means2d + (size_t)g * 2 + 0, acc[CH + 3]);
                atomicAdd(v_means2d + (size_t)g * 2 + 1, acc[CH + 4]);
                atomicAdd(v_opacities + g, acc[CH + 5]);
                atomicAdd(v_betas + g, acc[CH + 6]);
#pragma unroll
                for (int k = 0; k < NG; ++k) acc[k] = 0.f;
            }
        }
    }
}


// ---- RGB fast path ---------------------------------------------------------------------------------------------
// Same results as the generic kernel above for CH == 3, restructured for issue-slot count (the kernel is issue
// bound):  * array-of-structs staging, 16-bit byte offsets in the per-warp lists, explicit shared addresses;
//          * branch-free evaluation of a triple of pairs (invalid lanes contribute exact zeros), so the three
//            evaluations interleave and only the triple-level "anything valid?" test branches;
//          * the colour buffer of the reference (buffer[k], rasterize_to_pixels_bwd.cu:196-214) collapses to the
//            scalar B = sum_k buffer[k] v_rc[k], and the term T_final (v_ra - bg.v_rc) is hoisted per pixel;
//          * v_xy is linear in the moments Sx = sum v_sigma dx, Sy = sum v_sigma dy, so those are what the warp
//            reduces; the conic products, the factor 2 and ln 2 are applied once per (tile, pair) at flush time.
struct __align__(16) Staged3 {
    float4 xyob;   // mean2d.x, mean2d.y, opacity, beta
    float4 conic;  // a, 2b, c, (unused)
    float4 col;    // r, g, b, (unused)
};
constexpr int kAccStride = 12;  // 10 accumulators per pair in a 48-byte row: the byte offset of a pair's row equals that of its
                                // staged record (both 48 p), so the hot loop adds to s_acc + list offset + 4 * component

__global__ void __launch_bounds__(kTilePixels)
rasterize_bwd3_kernel(int C, int64_t N, const int64_t *__restrict__ n_isects_dev, int64_t isect_capacity,
                      const float2 *__restrict__ means2d, const float *__restrict__ conics,
                      const float *__restrict__ colors, const float *__restrict__ opacities,
                      const float *__restrict__ betas, const float *__restrict__ backgrounds,
                      const uint8_t *__restrict__ masks, uint32_t width, uint32_t height, uint32_t tile_width,
                      uint32_t tile_height, const int32_t *__restrict__ tile_offsets,
                      const int32_t *__restrict__ flatten_ids, const float *__restrict__ render_alphas,
                      const int32_t *__restrict__ last_ids, const float *__restrict__ v_render_colors,
                      const float *__restrict__ v_render_alphas, float *__restrict__ v_means2d,
                      float *__restrict__ v_conics, float *__restrict__ v_colors, float *__restrict__ v_opacities,
                      float *__restrict__ v_betas, const float4 *__restrict__ splats, bool splat_colors) {
    constexpr int kRec = (int)sizeof(Staged3);  // 48
    const uint32_t cam = blockIdx.z;
    const uint32_t tile_id = blockIdx.y * tile_width + blockIdx.x;
    const uint32_t tr = threadIdx.x, lane = tr & 31, warp = tr >> 5;
    const SubTile st = sub_tile_of(tr);
    const uint32_t i = blockIdx.y * kTile + st.py;
    const uint32_t j = blockIdx.x * kTile + st.px;
    const float px = (float)j + 0.5f, py = (float)i + 0.5f;
    const bool inside = (i < height && j < width);

    tile_offsets += (size_t)cam * tile_height * tile_width;
    if (backgrounds != nullptr) backgrounds += cam * 3;
    if (masks != nullptr && !masks[(size_t)cam * tile_height * tile_width + tile_id]) return;

    const int64_t n_isects = min(*n_isects_dev, isect_capacity);
    const int32_t range_start = tile_offsets[tile_id];
    const int32_t range_end = (cam == (uint32_t)C - 1 && tile_id == tile_width * tile_height - 1)
                                  ? (int32_t)n_isects
                                  : tile_offsets[tile_id + 1];
    const int32_t num_batches = (range_end - range_start + kTilePixels - 1) / kTilePixels;
    if (num_batches <= 0) return;

    __shared__ Staged3 s_rec[kTilePixels + 1];  // [kTilePixels] = sentinel (sigma = NaN) padding the lists
    __shared__ __align__(8) uint8_t s_mask[kTilePixels];  // sub-tiles the sigma < 1 box of each staged pair can touch (0 past the batch)
    __shared__ int32_t s_id[kTilePixels];
    __shared__ __align__(16) float s_acc[kTilePixels * kAccStride];
    __shared__ __align__(8) uint16_t s_list[kTilePixels / 32][kTilePixels + 4];

    const float tx0 = (float)(blockIdx.x * kTile) + 0.5f, ty0 = (float)(blockIdx.y * kTile) + 0.5f;  // first pixel centre
    uint16_t *my_list = s_list[warp];
    const uint32_t rec_addr = smem_addr(s_rec), list_addr = smem_addr(my_list);
    const float kNaN = __int_as_float(0x7fffffff);
    if (tr == 0) {
        s_rec[kTilePixels].xyob = make_float4(kNaN, kNaN, 0.f, 1.f);
        s_rec[kTilePixels].conic = make_float4(1.f, 0.f, 1.f, 0.f);
        s_rec[kTilePixels].col = make_float4(0.f, 0.f, 0.f, 0.f);
    }

    // per-pixel state
    const size_t pix = inside ? ((size_t)cam * height + i) * width + j : 0;
    const float T_final = inside ? 1.f - render_alphas[pix] : 1.f;
    float T = T_final;
    const float v_r = inside ? v_render_colors[pix * 3 + 0] : 0.f;
    const float v_g = inside ? v_render_colors[pix * 3 + 1] : 0.f;
    const float v_b = inside ? v_render_colors[pix * 3 + 2] : 0.f;
    float Kc = inside ? v_render_alphas[pix] : 0.f;  // becomes T_final (v_ra - bg . v_rc)
    if (backgrounds != nullptr) Kc -= backgrounds[0] * v_r + backgrounds[1] * v_g + backgrounds[2] * v_b;
    Kc *= T_final;
    float B = 0.f;  // sum_k buffer[k] v_rc[k]
    const int32_t bin_final = inside ? last_ids[pix] : -1;
    int32_t warp_bin_final = bin_final;
#pragma unroll
    for (int off = 16; off > 0; off >>= 1)
        warp_bin_final = max(warp_bin_final, __shfl_xor_sync(0xffffffffu, warp_bin_final, off));

    for (int k = tr; k < kTilePixels * kAccStride; k += kTilePixels) s_acc[k] = 0.f;

    // one (pixel, pair) evaluation; g[0..9] = rgb, (dx^2, dx dy, dy^2) v_sigma, Sx, Sy, v_opacity, v_beta / ln2.
    // Branch-free: an invalid lane sees vis = 0 -> alpha = 0, ra = 1, every output an exact zero.
    auto eval = [&](uint32_t off, uint32_t off_min, float *g) -> bool {
        const uint32_t rec = rec_addr + off;
        const float4 xyob = lds_f4(rec);
        const float4 conic = lds_f4(rec + 16);
        const float4 col = lds_f4(rec + 32);
        const float dx = xyob.x - px, dy = xyob.y - py;
        const float sigma = __fmaf_rn(dy, dx * conic.y, __fmaf_rn(dx, conic.x * dx, dy * (conic.z * dy)));
        // pairs behind this pixel's last contributor (rasterize_to_pixels_bwd.cu:166-168) and sigma outside [0,1)
        const bool valid = (__float_as_uint(sigma) < 0x3f800000u) && (off >= off_min);
        const float om = 1.f - (valid ? sigma : 0.f);
        const float lg = __log2f(om);
        const float vis = valid ? exp2f(xyob.w * lg) : 0.f;
        const float ov = xyob.z * vis;
        const float alpha = fminf(0.999f, ov);
        const float ra = fast_rcp(1.f - alpha);
        T *= ra;
        const float fac = alpha * T;
        g[0] = fac * v_r;
        g[1] = fac * v_g;
        g[2] = fac * v_b;
        const float cv = __fmaf_rn(col.z, v_b, __fmaf_rn(col.y, v_g, col.x * v_r));
        const float v_alpha = __fmaf_rn(T, cv, ra * (Kc - B));
        B = __fmaf_rn(fac, cv, B);
        const bool live = ov <= 0.999f;  // the clamp has zero slope above it (rasterize_to_pixels_bwd.cu:230)
        const float ov_g = live ? ov : 0.f;
        const float vis_g = live ? vis : 0.f;
        const float v_sigma = -(v_alpha * xyob.w) * (ov_g * fast_rcp(om));  // o beta (1-sigma)^(beta-1) = beta ov / (1-sigma)
        const float tx = dx * v_sigma, ty = dy * v_sigma;
        g[3] = tx * dx;
        g[4] = tx * dy;
        g[5] = ty * dy;
        g[6] = tx;
        g[7] = ty;
        g[8] = vis_g * v_alpha;
        g[9] = (v_alpha * ov_g) * lg;
        return valid;
    };

    for (int32_t b = 0; b < num_batches; ++b) {
        __syncthreads();  // previous batch fully consumed and flushed
        const int32_t batch_end = range_end - 1 - kTilePixels * b;  // pair index held by slot 0 (furthest back)
        const int32_t batch_size = min((int32_t)kTilePixels, batch_end + 1 - range_start);
        const int32_t idx = batch_end - (int32_t)tr;
        if (idx >= range_start) {
            const int32_t g = flatten_ids[idx];
            s_id[tr] = g;
            float4 xyob, col;
            float ca, cb, cc;
            if (splats != nullptr) {  // 48-byte rows of the fused projection kernel (see rasterize_fwd.cu)
                xyob = splats[(size_t)g * 3];
                const float4 cn = splats[(size_t)g * 3 + 1];
                ca = cn.x, cb = cn.y, cc = cn.z;
            } else {
                const float2 xy = means2d[g];
                xyob = make_float4(xy.x, xy.y, opacities[g], betas[g]);
                ca = conics[(size_t)g * 3], cb = conics[(size_t)g * 3 + 1], cc = conics[(size_t)g * 3 + 2];
            }
            if (splats != nullptr && splat_colors) {
                col = splats[(size_t)g * 3 + 2];
                col.w = 0.f;
            } else {
                col = make_float4(colors[(size_t)g * 3], colors[(size_t)g * 3 + 1], colors[(size_t)g * 3 + 2], 0.f);
            }
            s_rec[tr].xyob = xyob;
            s_rec[tr].conic = make_float4(ca, cb + cb, cc, 0.f);
            s_rec[tr].col = col;
            s_mask[tr] = (uint8_t)refine_sub_tile_mask(sub_tile_mask(support_bbox(xyob.x, xyob.y, ca, cb, cc), tx0, ty0),
                                                        xyob.x, xyob.y, ca, cb, cc, tx0, ty0);
        } else {
            s_mask[tr] = 0;
        }
        __syncthreads();

        // slot p holds pair batch_end - p; a pixel takes part iff batch_end - p <= bin_final
        const int32_t t_begin = max(0, batch_end - warp_bin_final);
        const int64_t pmin = inside ? max((int64_t)0, (int64_t)batch_end - bin_final) : (int64_t)kTilePixels + 1;
        const uint32_t off_min = (uint32_t)min(pmin, (int64_t)kTilePixels + 1) * kRec;
        // lane l takes the eight staged pairs 8 l .. 8 l + 7 (one 64-bit load of their masks); one warp scan of the per-lane
        // hit counts places them in order (see rasterize_fwd.cu)
        unsigned long long bits = (*reinterpret_cast<const unsigned long long *>(s_mask + 8 * lane) >> warp) & 0x0101010101010101ull;
        {
            const int32_t skip = t_begin - 8 * (int32_t)lane;  // leading pairs of this lane that lie before t_begin
            if (skip >= 8) bits = 0ull;
            else if (skip > 0) bits &= ~0ull << (8 * skip);
        }
        const uint32_t n_mine = (uint32_t)__popcll(bits);
        uint32_t incl = n_mine;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            const uint32_t t = __shfl_up_sync(0xffffffffu, incl, off);
            if ((int)lane >= off) incl += t;
        }
        const uint32_t cnt = __shfl_sync(0xffffffffu, incl, 31);
        {
            uint32_t pos = incl - n_mine;
            const uint32_t lo = (uint32_t)bits, hi = (uint32_t)(bits >> 32);
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const uint32_t word = k < 4 ? lo : hi;
                if ((word >> (8 * (k & 3))) & 1u) my_list[pos++] = (uint16_t)((8u * lane + (uint32_t)k) * (uint32_t)kRec);
            }
        }
        if (lane < 3) my_list[cnt + lane] = (uint16_t)(kTilePixels * kRec);  // pad the last triple with the sentinel
        __syncwarp();

        for (uint32_t t = 0; t < cnt; t += 3) {
            float v[32];
            v[30] = 0.f, v[31] = 0.f;
            const uint32_t o0 = lds_u16(list_addr + 2 * t), o1 = lds_u16(list_addr + 2 * t + 2),
                           o2 = lds_u16(list_addr + 2 * t + 4);
            bool any_valid = eval(o0, off_min, v);
            any_valid |= eval(o1, off_min, v + kGrad3);
            any_valid |= eval(o2, off_min, v + 2 * kGrad3);
            if (!__any_sync(0xffffffffu, any_valid)) continue;
            // transposing butterfly: afterwards v[0] of lane l is the warp-wide sum of component l
#pragma unroll
            for (int off = 16; off >= 1; off >>= 1) {
                const bool upper = (lane & off) != 0;
#pragma unroll
                for (int k = 0; k < off; ++k) {
                    const float send = upper ? v[k] : v[k + off];
                    const float keep = upper ? v[k + off] : v[k];
                    v[k] = keep + __shfl_xor_sync(0xffffffffu, send, off);
                }
            }
            const uint32_t q = lane / kGrad3, comp = lane - q * kGrad3;
            const uint32_t oq = q == 0 ? o0 : (q == 1 ? o1 : o2);
            static_assert(kAccStride * sizeof(float) == sizeof(Staged3), "accumulator rows mirror the staged records");
            if (lane < 3 * kGrad3 && oq < (uint32_t)(kTilePixels * kRec) && v[0] != 0.f)
                atomicAdd(reinterpret_cast<float *>(reinterpret_cast<unsigned char *>(s_acc) + oq) + comp, v[0]);
        }
        __syncthreads();

        // flush: one set of global atomics per (tile, pair); finish the moment form here
        if ((int32_t)tr < batch_size) {
            float4 *acc = reinterpret_cast<float4 *>(s_acc + tr * kAccStride);  // 48-byte rows: three vector loads
            const float4 q0 = acc[0], q1 = acc[1], q2 = acc[2];
            const float a[kGrad3] = {q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, q1.z, q1.w, q2.x, q2.y};
            bool nz = false;
#pragma unroll
            for (int k = 0; k < kGrad3; ++k) nz |= (a[k] != 0.f);
            if (nz) {
                const size_t g = (size_t)s_id[tr];
                const float4 conic = s_rec[tr].conic;  // a, 2b, c
                atomicAdd(v_colors + g * 3 + 0, a[0]);
                atomicAdd(v_colors + g * 3 + 1, a[1]);
                atomicAdd(v_colors + g * 3 + 2, a[2]);
                atomicAdd(v_conics + g * 3 + 0, a[3]);
                atomicAdd(v_conics + g * 3 + 1, a[4] + a[4]);
                atomicAdd(v_conics + g * 3 + 2, a[5]);
                // v_xy = 2 v_sigma (a dx + b dy, b dx + c dy) summed over pixels = (2a Sx + 2b Sy, 2b Sx + 2c Sy)
                atomicAdd(v_means2d + g * 2 + 0, __fmaf_rn(conic.x + conic.x, a[6], conic.y * a[7]));
                atomicAdd(v_means2d + g * 2 + 1, __fmaf_rn(conic.y, a[6], (conic.z + conic.z) * a[7]));
                atomicAdd(v_opacities + g, a[8]);
                atomicAdd(v_betas + g, a[9] * 0.693147180559945f);  // lg2 -> ln
                acc[0] = acc[1] = acc[2] = make_float4(0.f, 0.f, 0.f, 0.f);
            }
        }
    }
}


// ---- RGB path, lane = pair ("pair-lane") ---------------------------------------------------------------------------
// Same results as rasterize_bwd3_kernel, with the roles of lanes and loop swapped inside every (pixels x pairs) block of
// work.  There: lane = pixel, loop over the pairs of the warp's culled list, the per-pair sums over pixels need a
// transposing butterfly (31 SHFL + 62 SEL + 31 FADD per three pairs) and the sub-tile a warp culls against is tied to
// the warp width (8x4 pixels).  Here: lane = pair (its 48-byte record and its ten sums live in registers for a whole
// bucket of up to 32 pairs), the loop runs over the pixels of a 4x4 block, and what has to cross lanes is the
// per-pixel sequential state instead: walking a pixel's pairs back to front is the composition of the affine maps
//     (T, B) -> (ra T,  B + alpha ra cv T),     ra = 1 / (1 - alpha),  cv = colour . v_render_colour,
// so one inclusive warp scan over the monoid  (R1, K1) o (R2, K2) = (R1 R2, K1 + K2 R1)  -- five stages of two SHFL,
// one FMUL, one FFMA -- gives every lane the transmittance in front of its pair and the colour behind it
// (rasterize_to_pixels_bwd.cu:190-214 keeps both sequentially).  Nothing is reduced across lanes, the per-pixel state
// (T, B) sits in shared memory between buckets, and the cull granularity is free of the warp width: 4x4 blocks leave
// ~20 % fewer evaluations than 8x4 sub-tiles.  A bucket of <= 16 (<= 8) pairs runs two (four) pixels per step in
// 16-lane (8-lane) segments so that short lists do not idle lanes.
constexpr int kBlk = 4;                     // block edge in pixels
constexpr int kBlkPix = kBlk * kBlk;        // 16 pixels per block, 16 blocks per tile: bit (by * 4 + bx) of a pair's mask

// shfl.sync.up inside W-lane segments as a volatile asm statement: the compiler keeps volatile statements in source order,
// which is what interleaves the scans of independent pixels (left alone it finishes one scan before starting the next)
template <int W>
__device__ __forceinline__ float shfl_up_ordered(float v, int off) {
    float r;
    asm volatile("shfl.sync.up.b32 %0, %1, %2, %3, 0xffffffff;" : "=f"(r) : "f"(v), "r"(off), "r"((32 - W) << 8));
    return r;
}

// red.global.add.v4.f32 / .v2.f32 (sm_90+): one fire-and-forget reduction per 16 / 8 bytes
__device__ __forceinline__ void red_add_v4(float *p, float a, float b, float c, float d) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
__device__ __forceinline__ void red_add_v2(float *p, float a, float b) {
    asm volatile("red.global.add.v2.f32 [%0], {%1, %2};" ::"l"(p), "f"(a), "f"(b) : "memory");
}

// One bucket: lanes j = lane % W hold the pairs list[0 .. n), n <= W; 32 / W pixels of the block are processed per step and
// NCH steps (independent pixels) are advanced together: the shuffles of their scans are issued stage by stage for all of
// them, so that one chain's SHFL latency is covered by the others (the compiler keeps shuffles in source order).
// Per (pair, pixel), with R = 1 / (1 - alpha) and T' = T R the transmittance in front of the pair
// (rasterize_to_pixels_bwd.cu:190-245 in closed form):
//     v_alpha = T' cv - R (sum_behind - Kc)  =  R (T' cv + Kc - B_after),      B_after = B_before + alpha T' cv,
// since 1 + alpha R = R; and alpha R = R - 1 gives the scan element K = (R - 1) cv with one FFMA.
template <int W, int NCH_, bool ROWS_OUT>
__device__ __forceinline__ void pairlane_bucket(uint32_t lane, const uint8_t *__restrict__ list, int32_t n,
                                                const Staged3 *__restrict__ s_rec, float *__restrict__ s_acc,
                                                const int32_t *__restrict__ s_id, float *__restrict__ v_rows,
                                                const float4 *__restrict__ pix_const, float4 *__restrict__ pix_state,
                                                float pxb, float pyb, int32_t batch_end) {
    constexpr int G = 32 / W;                  // pixels per step
    constexpr int STEPS = kBlkPix / G;         // steps per block (W = 32: 16, 16: 8, 8: 4)
    constexpr int XS = kBlk / G;               // steps per pixel row (W = 32: 4, 16: 2, 8: 1)
    constexpr int NCH = NCH_ < STEPS ? NCH_ : STEPS;
    static_assert(STEPS % NCH == 0 && NCH % XS == 0, "an iteration covers whole pixel rows of the block");
    constexpr int ROWS = NCH / XS;             // pixel rows per iteration
    const uint32_t j = lane & (W - 1), ph = lane / W;
    const bool has = (int32_t)j < n;
    const uint32_t slot = has ? (uint32_t)list[j] : (uint32_t)kTilePixels;  // [kTilePixels] = the NaN sentinel
    const float4 xyob = s_rec[slot].xyob, conic = s_rec[slot].conic, col = s_rec[slot].col;
    const int32_t pair_idx = batch_end - (int32_t)slot;  // position of the pair in the sorted list
    const float neg_beta = -xyob.w;
    float acc[kGrad3];
#pragma unroll
    for (int k = 0; k < kGrad3; ++k) acc[k] = 0.f;
    // the columns of the steps of an iteration are the same in every iteration: their share of sigma is formed once
    // (sigma = fma(dy, dx (2b), fma(dx, a dx, dy (c dy))): the association of the forward pass, bit for bit)
    float dxc[XS], adx[XS], bdx[XS];
#pragma unroll
    for (int xs = 0; xs < XS; ++xs) {
        dxc[xs] = xyob.x - (pxb + (float)(ph + xs * G));
        adx[xs] = conic.x * dxc[xs];
        bdx[xs] = dxc[xs] * conic.y;
    }
    const float4 *pc_it = pix_const + ph;
    float4 *ps_it = pix_state + ph;
    float py0 = pyb;
#pragma unroll 1
    for (int it = 0; it < STEPS / NCH; ++it) {
        float4 pc[NCH];
        float T0[NCH], B0[NCH], dyr[ROWS], cdy2[ROWS], ra[NCH], cv[NCH], lg[NCH], vis[NCH], ov[NCH], rom[NCH];
        float R[NCH], K[NCH];
#pragma unroll
        for (int r = 0; r < ROWS; ++r) {
            dyr[r] = xyob.y - (py0 + (float)r);
            cdy2[r] = dyr[r] * (conic.z * dyr[r]);
        }
#pragma unroll
        for (int t = 0; t < NCH; ++t) {
            // step it * NCH + t covers pixel column (t % XS) * G + ph of row it * ROWS + t / XS of the block
            const int xs = t % XS, row = t / XS;
            const int pix = row * kBlk + xs * G;
            pc[t] = pc_it[pix];            // v_r, v_g, v_b, Kc
            const float4 ps = ps_it[pix];  // T, B, last_ids
            T0[t] = ps.x, B0[t] = ps.y;
            const float sigma = __fmaf_rn(dyr[row], bdx[xs], __fmaf_rn(dxc[xs], adx[xs], cdy2[row]));
            // sigma in [0, 1) and the pair not behind this pixel's last contributor (rasterize_to_pixels_bwd.cu:166-168)
            const bool valid = (__float_as_uint(sigma) < 0x3f800000u) && (pair_idx <= __float_as_int(ps.z));
            const float om = 1.f - (valid ? sigma : 0.f);
            lg[t] = __log2f(om);
            vis[t] = valid ? exp2f(xyob.w * lg[t]) : 0.f;
            ov[t] = xyob.z * vis[t];
            rom[t] = fast_rcp(om);
            ra[t] = fast_rcp(1.f - fminf(0.999f, ov[t]));
            cv[t] = __fmaf_rn(col.z, pc[t].z, __fmaf_rn(col.y, pc[t].y, col.x * pc[t].x));
            R[t] = ra[t], K[t] = __fmaf_rn(ra[t], cv[t], -cv[t]);  // alpha R cv
        }
        // inclusive scans of (R, K) over the pairs of the bucket, furthest back first; all chains stage by stage
#pragma unroll
        for (int off = 1; off < W; off <<= 1) {
            float Rp[NCH], Kp[NCH];
#pragma unroll
            for (int t = 0; t < NCH; ++t) {
                Rp[t] = shfl_up_ordered<W>(R[t], off);
                Kp[t] = shfl_up_ordered<W>(K[t], off);
            }
            if ((int)j >= off) {
#pragma unroll
                for (int t = 0; t < NCH; ++t) {
                    K[t] = __fmaf_rn(K[t], Rp[t], Kp[t]);
                    R[t] *= Rp[t];
                }
            }
        }
#pragma unroll
        for (int t = 0; t < NCH; ++t) {
            const int xs = t % XS, row = t / XS;
            const int pix = row * kBlk + xs * G;
            const float Tl = T0[t] * R[t];                        // transmittance in front of this pair
            const float B_after = __fmaf_rn(T0[t], K[t], B0[t]);  // sum_k buffer[k] v_rc[k] including this pair
            if (j == W - 1) *reinterpret_cast<float2 *>(ps_it + pix) = make_float2(Tl, B_after);
            const float fac = fminf(0.999f, ov[t]) * Tl;
            acc[0] = __fmaf_rn(fac, pc[t].x, acc[0]);
            acc[1] = __fmaf_rn(fac, pc[t].y, acc[1]);
            acc[2] = __fmaf_rn(fac, pc[t].z, acc[2]);
            float v_alpha = ra[t] * __fmaf_rn(Tl, cv[t], pc[t].w - B_after);
            v_alpha = ov[t] <= 0.999f ? v_alpha : 0.f;  // the clamp has zero slope above it (rasterize_to_pixels_bwd.cu:230)
            acc[8] = __fmaf_rn(vis[t], v_alpha, acc[8]);
            const float u = v_alpha * ov[t];
            acc[9] = __fmaf_rn(u, lg[t], acc[9]);
            const float v_sigma = (u * rom[t]) * neg_beta;  // d alpha / d sigma = -o beta (1 - sigma)^(beta - 1)
            const float tx = dxc[xs] * v_sigma, ty = dyr[row] * v_sigma;
            acc[3] = __fmaf_rn(tx, dxc[xs], acc[3]);
            acc[4] = __fmaf_rn(tx, dyr[row], acc[4]);
            acc[5] = __fmaf_rn(ty, dyr[row], acc[5]);
            acc[6] += tx;
            acc[7] += ty;
        }
        py0 += (float)ROWS;
        pc_it += ROWS * kBlk, ps_it += ROWS * kBlk;
    }
    if constexpr (G > 1) {  // the segments hold the same pairs: add their sums
#pragma unroll
        for (int off = W; off < 32; off <<= 1) {
#pragma unroll
            for (int k = 0; k < kGrad3; ++k) acc[k] += __shfl_xor_sync(0xffffffffu, acc[k], off);
        }
    }
    if (has && ph == 0) {
        if constexpr (ROWS_OUT) {
            // the ten sums of this (block, pair) go straight to the primitive's 48-byte gradient row: three vector
            // reductions -- unless no pixel of the block was inside the support (every sum is then a signed zero)
            uint32_t bits = 0;
#pragma unroll
            for (int k = 0; k < kGrad3; ++k) bits |= __float_as_uint(acc[k]);
            if ((bits << 1) != 0u) {
                float *row = v_rows + (size_t)s_id[slot] * kAccStride;
                red_add_v4(row, acc[0], acc[1], acc[2], acc[3]);
                red_add_v4(row + 4, acc[4], acc[5], acc[6], acc[7]);
                red_add_v2(row + 8, acc[8], acc[9]);
            }
        } else {
            float *row = s_acc + slot * kAccStride;
#pragma unroll
            for (int k = 0; k < kGrad3; ++k) atomicAdd(row + k, acc[k]);
        }
    }
}

template <int NCH, int MINB, bool ROWS_OUT>
__global__ void __launch_bounds__(kTilePixels, MINB)
rasterize_bwd3_pairlane_kernel(int C, int64_t N, const int64_t *__restrict__ n_isects_dev, int64_t isect_capacity,
                               const float2 *__restrict__ means2d, const float *__restrict__ conics,
                               const float *__restrict__ colors, const float *__restrict__ opacities,
                               const float *__restrict__ betas, const float *__restrict__ backgrounds,
                               const uint8_t *__restrict__ masks, uint32_t width, uint32_t height, uint32_t tile_width,
                               uint32_t tile_height, const int32_t *__restrict__ tile_offsets,
                               const int32_t *__restrict__ flatten_ids, const float *__restrict__ render_alphas,
                               const int32_t *__restrict__ last_ids, const float *__restrict__ v_render_colors,
                               const float *__restrict__ v_render_alphas, float *__restrict__ v_means2d,
                               float *__restrict__ v_conics, float *__restrict__ v_colors,
                               float *__restrict__ v_opacities, float *__restrict__ v_betas,
                               const float4 *__restrict__ splats, bool splat_colors, float *__restrict__ v_rows,
                               const int32_t *__restrict__ skip_flag) {
    // the frame lost pairs to the capacity bound (isect.cuh: report_truncation): it contributes no gradient
    if (skip_flag != nullptr && *skip_flag != 0) return;
    const uint32_t cam = blockIdx.z;
    const uint32_t tile_id = blockIdx.y * tile_width + blockIdx.x;
    const uint32_t tr = threadIdx.x, lane = tr & 31, warp = tr >> 5;
    // pixels in block order: thread tr initialises pixel (tr & 15) of block (tr >> 4); warp w then owns blocks 2w, 2w + 1,
    // i.e. exactly the pixels its own threads initialised (the state never crosses warps)
    const uint32_t blk = tr >> 4, sp = tr & 15;
    const uint32_t i = blockIdx.y * kTile + (blk >> 2) * kBlk + (sp >> 2);
    const uint32_t j = blockIdx.x * kTile + (blk & 3) * kBlk + (sp & 3);
    const bool inside = (i < height && j < width);

    tile_offsets += (size_t)cam * tile_height * tile_width;
    if (backgrounds != nullptr) backgrounds += cam * 3;
    if (masks != nullptr && !masks[(size_t)cam * tile_height * tile_width + tile_id]) return;

    const int64_t n_isects = min(*n_isects_dev, isect_capacity);
    const int32_t range_start = tile_offsets[tile_id];
    const int32_t range_end = (cam == (uint32_t)C - 1 && tile_id == tile_width * tile_height - 1)
                                  ? (int32_t)n_isects
                                  : tile_offsets[tile_id + 1];
    const int32_t num_batches = (range_end - range_start + kTilePixels - 1) / kTilePixels;
    if (num_batches <= 0) return;

    __shared__ Staged3 s_rec[kTilePixels + 1];  // [kTilePixels] = sentinel (sigma = NaN) for lanes without a pair
    __shared__ __align__(16) uint16_t s_mask[kTilePixels];  // blocks the support of each staged pair can touch (0 past the batch)
    __shared__ int32_t s_id[kTilePixels];
    __shared__ __align__(16) float s_acc[ROWS_OUT ? 4 : kTilePixels * kAccStride];
    __shared__ float4 s_pix_const[kTilePixels];  // v_r, v_g, v_b, T_final (v_ra - bg . v_rc)
    __shared__ float4 s_pix_state[kTilePixels];  // T, B = sum_k buffer[k] v_rc[k], last_ids
    __shared__ uint8_t s_list[kTilePixels / 32][2][kTilePixels];

    const float tx0 = (float)(blockIdx.x * kTile) + 0.5f, ty0 = (float)(blockIdx.y * kTile) + 0.5f;  // first pixel centre
    const float kNaN = __int_as_float(0x7fffffff);
    if (tr == 0) {
        s_rec[kTilePixels].xyob = make_float4(kNaN, kNaN, 0.f, 1.f);
        s_rec[kTilePixels].conic = make_float4(1.f, 0.f, 1.f, 0.f);
        s_rec[kTilePixels].col = make_float4(0.f, 0.f, 0.f, 0.f);
    }

    // per-pixel state
    int32_t bin_final = -1;
    {
        const size_t pix = inside ? ((size_t)cam * height + i) * width + j : 0;
        const float T_final = inside ? 1.f - render_alphas[pix] : 1.f;
        const float v_r = inside ? v_render_colors[pix * 3 + 0] : 0.f;
        const float v_g = inside ? v_render_colors[pix * 3 + 1] : 0.f;
        const float v_b = inside ? v_render_colors[pix * 3 + 2] : 0.f;
        float Kc = inside ? v_render_alphas[pix] : 0.f;
        if (backgrounds != nullptr) Kc -= backgrounds[0] * v_r + backgrounds[1] * v_g + backgrounds[2] * v_b;
        Kc *= T_final;
        if (inside) bin_final = last_ids[pix];
        s_pix_const[tr] = make_float4(v_r, v_g, v_b, Kc);
        s_pix_state[tr] = make_float4(T_final, 0.f, __int_as_float(bin_final), 0.f);
    }
    int32_t blk_bin_final = bin_final;  // furthest-front contributor of any pixel of the thread's block
#pragma unroll
    for (int off = 8; off > 0; off >>= 1)
        blk_bin_final = max(blk_bin_final, __shfl_xor_sync(0xffffffffu, blk_bin_final, off));
    const int32_t bin_final_a = __shfl_sync(0xffffffffu, blk_bin_final, 0);
    const int32_t bin_final_b = __shfl_sync(0xffffffffu, blk_bin_final, 16);
    const uint32_t bit_a = 2 * warp;  // mask bit of the warp's first block; the second is bit_a + 1
    const float pxb_a = tx0 + (float)((bit_a & 3) * kBlk), pyb = ty0 + (float)((warp >> 1) * kBlk);

    if constexpr (!ROWS_OUT)
        for (int k = tr; k < kTilePixels * kAccStride; k += kTilePixels) s_acc[k] = 0.f;

    for (int32_t b = 0; b < num_batches; ++b) {
        __syncthreads();  // previous batch fully consumed and flushed
        const int32_t batch_end = range_end - 1 - kTilePixels * b;  // pair index held by slot 0 (furthest back)
        const int32_t batch_size = min((int32_t)kTilePixels, batch_end + 1 - range_start);
        const int32_t idx = batch_end - (int32_t)tr;
        if (idx >= range_start) {
            const int32_t g = flatten_ids[idx];
            s_id[tr] = g;
            float4 xyob, col;
            float ca, cb, cc;
            if (splats != nullptr) {  // 48-byte rows of the fused projection kernel (see rasterize_fwd.cu)
                xyob = splats[(size_t)g * 3];
                const float4 cn = splats[(size_t)g * 3 + 1];
                ca = cn.x, cb = cn.y, cc = cn.z;
            } else {
                const float2 xy = means2d[g];
                xyob = make_float4(xy.x, xy.y, opacities[g], betas[g]);
                ca = conics[(size_t)g * 3], cb = conics[(size_t)g * 3 + 1], cc = conics[(size_t)g * 3 + 2];
            }
            if (splats != nullptr && splat_colors) {
                col = splats[(size_t)g * 3 + 2];
                col.w = 0.f;
            } else {
                col = make_float4(colors[(size_t)g * 3], colors[(size_t)g * 3 + 1], colors[(size_t)g * 3 + 2], 0.f);
            }
            s_rec[tr].xyob = xyob;
            s_rec[tr].conic = make_float4(ca, cb + cb, cc, 0.f);
            s_rec[tr].col = col;
            s_mask[tr] = (uint16_t)slab_mask<4>(xyob.x, xyob.y, ca, cb, cc, tx0, ty0);
        } else {
            s_mask[tr] = 0;
        }
        __syncthreads();

        // per-warp lists of the staged pairs that can touch block a / block b, in order.  Lane l takes the staged pairs
        // 8 l .. 8 l + 7 (one 128-bit load of their masks); one warp scan of the packed per-lane counts places them.
        // Slot p holds pair batch_end - p; pairs in front of every pixel's last contributor are left out.
        uint32_t cnt_a, cnt_b;
        {
            const uint4 m4 = *reinterpret_cast<const uint4 *>(s_mask + 8 * lane);
            const uint32_t w[4] = {m4.x >> bit_a, m4.y >> bit_a, m4.z >> bit_a, m4.w >> bit_a};
            uint32_t hits_a = 0, hits_b = 0;  // bit k: staged pair 8 lane + k
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                hits_a |= ((w[q] & 1u) | ((w[q] >> 15) & 2u)) << (2 * q);
                hits_b |= (((w[q] >> 1) & 1u) | ((w[q] >> 16) & 2u)) << (2 * q);
            }
            const int32_t skip_a = max(0, batch_end - bin_final_a) - 8 * (int32_t)lane;  // leading pairs to leave out
            const int32_t skip_b = max(0, batch_end - bin_final_b) - 8 * (int32_t)lane;
            hits_a = skip_a >= 8 ? 0u : (skip_a > 0 ? hits_a & (0xFFu << skip_a) : hits_a);
            hits_b = skip_b >= 8 ? 0u : (skip_b > 0 ? hits_b & (0xFFu << skip_b) : hits_b);
            const uint32_t mine = (uint32_t)__popc(hits_a) | ((uint32_t)__popc(hits_b) << 16);
            uint32_t incl = mine;
#pragma unroll
            for (int off = 1; off < 32; off <<= 1) {
                const uint32_t t = __shfl_up_sync(0xffffffffu, incl, off);
                if ((int)lane >= off) incl += t;
            }
            const uint32_t total = __shfl_sync(0xffffffffu, incl, 31);
            cnt_a = total & 0xFFFFu, cnt_b = total >> 16;
            uint32_t pos_a = (incl - mine) & 0xFFFFu, pos_b = (incl - mine) >> 16;
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                if ((hits_a >> k) & 1u) s_list[warp][0][pos_a++] = (uint8_t)(8u * lane + (uint32_t)k);
                if ((hits_b >> k) & 1u) s_list[warp][1][pos_b++] = (uint8_t)(8u * lane + (uint32_t)k);
            }
        }
        __syncwarp();

#pragma unroll 1
        for (int h = 0; h < 2; ++h) {
            const uint8_t *list = s_list[warp][h];
            const int32_t cnt = (int32_t)(h == 0 ? cnt_a : cnt_b);
            const float4 *pc = s_pix_const + 32 * warp + kBlkPix * h;
            float4 *ps = s_pix_state + 32 * warp + kBlkPix * h;
            const float pxb = pxb_a + (float)(kBlk * h);
#pragma unroll 1
            for (int32_t base = 0; base < cnt; base += 32) {
                const int32_t n = min(32, cnt - base);
                if (n > 16) pairlane_bucket<32, NCH, ROWS_OUT>(lane, list + base, n, s_rec, s_acc, s_id, v_rows, pc, ps, pxb, pyb, batch_end);
                else if (n > 8) pairlane_bucket<16, NCH, ROWS_OUT>(lane, list + base, n, s_rec, s_acc, s_id, v_rows, pc, ps, pxb, pyb, batch_end);
                else pairlane_bucket<8, NCH, ROWS_OUT>(lane, list + base, n, s_rec, s_acc, s_id, v_rows, pc, ps, pxb, pyb, batch_end);
                __syncwarp();  // the bucket's state writes are visible to the next bucket's reads
            }
        }
        if constexpr (ROWS_OUT) continue;
        __syncthreads();

        // flush: one set of global atomics per (tile, pair); finish the moment form here
        if ((int32_t)tr < batch_size) {
            float4 *acc = reinterpret_cast<float4 *>(s_acc + tr * kAccStride);  // 48-byte rows: three vector loads
            const float4 q0 = acc[0], q1 = acc[1], q2 = acc[2];
            const float a[kGrad3] = {q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, q1.z, q1.w, q2.x, q2.y};
            bool nz = false;
#pragma unroll
            for (int k = 0; k < kGrad3; ++k) nz |= (a[k] != 0.f);
            if (nz) {
                const size_t g = (size_t)s_id[tr];
                const float4 conic = s_rec[tr].conic;  // a, 2b, c
                atomicAdd(v_colors + g * 3 + 0, a[0]);
                atomicAdd(v_colors + g * 3 + 1, a[1]);
                atomicAdd(v_colors + g * 3 + 2, a[2]);
                atomicAdd(v_conics + g * 3 + 0, a[3]);
                atomicAdd(v_conics + g * 3 + 1, a[4] + a[4]);
                atomicAdd(v_conics + g * 3 + 2, a[5]);
                atomicAdd(v_means2d + g * 2 + 0, __fmaf_rn(conic.x + conic.x, a[6], conic.y * a[7]));
                atomicAdd(v_means2d + g * 2 + 1, __fmaf_rn(conic.y, a[6], (conic.z + conic.z) * a[7]));
                atomicAdd(v_opacities + g, a[8]);
                atomicAdd(v_betas + g, a[9] * 0.693147180559945f);  // lg2 -> ln
                acc[0] = acc[1] = acc[2] = make_float4(0.f, 0.f, 0.f, 0.f);
            }
        }
    }
}

// 0 = lane-per-pixel kernel with the transposing butterfly, 1 = lane-per-pair kernel (default); UBS_BWD3_VARIANT
// overrides it for A/B measurements (read once).
static int bwd3_variant() {
    static const int v = [] {
        const char *e = getenv("UBS_BWD3_VARIANT");
        return e != nullptr ? atoi(e) : 1;
    }();
    return v;
}

template <int CH>
int launch_bwd(int C, int64_t N, const int64_t *n_isects, int64_t cap, const float *means2d, const float *conics,
               const float *colors, const float *opacities, const float *betas, const float *backgrounds,
               const uint8_t *masks, int width, int height, const int32_t *offsets, const int32_t *flatten_ids,
               const float *render_alphas, const int32_t *last_ids, const float *v_render_colors,
               const float *v_render_alphas, float *v_means2d, float *v_conics, float *v_colors, float *v_opacities,
               float *v_betas, const float *splats, int splat_colors, float *v_depths, cudaStream_t s) {
    const uint32_t tw = (uint32_t)ceil_div(width, kTile), th = (uint32_t)ceil_div(height, kTile);
    dim3 grid(tw, th, (unsigned)C), block(kTilePixels, 1, 1);
    if constexpr (CH == 3) {
        const int variant = bwd3_variant();
        if (variant != 0) {
            auto *k1 = rasterize_bwd3_pairlane_kernel<4, 3, false>;
            k1<<<grid, block, 0, s>>>(
                C, N, n_isects, cap, (const float2 *)means2d, conics, colors, opacities, betas, backgrounds, masks,
                (uint32_t)width, (uint32_t)height, tw, th, offsets, flatten_ids, render_alphas, last_ids,
                v_render_colors, v_render_alphas, v_means2d, v_conics, v_colors, v_opacities, v_betas,
                (const float4 *)splats, splat_colors != 0, nullptr, nullptr);
            UBS_LAUNCH_CHECK("rasterize_bwd_kernel");
            return UBS_OK;
        }
        rasterize_bwd3_kernel<<<grid, block, 0, s>>>(
            C, N, n_isects, cap, (const float2 *)means2d, conics, colors, opacities, betas, backgrounds, masks,
            (uint32_t)width, (uint32_t)height, tw, th, offsets, flatten_ids, render_alphas, last_ids,
            v_render_colors, v_render_alphas, v_means2d, v_conics, v_colors, v_opacities, v_betas,
            (const float4 *)splats, splat_colors != 0);
    } else {
        rasterize_bwd_kernel<CH><<<grid, block, 0, s>>>(
            C, N, n_isects, cap, (const float2 *)means2d, conics, colors, opacities, betas, backgrounds, masks,
            (uint32_t)width, (uint32_t)height, tw, th, offsets, flatten_ids, render_alphas, last_ids,
            v_render_colors, v_render_alphas, v_means2d, v_conics, v_colors, v_opacities, v_betas,
            (const float4 *)splats, splat_colors != 0, v_depths);
    }
    UBS_LAUNCH_CHECK("rasterize_bwd_kernel");
    return UBS_OK;
}

// RGB from the 48-byte splat rows, gradients into 48-byte rows (ubs_rasterize_bwd_rows)
int launch_bwd_rows(int C, int64_t N, const int64_t *n_isects, int64_t cap, const float *splats,
                    const float *backgrounds, const uint8_t *masks, int width, int height, const int32_t *offsets,
                    const int32_t *flatten_ids, const float *render_alphas, const int32_t *last_ids,
                    const float *v_render_colors, const float *v_render_alphas, float *v_rows, const int32_t *skip_flag,
                    cudaStream_t s) {
    const uint32_t tw = (uint32_t)ceil_div(width, kTile), th = (uint32_t)ceil_div(height, kTile);
    dim3 grid(tw, th, (unsigned)C), block(kTilePixels, 1, 1);
    // four pixels in flight per lane at 3 CTAs / SM measured faster than two at 4 CTAs / SM (1.124 against 1.156 ms, cfg3)
    // (eight in flight at 2 CTAs / SM: 2.2 ms -- 128 registers and spills)
    rasterize_bwd3_pairlane_kernel<4, 3, true><<<grid, block, 0, s>>>(
        C, N, n_isects, cap, nullptr, nullptr, nullptr, nullptr, nullptr, backgrounds, masks, (uint32_t)width,
        (uint32_t)height, tw, th, offsets, flatten_ids, render_alphas, last_ids, v_render_colors, v_render_alphas,
        nullptr, nullptr, nullptr, nullptr, nullptr, (const float4 *)splats, true, v_rows, skip_flag);
    UBS_LAUNCH_CHECK("rasterize_bwd_rows_kernel");
    return UBS_OK;
}

}  // namespace
}  // namespace ubs

static int rasterize_bwd_impl(int C, int64_t N, const int64_t *n_isects, int64_t isect_capacity,
                                 const float *means2d, const float *conics, const float *colors,
                                 const float *opacities, const float *betas, const float *backgrounds,
                                 const uint8_t *masks, int channels, int width, int height, int tile_size,
                                 const int32_t *offsets, const int32_t *flatten_ids, const float *render_alphas,
                                 const int32_t *last_ids, const float *v_render_colors, const float *v_render_alphas,
                                 float *v_means2d, float *v_conics, float *v_colors, float *v_opacities,
                                 float *v_betas, const float *splats, int splat_colors, float *v_depths, void *stream) {
    using namespace ubs;
    UBS_CHECK_ARG(C >= 0 && N >= 0 && width > 0 && height > 0, "rasterize_bwd: bad sizes");
    UBS_CHECK_ARG(tile_size == kTile, "rasterize_bwd: tile_size must be %d (got %d)", kTile, tile_size);
    if (C == 0 || N == 0 || isect_capacity == 0) return UBS_OK;  // no pairs: every gradient stays zero
    UBS_CHECK_ARG(n_isects && offsets && ((means2d && conics && opacities && betas) || splats) &&
                      (colors || (splats && splat_colors)) && flatten_ids &&
                      render_alphas && last_ids && v_render_colors && v_render_alphas && v_means2d && v_conics &&
                      v_colors && v_opacities && v_betas,
                  "rasterize_bwd: null pointer");
    UBS_CHECK_ARG(((uintptr_t)splats & 15) == 0, "rasterize_bwd: splats must be 16-byte aligned");
    UBS_CHECK_ARG(!splat_colors || channels == 3 || ((channels == 4 || channels == 1) && v_depths != nullptr),
                  "rasterize_bwd: splat colours are RGB, or RGB+depth / depth with v_depths given (channels = %d)", channels);
    UBS_CHECK_ARG(v_depths == nullptr || (splat_colors && (channels == 4 || channels == 1)),
                  "rasterize_bwd: v_depths goes with splat colours of 4 or 1 channels");
    cudaStream_t s = (cudaStream_t)stream;
#define UBS_BWD_CASE(CH)                                                                                               \
    case CH:                                                                                                           \
        return launch_bwd<CH>(C, N, n_isects, isect_capacity, means2d, conics, colors, opacities, betas, backgrounds,  \
                              masks, width, height, offsets, flatten_ids, render_alphas, last_ids, v_render_colors,    \
                              v_render_alphas, v_means2d, v_conics, v_colors, v_opacities, v_betas, splats, splat_colors,      \
                              v_depths, s);
    switch (channels) {
        UBS_BWD_CASE(1)
        UBS_BWD_CASE(2)
        UBS_BWD_CASE(3)
        UBS_BWD_CASE(4)
        UBS_BWD_CASE(8)
        UBS_BWD_CASE(16)
        default:
            set_error("rasterize_bwd: unsupported channel count %d (supported: 1,2,3,4,8,16)", channels);
            return UBS_EUNSUPPORTED;
    }
#undef UBS_BWD_CASE
}

extern "C" int ubs_rasterize_bwd(int C, int64_t N, const int64_t *n_isects, int64_t isect_capacity,
                                 const float *means2d, const float *conics, const float *colors,
                                 const float *opacities, const float *betas, const float *backgrounds,
                                 const uint8_t *masks, int channels, int width, int height, int tile_size,
                                 const int32_t *offsets, const int32_t *flatten_ids, const float *render_alphas,
                                 const int32_t *last_ids, const float *v_render_colors, const float *v_render_alphas,
                                 float *v_means2d, float *v_conics, float *v_colors, float *v_opacities,
                                 float *v_betas, void *stream) {
    return rasterize_bwd_impl(C, N, n_isects, isect_capacity, means2d, conics, colors, opacities, betas, backgrounds, masks,
                              channels, width, height, tile_size, offsets, flatten_ids, render_alphas, last_ids,
                              v_render_colors, v_render_alphas, v_means2d, v_conics, v_colors, v_opacities, v_betas,
                              nullptr, 0, nullptr, stream);
}

extern "C" int ubs_rasterize_bwd_splats(int C, int64_t N, const int64_t *n_isects, int64_t isect_capacity,
                                        const float *splats, const float *colors, const float *backgrounds,
                                        const uint8_t *masks, int channels, int width, int height, int tile_size,
                                        const int32_t *offsets, const int32_t *flatten_ids,
                                        const float *render_alphas, const int32_t *last_ids,
                                        const float *v_render_colors, const float *v_render_alphas, float *v_means2d,
                                        float *v_conics, float *v_colors, float *v_opacities, float *v_betas,
                                        float *v_depths, void *stream) {
    using namespace ubs;
    UBS_CHECK_ARG(splats != nullptr || N == 0 || C == 0 || isect_capacity == 0, "rasterize_bwd_splats: splats is null");
    return rasterize_bwd_impl(C, N, n_isects, isect_capacity, nullptr, nullptr, colors, nullptr, nullptr, backgrounds, masks,
                              channels, width, height, tile_size, offsets, flatten_ids, render_alphas, last_ids,
                              v_render_colors, v_render_alphas, v_means2d, v_conics, v_colors, v_opacities, v_betas,
                              splats, colors == nullptr ? 1 : 0, colors == nullptr ? v_depths : nullptr, stream);
}

extern "C" int ubs_rasterize_bwd_rows(int C, int64_t N, const int64_t *n_isects, int64_t isect_capacity,
                                      const float *splats, const float *backgrounds, const uint8_t *masks, int width,
                                      int height, int tile_size, const int32_t *offsets, const int32_t *flatten_ids,
                                      const float *render_alphas, const int32_t *last_ids,
                                      const float *v_render_colors, const float *v_render_alphas, float *v_rows,
                                      const int32_t *skip_flag, void *stream) {
    using namespace ubs;
    UBS_CHECK_ARG(C >= 0 && N >= 0 && width > 0 && height > 0, "rasterize_bwd_rows: bad sizes");
    UBS_CHECK_ARG(tile_size == kTile, "rasterize_bwd_rows: tile_size must be %d (got %d)", kTile, tile_size);
    if (C == 0 || N == 0 || isect_capacity == 0) return UBS_OK;  // no pairs: every gradient stays zero
    UBS_CHECK_ARG(n_isects && offsets && splats && flatten_ids && render_alphas && last_ids && v_render_colors &&
                      v_render_alphas && v_rows,
                  "rasterize_bwd_rows: null pointer");
    UBS_CHECK_ARG((((uintptr_t)splats | (uintptr_t)v_rows) & 15) == 0,
                  "rasterize_bwd_rows: splats / v_rows must be 16-byte aligned");
    return launch_bwd_rows(C, N, n_isects, isect_capacity, splats, backgrounds, masks, width, height, offsets, flatten_ids,
                           render_alphas, last_ids, v_render_colors, v_render_alphas, v_rows, skip_flag, (cudaStream_t)stream);
}
