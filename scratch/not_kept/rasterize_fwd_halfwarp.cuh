// NOT COMPILED: the 4x4 half-warp-list forward kernel measured in round 2 (cfg3: 0.542 ms against 0.534 ms for the
// 8x4 warp-list kernel, images / last_ids bit-identical; profiles/r2_experiments.md).  Kept for the record.
// ---- half-warp variant -------------------------------------------------------------------------------------------
// Same results as rasterize_fwd_kernel for CH <= 4, with the cull at 4x4-pixel granularity: the two half-warps of a warp
// own the left / right 4x4 block of its 8x4 pixels and walk their OWN compacted lists (a shared-memory load whose lanes
// name two different addresses costs the same as a broadcast on sm_100a: scratch/ub/lds_bcast.cu), so a pair that
// touches only one of the two blocks is evaluated by 16 lanes instead of 32.  A staged batch of 256 pairs is compacted
// in two passes of 128 (lists of <= 128 + padding entries keep the CTA at the shared-memory footprint of the 8x4
// kernel: what its per-pair gather finds in L1 is what bounds it).
template <int CH>
__global__ void __launch_bounds__(kTilePixels, 5)
rasterize_fwd_hw_kernel(int C, int64_t N, const int64_t *__restrict__ n_isects_dev, int64_t isect_capacity,
                        const float *__restrict__ colors, const float *__restrict__ backgrounds,
                        const uint8_t *__restrict__ masks, uint32_t width, uint32_t height, uint32_t tile_width,
                        uint32_t tile_height, const int32_t *__restrict__ tile_offsets,
                        const int32_t *__restrict__ flatten_ids, float *__restrict__ render_colors,
                        float *__restrict__ render_alphas, int32_t *__restrict__ last_ids,
                        const float4 *__restrict__ splats, bool splat_colors) {
    static_assert(CH <= 4, "packed colours only");
    constexpr int kSub = kTilePixels / 2;  // pairs per compaction pass
    constexpr int kUnroll = 4;
    constexpr int kListCap = kSub + kUnroll;
    const uint32_t cam = blockIdx.z;
    const uint32_t tile_id = blockIdx.y * tile_width + blockIdx.x;
    const uint32_t tr = threadIdx.x, lane = tr & 31, warp = tr >> 5, half = lane >> 4, l16 = lane & 15;
    // warp w: 8x4 pixels at (8 (w & 1), 4 (w >> 1)); half-warp h: its 4x4 block at x offset 4 h
    const uint32_t pxl = (warp & 1) * 8 + half * 4 + (l16 & 3), pyl = (warp >> 1) * 4 + (l16 >> 2);
    const uint32_t i = blockIdx.y * kTile + pyl, j = blockIdx.x * kTile + pxl;
    const float px = (float)j + 0.5f, py = (float)i + 0.5f;
    const bool inside = (i < height && j < width);
    const size_t pix = ((size_t)cam * height + i) * width + j;

    tile_offsets += (size_t)cam * tile_height * tile_width;
    if (backgrounds != nullptr) backgrounds += cam * CH;
    if (masks != nullptr && !masks[(size_t)cam * tile_height * tile_width + tile_id]) {
        if (inside) {
#pragma unroll
            for (int k = 0; k < CH; ++k) render_colors[pix * CH + k] = backgrounds == nullptr ? 0.f : backgrounds[k];
        }
        return;
    }
    const int64_t n_isects = min(*n_isects_dev, isect_capacity);
    const int32_t range_start = tile_offsets[tile_id];
    const int32_t range_end = (cam == (uint32_t)C - 1 && tile_id == tile_width * tile_height - 1)
                                  ? (int32_t)n_isects
                                  : tile_offsets[tile_id + 1];
    const int32_t num_batches = (range_end - range_start + kTilePixels - 1) / kTilePixels;

    struct __align__(16) Staged {
        float4 xyob;   // mean2d.x, mean2d.y, opacity, beta
        float4 conic;  // conic a, 2b, c, (unused)
        float4 col;
    };
    __shared__ Staged s_rec[kTilePixels + 1];               // [kTilePixels] = sentinel whose sigma is NaN
    __shared__ __align__(16) uint16_t s_mask[kTilePixels];  // bit 4 by + bx: 4x4 block (bx, by) of the tile is touched
    __shared__ __align__(16) uint32_t s_list[kTilePixels / 32][2][kListCap];

    const float tx0 = (float)(blockIdx.x * kTile) + 0.5f, ty0 = (float)(blockIdx.y * kTile) + 0.5f;
    const uint32_t rec_addr = smem_addr(s_rec);
    const uint32_t list_lo = smem_addr(s_list[warp][0]), list_mine = list_lo + half * (uint32_t)(kListCap * 4);
    const uint32_t sentinel = rec_addr + (uint32_t)(kTilePixels * sizeof(Staged));
    const uint32_t bit0 = (warp >> 1) * 4 + (warp & 1) * 2;  // mask bit of this warp's left block; the right one follows
    const float kNaN = __int_as_float(0x7fffffff);
    if (tr == 0) {
        s_rec[kTilePixels].xyob = make_float4(kNaN, kNaN, 0.f, 1.f);
        s_rec[kTilePixels].conic = make_float4(1.f, 0.f, 1.f, 0.f);
        s_rec[kTilePixels].col = make_float4(0.f, 0.f, 0.f, 0.f);
    }

    float T = inside ? 1.f : -1.f;  // sign = done (see rasterize_fwd_kernel)
    int32_t cur_idx = 0;
    float pix_out[CH];
#pragma unroll
    for (int k = 0; k < CH; ++k) pix_out[k] = 0.f;

    float4 r_xyob = make_float4(0.f, 0.f, 0.f, 0.f), r_conic = r_xyob, r_col = r_xyob;
    auto gather = [&](int32_t batch) {
        const int32_t idx = range_start + batch * kTilePixels + (int32_t)tr;
        if (idx < range_end) {
            const int32_t g = flatten_ids[idx];
            const float4 *sp = splats + (size_t)g * 3;
            r_xyob = sp[0];
            r_conic = sp[1];
            if (splat_colors) {
                if constexpr (CH == 1) r_col = make_float4(r_conic.w, 0.f, 0.f, 0.f);
                else r_col = sp[2];
            } else {
                float c[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
                for (int k = 0; k < CH; ++k) c[k] = colors[(size_t)g * CH + k];
                r_col = make_float4(c[0], c[1], c[2], c[3]);
            }
        }
    };
    if (num_batches > 0) gather(0);

    for (int32_t b = 0; b < num_batches; ++b) {
        if (__syncthreads_count(T < 0.f) >= kTilePixels) break;
        const int32_t batch_start = range_start + b * kTilePixels;
        s_rec[tr].xyob = r_xyob;
        {
            uint32_t m = 0;
            if ((int32_t)tr < range_end - batch_start) {
                const float4 bb = support_bbox(r_xyob.x, r_xyob.y, r_conic.x, r_conic.y, r_conic.z);
                uint32_t mx = 0, my = 0;
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    if (bb.x <= tx0 + (float)(4 * k + 3) && bb.y >= tx0 + (float)(4 * k)) mx |= 1u << k;
                    if (bb.z <= ty0 + (float)(4 * k + 3) && bb.w >= ty0 + (float)(4 * k)) my |= 1u << (4 * k);
                }
                m = mx * my;  // bit 4 by + bx = mx[bx] & my[by]
            }
            s_mask[tr] = (uint16_t)m;
        }
        s_rec[tr].conic = make_float4(r_conic.x, r_conic.y + r_conic.y, r_conic.z, 0.f);
        s_rec[tr].col = r_col;
        __syncthreads();
        if (b + 1 < num_batches) gather(b + 1);
        if (__all_sync(0xffffffffu, T < 0.f)) continue;

        uint32_t last_rec = 0;
#pragma unroll 1
        for (int sub = 0; sub < 2; ++sub) {
            // lane l takes the four staged pairs sub * 128 + 4 l .. + 3 (one 64-bit load of their masks); one warp scan of
            // the packed (left, right) hit counts places them in both lists, order preserved
            const uint2 m4 = *reinterpret_cast<const uint2 *>(s_mask + sub * kSub + 4 * lane);
            const uint32_t w0 = m4.x >> bit0, w1 = m4.y >> bit0;  // pairs (0, 1) and (2, 3): bits 0/1 and 16/17
            const uint32_t n_mine = (uint32_t)__popc(w0 & 0x00010001u) + (uint32_t)__popc(w1 & 0x00010001u) +
                                    (((uint32_t)__popc(w0 & 0x00020002u) + (uint32_t)__popc(w1 & 0x00020002u)) << 16);
            uint32_t incl = n_mine;
#pragma unroll
            for (int off = 1; off < 32; off <<= 1) {
                const uint32_t t = __shfl_up_sync(0xffffffffu, incl, off);
                if ((int)lane >= off) incl += t;
            }
            const uint32_t total = __shfl_sync(0xffffffffu, incl, 31);
            const uint32_t cnt_l = total & 0xffffu, cnt_r = total >> 16;
            const uint32_t n_it = (max(cnt_l, cnt_r) + kUnroll - 1) & ~(uint32_t)(kUnroll - 1);
            {
                const uint32_t excl = incl - n_mine;
                uint32_t pl = excl & 0xffffu, pr = excl >> 16;
                const uint32_t first = rec_addr + (uint32_t)(sub * kSub + 4 * lane) * (uint32_t)sizeof(Staged);
                uint32_t *ll = s_list[warp][0], *lr = s_list[warp][1];
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const uint32_t w = (k < 2 ? w0 : w1) >> (16 * (k & 1));
                    const uint32_t a = first + (uint32_t)k * (uint32_t)sizeof(Staged);
                    if (w & 1u) ll[pl++] = a;
                    if (w & 2u) lr[pr++] = a;
                }
                // pad both lists with the sentinel up to the common (unrolled) trip count
                for (uint32_t q = cnt_l + lane; q < n_it; q += 32) ll[q] = sentinel;
                for (uint32_t q = cnt_r + lane; q < n_it; q += 32) lr[q] = sentinel;
            }
            __syncwarp();

            for (uint32_t t = 0; t < n_it; t += kUnroll) {
                const uint4 recs = lds_u4(list_mine + 4 * t);
#pragma unroll
                for (int u = 0; u < kUnroll; ++u) {
                    const uint32_t rec = u == 0 ? recs.x : u == 1 ? recs.y : u == 2 ? recs.z : recs.w;
                    const float4 xyob = lds_f4(rec);
                    const float4 conic = lds_f4(rec + 16);
                    const float dx = xyob.x - px, dy = xyob.y - py;
                    const float sigma = __fmaf_rn(dy, dx * conic.y, __fmaf_rn(dx, conic.x * dx, dy * (conic.z * dy)));
                    if (__float_as_uint(sigma) >= 0x3f800000u) continue;
                    const float alpha = fminf(0.999f, xyob.z * __powf(1.f - sigma, xyob.w));
                    const float next_T = T * (1.f - alpha);
                    if (!(next_T > 1e-4f)) {
                        T = set_sign(T);
                        continue;
                    }
                    const float vis = alpha * T;
                    const float4 col = lds_f4(rec + 32);
                    pix_out[0] += col.x * vis;
                    if constexpr (CH > 1) pix_out[1] += col.y * vis;
                    if constexpr (CH > 2) pix_out[2] += col.z * vis;
                    if constexpr (CH > 3) pix_out[3] += col.w * vis;
                    last_rec = rec;
                    T = next_T;
                }
            }
            __syncwarp();  // the lists are rewritten by the next pass
        }
        if (last_rec != 0) cur_idx = batch_start + (int32_t)((last_rec - rec_addr) / (uint32_t)sizeof(Staged));
    }

    if (inside) {
        T = fabsf(T);
        render_alphas[pix] = 1.f - T;
#pragma unroll
        for (int k = 0; k < CH; ++k)
            render_colors[pix * CH + k] = backgrounds == nullptr ? pix_out[k] : (pix_out[k] + T * backgrounds[k]);
        last_ids[pix] = cur_idx;
    }
}

