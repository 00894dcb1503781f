// K10: per-tile front-to-back alpha compositing with the Beta falloff (forward).
// Drop-in for rasterize_to_pixels_fwd (rasterize_to_pixels_fwd.cu:16-191): same sigma, alpha, early-termination
// and last_ids semantics.  Differences in *how*: colours are staged through shared memory with the rest of the
// 2-D record (the reference re-reads them from global memory per accepted pair), the gather of batch b+1 is
// issued before batch b is composited (register double-buffering), each warp owns an 8x4 pixel sub-tile and first
// compacts the batch to the pairs whose sigma < 1 ellipse can touch those 32 pixels (the tile lists themselves stay
// bit-identical to the reference's -- the cull is internal), and a warp whose pixels are all done skips the batch.
#include "common.cuh"
#include "raster_common.cuh"

namespace ubs {

namespace {

template <int CH>
__global__ void __launch_bounds__(kTilePixels, CH <= 4 ? 5 : 2)
rasterize_fwd_kernel(int C, int64_t N, const int64_t *__restrict__ n_isects_dev, int64_t isect_capacity,
                     const float2 *__restrict__ means2d, const float *__restrict__ conics,
                     const float *__restrict__ colors, const float *__restrict__ opacities,
                     const float *__restrict__ betas, const float *__restrict__ backgrounds,
                     const uint8_t *__restrict__ masks, uint32_t width, uint32_t height, uint32_t tile_width,
                     uint32_t tile_height, const int32_t *__restrict__ tile_offsets,
                     const int32_t *__restrict__ flatten_ids, float *__restrict__ render_colors,
                     float *__restrict__ render_alphas, int32_t *__restrict__ last_ids,
                     const float4 *__restrict__ splats, bool splat_colors) {
    constexpr bool kPacked = CH <= 4;        // colour rides in one float4 of the staged record
    constexpr int kColW = kPacked ? 4 : CH;  // floats of colour per staged pair
    const uint32_t cam = blockIdx.z;
    const uint32_t tile_id = blockIdx.y * tile_width + blockIdx.x;
    const uint32_t tr = threadIdx.x;
    const SubTile st = sub_tile_of(tr);
    const uint32_t i = blockIdx.y * kTile + st.py;
    const uint32_t j = blockIdx.x * kTile + st.px;
    const float px = (float)j + 0.5f, py = (float)i + 0.5f;
    const bool inside = (i < height && j < width);
    const size_t pix = ((size_t)cam * height + i) * width + j;

    tile_offsets += (size_t)cam * tile_height * tile_width;
    if (backgrounds != nullptr) backgrounds += cam * CH;

    if (masks != nullptr && !masks[(size_t)cam * tile_height * tile_width + tile_id]) {
        // masked-out tile: background colour only (rasterize_to_pixels_fwd.cu:73-79)
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
    const int32_t n_pairs = range_end - range_start;
    const int32_t num_batches = (n_pairs + kTilePixels - 1) / kTilePixels;

    // staged 2-D records of one 256-pair batch (array of structs: one address computation per pair in the hot loop)
    struct __align__(16) Staged {
        float4 xyob;   // mean2d.x, mean2d.y, opacity, beta
        float4 conic;  // conic a, 2b, c, (unused)
        float col[kColW];
    };
    constexpr int kUnroll = 4;
    __shared__ Staged s_rec[kTilePixels + 1];  // [kTilePixels] = sentinel whose sigma is NaN (pads the lists)
    __shared__ __align__(8) uint8_t s_mask[kTilePixels];  // sub-tiles the (inflated) sigma < 1 box of each staged pair can touch (0 past the batch)
    // per-warp compacted lists of shared-window ADDRESSES of the staged records (one LDS.128 fetches the four addresses
    // of an unrolled round), padded to a multiple of kUnroll with the sentinel
    __shared__ __align__(16) uint32_t s_list[kTilePixels / 32][kTilePixels + kUnroll];

    const float tx0 = (float)(blockIdx.x * kTile) + 0.5f, ty0 = (float)(blockIdx.y * kTile) + 0.5f;  // first pixel centre
    const uint32_t lane = tr & 31, warp = tr >> 5;
    uint32_t *my_list = s_list[warp];
    const uint32_t rec_addr = smem_addr(s_rec), list_addr = smem_addr(my_list);
    const float kNaN = __int_as_float(0x7fffffff);
    if (tr == 0) {
        s_rec[kTilePixels].xyob = make_float4(kNaN, kNaN, 0.f, 1.f);
        s_rec[kTilePixels].conic = make_float4(1.f, 0.f, 1.f, 0.f);
    }

    // "done" is carried in the sign of T: a finished pixel (or one outside the image) holds -T, so that
    // next_T = T (1 - alpha) is negative, fails `next_T > 1e-4`, and the lane idles through the rest of the list
    // without a separate predicate in the loop; |T| is the transmittance before the primitive that tripped it.
    float T = inside ? 1.f : -1.f;
    int32_t cur_idx = 0;
    float pix_out[CH];
#pragma unroll
    for (int k = 0; k < CH; ++k) pix_out[k] = 0.f;

    // register double buffer: the record this thread will publish for the next batch
    float4 r_xyob = make_float4(0.f, 0.f, 0.f, 0.f), r_conic = r_xyob;
    float r_color[kColW];
#pragma unroll
    for (int k = 0; k < kColW; ++k) r_color[k] = 0.f;
    auto gather = [&](int32_t batch) {
        const int32_t idx = range_start + batch * kTilePixels + (int32_t)tr;
        if (idx < range_end) {
            const int32_t g = flatten_ids[idx];
            if (splats != nullptr) {
                // one 48-byte row per primitive (written by the fused projection kernel): two sectors per gather
                // instead of the five of the separate arrays -- the kernel is sensitive to what fits in L1
                const float4 *sp = splats + (size_t)g * 3;
                r_xyob = sp[0];
                r_conic = sp[1];
                if ((CH == 1 || CH == 3 || CH == 4) && splat_colors) {
                    // colours out of the row itself: RGB (CH = 3), RGB + depth (CH = 4, render modes "RGB+D" / "RGB+ED"),
                    // depth alone (CH = 1, "Depth" / "EDepth" / "Normal") -- reference: rendering.py:131-142
                    if constexpr (CH == 1) {
                        r_color[0] = r_conic.w;
                    } else {
                        const float4 c4 = sp[2];
                        r_color[0] = c4.x;
                        if constexpr (CH > 1) r_color[1] = c4.y;
                        if constexpr (CH > 2) r_color[2] = c4.z;
                        if constexpr (CH > 3) r_color[3] = c4.w;
                    }
                } else {
#pragma unroll
                    for (int k = 0; k < CH; ++k) r_color[k] = colors[(size_t)g * CH + k];
                }
            } else {
                const float2 xy = means2d[g];
                r_xyob = make_float4(xy.x, xy.y, opacities[g], betas[g]);
                r_conic = make_float4(conics[(size_t)g * 3], conics[(size_t)g * 3 + 1], conics[(size_t)g * 3 + 2], 0.f);
#pragma unroll
                for (int k = 0; k < CH; ++k) r_color[k] = colors[(size_t)g * CH + k];
            }
        }
    };
    if (num_batches > 0) gather(0);

    for (int32_t b = 0; b < num_batches; ++b) {
        // everyone has finished reading the previous batch; stop when every pixel of the tile is done
        if (__syncthreads_count(T < 0.f) >= kTilePixels) break;
        s_rec[tr].xyob = r_xyob;
        s_mask[tr] = (int32_t)tr < range_end - (range_start + b * kTilePixels)
                         ? (uint8_t)sub_tile_mask(support_bbox(r_xyob.x, r_xyob.y, r_conic.x, r_conic.y, r_conic.z), tx0, ty0)
                         : (uint8_t)0;
        s_rec[tr].conic = make_float4(r_conic.x, r_conic.y + r_conic.y, r_conic.z, 0.f);  // b + b as the reference forms it
        if constexpr (kPacked) {
            *reinterpret_cast<float4 *>(s_rec[tr].col) = make_float4(r_color[0], r_color[1], r_color[2], r_color[3]);
        } else {
#pragma unroll
            for (int k = 0; k < CH; ++k) s_rec[tr].col[k] = r_color[k];
        }
        __syncthreads();
        if (b + 1 < num_batches) gather(b + 1);  // in flight while this batch is composited

        const int32_t batch_start = range_start + b * kTilePixels;
        const int32_t batch_size = min((int32_t)kTilePixels, range_end - batch_start);
        if (__all_sync(0xffffffffu, T < 0.f)) continue;  // whole warp finished: nothing to composite

        // warp-level cull: keep (in order) only the pairs whose support touches this warp's 8x4 pixels.  Lane l takes the
        // eight staged pairs 8 l .. 8 l + 7 (one 64-bit load of their masks), one warp scan of the per-lane hit counts
        // places them: ~50 instructions per 256-pair batch instead of eight ballot rounds of ~29
        const unsigned long long m8 = *reinterpret_cast<const unsigned long long *>(s_mask + 8 * lane);
        const unsigned long long bits = (m8 >> warp) & 0x0101010101010101ull;
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
            const uint32_t first = rec_addr + 8u * lane * (uint32_t)sizeof(Staged);
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const uint32_t word = k < 4 ? lo : hi;
                if ((word >> (8 * (k & 3))) & 1u) my_list[pos++] = first + (uint32_t)k * (uint32_t)sizeof(Staged);
            }
        }
        if (lane < kUnroll) my_list[cnt + lane] = rec_addr + (uint32_t)(kTilePixels * sizeof(Staged));
        __syncwarp();

        uint32_t last_rec = 0;
        for (uint32_t t = 0; t < cnt; t += kUnroll) {
          const uint4 recs = lds_u4(list_addr + 4 * t);
#pragma unroll
          for (int u = 0; u < kUnroll; ++u) {
            const uint32_t rec = u == 0 ? recs.x : u == 1 ? recs.y : u == 2 ? recs.z : recs.w;
            const float4 xyob = lds_f4(rec);
            const float4 conic = lds_f4(rec + 16);
            const float dx = xyob.x - px, dy = xyob.y - py;
            // same association as the reference's compiled form (rasterize_to_pixels_fwd.cu:143-145):
            // sigma = fma(dy, dx * (b + b), fma(dx, a * dx, dy * (c * dy)))
            const float sigma = __fmaf_rn(dy, dx * conic.y, __fmaf_rn(dx, conic.x * dx, dy * (conic.z * dy)));
            // sigma in [0, 1)  <=>  its bit pattern is below that of 1.0f (negatives and NaN compare above)
            if (__float_as_uint(sigma) >= 0x3f800000u) continue;
            const float alpha = fminf(0.999f, xyob.z * __powf(1.f - sigma, xyob.w));
            const float next_T = T * (1.f - alpha);
            if (!(next_T > 1e-4f)) {  // this pixel is done (now or earlier); the primitive is NOT accumulated
                T = set_sign(T);
                continue;
            }
            const float vis = alpha * T;
            if constexpr (kPacked) {
                const float4 col = lds_f4(rec + 32);
                pix_out[0] += col.x * vis;
                if constexpr (CH > 1) pix_out[1] += col.y * vis;
                if constexpr (CH > 2) pix_out[2] += col.z * vis;
                if constexpr (CH > 3) pix_out[3] += col.w * vis;
            } else {
#pragma unroll
                for (int k = 0; k < CH; ++k)
                    pix_out[k] += reinterpret_cast<const Staged *>(reinterpret_cast<const unsigned char *>(s_rec) + (rec - rec_addr))->col[k] * vis;
            }
            last_rec = rec;
            T = next_T;
          }
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

// Diagnostic (not on the hot path): work counters for the roofline bookkeeping of SURVEY.md 8(d).
//   counts[0] = E_test: (pixel, pair) evaluations the reference algorithm performs (every pair of the tile list
//               until the pixel is done, rasterize_to_pixels_fwd.cu:135-173)
//   counts[1] = E_acc : evaluations that pass the sigma test (alpha is computed)
//   counts[2] = E_cull: evaluations left after this library's per-warp 8x4 sub-tile cull (what it executes)
//   counts[3] = pairs staged: 256-pair batches loaded until the whole tile is done, in pairs
//   counts[4] = E_any : lane evaluations if a warp's 8x4 sub-tile evaluated exactly the pairs that some pixel of
//               it accepts (floor of E_cull at this granularity); [5]/[6] = the same two numbers for 4x4 blocks
//               (bbox cull / exact), [7] = exact for 8x2 half-warps.  Design exploration only.
__global__ void __launch_bounds__(kTilePixels)
rasterize_count_kernel(int C, const int64_t *__restrict__ n_isects_dev, int64_t isect_capacity,
                       const float2 *__restrict__ means2d, const float *__restrict__ conics,
                       const float *__restrict__ opacities, const float *__restrict__ betas, uint32_t width,
                       uint32_t height, uint32_t tile_width, uint32_t tile_height,
                       const int32_t *__restrict__ tile_offsets, const int32_t *__restrict__ flatten_ids,
                       unsigned long long *__restrict__ counts) {
    const uint32_t cam = blockIdx.z;
    const uint32_t tile_id = blockIdx.y * tile_width + blockIdx.x;
    const uint32_t tr = threadIdx.x;
    const SubTile st = sub_tile_of(tr);
    const uint32_t i = blockIdx.y * kTile + st.py, j = blockIdx.x * kTile + st.px;
    const float px = (float)j + 0.5f, py = (float)i + 0.5f;
    const bool inside = (i < height && j < width);
    tile_offsets += (size_t)cam * tile_height * tile_width;
    const int64_t n_isects = min(*n_isects_dev, isect_capacity);
    const int32_t range_start = tile_offsets[tile_id];
    const int32_t range_end = (cam == (uint32_t)C - 1 && tile_id == tile_width * tile_height - 1)
                                  ? (int32_t)n_isects
                                  : tile_offsets[tile_id + 1];
    const float wx0 = (float)(blockIdx.x * kTile + st.bx * kSubW) + 0.5f, wx1 = wx0 + (float)(kSubW - 1);
    const float wy0 = (float)(blockIdx.y * kTile + st.by * kSubH) + 0.5f, wy1 = wy0 + (float)(kSubH - 1);
    unsigned long long n_test = 0, n_acc = 0, n_cull = 0, n_any = 0, n_cull4 = 0, n_any4 = 0, n_any8x2 = 0;
    int32_t stop = range_start;  // one past the last pair this pixel looks at
    {
        const uint32_t lane = tr & 31;
        const uint32_t bx4 = (lane & 7) >> 2;
        const uint32_t mask4 = bx4 ? 0xF0F0F0F0u : 0x0F0F0F0Fu;
        const uint32_t mask8x2 = lane < 16 ? 0x0000FFFFu : 0xFFFF0000u;
        const float qx0 = wx0 + 4.f * (float)bx4, qx1 = qx0 + 3.f;
        float T = 1.f;
        bool done = !inside;
        if (inside) stop = range_end;
        for (int32_t idx = range_start; idx < range_end; ++idx) {
            if (__all_sync(0xffffffffu, done)) break;
            const int32_t g = flatten_ids[idx];
            const float2 xy = means2d[g];
            const float a = conics[(size_t)g * 3], b = conics[(size_t)g * 3 + 1], c = conics[(size_t)g * 3 + 2];
            const float4 bb = support_bbox(xy.x, xy.y, a, b, c);
            const float dx = xy.x - px, dy = xy.y - py;
            const float sigma = (a * dx * dx + c * dy * dy) + 2.f * b * dx * dy;
            const bool acc = !done && !(sigma < 0.f || sigma >= 1.f);
            const uint32_t m = __ballot_sync(0xffffffffu, acc);
            if (done) continue;
            ++n_test;
            if ((bb.x <= wx1) && (bb.y >= wx0) && (bb.z <= wy1) && (bb.w >= wy0)) ++n_cull;
            if ((bb.x <= qx1) && (bb.y >= qx0) && (bb.z <= wy1) && (bb.w >= wy0)) ++n_cull4;
            if (m) ++n_any;
            if (m & mask4) ++n_any4;
            if (m & mask8x2) ++n_any8x2;
            if (!acc) continue;
            ++n_acc;
            const float alpha = fminf(0.999f, opacities[g] * __powf(1.f - sigma, betas[g]));
            const float next_T = T * (1.f - alpha);
            if (next_T <= 1e-4f) {
                stop = idx + 1;
                done = true;
                continue;
            }
            T = next_T;
        }
    }
    __shared__ unsigned long long s_cnt[7];
    __shared__ int32_t s_stop;
    if (tr == 0) {
        for (int k = 0; k < 7; ++k) s_cnt[k] = 0ull;
        s_stop = range_start;
    }
    __syncthreads();
    atomicAdd(&s_cnt[0], n_test);
    atomicAdd(&s_cnt[1], n_acc);
    atomicAdd(&s_cnt[2], n_cull);
    atomicAdd(&s_cnt[3], n_any);
    atomicAdd(&s_cnt[4], n_cull4);
    atomicAdd(&s_cnt[5], n_any4);
    atomicAdd(&s_cnt[6], n_any8x2);
    atomicMax(&s_stop, stop);
    __syncthreads();
    if (tr == 0) {
        atomicAdd(counts + 0, s_cnt[0]);
        atomicAdd(counts + 1, s_cnt[1]);
        atomicAdd(counts + 2, s_cnt[2]);
        const int32_t staged = min(range_end - range_start, (s_stop - range_start + kTilePixels - 1) / kTilePixels * kTilePixels);
        atomicAdd(counts + 3, (unsigned long long)max(staged, 0));
        atomicAdd(counts + 4, s_cnt[3]);
        atomicAdd(counts + 5, s_cnt[4]);
        atomicAdd(counts + 6, s_cnt[5]);
        atomicAdd(counts + 7, s_cnt[6]);
    }
}

template <int CH>
int launch_fwd(int C, int64_t N, const int64_t *n_isects, int64_t cap, const float *means2d, const float *conics,
               const float *colors, const float *opacities, const float *betas, const float *backgrounds,
               const uint8_t *masks, int width, int height, const int32_t *offsets, const int32_t *flatten_ids,
               float *render_colors, float *render_alphas, int32_t *last_ids, const float *splats, int splat_colors,
               cudaStream_t s) {
    const uint32_t tw = (uint32_t)ceil_div(width, kTile), th = (uint32_t)ceil_div(height, kTile);
    dim3 grid(tw, th, (unsigned)C), block(kTilePixels, 1, 1);
    rasterize_fwd_kernel<CH><<<grid, block, 0, s>>>(C, N, n_isects, cap, (const float2 *)means2d, conics, colors, opacities,
                                                    betas, backgrounds, masks, (uint32_t)width, (uint32_t)height, tw,
                                                    th, offsets, flatten_ids, render_colors, render_alphas, last_ids,
                                                    (const float4 *)splats, splat_colors != 0);
    UBS_LAUNCH_CHECK("rasterize_fwd_kernel");
    return UBS_OK;
}

}  // namespace
}  // namespace ubs

static int rasterize_fwd_impl(int C, int64_t N, const int64_t *n_isects, int64_t isect_capacity,
                                 const float *means2d, const float *conics,
                                 const float *colors, const float *opacities, const float *betas,
                                 const float *backgrounds, const uint8_t *masks, int channels, int width, int height,
                                 int tile_size, const int32_t *offsets, const int32_t *flatten_ids,
                                 float *render_colors, float *render_alphas, int32_t *last_ids, const float *splats,
                             int splat_colors, void *stream) {
    using namespace ubs;
    UBS_CHECK_ARG(C >= 0 && N >= 0 && width > 0 && height > 0, "rasterize_fwd: bad sizes");
    UBS_CHECK_ARG(tile_size == kTile, "rasterize_fwd: tile_size must be %d (got %d)", kTile, tile_size);
    if (C == 0) return UBS_OK;
    UBS_CHECK_ARG(n_isects && offsets && render_colors && render_alphas && last_ids, "rasterize_fwd: null pointer");
    UBS_CHECK_ARG(N == 0 || (((means2d && conics && opacities && betas) || splats) && (colors || (splats && splat_colors)) &&
                             (flatten_ids || isect_capacity == 0)),
                  "rasterize_fwd: null primitive arrays");
    UBS_CHECK_ARG(C <= 65535, "rasterize_fwd: C=%d exceeds 65535", C);
    UBS_CHECK_ARG(((uintptr_t)splats & 15) == 0, "rasterize_fwd: splats must be 16-byte aligned");
    UBS_CHECK_ARG(!splat_colors || channels == 3 || channels == 4 || channels == 1,
                  "rasterize_fwd: splat colours are RGB, RGB+depth or depth (channels = %d)", channels);
    cudaStream_t s = (cudaStream_t)stream;
#define UBS_FWD_CASE(CH)                                                                                               \
    case CH:                                                                                                           \
        return launch_fwd<CH>(C, N, n_isects, isect_capacity, means2d, conics, colors, opacities, betas, backgrounds, masks, width,    \
                              height, offsets, flatten_ids, render_colors, render_alphas, last_ids, splats, splat_colors, s);
    switch (channels) {
        UBS_FWD_CASE(1)
        UBS_FWD_CASE(2)
        UBS_FWD_CASE(3)
        UBS_FWD_CASE(4)
        UBS_FWD_CASE(8)
        UBS_FWD_CASE(16)
        default:
            set_error("rasterize_fwd: unsupported channel count %d (supported: 1,2,3,4,8,16; pad on the host)",
                      channels);
            return UBS_EUNSUPPORTED;
    }
#undef UBS_FWD_CASE
}

extern "C" int ubs_rasterize_fwd(int C, int64_t N, const int64_t *n_isects, int64_t isect_capacity,
                                 const float *means2d, const float *conics,
                                 const float *colors, const float *opacities, const float *betas,
                                 const float *backgrounds, const uint8_t *masks, int channels, int width, int height,
                                 int tile_size, const int32_t *offsets, const int32_t *flatten_ids,
                                 float *render_colors, float *render_alphas, int32_t *last_ids, void *stream) {
    return rasterize_fwd_impl(C, N, n_isects, isect_capacity, means2d, conics, colors, opacities, betas, backgrounds, masks,
                              channels, width, height, tile_size, offsets, flatten_ids, render_colors, render_alphas,
                              last_ids, nullptr, 0, stream);
}

extern "C" int ubs_rasterize_fwd_splats(int C, int64_t N, const int64_t *n_isects, int64_t isect_capacity,
                                        const float *splats, const float *colors, const float *backgrounds,
                                        const uint8_t *masks, int channels, int width, int height, int tile_size,
                                        const int32_t *offsets, const int32_t *flatten_ids, float *render_colors,
                                        float *render_alphas, int32_t *last_ids, void *stream) {
    using namespace ubs;
    UBS_CHECK_ARG(splats != nullptr || N == 0 || C == 0, "rasterize_fwd_splats: splats is null");
    return rasterize_fwd_impl(C, N, n_isects, isect_capacity, nullptr, nullptr, colors, nullptr, nullptr, backgrounds, masks,
                              channels, width, height, tile_size, offsets, flatten_ids, render_colors, render_alphas,
                              last_ids, splats, colors == nullptr ? 1 : 0, stream);
}

extern "C" int ubs_rasterize_count(int C, const int64_t *n_isects, int64_t isect_capacity, const float *means2d,
                                   const float *conics, const float *opacities, const float *betas, int width,
                                   int height, int tile_size, const int32_t *offsets, const int32_t *flatten_ids,
                                   unsigned long long *counts, void *stream) {
    using namespace ubs;
    UBS_CHECK_ARG(C > 0 && width > 0 && height > 0 && tile_size == kTile, "rasterize_count: bad sizes");
    UBS_CHECK_ARG(n_isects && means2d && conics && opacities && betas && offsets && flatten_ids && counts,
                  "rasterize_count: null pointer");
    cudaStream_t s = (cudaStream_t)stream;
    UBS_CUDA_TRY(cudaMemsetAsync(counts, 0, 8 * sizeof(unsigned long long), s));
    const uint32_t tw = (uint32_t)ceil_div(width, kTile), th = (uint32_t)ceil_div(height, kTile);
    rasterize_count_kernel<<<dim3(tw, th, (unsigned)C), kTilePixels, 0, s>>>(
        C, n_isects, isect_capacity, (const float2 *)means2d, conics, opacities, betas, (uint32_t)width,
        (uint32_t)height, tw, th, offsets, flatten_ids, counts);
    UBS_LAUNCH_CHECK("rasterize_count_kernel");
    return UBS_OK;
}
