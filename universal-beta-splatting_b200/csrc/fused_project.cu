// North-star kernel (1): fused activation + covariance build + conditioning + projection + tile count, straight
// from the PACKED primitive records.  No single reference counterpart: it replaces the chain
//   softplus/sigmoid/exp/cat/contiguous (scene/beta_model.py:103-121) -> K1 -> K2 -> view-dir glue (:675-690)
//   -> K3 -> 3x3->6 gather (rendering.py:55-56) -> K5 -> first pass of K7
// which in the reference is ~25 launches and ~1.3 KB of HBM traffic per primitive, with one pass that reads the
// 144 B (D=6) / 176 B (D=7) record once and writes 52 B per (camera, primitive).
//
// Record staging: each CTA owns kFusedThreads consecutive records = one contiguous, 16-byte aligned span of
// global memory, fetched with ONE TMA bulk copy (cp.async.bulk ... mbarrier::complete_tx) into shared memory;
// threads then read their own record with conflict-free 128-bit shared loads (row strides of 36 / 44 floats map
// each quarter-warp onto 32 distinct banks).
#include <algorithm>

#include "common.cuh"
#include "cond_math.cuh"
#include "isect.cuh"
#include "optim.cuh"
#include "pack.cuh"
#include "proj_math.cuh"

namespace ubs {
namespace {

constexpr int kFusedThreads = 128;

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_bulk_g2s(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t phase) {
    uint32_t done = 0;
    while (!done) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(smem_u32(bar)), "r"(phase)
            : "memory");
    }
}

// Everything about one primitive that does not depend on the camera.
template <int D>
struct PrimState {
    static constexpr int C = D - 3;
    float xyz[3];
    float mu2[C];
    float rgb[3];
    float opacity;       // sigmoid
    float beta0;         // spatial beta
    float beta_c[C];     // conditional betas
    CondPrep<C> prep;
};

// |v| exactly as torch computes `view_dir.norm(dim=-1)` on CUDA (scene/beta_model.py:677): its reduction kernel gives a
// size-3 row to TWO threads -- one squares and adds elements 0 and 2, the other squares element 1 -- and then adds the
// partial sums: sqrt((x^2 + z^2) + y^2), every operation rounded to FP32 (3,000,000 of 3,000,000 norms bit-equal on
// the B200 box against 88.7 % for the left-to-right order, scratch/norm_order.py); IEEE sqrt and division, as ATen's.
__device__ __forceinline__ float view_norm(float dx, float dy, float dz) {
    return __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dz, dz)), __fmul_rn(dy, dy)));
}

__device__ __forceinline__ void prefetch_l1(const void *p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }

// `activated`: the record already holds the ACTIVATED values (softplus'd scales, sigmoid'd opacity, 4 exp'd betas), as
// the reference's operator chain sees them (scene/beta_model.py:103-121) -- the zero-edit drop-in route packs those.
template <int D>
__device__ __forceinline__ void decode_record(const float *rec, PrimState<D> &ps, bool activated) {
    constexpr int C = D - 3, M = NdDims<D>::M;
    // layout (ubs_b200.h): xyz | mean | rgb | opacity | beta(D-2) | scale(D) | l_triangle(M)
#pragma unroll
    for (int k = 0; k < 3; ++k) ps.xyz[k] = rec[k];
#pragma unroll
    for (int k = 0; k < C; ++k) ps.mu2[k] = rec[3 + k];
#pragma unroll
    for (int k = 0; k < 3; ++k) ps.rgb[k] = rec[D + k];
    ps.opacity = activated ? rec[D + 3] : sigmoid_f(rec[D + 3]);
    ps.beta0 = activated ? rec[D + 4] : beta_act_f(rec[D + 4]);
#pragma unroll
    for (int k = 0; k < C; ++k) ps.beta_c[k] = activated ? rec[D + 5 + k] : beta_act_f(rec[D + 5 + k]);
    float s[D], lt[M];
#pragma unroll
    for (int k = 0; k < D; ++k) s[k] = activated ? rec[2 * D + 2 + k] : softplus_f(rec[2 * D + 2 + k]);
#pragma unroll
    for (int k = 0; k < M; ++k) lt[k] = rec[3 * D + 2 + k];
    const float R[9] = {1.f, lt[0], lt[1], -lt[0], 1.f, lt[2], -lt[1], -lt[2], 1.f};  // K1: I + skew
    float L[D * D], S[D * D];
    build_L<D>(R, s, lt, L);
    covar_from_L<D>(L, s, S, D);
    float V11[9], V12[3 * C], V21[C * 3], V22[C * C];
#pragma unroll
    for (int r = 0; r < 3; ++r) {
#pragma unroll
        for (int c = 0; c < 3; ++c) V11[r * 3 + c] = S[r * D + c];
#pragma unroll
        for (int c = 0; c < C; ++c) {
            V12[r * C + c] = S[r * D + 3 + c];
            V21[c * 3 + r] = S[(3 + c) * D + r];
        }
    }
#pragma unroll
    for (int r = 0; r < C; ++r)
#pragma unroll
        for (int c = 0; c < C; ++c) V22[r * C + c] = S[(3 + r) * D + 3 + c];
    cond_prepare<C>(V11, V12, V21, V22, ps.beta_c, ps.prep);
}

template <int D>
__global__ void __launch_bounds__(kFusedThreads)
fused_project_fwd_kernel(int C, int64_t N, const float *__restrict__ records, const float *__restrict__ viewmats,
                         const float *__restrict__ Ks, const float *__restrict__ cam_pos,
                         const float *__restrict__ timestamps, const uint8_t *__restrict__ prim_mask, uint32_t width,
                         uint32_t height, float eps2d, float near_plane, float far_plane, float radius_clip,
                         int calc_comp, uint32_t tile_size, uint32_t tile_width, uint32_t tile_height,
                         int32_t *__restrict__ radii, float *__restrict__ means2d, float *__restrict__ depths,
                         float *__restrict__ conics, float *__restrict__ opacities, float *__restrict__ betas,
                         float *__restrict__ colors, int32_t *__restrict__ tiles_per_gauss,
                         float4 *__restrict__ splats, int32_t *__restrict__ tile_delta, int activated,
                         const float *__restrict__ query) {
    constexpr int Cd = D - 3;
    constexpr int STRIDE = UBS_RECORD_STRIDE(D);
    __shared__ __align__(128) float s_rec[kFusedThreads * STRIDE];
    __shared__ __align__(8) uint64_t s_bar;

    const int64_t base = (int64_t)blockIdx.x * kFusedThreads;
    const int n_here = (int)min((int64_t)kFusedThreads, N - base);
    if (threadIdx.x == 0) mbar_init(&s_bar, 1);
    __syncthreads();
    if (threadIdx.x == 0) {
        const uint32_t bytes = (uint32_t)n_here * STRIDE * sizeof(float);
        mbar_expect_tx(&s_bar, bytes);
        tma_bulk_g2s(s_rec, records + base * STRIDE, bytes, &s_bar);
    }
    mbar_wait(&s_bar, 0);

    const int64_t gid = base + threadIdx.x;
    const bool active = threadIdx.x < n_here && (prim_mask == nullptr || prim_mask[gid]);
    PrimState<D> ps;
    if (active) {
        float rec[STRIDE];
        const float4 *src = reinterpret_cast<const float4 *>(s_rec + threadIdx.x * STRIDE);
#pragma unroll
        for (int k = 0; k < STRIDE / 4; ++k) {
            const float4 v = src[k];
            rec[4 * k + 0] = v.x, rec[4 * k + 1] = v.y, rec[4 * k + 2] = v.z, rec[4 * k + 3] = v.w;
        }
        decode_record<D>(rec, ps, activated != 0);
    }
    // query given by the caller (the drop-in route: scene/beta_model.py:675-690 builds it with torch): one row per
    // primitive, the same for every camera
    float xq[Cd];
    if (query != nullptr && active) {
#pragma unroll
        for (int k = 0; k < Cd; ++k) xq[k] = query[gid * Cd + k] - ps.mu2[k];
    }

    for (int cid = blockIdx.y; cid < C; cid += gridDim.y) {
        if (threadIdx.x >= n_here) continue;
        const int64_t idx = (int64_t)cid * N + gid;
        Splat2D o;
        o.radius = 0;
        o.mean2d[0] = o.mean2d[1] = o.depth = o.conic[0] = o.conic[1] = o.conic[2] = o.compensation = 0.f;
        float opac = 0.f;
        if (active) {
            const Cam cam = load_cam(viewmats + cid * 16, Ks + cid * 9);
            // query: unit view direction (+ timestamp)   (scene/beta_model.py:675-690)
            float x[Cd];
            if (query != nullptr) {
#pragma unroll
                for (int k = 0; k < Cd; ++k) x[k] = xq[k];
            } else {
                const float dx = ps.xyz[0] - cam_pos[cid * 3 + 0], dy = ps.xyz[1] - cam_pos[cid * 3 + 1],
                            dz = ps.xyz[2] - cam_pos[cid * 3 + 2];
                const float nrm = view_norm(dx, dy, dz);
                x[0] = __fdiv_rn(dx, nrm) - ps.mu2[0];
                x[1] = __fdiv_rn(dy, nrm) - ps.mu2[1];
                x[2] = __fdiv_rn(dz, nrm) - ps.mu2[2];
                if constexpr (Cd > 3) {
                    x[3] = timestamps[cid] - ps.mu2[3];
#pragma unroll
                    for (int k = 4; k < Cd; ++k) x[k] = -ps.mu2[k];
                }
            }
            float mean[3];
            cond_apply<Cd>(ps.prep, ps.xyz, x, ps.opacity, ps.beta_c, mean, opac);
            // upper triangle of the (unsymmetrised) conditional covariance, as rendering.py:55-56 gathers it
            const float s6[6] = {ps.prep.cov[0], ps.prep.cov[1], ps.prep.cov[2], ps.prep.cov[4], ps.prep.cov[5], ps.prep.cov[8]};
            o = project_splat(cam, mean, s6, width, height, eps2d, near_plane, far_plane, radius_clip);
            if (calc_comp) opac *= o.compensation;
        }
        int32_t cnt = 0;
        if (o.radius > 0) {
            const TileRect t = tile_rect(o.mean2d[0], o.mean2d[1], o.radius, tile_size, tile_width, tile_height);
            cnt = (int32_t)((t.y1 - t.y0) * (t.x1 - t.x0));
            if (tile_delta != nullptr && cnt > 0)
                add_tile_deltas(tile_delta + (size_t)cid * ((tile_height + 1) * (tile_width + 1)) * kDeltaStride, t,
                                tile_width);
        }
        radii[idx] = o.radius;
        reinterpret_cast<float2 *>(means2d)[idx] = make_float2(o.mean2d[0], o.mean2d[1]);
        depths[idx] = o.depth;
        // the rest of the separate arrays is what `meta` and the projection backward read; a render-only frame
        // (conics == NULL) composites from the splat rows alone and saves 36 bytes of writes per primitive
        if (conics != nullptr) {
            conics[idx * 3 + 0] = o.conic[0];
            conics[idx * 3 + 1] = o.conic[1];
            conics[idx * 3 + 2] = o.conic[2];
            opacities[idx] = o.radius > 0 ? opac : 0.f;
            betas[idx] = active ? ps.beta0 : 0.f;
            if (colors != nullptr) {
                colors[idx * 3 + 0] = active ? ps.rgb[0] : 0.f;
                colors[idx * 3 + 1] = active ? ps.rgb[1] : 0.f;
                colors[idx * 3 + 2] = active ? ps.rgb[2] : 0.f;
            }
            tiles_per_gauss[idx] = cnt;
        }
        // the same screen-space record once more as ONE 48-byte row for the compositing kernels: their per-pair gather
        // then touches two sectors instead of five (the compositing forward is L1-sensitive: 0.62 -> 0.53 ms)
        if (splats != nullptr && o.radius > 0) {
            float4 *sp = splats + idx * 3;
            sp[0] = make_float4(o.mean2d[0], o.mean2d[1], opac, ps.beta0);
            sp[1] = make_float4(o.conic[0], o.conic[1], o.conic[2], o.depth);
            sp[2] = make_float4(ps.rgb[0], ps.rgb[1], ps.rgb[2], o.depth);  // .w: the 4th colour channel of "RGB+D"
        }
    }
}


__device__ __forceinline__ void tma_bulk_s2g(void *dst_gmem, const void *src_smem, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst_gmem), "r"(smem_u32(src_smem)),
                 "r"(bytes)
                 : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}

__device__ __forceinline__ void tma_bulk_s2g_issue(void *dst_gmem, const void *src_smem, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst_gmem), "r"(smem_u32(src_smem)),
                 "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void tma_bulk_commit_wait() {
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}

// Destination of the gradient tiles when the rows are sharded over the ranks of one NVSwitch box.
struct ScatterDst {
    float *staging[UBS_MAX_RANKS];  // staging[g] = rank g's [world][shard_rows][stride] buffer (peer-mapped)
    int64_t shard_rows;             // 0 = plain v_records
    int rank;
};

// Sharded train step, pull form: the owner of a shard of rows reads every rank's screen-space gradient rows of that shard
// straight from the ranks' (peer-mapped) v_rows buffers and stores the updated parameter tiles into every rank's records.
struct PullSrc {
    const float *rows[UBS_MAX_RANKS];  // rows[c] = rank c's [N, 12] gradient rows (camera c of the step)
    float *records[UBS_MAX_RANKS];     // records[g] = rank g's [N, stride] parameters
    int world;
    int64_t row0;                      // first row of this rank's shard
};

// Options shared by the entry points of the backward kernel.
struct BwdOpts {
    int activated;             // records hold activated values (see decode_record): no activation derivative
    const float *query;        // NULL, or [N, D-3] queries given by the caller instead of the view direction
    const int32_t *skip_flag;  // NULL, or the `status` word of the frame's tile-list build
    const float *v_rows;       // [C, N, 12] screen-space gradient rows (layout: include/ubs_b200.h)
    int moment_form;           // rows of ubs_rasterize_bwd_rows (1) or of ubs_pack_gradient_rows (0)
    float *v_viewmats;         // POSE instantiation: [C, 4, 4] gradient of the world-to-camera matrices (zeroed by the host)
    PackSegs v_segs;           // plain form: when v_records is NULL, the gradient tile leaves as the seven separate arrays
};

// a 48-byte gradient row with all ten gradient slots +-0: the primitive received nothing from that view
__device__ __forceinline__ bool pull_row_nonzero(const float *row) {
    const float4 *r = reinterpret_cast<const float4 *>(row);
    const float4 a = r[0], b = r[1];
    const float2 c = *reinterpret_cast<const float2 *>(row + 8);
    const uint32_t bits = __float_as_uint(a.x) | __float_as_uint(a.y) | __float_as_uint(a.z) | __float_as_uint(a.w) |
                          __float_as_uint(b.x) | __float_as_uint(b.y) | __float_as_uint(b.z) | __float_as_uint(b.w) |
                          __float_as_uint(c.x) | __float_as_uint(c.y);
    return (bits << 1) != 0u;
}

// Backward of fused_project_fwd_kernel: one thread per primitive, looping over cameras, so the per-primitive sums
// need no atomics.  Recomputes the cheap forward intermediates from the record instead of storing them.
//
// ADAM = true (single-GPU, batch-1 training: nothing has to be summed over ranks or views first): the Adam moment
// tiles of the CTA's records are fetched with two more bulk copies at kernel start, so they land during the long
// gradient computation; every thread then applies torch.optim.Adam to its own row straight from the gradient in
// its registers and the updated parameters and moments go back with three bulk stores.  The gradient records are
// never written: 6 record-sized HBM streams instead of the 2 + 7 of backward-then-optimiser.
//
// PULL = true (with ADAM; sharded train step): the "cameras" are the `world` ranks' views, N counts the rows of this
// rank's shard, `records_in` / the moments point at the shard; the gradient rows of the CTA's 128 primitives are fetched
// from every rank's v_rows with one bulk copy per rank (6 KB each, over NVLink) next to the record and moment tiles;
// visibility is "the row is non-zero", the conic is recomputed (the forward outputs of the other ranks' cameras are not
// here), and the updated record tile goes to every rank with one bulk store each.
//
// POSE = true (plain form only): also the gradient of the C world-to-camera matrices through the projection
// (fully_fused_projection_bwd.cu:178-201 -- not through the view direction of the conditioning, which the reference
// detaches): per-(camera, primitive) contributions are summed in shared memory and leave as one atomic per CTA and entry.
constexpr int kPoseCams = 16;  // cameras with a shared-memory accumulator per CTA (more: global atomics per thread)
template <int D, int MINB, bool ADAM, bool PULL = false, bool POSE = false>
__global__ void __launch_bounds__(kFusedThreads, MINB)
fused_project_bwd_kernel(int C, int64_t N, const float *__restrict__ records_in, const float *__restrict__ viewmats,
                         const float *__restrict__ Ks, const float *__restrict__ cam_pos,
                         const float *__restrict__ timestamps, uint32_t width, uint32_t height, float eps2d,
                         int calc_comp, const int32_t *__restrict__ radii, const float *__restrict__ conics,
                         float *__restrict__ v_records, float *__restrict__ exp_avg,
                         float *__restrict__ exp_avg_sq, const AdamParams adam, const ScatterDst scatter,
                         const BwdOpts opts, const PullSrc pull) {
    static_assert(!PULL || ADAM, "the pull form updates the parameters");
    static_assert(!POSE || (!ADAM && !PULL), "camera gradients come out of the plain form");
    __shared__ float s_pose[POSE ? kPoseCams * 12 : 1];
    if constexpr (POSE) {
        for (int k = threadIdx.x; k < kPoseCams * 12; k += kFusedThreads) s_pose[k] = 0.f;  // visible after the barriers below
    }
    constexpr int Cd = D - 3, M = NdDims<D>::M;
    // the frame this gradient belongs to lost pairs to the capacity bound (isect.cuh: report_truncation): with ADAM the
    // update is not applied at all, otherwise the view contributes a zero gradient (it is dropped from the batch)
    const bool skip = opts.skip_flag != nullptr && *opts.skip_flag != 0;
    if (ADAM && skip) return;
    const bool activated = opts.activated != 0;
    constexpr int STRIDE = UBS_RECORD_STRIDE(D);
    extern __shared__ __align__(128) float s_tiles[];  // [records][exp_avg][exp_avg_sq] (the last two with ADAM)
    float *const s_rec = s_tiles;
    float *const s_m = s_tiles + kFusedThreads * STRIDE;
    float *const s_v = s_tiles + 2 * kFusedThreads * STRIDE;
    float *const s_rows = s_tiles + 3 * kFusedThreads * STRIDE;  // PULL: [world][128][12] gradient rows
    __shared__ __align__(8) uint64_t s_bar, s_bar2;
    const float *records = records_in;

    const int64_t base = (int64_t)blockIdx.x * kFusedThreads;
    const int n_here = (int)min((int64_t)kFusedThreads, N - base);
    const uint32_t bytes = (uint32_t)n_here * STRIDE * sizeof(float);
    if (threadIdx.x == 0) {
        mbar_init(&s_bar, 1);
        if constexpr (ADAM) mbar_init(&s_bar2, 1);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        const uint32_t row_bytes = (uint32_t)n_here * 12u * sizeof(float);
        mbar_expect_tx(&s_bar, PULL ? bytes + (uint32_t)C * row_bytes : bytes);
        tma_bulk_g2s(s_rec, records + base * STRIDE, bytes, &s_bar);
        if constexpr (PULL) {
            for (int c = 0; c < C; ++c)
                tma_bulk_g2s(s_rows + c * kFusedThreads * 12, pull.rows[c] + (pull.row0 + base) * 12, row_bytes, &s_bar);
        }
        if constexpr (ADAM) {
            mbar_expect_tx(&s_bar2, 2 * bytes);
            tma_bulk_g2s(s_m, exp_avg + base * STRIDE, bytes, &s_bar2);
            tma_bulk_g2s(s_v, exp_avg_sq + base * STRIDE, bytes, &s_bar2);
        }
    }

    // Compaction: only primitives visible in some camera have a non-zero gradient.  Their row indices are packed to
    // the front, so the long recompute + backward chain below runs in ceil(n_vis / 32) fully populated warps and
    // the remaining warps go straight to the zero fill (about half the primitives of a frame are culled).
    __shared__ int s_row[kFusedThreads];
    __shared__ int s_wcnt[kFusedThreads / 32];
    int my_row = -1, any_row = -1;
    {
        bool vis = false;
        if constexpr (PULL) {
            mbar_wait(&s_bar, 0);  // records and gradient rows have landed
            if (threadIdx.x < n_here) {
                for (int cid = 0; cid < C && !vis; ++cid)
                    vis = pull_row_nonzero(s_rows + (cid * kFusedThreads + threadIdx.x) * 12);
            }
        } else if (threadIdx.x < n_here) {
            int cid = 0;
            for (; cid < C && !vis && !skip; ++cid) vis = radii[(int64_t)cid * N + base + threadIdx.x] > 0;
            if (vis) {
                // The screen-space record and its gradients are gathered per visible row far below, behind the bulk-copy
                // wait and the forward recompute; start those lines towards L1 now (first camera that sees the row).
                // The kernel is latency bound: 0.69 -> 0.63 ms with the Adam epilogue at 3M primitives.
                const int64_t idx = (int64_t)(cid - 1) * N + base + threadIdx.x;
                prefetch_l1(conics + idx * 3);
                prefetch_l1(conics + idx * 3 + 2);
                prefetch_l1(opts.v_rows + idx * 12);  // one 48-byte row: two sectors
                prefetch_l1(opts.v_rows + idx * 12 + 8);
            }
        }
        const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
        const uint32_t m = __ballot_sync(0xffffffffu, vis);
        if (lane == 0) s_wcnt[warp] = __popc(m);
        __syncthreads();
        int before = 0, n_vis = 0;
#pragma unroll
        for (int w = 0; w < kFusedThreads / 32; ++w) {
            const int c = s_wcnt[w];
            if (w < (int)warp) before += c;
            n_vis += c;
        }
        const int vis_before = before + __popc(m & ((1u << lane) - 1u));
        if (vis) s_row[vis_before] = (int)threadIdx.x;
        // with ADAM the culled rows are listed behind the visible ones: every row is updated by exactly one thread
        if (ADAM && !vis && (int)threadIdx.x < n_here) s_row[n_vis + ((int)threadIdx.x - vis_before)] = (int)threadIdx.x;
        __syncthreads();
        if ((int)threadIdx.x < n_vis) my_row = s_row[threadIdx.x];
        if (ADAM && (int)threadIdx.x < n_here) any_row = s_row[threadIdx.x];
    }
    if constexpr (!PULL) mbar_wait(&s_bar, 0);  // the records have landed (the visibility loads above overlapped the bulk copy)
    const int64_t gid = base + my_row;
    const bool active = my_row >= 0;
    float grad[STRIDE];
#pragma unroll
    for (int k = 0; k < STRIDE; ++k) grad[k] = 0.f;

    if (active) {
        float rec[STRIDE];
        const float4 *src = reinterpret_cast<const float4 *>(s_rec + my_row * STRIDE);
#pragma unroll
        for (int k = 0; k < STRIDE / 4; ++k) {
            const float4 v = src[k];
            rec[4 * k + 0] = v.x, rec[4 * k + 1] = v.y, rec[4 * k + 2] = v.z, rec[4 * k + 3] = v.w;
        }
        // ---- recompute the camera-independent forward state ------------------------------------------------
        float xyz[3], mu2[Cd], beta_c[Cd], s[D], lt[M];
#pragma unroll
        for (int k = 0; k < 3; ++k) xyz[k] = rec[k];
#pragma unroll
        for (int k = 0; k < Cd; ++k) mu2[k] = rec[3 + k];
        const float o_in = activated ? rec[D + 3] : sigmoid_fast(rec[D + 3]);
        const float beta0 = activated ? rec[D + 4] : beta_act_fast(rec[D + 4]);
#pragma unroll
        for (int k = 0; k < Cd; ++k) beta_c[k] = activated ? rec[D + 5 + k] : beta_act_fast(rec[D + 5 + k]);
#pragma unroll
        for (int k = 0; k < D; ++k) s[k] = activated ? rec[2 * D + 2 + k] : softplus_fast(rec[2 * D + 2 + k]);
#pragma unroll
        for (int k = 0; k < M; ++k) lt[k] = rec[3 * D + 2 + k];
        const float R[9] = {1.f, lt[0], lt[1], -lt[0], 1.f, lt[2], -lt[1], -lt[2], 1.f};
        float L[D * D], S[D * D];
        build_L<D>(R, s, lt, L);
        covar_from_L<D>(L, s, S, D);
        float V11[9], V12[3 * Cd], V21[Cd * 3], V22[Cd * Cd];
#pragma unroll
        for (int r = 0; r < 3; ++r) {
#pragma unroll
            for (int c = 0; c < 3; ++c) V11[r * 3 + c] = S[r * D + c];
#pragma unroll
            for (int c = 0; c < Cd; ++c) {
                V12[r * Cd + c] = S[r * D + 3 + c];
                V21[c * 3 + r] = S[(3 + c) * D + r];
            }
        }
#pragma unroll
        for (int r = 0; r < Cd; ++r)
#pragma unroll
            for (int c = 0; c < Cd; ++c) V22[r * Cd + c] = S[(3 + r) * D + 3 + c];
        CondPrep<Cd> prep;
        cond_prepare<Cd>(V11, V12, V21, V22, beta_c, prep);
        const float s6[6] = {prep.cov[0], prep.cov[1], prep.cov[2], prep.cov[4], prep.cov[5], prep.cov[8]};

        // ---- accumulate over cameras ------------------------------------------------------------------------
        float G[D * D];  // gradient w.r.t. the full D x D covariance
#pragma unroll
        for (int k = 0; k < D * D; ++k) G[k] = 0.f;
        float g_mu[D], g_o = 0.f, g_beta0 = 0.f, g_beta_c[Cd], g_rgb[3] = {0.f, 0.f, 0.f};
#pragma unroll
        for (int k = 0; k < D; ++k) g_mu[k] = 0.f;
#pragma unroll
        for (int k = 0; k < Cd; ++k) g_beta_c[k] = 0.f;

        for (int cid = 0; cid < C; ++cid) {
            const int64_t idx = (int64_t)cid * N + gid;
            const float4 *row = PULL ? reinterpret_cast<const float4 *>(s_rows + (cid * kFusedThreads + my_row) * 12)
                                     : reinterpret_cast<const float4 *>(opts.v_rows + idx * 12);
            if constexpr (PULL) {
                if (!pull_row_nonzero(reinterpret_cast<const float *>(row))) continue;
            } else {
                if (radii[idx] <= 0) continue;
            }
            const Cam cam = load_cam(viewmats + cid * 16, Ks + cid * 9);
            float x[Cd];
            if (opts.query != nullptr) {
#pragma unroll
                for (int k = 0; k < Cd; ++k) x[k] = opts.query[gid * Cd + k] - mu2[k];
            } else {
                const float dx = xyz[0] - cam_pos[cid * 3 + 0], dy = xyz[1] - cam_pos[cid * 3 + 1],
                            dz = xyz[2] - cam_pos[cid * 3 + 2];
                const float nrm = view_norm(dx, dy, dz);
                x[0] = __fdiv_rn(dx, nrm) - mu2[0];
                x[1] = __fdiv_rn(dy, nrm) - mu2[1];
                x[2] = __fdiv_rn(dz, nrm) - mu2[2];
                if constexpr (Cd > 3) {
                    x[3] = timestamps[cid] - mu2[3];
#pragma unroll
                    for (int k = 4; k < Cd; ++k) x[k] = -mu2[k];
                }
            }
            float mean[3], o_cond;
            cond_apply<Cd>(prep, xyz, x, o_in, beta_c, mean, o_cond);

            float conic[3], comp = 0.f;
            if constexpr (PULL) {
                // the conic of the forward pass, recomputed (same function, same inputs as fused_project_fwd_kernel)
                const Splat2D f = project_splat(cam, mean, s6, width, height, eps2d, -3.0e38f, 3.0e38f, -1.f);
                conic[0] = f.conic[0], conic[1] = f.conic[1], conic[2] = f.conic[2];
                comp = f.compensation;
            } else {
                conic[0] = conics[idx * 3], conic[1] = conics[idx * 3 + 1], conic[2] = conics[idx * 3 + 2];
            }
            // one 48-byte gradient row; the two forms differ in three slots (selects, no branch: the code below is the
            // same in every instantiation, which keeps the ADAM and plain kernels bit-identical)
            const float4 q0 = row[0], q1 = row[1], q2 = row[2];
            const bool moment = opts.moment_form != 0;
            const float v_col[3] = {q0.x, q0.y, q0.z};
            const float vc[3] = {q0.w, moment ? q1.x + q1.x : q1.x, q1.y};
            const float a2 = conic[0] + conic[0], b2 = conic[1] + conic[1], c2 = conic[2] + conic[2];
            const float vm[2] = {moment ? __fmaf_rn(a2, q1.z, __fmul_rn(b2, q1.w)) : q1.z,
                                 moment ? __fmaf_rn(b2, q1.z, __fmul_rn(c2, q1.w)) : q1.w};
            float v_o = q2.x;
            const float v_b = moment ? __fmul_rn(q2.y, 0.693147180559945f) : q2.y, v_d = q2.z;
            float v_comp = 0.f;
            if (calc_comp) {
                if constexpr (!PULL) {
                    const Splat2D f = project_splat(cam, mean, s6, width, height, eps2d, -3.0e38f, 3.0e38f, -1.f);
                    comp = f.compensation;
                }
                v_comp = v_o * o_cond;
                v_o = v_o * comp;
            }
            float v_mean[3] = {0.f, 0.f, 0.f}, v_s6[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
            if constexpr (POSE) {
                float vR[9] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f}, vt[3] = {0.f, 0.f, 0.f};
                project_splat_vjp(cam, mean, s6, width, height, eps2d, conic, calc_comp ? &comp : nullptr, vm, v_d, vc,
                                  calc_comp ? &v_comp : nullptr, v_mean, v_s6, vR, vt);
#pragma unroll
                for (int k = 0; k < 12; ++k) {
                    const float v = k < 9 ? vR[k] : vt[k - 9];
                    if (cid < kPoseCams) atomicAdd(&s_pose[cid * 12 + k], v);
                    else atomicAdd(opts.v_viewmats + cid * 16 + (k < 9 ? (k / 3) * 4 + k % 3 : (k - 9) * 4 + 3), v);
                }
            } else {
                project_splat_vjp(cam, mean, s6, width, height, eps2d, conic, calc_comp ? &comp : nullptr, vm, v_d, vc,
                                  calc_comp ? &v_comp : nullptr, v_mean, v_s6, nullptr, nullptr);
            }
            // index-backward of the 3x3 -> 6 gather: only the upper triangle receives gradient (rendering.py:55-56)
            const float gV[9] = {v_s6[0], v_s6[1], v_s6[2], 0.f, v_s6[3], v_s6[4], 0.f, 0.f, v_s6[5]};
            float g_mu1[3], g_mu2[Cd], g11[9], g12[3 * Cd], g21[Cd * 3], g22[Cd * Cd], go, gb[Cd];
            cond_backward<Cd>(x, V12, V21, V22, o_in, beta_c, v_mean, gV, v_o, g_mu1, g_mu2, g11, g12, g21, g22, go, gb);
#pragma unroll
            for (int k = 0; k < 3; ++k) g_mu[k] += g_mu1[k];
#pragma unroll
            for (int k = 0; k < Cd; ++k) {
                g_mu[3 + k] += g_mu2[k];
                g_beta_c[k] += gb[k];
            }
            g_o += go;
            g_beta0 += v_b;
#pragma unroll
            for (int k = 0; k < 3; ++k) g_rgb[k] += v_col[k];
#pragma unroll
            for (int r = 0; r < 3; ++r) {
#pragma unroll
                for (int c = 0; c < 3; ++c) G[r * D + c] += g11[r * 3 + c];
#pragma unroll
                for (int c = 0; c < Cd; ++c) {
                    G[r * D + 3 + c] += g12[r * Cd + c];
                    G[(3 + c) * D + r] += g21[c * 3 + r];
                }
            }
#pragma unroll
            for (int r = 0; r < Cd; ++r)
#pragma unroll
                for (int c = 0; c < Cd; ++c) G[(3 + r) * D + 3 + c] += g22[r * Cd + c];
        }

        // ---- covariance build, K1 and activation backward ---------------------------------------------------
        float vR[9], vs[D], vlt[M];
        covar_vjp<D>(R, s, L, G, D, vR, vs, vlt);
        vlt[0] += vR[1] - vR[3];
        vlt[1] += vR[2] - vR[6];
        vlt[2] += vR[5] - vR[7];
#pragma unroll
        for (int k = 0; k < D; ++k) grad[k] = g_mu[k];  // xyz | mean (no gradient through the view direction)
#pragma unroll
        for (int k = 0; k < 3; ++k) grad[D + k] = g_rgb[k];
        grad[D + 3] = activated ? g_o : g_o * o_in * (1.f - o_in);  // sigmoid'
        grad[D + 4] = activated ? g_beta0 : g_beta0 * beta0;         // d(4 e^x)/dx = 4 e^x
#pragma unroll
        for (int k = 0; k < Cd; ++k) grad[D + 5 + k] = activated ? g_beta_c[k] : g_beta_c[k] * beta_c[k];
#pragma unroll
        for (int k = 0; k < D; ++k) {
            const float raw = rec[2 * D + 2 + k];
            grad[2 * D + 2 + k] = activated ? vs[k] : vs[k] * (raw > 20.f ? 1.f : sigmoid_fast(raw));  // softplus'
        }
#pragma unroll
        for (int k = 0; k < M; ++k) grad[3 * D + 2 + k] = vlt[k];
    }

    if constexpr (ADAM) {
        mbar_wait(&s_bar2, 0);  // the moment tiles landed long ago
        if (any_row >= 0) {
            const int64_t row = (PULL ? pull.row0 : 0) + base + any_row;
            float4 *p4 = reinterpret_cast<float4 *>(s_rec + any_row * STRIDE);
            float4 *m4 = reinterpret_cast<float4 *>(s_m + any_row * STRIDE);
            float4 *v4 = reinterpret_cast<float4 *>(s_v + any_row * STRIDE);
            // regularisers (train.py:122-124): statically indexed columns, so the unrolled update below stays lean
            if (adam.reg_opacity != 0.f) grad[D + 3] += adam_reg_grad(adam, row, D + 3, s_rec[any_row * STRIDE + D + 3]);
            if (adam.reg_scale != 0.f && row < 3) {
#pragma unroll
                for (int k = 0; k < D; ++k)
                    grad[2 * D + 2 + k] += adam_reg_grad(adam, row, 2 * D + 2 + k, s_rec[any_row * STRIDE + 2 * D + 2 + k]);
            }
            // moments from the gradient in registers (unrolled, a few instructions per element) ...
#pragma unroll
            for (int k = 0; k < STRIDE / 4; ++k) {
                float4 m = m4[k], v = v4[k];
                adam_moments(adam, grad[4 * k + 0], m.x, v.x);
                adam_moments(adam, grad[4 * k + 1], m.y, v.y);
                adam_moments(adam, grad[4 * k + 2], m.z, v.z);
                adam_moments(adam, grad[4 * k + 3], m.w, v.w);
                m4[k] = m, v4[k] = v;
            }
            // ... then the parameter update as a rolled loop (IEEE sqrt and division: kept out of the unrolled code)
#pragma unroll 1
            for (int k = 0; k < STRIDE / 4; ++k) {
                float4 p = p4[k];
                const float4 m = m4[k], v = v4[k];
                p.x = adam_apply(adam, adam.step_size[4 * k + 0], p.x, m.x, v.x);
                p.y = adam_apply(adam, adam.step_size[4 * k + 1], p.y, m.y, v.y);
                p.z = adam_apply(adam, adam.step_size[4 * k + 2], p.z, m.z, v.z);
                p.w = adam_apply(adam, adam.step_size[4 * k + 3], p.w, m.w, v.w);
                p4[k] = p;
            }
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncthreads();
        if (threadIdx.x == 0) {
            if constexpr (PULL) {
                for (int g = 0; g < pull.world; ++g)
                    tma_bulk_s2g_issue(pull.records[g] + (pull.row0 + base) * STRIDE, s_rec, bytes);
            } else {
                tma_bulk_s2g_issue(const_cast<float *>(records_in) + base * STRIDE, s_rec, bytes);
            }
            tma_bulk_s2g_issue(exp_avg + base * STRIDE, s_m, bytes);
            tma_bulk_s2g_issue(exp_avg_sq + base * STRIDE, s_v, bytes);
            tma_bulk_commit_wait();
        }
        return;
    }

    // ---- stage the gradient records in shared memory and write them with one TMA bulk store --------------------
    __syncthreads();  // all record reads done; reuse s_rec
    if constexpr (POSE) {
        for (int k = threadIdx.x; k < min(C, kPoseCams) * 12; k += kFusedThreads) {
            const int cid = k / 12, e = k - cid * 12;
            const float v = s_pose[k];
            if (v != 0.f) atomicAdd(opts.v_viewmats + cid * 16 + (e < 9 ? (e / 3) * 4 + e % 3 : (e - 9) * 4 + 3), v);
        }
    }
    if ((int)threadIdx.x < n_here) {  // zero rows for the culled primitives (every row, then the live ones overwrite)
        float4 *dst = reinterpret_cast<float4 *>(s_rec + threadIdx.x * STRIDE);
#pragma unroll
        for (int k = 0; k < STRIDE / 4; ++k) dst[k] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    __syncthreads();
    if (active) {
        float4 *dst = reinterpret_cast<float4 *>(s_rec + my_row * STRIDE);
#pragma unroll
        for (int k = 0; k < STRIDE / 4; ++k) dst[k] = make_float4(grad[4 * k], grad[4 * k + 1], grad[4 * k + 2], grad[4 * k + 3]);
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    if (v_records == nullptr && scatter.shard_rows == 0) {
        // the drop-in route: gradients straight into the reference's seven tensors (ubs_unpack_records fused in)
        unpack_tile<D, kFusedThreads>(s_rec, opts.v_segs, base, n_here);
        return;
    }
    if (threadIdx.x == 0) {
        float *dst = v_records + base * STRIDE;
        if (scatter.shard_rows > 0) {
            // multi-GPU scatter: this CTA's 128 rows belong to one owner rank (shard_rows is a multiple of 128); the
            // tile goes straight into this rank's slot of the owner's staging buffer -- a peer address over NVLink
            const int64_t owner = base / scatter.shard_rows, local = base - owner * scatter.shard_rows;
            dst = scatter.staging[owner] + ((int64_t)scatter.rank * scatter.shard_rows + local) * STRIDE;
        }
        tma_bulk_s2g(dst, s_rec, bytes);
    }
}

}  // namespace
}  // namespace ubs

extern "C" int ubs_fused_project_fwd(int C, int64_t N, int D, const float *records, const float *viewmats,
                                     const float *Ks, const float *cam_pos, const float *timestamps,
                                     const uint8_t *prim_mask, int width, int height, float eps2d, float near_plane,
                                     float far_plane, float radius_clip, int calc_compensations, int tile_size,
                                     int tile_width, int tile_height, int32_t *radii, float *means2d, float *depths,
                                     float *conics, float *opacities, float *betas, float *colors,
                                     int32_t *tiles_per_gauss, float *splats, int32_t *tile_delta, int64_t *n_isects,
                                     void *workspace, size_t workspace_bytes, int activated, const float *query,
                                     void *stream) {
    using namespace ubs;
    UBS_CHECK_ARG(C >= 0 && N >= 0 && width > 0 && height > 0 && tile_size > 0, "fused_project_fwd: bad sizes");
    UBS_CHECK_ARG(((uintptr_t)splats & 15) == 0, "fused_project_fwd: splats must be 16-byte aligned");
    UBS_CHECK_ARG(D == 6 || D == 7, "fused_project_fwd: D must be 6 or 7 (got %d)", D);
    UBS_CHECK_ARG(n_isects != nullptr || tile_delta != nullptr, "fused_project_fwd: n_isects is null");
    UBS_CHECK_ARG(tile_width > 0 && tile_height > 0, "fused_project_fwd: bad tile grid");
    cudaStream_t s = (cudaStream_t)stream;
    const int64_t CN = (int64_t)C * N;
    // tile binning route (ubs_isect_bin_sort with deltas_ready = 1): zero this frame's corner-delta grid
    if (tile_delta != nullptr && C > 0)
        UBS_CUDA_TRY(cudaMemsetAsync(tile_delta, 0, bin_delta_bytes(C, tile_width, tile_height), s));
    if (CN == 0) {
        if (n_isects != nullptr) UBS_CUDA_TRY(cudaMemsetAsync(n_isects, 0, sizeof(int64_t), s));
        return UBS_OK;
    }
    UBS_CHECK_ARG(records && viewmats && Ks && (cam_pos || query) && radii && means2d && depths &&
                      (workspace || tile_delta),
                  "fused_project_fwd: null pointer");
    UBS_CHECK_ARG((conics && opacities && betas && tiles_per_gauss) || (!conics && splats && tile_delta),
                  "fused_project_fwd: conics / opacities / betas / tiles_per_gauss go together; a render-only frame "
                  "(conics == NULL) needs the splat rows and the tile-binning route");
    UBS_CHECK_ARG(D != 7 || timestamps != nullptr || query != nullptr, "fused_project_fwd: D=7 needs timestamps");
    UBS_CHECK_ARG(((uintptr_t)records & 15) == 0, "fused_project_fwd: records must be 16-byte aligned");
    UBS_CHECK_ARG(CN < ((int64_t)1 << 31), "fused_project_fwd: C*N must fit int32 flatten ids");
    if (tile_delta == nullptr && workspace_bytes < ubs_isect_workspace_bytes(CN, 0)) {
        set_error("fused_project_fwd: workspace %zu < %zu", workspace_bytes, ubs_isect_workspace_bytes(CN, 0));
        return UBS_ENOSPC;
    }
    const unsigned gx = (unsigned)ceil_div(N, kFusedThreads);
    // enough CTAs to fill the machine: split the camera loop over grid.y only when there are few primitives
    int sm = ubs_device_sm_count();
    if (sm <= 0) sm = 148;
    unsigned gy = 1;
    while ((int64_t)gx * gy < (int64_t)sm * 8 && (int)gy < C) gy *= 2;
    if ((int)gy > C) gy = (unsigned)C;
    dim3 grid(gx, gy);
#define UBS_FUSED_LAUNCH(DD)                                                                                           \
    fused_project_fwd_kernel<DD><<<grid, kFusedThreads, 0, s>>>(                                                       \
        C, N, records, viewmats, Ks, cam_pos, timestamps, prim_mask, (uint32_t)width, (uint32_t)height, eps2d,         \
        near_plane, far_plane, radius_clip, calc_compensations, (uint32_t)tile_size, (uint32_t)tile_width,             \
        (uint32_t)tile_height, radii, means2d, depths, conics, opacities, betas, colors, tiles_per_gauss,             \
        (float4 *)splats, tile_delta, activated, query)
    if (D == 6) UBS_FUSED_LAUNCH(6);
    else UBS_FUSED_LAUNCH(7);
#undef UBS_FUSED_LAUNCH
    UBS_LAUNCH_CHECK("fused_project_fwd_kernel");
    if (tile_delta != nullptr) return UBS_OK;  // the pair count comes out of ubs_isect_bin_sort's scan
    return isect_blocksums_from_counts(CN, tiles_per_gauss, workspace, n_isects, s);
}

static int fused_project_bwd_impl(int C, int64_t N, int D, const float *records, const float *viewmats,
                                     const float *Ks, const float *cam_pos, const float *timestamps, int width,
                                     int height, float eps2d, int calc_compensations, const int32_t *radii,
                                     const float *conics, const float *v_rows, int rows_form, float *v_records, float *v_viewmats, int activated, const float *query,
                                     const int32_t *skip_flag, const ubs::PackSegs &v_segs, void *stream) {
    using namespace ubs;
    UBS_CHECK_ARG(C >= 0 && N >= 0 && width > 0 && height > 0, "fused_project_bwd: bad sizes");
    UBS_CHECK_ARG(D == 6 || D == 7, "fused_project_bwd: D must be 6 or 7 (got %d)", D);
    if (N == 0) return UBS_OK;
    UBS_CHECK_ARG(records && viewmats && Ks && (cam_pos || query) && radii && conics &&
                      v_rows && (v_records || v_segs.ptr[0] || v_segs.ptr[1] || v_segs.ptr[2] || v_segs.ptr[3] ||
                                 v_segs.ptr[4] || v_segs.ptr[5] || v_segs.ptr[6]),
                  "fused_project_bwd: null pointer");
    UBS_CHECK_ARG(D != 7 || timestamps != nullptr || query != nullptr, "fused_project_bwd: D=7 needs timestamps");
    UBS_CHECK_ARG((((uintptr_t)records | (uintptr_t)v_records) & 15) == 0,
                  "fused_project_bwd: records / v_records must be 16-byte aligned");
    cudaStream_t s = (cudaStream_t)stream;
    const unsigned gx = (unsigned)ceil_div(N, kFusedThreads);
    // compiled for 3 resident CTAs per SM (168 registers, ~0.6 KB of spills): measured 0.32 ms against 0.38 ms for the
    // spill-free 255-register build at 3M primitives -- the kernel is latency bound, occupancy wins
    const size_t smem = (size_t)kFusedThreads * UBS_RECORD_STRIDE(D) * sizeof(float);
    const AdamParams unused{};
    const ScatterDst no_scatter{};
    const BwdOpts opts{activated, query, skip_flag, v_rows, rows_form, v_viewmats, v_segs};
    if (v_viewmats != nullptr) {
        UBS_CUDA_TRY(cudaMemsetAsync(v_viewmats, 0, sizeof(float) * 16 * C, s));
        if (D == 6)
            fused_project_bwd_kernel<6, 2, false, false, true><<<gx, kFusedThreads, smem, s>>>(
                C, N, records, viewmats, Ks, cam_pos, timestamps, (uint32_t)width, (uint32_t)height, eps2d,
                calc_compensations, radii, conics, v_records, nullptr, nullptr, unused, no_scatter, opts, PullSrc{});
        else
            fused_project_bwd_kernel<7, 2, false, false, true><<<gx, kFusedThreads, smem, s>>>(
                C, N, records, viewmats, Ks, cam_pos, timestamps, (uint32_t)width, (uint32_t)height, eps2d,
                calc_compensations, radii, conics, v_records, nullptr, nullptr, unused, no_scatter, opts, PullSrc{});
        UBS_LAUNCH_CHECK("fused_project_bwd_pose_kernel");
        return UBS_OK;
    }
    if (D == 6)
        fused_project_bwd_kernel<6, 3, false><<<gx, kFusedThreads, smem, s>>>(
            C, N, records, viewmats, Ks, cam_pos, timestamps, (uint32_t)width, (uint32_t)height, eps2d,
            calc_compensations, radii, conics, v_records,
            nullptr, nullptr, unused, no_scatter, opts, PullSrc{});
    else
        fused_project_bwd_kernel<7, 3, false><<<gx, kFusedThreads, smem, s>>>(
            C, N, records, viewmats, Ks, cam_pos, timestamps, (uint32_t)width, (uint32_t)height, eps2d,
            calc_compensations, radii, conics, v_records,
            nullptr, nullptr, unused, no_scatter, opts, PullSrc{});
    UBS_LAUNCH_CHECK("fused_project_bwd_kernel");
    return UBS_OK;
}

extern "C" int ubs_fused_project_bwd(int C, int64_t N, int D, const float *records, const float *viewmats,
                                     const float *Ks, const float *cam_pos, const float *timestamps, int width,
                                     int height, float eps2d, int calc_compensations, const int32_t *radii,
                                     const float *conics, const float *v_rows, int rows_form, float *v_records,
                                     float *v_viewmats, int activated, const float *query, const int32_t *skip_flag,
                                     void *stream) {
    using namespace ubs;
    UBS_CHECK_ARG(v_records != nullptr || N == 0, "fused_project_bwd: v_records is null");
    return fused_project_bwd_impl(C, N, D, records, viewmats, Ks, cam_pos, timestamps, width, height, eps2d,
                                  calc_compensations, radii, conics, v_rows, rows_form, v_records, v_viewmats, activated,
                                  query, skip_flag, PackSegs{}, stream);
}

extern "C" int ubs_fused_project_bwd_unpacked(int C, int64_t N, int D, const float *records, const float *viewmats,
                                              const float *Ks, const float *cam_pos, const float *timestamps, int width,
                                              int height, float eps2d, int calc_compensations, const int32_t *radii,
                                              const float *conics, const float *v_rows, int rows_form, float *v_mean,
                                              float *v_rgb, float *v_opacity, float *v_beta0, float *v_beta_c,
                                              float *v_scale, float *v_l_triangle, float *v_viewmats, int activated,
                                              const float *query, const int32_t *skip_flag, void *stream) {
    using namespace ubs;
    const PackSegs segs{{v_mean, v_rgb, v_opacity, v_beta0, v_beta_c, v_scale, v_l_triangle}};
    UBS_CHECK_ARG(v_mean || v_rgb || v_opacity || v_beta0 || v_beta_c || v_scale || v_l_triangle || N == 0,
                  "fused_project_bwd_unpacked: every destination is null");
    return fused_project_bwd_impl(C, N, D, records, viewmats, Ks, cam_pos, timestamps, width, height, eps2d,
                                  calc_compensations, radii, conics, v_rows, rows_form, nullptr, v_viewmats, activated,
                                  query, skip_flag, segs, stream);
}

extern "C" int ubs_fused_project_bwd_adam(int C, int64_t N, int D, float *records, const float *viewmats,
                                          const float *Ks, const float *cam_pos, const float *timestamps, int width,
                                          int height, float eps2d, int calc_compensations, const int32_t *radii,
                                          const float *conics, const float *v_rows, int rows_form, float *exp_avg, float *exp_avg_sq, const double *h_lr,
                                          double beta1, double beta2, double eps, int64_t step, double opacity_reg,
                                          double scale_reg, const int32_t *skip_flag, void *stream) {
    using namespace ubs;
    UBS_CHECK_ARG(C >= 0 && N >= 0 && width > 0 && height > 0, "fused_project_bwd_adam: bad sizes");
    UBS_CHECK_ARG(D == 6 || D == 7, "fused_project_bwd_adam: D must be 6 or 7 (got %d)", D);
    if (N == 0) return UBS_OK;
    UBS_CHECK_ARG(records && viewmats && Ks && cam_pos && radii && conics &&
                      v_rows && exp_avg && exp_avg_sq && h_lr,
                  "fused_project_bwd_adam: null pointer");
    UBS_CHECK_ARG(D != 7 || timestamps != nullptr, "fused_project_bwd_adam: D=7 needs timestamps");
    UBS_CHECK_ARG(step >= 1, "fused_project_bwd_adam: step counts from 1 (got %lld)", (long long)step);
    UBS_CHECK_ARG((((uintptr_t)records | (uintptr_t)exp_avg | (uintptr_t)exp_avg_sq) & 15) == 0,
                  "fused_project_bwd_adam: records / moments must be 16-byte aligned");
    cudaStream_t s = (cudaStream_t)stream;
    const unsigned gx = (unsigned)ceil_div(N, kFusedThreads);
    const size_t smem = (size_t)3 * kFusedThreads * UBS_RECORD_STRIDE(D) * sizeof(float);
    const AdamParams a = make_adam_params(N, D, h_lr, beta1, beta2, eps, step, opacity_reg, scale_reg);
    const BwdOpts opts{0, nullptr, skip_flag, v_rows, rows_form, nullptr, PackSegs{}};
    if (D == 6) {
        UBS_CUDA_TRY(cudaFuncSetAttribute(fused_project_bwd_kernel<6, 3, true>,
                                          cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        UBS_CUDA_TRY(cudaFuncSetAttribute(fused_project_bwd_kernel<6, 3, true>,
                                          cudaFuncAttributePreferredSharedMemoryCarveout, 100));
        fused_project_bwd_kernel<6, 3, true><<<gx, kFusedThreads, smem, s>>>(
            C, N, records, viewmats, Ks, cam_pos, timestamps, (uint32_t)width, (uint32_t)height, eps2d,
            calc_compensations, radii, conics, nullptr,
            exp_avg, exp_avg_sq, a, ScatterDst{}, opts, PullSrc{});
    } else {
        UBS_CUDA_TRY(cudaFuncSetAttribute(fused_project_bwd_kernel<7, 3, true>,
                                          cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        UBS_CUDA_TRY(cudaFuncSetAttribute(fused_project_bwd_kernel<7, 3, true>,
                                          cudaFuncAttributePreferredSharedMemoryCarveout, 100));
        fused_project_bwd_kernel<7, 3, true><<<gx, kFusedThreads, smem, s>>>(
            C, N, records, viewmats, Ks, cam_pos, timestamps, (uint32_t)width, (uint32_t)height, eps2d,
            calc_compensations, radii, conics, nullptr,
            exp_avg, exp_avg_sq, a, ScatterDst{}, opts, PullSrc{});
    }
    UBS_LAUNCH_CHECK("fused_project_bwd_adam_kernel");
    return UBS_OK;
}

extern "C" int ubs_fused_project_bwd_scatter(int64_t N, int D, const float *records, const float *viewmats,
                                             const float *Ks, const float *cam_pos, const float *timestamps,
                                             int width, int height, float eps2d, int calc_compensations,
                                             const int32_t *radii, const float *conics, const float *v_rows, int rows_form, int world, int rank,
                                             int64_t shard_rows, float *const *h_staging, const int32_t *skip_flag,
                                             void *stream) {
    using namespace ubs;
    UBS_CHECK_ARG(N >= 0 && width > 0 && height > 0, "fused_project_bwd_scatter: bad sizes");
    UBS_CHECK_ARG(D == 6 || D == 7, "fused_project_bwd_scatter: D must be 6 or 7 (got %d)", D);
    UBS_CHECK_ARG(world >= 1 && world <= UBS_MAX_RANKS && rank >= 0 && rank < world,
                  "fused_project_bwd_scatter: rank %d of %d (at most %d ranks)", rank, world, UBS_MAX_RANKS);
    UBS_CHECK_ARG(shard_rows > 0 && shard_rows % kFusedThreads == 0 && shard_rows * world >= N,
                  "fused_project_bwd_scatter: shard_rows must be a positive multiple of %d covering N", kFusedThreads);
    if (N == 0) return UBS_OK;
    UBS_CHECK_ARG(records && viewmats && Ks && cam_pos && radii && conics &&
                      v_rows && h_staging,
                  "fused_project_bwd_scatter: null pointer");
    UBS_CHECK_ARG(D != 7 || timestamps != nullptr, "fused_project_bwd_scatter: D=7 needs timestamps");
    ScatterDst sc{};
    for (int g = 0; g < world; ++g) {
        UBS_CHECK_ARG(h_staging[g] != nullptr && ((uintptr_t)h_staging[g] & 15) == 0,
                      "fused_project_bwd_scatter: staging[%d] null or not 16-byte aligned", g);
        sc.staging[g] = h_staging[g];
    }
    sc.shard_rows = shard_rows;
    sc.rank = rank;
    cudaStream_t s = (cudaStream_t)stream;
    const unsigned gx = (unsigned)ceil_div(N, kFusedThreads);
    const size_t smem = (size_t)kFusedThreads * UBS_RECORD_STRIDE(D) * sizeof(float);
    const AdamParams unused{};
    const BwdOpts opts{0, nullptr, skip_flag, v_rows, rows_form, nullptr, PackSegs{}};
    if (D == 6)
        fused_project_bwd_kernel<6, 3, false><<<gx, kFusedThreads, smem, s>>>(
            1, N, records, viewmats, Ks, cam_pos, timestamps, (uint32_t)width, (uint32_t)height, eps2d,
            calc_compensations, radii, conics, nullptr,
            nullptr, nullptr, unused, sc, opts, PullSrc{});
    else
        fused_project_bwd_kernel<7, 3, false><<<gx, kFusedThreads, smem, s>>>(
            1, N, records, viewmats, Ks, cam_pos, timestamps, (uint32_t)width, (uint32_t)height, eps2d,
            calc_compensations, radii, conics, nullptr,
            nullptr, nullptr, unused, sc, opts, PullSrc{});
    UBS_LAUNCH_CHECK("fused_project_bwd_scatter_kernel");
    return UBS_OK;
}

extern "C" int ubs_fused_project_bwd_adam_pull(int64_t N, int D, int world, int rank, int64_t shard_rows,
                                               float *const *h_peer_records, const float *const *h_peer_rows,
                                               const float *viewmats, const float *Ks, const float *cam_pos,
                                               const float *timestamps, int width, int height, float eps2d,
                                               int calc_compensations, float *exp_avg_shard, float *exp_avg_sq_shard,
                                               const double *h_lr, double beta1, double beta2, double eps, int64_t step,
                                               double opacity_reg, double scale_reg, void *stream) {
    using namespace ubs;
    UBS_CHECK_ARG(N >= 0 && width > 0 && height > 0, "fused_project_bwd_adam_pull: bad sizes");
    UBS_CHECK_ARG(D == 6 || D == 7, "fused_project_bwd_adam_pull: D must be 6 or 7 (got %d)", D);
    UBS_CHECK_ARG(world >= 1 && world <= UBS_MAX_RANKS && rank >= 0 && rank < world,
                  "fused_project_bwd_adam_pull: rank %d of %d (at most %d ranks)", rank, world, UBS_MAX_RANKS);
    UBS_CHECK_ARG(shard_rows > 0 && shard_rows % kFusedThreads == 0 && shard_rows * world >= N,
                  "fused_project_bwd_adam_pull: shard_rows must be a positive multiple of %d covering N", kFusedThreads);
    UBS_CHECK_ARG(step >= 1, "fused_project_bwd_adam_pull: step counts from 1 (got %lld)", (long long)step);
    const int64_t row0 = (int64_t)rank * shard_rows;
    const int64_t count = std::max<int64_t>(0, std::min<int64_t>(shard_rows, N - row0));
    if (count == 0) return UBS_OK;
    UBS_CHECK_ARG(h_peer_records && h_peer_rows && viewmats && Ks && cam_pos && exp_avg_shard && exp_avg_sq_shard && h_lr,
                  "fused_project_bwd_adam_pull: null pointer");
    UBS_CHECK_ARG(D != 7 || timestamps != nullptr, "fused_project_bwd_adam_pull: D=7 needs timestamps");
    PullSrc pull{};
    for (int g = 0; g < world; ++g) {
        UBS_CHECK_ARG(h_peer_records[g] && h_peer_rows[g] &&
                          ((((uintptr_t)h_peer_records[g]) | ((uintptr_t)h_peer_rows[g])) & 15) == 0,
                      "fused_project_bwd_adam_pull: records / rows of rank %d null or not 16-byte aligned", g);
        pull.records[g] = h_peer_records[g];
        pull.rows[g] = h_peer_rows[g];
    }
    UBS_CHECK_ARG((((uintptr_t)exp_avg_shard | (uintptr_t)exp_avg_sq_shard) & 15) == 0,
                  "fused_project_bwd_adam_pull: moments must be 16-byte aligned");
    pull.world = world;
    pull.row0 = row0;
    cudaStream_t s = (cudaStream_t)stream;
    const unsigned gx = (unsigned)ceil_div(count, kFusedThreads);
    const int stride = UBS_RECORD_STRIDE(D);
    const size_t smem = ((size_t)3 * kFusedThreads * stride + (size_t)world * kFusedThreads * 12) * sizeof(float);
    const AdamParams a = make_adam_params(N, D, h_lr, beta1, beta2, eps, step, opacity_reg, scale_reg);
    const BwdOpts opts{0, nullptr, nullptr, nullptr, 1, nullptr, PackSegs{}};
    const float *records = pull.records[rank] + row0 * stride;
#define UBS_PULL_LAUNCH(DD)                                                                                            \
    do {                                                                                                               \
        UBS_CUDA_TRY(cudaFuncSetAttribute(fused_project_bwd_kernel<DD, 2, true, true>,                                 \
                                          cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));                    \
        UBS_CUDA_TRY(cudaFuncSetAttribute(fused_project_bwd_kernel<DD, 2, true, true>,                                 \
                                          cudaFuncAttributePreferredSharedMemoryCarveout, 100));                       \
        fused_project_bwd_kernel<DD, 2, true, true><<<gx, kFusedThreads, smem, s>>>(                                   \
            world, count, records, viewmats, Ks, cam_pos, timestamps, (uint32_t)width, (uint32_t)height, eps2d,        \
            calc_compensations, nullptr, nullptr, nullptr, exp_avg_shard, exp_avg_sq_shard, a, ScatterDst{}, opts,     \
            pull);                                                                                                     \
    } while (0)
    if (D == 6) UBS_PULL_LAUNCH(6);
    else UBS_PULL_LAUNCH(7);
#undef UBS_PULL_LAUNCH
    UBS_LAUNCH_CHECK("fused_project_bwd_adam_pull_kernel");
    return UBS_OK;
}
