// K7-K9: tile intersection (count -> scan -> emit), onesweep sort of the (tile|depth) keys, tile offsets.
// Drop-in for the reference's isect_tiles / isect_offset_encode (isect_tiles.cu:16-333).  Integer results are
// bit-exact with the reference by construction: same tile-bound arithmetic, same key packing, stable sort.
#include "common.cuh"
#include "isect.cuh"
#include "radix_sort.cuh"

namespace ubs {

namespace {

// ---- phase 1: per-(camera, primitive) tile count + per-block sums -------------------------------------------
__global__ void __launch_bounds__(kIsectThreads)
isect_count_kernel(int64_t CN, const float *__restrict__ means2d, const int32_t *__restrict__ radii,
                   uint32_t tile_size, uint32_t tile_width, uint32_t tile_height,
                   int32_t *__restrict__ tiles_per_gauss, int64_t *__restrict__ block_sums) {
    const int64_t idx = (int64_t)blockIdx.x * kIsectThreads + threadIdx.x;
    int32_t cnt = 0;
    if (idx < CN) {
        const int32_t r = radii[idx];
        if (r > 0) {
            const float2 m = reinterpret_cast<const float2 *>(means2d)[idx];
            const TileRect t = tile_rect(m.x, m.y, r, tile_size, tile_width, tile_height);
            cnt = (int32_t)((t.y1 - t.y0) * (t.x1 - t.x0));
        }
        tiles_per_gauss[idx] = cnt;
    }
    const int64_t total = block_reduce_sum_i64((int64_t)cnt);
    if (threadIdx.x == 0) block_sums[blockIdx.x] = total;
}

// phase 1 variant for the fused projection kernel, which already wrote the per-primitive tile counts
__global__ void __launch_bounds__(kIsectThreads)
isect_blocksum_kernel(int64_t CN, const int32_t *__restrict__ tiles_per_gauss, int64_t *__restrict__ block_sums) {
    const int64_t idx = (int64_t)blockIdx.x * kIsectThreads + threadIdx.x;
    const int64_t total = block_reduce_sum_i64(idx < CN ? (int64_t)tiles_per_gauss[idx] : 0);
    if (threadIdx.x == 0) block_sums[blockIdx.x] = total;
}

// ---- phase 3: emit pairs; block-local exclusive scan of the counts + scanned block base ---------------------
__global__ void __launch_bounds__(kIsectThreads)
isect_emit_kernel(int64_t CN, int64_t N, const float *__restrict__ means2d, const int32_t *__restrict__ radii,
                  const float *__restrict__ depths, const int32_t *__restrict__ tiles_per_gauss,
                  const int64_t *__restrict__ block_offsets, uint32_t tile_size, uint32_t tile_width,
                  uint32_t tile_height, uint32_t tile_n_bits, int64_t capacity, int64_t *__restrict__ isect_ids,
                  int32_t *__restrict__ flatten_ids) {
    const int64_t idx = (int64_t)blockIdx.x * kIsectThreads + threadIdx.x;
    int32_t cnt = 0;
    if (idx < CN) cnt = tiles_per_gauss[idx];
    const int64_t local = block_exclusive_scan_i64((int64_t)cnt);
    if (cnt == 0) return;
    int64_t cur = block_offsets[blockIdx.x] + local;

    const float2 m = reinterpret_cast<const float2 *>(means2d)[idx];
    const TileRect t = tile_rect(m.x, m.y, radii[idx], tile_size, tile_width, tile_height);
    const int64_t cid = idx / N;
    const int64_t cid_enc = cid << (32 + tile_n_bits);
    const int64_t depth_enc = (int64_t)__float_as_int(depths[idx]);  // sign-extending, as the reference does
    for (uint32_t i = t.y0; i < t.y1; ++i) {
        for (uint32_t j = t.x0; j < t.x1; ++j) {
            if (cur < capacity) {
                const int64_t tile_id = (int64_t)(i * tile_width + j);
                isect_ids[cur] = cid_enc | (tile_id << 32) | depth_enc;
                flatten_ids[cur] = (int32_t)idx;
            }
            ++cur;
        }
    }
}

// ---- phase 2: scan of the block sums (single CTA; a few thousand entries) -----------------------------------
__global__ void __launch_bounds__(1024)
isect_scan_blocks_kernel(int64_t n_blocks, int64_t *__restrict__ block_sums, int64_t *__restrict__ n_isects) {
    __shared__ int64_t s_warp[32];
    __shared__ int64_t s_carry;
    if (threadIdx.x == 0) s_carry = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int64_t base = 0; base < n_blocks; base += 1024) {
        const int64_t i = base + threadIdx.x;
        const int64_t v = i < n_blocks ? block_sums[i] : 0;
        int64_t incl = v;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            const int64_t t = __shfl_up_sync(0xffffffffu, incl, off);
            if (lane >= off) incl += t;
        }
        if (lane == 31) s_warp[warp] = incl;
        __syncthreads();
        if (warp == 0) {
            int64_t w = s_warp[lane];
#pragma unroll
            for (int off = 1; off < 32; off <<= 1) {
                const int64_t t = __shfl_up_sync(0xffffffffu, w, off);
                if (lane >= off) w += t;
            }
            s_warp[lane] = w;  // inclusive over warps
        }
        __syncthreads();
        const int64_t carry = s_carry;
        const int64_t warp_base = warp > 0 ? s_warp[warp - 1] : 0;
        if (i < n_blocks) block_sums[i] = carry + warp_base + incl - v;  // exclusive
        __syncthreads();
        if (threadIdx.x == 1023) s_carry = carry + warp_base + incl;
        __syncthreads();
    }
    if (threadIdx.x == 0) *n_isects = s_carry;
}

// ---- offsets (isect_tiles.cu:287-333) ------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
isect_offsets_kernel(const int64_t *__restrict__ n_dev, int64_t n_host, int64_t capacity,
                     const int64_t *__restrict__ isect_ids,
                     uint32_t n_slots /* C * n_tiles */, uint32_t n_tiles, uint32_t tile_n_bits,
                     int32_t *__restrict__ offsets, int32_t *__restrict__ status) {
    int64_t n = n_dev != nullptr ? *n_dev : n_host;
    const int64_t tile_mask = ((int64_t)1 << tile_n_bits) - 1;
    if (blockIdx.x == 0 && threadIdx.x == 0) report_truncation(status, n > capacity);
    if (n > capacity) n = capacity;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const int64_t first = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (n == 0) {
        for (int64_t idx = first; idx < n_slots; idx += stride) offsets[idx] = 0;
        return;
    }
    // grid-stride over the n pairs actually present (the grid is bounded, not sized by the capacity)
    for (int64_t idx = first; idx < n; idx += stride) {
        const int64_t hi = isect_ids[idx] >> 32;
        const int64_t slot = (hi >> tile_n_bits) * n_tiles + (hi & tile_mask);
        if (idx == 0) {
            for (int64_t i = 0; i <= slot; ++i) offsets[i] = 0;
        } else {
            const int64_t hp = isect_ids[idx - 1] >> 32;
            if (hp != hi) {
                const int64_t slot_prev = (hp >> tile_n_bits) * n_tiles + (hp & tile_mask);
                for (int64_t i = slot_prev + 1; i <= slot; ++i) offsets[i] = (int32_t)idx;
            }
        }
        if (idx == n - 1) {
            for (int64_t i = slot + 1; i < n_slots; ++i) offsets[i] = (int32_t)n;
        }
    }
}

struct IsectWorkspace {
    int64_t *block_sums;
    int64_t n_blocks;
    void *sort_base;
};

size_t isect_ws_header_bytes(int64_t CN) { return align_up(sizeof(int64_t) * (size_t)(ceil_div(CN, kIsectThreads) + 1), 256); }

IsectWorkspace carve(void *base, int64_t CN) {
    IsectWorkspace w;
    w.block_sums = (int64_t *)base;
    w.n_blocks = ceil_div(CN, kIsectThreads);
    w.sort_base = (unsigned char *)base + isect_ws_header_bytes(CN);
    return w;
}

}  // namespace

int isect_blocksums_from_counts(int64_t CN, const int32_t *tiles_per_gauss, void *workspace, int64_t *n_isects,
                                cudaStream_t s) {
    IsectWorkspace w = carve(workspace, CN);
    isect_blocksum_kernel<<<(unsigned)w.n_blocks, kIsectThreads, 0, s>>>(CN, tiles_per_gauss, w.block_sums);
    UBS_LAUNCH_CHECK("isect_blocksum_kernel");
    return isect_scan_and_total(w.n_blocks, w.block_sums, n_isects, s);
}

int isect_scan_and_total(int64_t n_blocks, int64_t *block_sums, int64_t *n_isects, cudaStream_t s) {
    isect_scan_blocks_kernel<<<1, 1024, 0, s>>>(n_blocks, block_sums, n_isects);
    UBS_LAUNCH_CHECK("isect_scan_blocks_kernel");
    return UBS_OK;
}

}  // namespace ubs

extern "C" size_t ubs_isect_workspace_bytes(int64_t CN, int64_t capacity) {
    using namespace ubs;
    if (CN < 0) CN = 0;
    if (capacity < 0) capacity = 0;
    return isect_ws_header_bytes(CN) + sort_workspace_bytes(capacity);
}

extern "C" int ubs_isect_count(int C, int64_t N, const float *means2d, const int32_t *radii, int tile_size,
                               int tile_width, int tile_height, int32_t *tiles_per_gauss, int64_t *n_isects,
                               void *workspace, size_t workspace_bytes, void *stream) {
    using namespace ubs;
    UBS_CHECK_ARG(C >= 0 && N >= 0 && tile_size > 0 && tile_width > 0 && tile_height > 0, "isect_count: bad sizes");
    UBS_CHECK_ARG(n_isects != nullptr, "isect_count: n_isects is null");
    cudaStream_t s = (cudaStream_t)stream;
    const int64_t CN = (int64_t)C * N;
    if (CN == 0) {
        UBS_CUDA_TRY(cudaMemsetAsync(n_isects, 0, sizeof(int64_t), s));
        return UBS_OK;
    }
    UBS_CHECK_ARG(means2d && radii && tiles_per_gauss && workspace, "isect_count: null pointer");
    UBS_CHECK_ARG(CN < ((int64_t)1 << 31), "isect_count: C*N must fit int32 flatten ids");
    if (workspace_bytes < isect_ws_header_bytes(CN)) {
        set_error("isect_count: workspace %zu < %zu", workspace_bytes, isect_ws_header_bytes(CN));
        return UBS_ENOSPC;
    }
    IsectWorkspace w = carve(workspace, CN);
    isect_count_kernel<<<(unsigned)w.n_blocks, kIsectThreads, 0, s>>>(CN, means2d, radii, (uint32_t)tile_size,
                                                                      (uint32_t)tile_width, (uint32_t)tile_height,
                                                                      tiles_per_gauss, w.block_sums);
    UBS_LAUNCH_CHECK("isect_count_kernel");
    return isect_scan_and_total(w.n_blocks, w.block_sums, n_isects, s);
}

extern "C" int ubs_isect_emit_sort(int C, int64_t N, const float *means2d, const int32_t *radii, const float *depths,
                                   int tile_size, int tile_width, int tile_height, int do_sort,
                                   const int32_t *tiles_per_gauss, const int64_t *n_isects, int64_t capacity,
                                   int64_t *isect_ids, int32_t *flatten_ids, int32_t *offsets, int32_t *status,
                                   void *workspace, size_t workspace_bytes, void *stream) {
    using namespace ubs;
    UBS_CHECK_ARG(C >= 0 && N >= 0 && tile_size > 0 && tile_width > 0 && tile_height > 0 && capacity >= 0,
                  "isect_emit_sort: bad sizes");
    UBS_CHECK_ARG(n_isects != nullptr, "isect_emit_sort: n_isects is null");
    cudaStream_t s = (cudaStream_t)stream;
    const int64_t CN = (int64_t)C * N;
    const uint32_t n_tiles = (uint32_t)tile_width * (uint32_t)tile_height;
    const int tile_n_bits = id_bits(n_tiles), cam_n_bits = id_bits((uint32_t)(C > 0 ? C : 1));
    UBS_CHECK_ARG(tile_n_bits + cam_n_bits <= 32, "isect_emit_sort: tile+camera ids need more than 32 bits");
    const uint32_t n_slots = (uint32_t)C * n_tiles;

    if (CN > 0 && capacity > 0) {
        UBS_CHECK_ARG(means2d && radii && depths && tiles_per_gauss && isect_ids && flatten_ids && workspace,
                      "isect_emit_sort: null pointer");
        if (workspace_bytes < ubs_isect_workspace_bytes(CN, capacity)) {
            set_error("isect_emit_sort: workspace %zu < %zu", workspace_bytes, ubs_isect_workspace_bytes(CN, capacity));
            return UBS_ENOSPC;
        }
        IsectWorkspace w = carve(workspace, CN);
        const SortWorkspace sw = sort_workspace_carve(w.sort_base, capacity);
        const int end_bit = 32 + tile_n_bits + cam_n_bits;
        const int passes = do_sort ? sort_num_passes(0, end_bit) : 0;
        // even pass count: emit straight into the caller's arrays (the ping-pong ends where it started);
        // odd: emit into the workspace so the last pass lands in the caller's arrays.
        int64_t *emit_keys = (passes % 2 == 0) ? isect_ids : sw.alt_keys;
        int32_t *emit_vals = (passes % 2 == 0) ? flatten_ids : sw.alt_vals;
        isect_emit_kernel<<<(unsigned)w.n_blocks, kIsectThreads, 0, s>>>(
            CN, N, means2d, radii, depths, tiles_per_gauss, w.block_sums, (uint32_t)tile_size, (uint32_t)tile_width,
            (uint32_t)tile_height, (uint32_t)tile_n_bits, capacity, emit_keys, emit_vals);
        UBS_LAUNCH_CHECK("isect_emit_kernel");
        if (do_sort) {
            int64_t *other_keys = (passes % 2 == 0) ? sw.alt_keys : isect_ids;
            int32_t *other_vals = (passes % 2 == 0) ? sw.alt_vals : flatten_ids;
            const int rc = radix_sort_pairs_pingpong(n_isects, capacity, emit_keys, emit_vals, other_keys, other_vals,
                                                     0, end_bit, sw, s);
            if (rc < 0) return rc;
        }
    }
    if (offsets != nullptr && n_slots > 0) {
        const int64_t work = capacity > (int64_t)n_slots ? capacity : (int64_t)n_slots;
        int sm = ubs_device_sm_count();
        if (sm <= 0) sm = 148;
        const int64_t max_blocks = (int64_t)sm * 16;
        const int64_t blocks = ceil_div(work, 256) < max_blocks ? ceil_div(work, 256) : max_blocks;
        isect_offsets_kernel<<<(unsigned)blocks, 256, 0, s>>>(n_isects, 0, capacity, isect_ids, n_slots,
                                                                           n_tiles, (uint32_t)tile_n_bits, offsets,
                                                                           status);
        UBS_LAUNCH_CHECK("isect_offsets_kernel");
    }
    return UBS_OK;
}

extern "C" int ubs_isect_offset_encode(int64_t n_isects, const int64_t *isect_ids, int C, int tile_width,
                                       int tile_height, int32_t *offsets, void *stream) {
    using namespace ubs;
    UBS_CHECK_ARG(n_isects >= 0 && C >= 0 && tile_width > 0 && tile_height > 0, "isect_offset_encode: bad sizes");
    UBS_CHECK_ARG(offsets != nullptr || C == 0, "isect_offset_encode: null offsets");
    cudaStream_t s = (cudaStream_t)stream;
    const uint32_t n_tiles = (uint32_t)tile_width * (uint32_t)tile_height;
    const uint32_t n_slots = (uint32_t)C * n_tiles;
    if (n_slots == 0) return UBS_OK;
    if (n_isects == 0) {
        UBS_CUDA_TRY(cudaMemsetAsync(offsets, 0, sizeof(int32_t) * n_slots, s));
        return UBS_OK;
    }
    UBS_CHECK_ARG(isect_ids != nullptr, "isect_offset_encode: null isect_ids");
    const int64_t work = n_isects > (int64_t)n_slots ? n_isects : (int64_t)n_slots;
    isect_offsets_kernel<<<(unsigned)ceil_div(work, 256), 256, 0, s>>>(nullptr, n_isects, n_isects, isect_ids, n_slots,
                                                                       n_tiles, (uint32_t)id_bits(n_tiles), offsets,
                                                                       nullptr);
    UBS_LAUNCH_CHECK("isect_offsets_kernel");
    return UBS_OK;
}
