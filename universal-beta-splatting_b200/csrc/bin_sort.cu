// K7-K9, B200-first: tile binning + per-tile segment sort.  Produces exactly what the reference's
// isect_tiles(sort=True) + isect_offset_encode produce (isect_tiles.cu:99-333) -- the pair list ordered by
// (camera, tile, depth bits) with ties in primitive order, and the CSR tile offsets -- without ever running a
// global sort:
//
//   1. every visible primitive adds four corner deltas of its tile rectangle to a (th+1) x (tw+1) grid per
//      camera (done inside the fused projection kernel, or by bin_count_kernel for the stand-alone operator);
//      a 2-D prefix sum of that grid is the number of pairs per tile, whose exclusive scan IS the reference's
//      `offsets` tensor -- it exists before any pair has been written;
//   2. bin_emit_kernel writes each pair straight into its tile's segment (slot from an atomic per-tile cursor)
//      as one 64-bit word  depth_bits << 32 | flatten_id;
//   3. segment_sort_kernel sorts each tile's segment in shared memory: stable 8-bit LSD radix passes over only the
//      depth bits that vary inside the tile, then primitives with bit-identical depth are ordered by flatten id.
//      (depth, flatten_id) is a total order and is the order the reference's stable sort yields, because it
//      emits pairs in ascending flatten id (isect_tiles.cu:82-95); the result is therefore bit-exact although the
//      arrival order inside a segment is not deterministic.
//
// HBM traffic per pair: 8 B written + 8 B read + 12 B written (the 50 MB segment buffer of a 3M-primitive frame
// stays in the 126 MB L2) against 8 + 24 * 6 = 152 B for the six onesweep passes over 46-bit keys it replaces.
// Segments longer than kSegMax are sorted by segment_sort_big_kernel in global memory (any length).
//
// Precondition: depths of visible primitives are >= +0 (near_plane > 0), so the reference's sign extension of the
// depth bits (isect_tiles.cu:91) is the identity; callers with near_plane <= 0 use ubs_isect_emit_sort.
#include "common.cuh"
#include "isect.cuh"

namespace ubs {
namespace {

constexpr int kSegThreads = 256;
constexpr int kSegWarps = kSegThreads / 32;
constexpr int kSegItems = 8;                       // elements per thread in the shared-memory sort
constexpr int kSegMax = kSegThreads * kSegItems;   // longest segment sorted in shared memory
constexpr int kScanThreads = 1024;

struct BinWorkspace {
    int32_t *delta;      // [C][(th+1)*(tw+1)] corner deltas (zeroed per frame)
    int32_t *cursor;     // [C*n_tiles] next free slot of every tile segment
    int64_t *cam_total;  // [C] pairs per camera
    uint64_t *keyval;    // [capacity] depth_bits << 32 | flatten_id, grouped by tile
    uint64_t *alt;       // [capacity] ping-pong buffer of the long-segment sort
};

size_t delta_bytes(int C, int tw, int th) { return align_up(sizeof(int32_t) * (size_t)C * (th + 1) * (tw + 1), 256); }

size_t bin_ws_bytes(int C, int tw, int th, int64_t capacity) {
    size_t b = delta_bytes(C, tw, th);
    b += align_up(sizeof(int32_t) * (size_t)C * tw * th, 256);
    b += align_up(sizeof(int64_t) * (size_t)(C > 0 ? C : 1), 256);
    b += 2 * align_up(sizeof(uint64_t) * (size_t)capacity, 256);
    return b;
}

BinWorkspace bin_carve(void *base, int C, int tw, int th, int64_t capacity) {
    BinWorkspace w;
    unsigned char *p = (unsigned char *)base;
    w.delta = (int32_t *)p;
    p += delta_bytes(C, tw, th);
    w.cursor = (int32_t *)p;
    p += align_up(sizeof(int32_t) * (size_t)C * tw * th, 256);
    w.cam_total = (int64_t *)p;
    p += align_up(sizeof(int64_t) * (size_t)(C > 0 ? C : 1), 256);
    w.keyval = (uint64_t *)p;
    p += align_up(sizeof(uint64_t) * (size_t)capacity, 256);
    w.alt = (uint64_t *)p;
    return w;
}

// ---- corner deltas for the stand-alone operator (the fused projection kernel does this itself) ----------------
__global__ void __launch_bounds__(kIsectThreads)
bin_count_kernel(int64_t CN, int64_t N, const float *__restrict__ means2d, const int32_t *__restrict__ radii,
                 uint32_t tile_size, uint32_t tile_width, uint32_t tile_height, int32_t *__restrict__ tiles_per_gauss,
                 int32_t *__restrict__ delta) {
    const int64_t idx = (int64_t)blockIdx.x * kIsectThreads + threadIdx.x;
    if (idx >= CN) return;
    int32_t cnt = 0;
    const int32_t r = radii[idx];
    if (r > 0) {
        const float2 m = reinterpret_cast<const float2 *>(means2d)[idx];
        const TileRect t = tile_rect(m.x, m.y, r, tile_size, tile_width, tile_height);
        cnt = (int32_t)((t.y1 - t.y0) * (t.x1 - t.x0));
        if (cnt > 0) add_tile_deltas(delta + (idx / N) * (int64_t)((tile_height + 1) * (tile_width + 1)), t, tile_width);
    }
    if (tiles_per_gauss != nullptr) tiles_per_gauss[idx] = cnt;
}

// ---- 2-D prefix sum of the deltas -> pairs per tile; exclusive scan -> offsets -------------------------------
// Block-wide inclusive scan of one int64 per thread (kScanThreads threads); returns the inclusive value and the
// block total.
__device__ __forceinline__ int64_t block_inclusive_scan_i64(int64_t v, int64_t *s_warp, int64_t *total) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int64_t incl = v;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
        const int64_t t = __shfl_up_sync(0xffffffffu, incl, off);
        if (lane >= off) incl += t;
    }
    __syncthreads();  // s_warp may still be read from a previous call
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    int64_t base = 0, tot = 0;
#pragma unroll
    for (int w = 0; w < kScanThreads / 32; ++w) {
        const int64_t t = s_warp[w];
        if (w < warp) base += t;
        tot += t;
    }
    *total = tot;
    return base + incl;
}

// Phase 1 (one CTA per camera): delta grid -> per-tile counts (written into `offsets`), camera total.
__device__ void bin_counts_of_camera(int32_t *delta, uint32_t tw, uint32_t th, int32_t *counts, int64_t *cam_total,
                                     int64_t *s_warp) {
    const uint32_t gw = tw + 1;
    // rows: inclusive scan along x, one warp per row
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (uint32_t y = warp; y < th; y += kScanThreads / 32) {
        int32_t carry = 0;
        for (uint32_t x0 = 0; x0 < tw; x0 += 32) {
            const uint32_t x = x0 + lane;
            int32_t v = x < tw ? delta[y * gw + x] : 0;
#pragma unroll
            for (int off = 1; off < 32; off <<= 1) {
                const int32_t t = __shfl_up_sync(0xffffffffu, v, off);
                if (lane >= off) v += t;
            }
            v += carry;
            if (x < tw) delta[y * gw + x] = v;
            carry = __shfl_sync(0xffffffffu, v, 31);
        }
    }
    __syncthreads();
    // columns: running sum down y, one thread per column; the result is the number of pairs of tile (y, x)
    int64_t mine = 0;
    for (uint32_t x = threadIdx.x; x < tw; x += kScanThreads) {
        int32_t acc = 0;
        for (uint32_t y0 = 0; y0 < th; y0 += 8) {
            int32_t d[8];  // independent loads first: the running sum must not serialise eight L2 round trips
#pragma unroll
            for (int k = 0; k < 8; ++k) d[k] = y0 + k < th ? delta[(y0 + k) * gw + x] : 0;
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                if (y0 + k < th) {
                    acc += d[k];
                    counts[(y0 + k) * tw + x] = acc;
                    mine += acc;
                }
            }
        }
    }
    int64_t total;
    block_inclusive_scan_i64(mine, s_warp, &total);
    if (threadIdx.x == 0) *cam_total = total;
    __syncthreads();
}

// Phase 2 (one CTA per camera): counts -> exclusive offsets (clamped to the capacity) and cursors.
__device__ void bin_offsets_of_camera(uint32_t cam, uint32_t C, uint32_t n_tiles, const int64_t *cam_total,
                                      int64_t capacity, int32_t *offsets /* counts in, offsets out */,
                                      int32_t *cursor, int64_t *n_isects, int32_t *status, int64_t *s_warp) {
    int64_t mine = 0, base = 0, tot;
    for (uint32_t c = threadIdx.x; c < cam; c += kScanThreads) mine += cam_total[c];
    block_inclusive_scan_i64(mine, s_warp, &base);
    for (uint32_t t0 = 0; t0 < n_tiles; t0 += kScanThreads) {
        const uint32_t t = t0 + threadIdx.x;
        const int64_t cnt = t < n_tiles ? (int64_t)offsets[t] : 0;
        const int64_t incl = block_inclusive_scan_i64(cnt, s_warp, &tot);
        if (t < n_tiles) {
            const int64_t excl = base + incl - cnt;
            const int32_t o = (int32_t)(excl < capacity ? excl : capacity);
            offsets[t] = o;
            cursor[t] = o;
        }
        base += tot;
    }
    if (cam == C - 1 && threadIdx.x == 0) {
        *n_isects = base;
        if (base > capacity && status != nullptr) atomicOr(status, 1);
    }
}

__global__ void __launch_bounds__(kScanThreads)
bin_scan_kernel(uint32_t C, uint32_t tw, uint32_t th, int64_t capacity, int32_t *__restrict__ delta,
                int32_t *__restrict__ offsets, int32_t *__restrict__ cursor, int64_t *__restrict__ cam_total,
                int64_t *__restrict__ n_isects, int32_t *__restrict__ status, int phase) {
    __shared__ int64_t s_warp[kScanThreads / 32];
    const uint32_t cam = blockIdx.x, n_tiles = tw * th;
    if (phase != 2) bin_counts_of_camera(delta + (size_t)cam * (th + 1) * (tw + 1), tw, th, offsets + (size_t)cam * n_tiles,
                                         cam_total + cam, s_warp);
    // a single camera needs no second launch: its total is already visible to this CTA
    if (phase == 2 || C == 1)
        bin_offsets_of_camera(cam, C, n_tiles, cam_total, capacity, offsets + (size_t)cam * n_tiles,
                              cursor + (size_t)cam * n_tiles, n_isects, status, s_warp);
}

// ---- emit: every pair goes straight into its tile's segment --------------------------------------------------
__global__ void __launch_bounds__(kIsectThreads)
bin_emit_kernel(int64_t CN, int64_t N, const float *__restrict__ means2d, const int32_t *__restrict__ radii,
                const float *__restrict__ depths, uint32_t tile_size, uint32_t tile_width, uint32_t tile_height,
                int64_t capacity, int32_t *__restrict__ cursor, uint64_t *__restrict__ keyval) {
    const int64_t idx = (int64_t)blockIdx.x * kIsectThreads + threadIdx.x;
    if (idx >= CN) return;
    const int32_t r = radii[idx];
    if (r <= 0) return;
    const float2 m = reinterpret_cast<const float2 *>(means2d)[idx];
    const TileRect t = tile_rect(m.x, m.y, r, tile_size, tile_width, tile_height);
    if (t.x1 <= t.x0 || t.y1 <= t.y0) return;
    const uint64_t kv = ((uint64_t)__float_as_uint(depths[idx]) << 32) | (uint64_t)(uint32_t)idx;
    int32_t *cur = cursor + (idx / N) * (int64_t)(tile_width * tile_height);
    for (uint32_t i = t.y0; i < t.y1; ++i) {
        for (uint32_t j = t.x0; j < t.x1; ++j) {
            const int32_t pos = atomicAdd(cur + i * tile_width + j, 1);
            if ((int64_t)(uint32_t)pos < capacity) keyval[(uint32_t)pos] = kv;
        }
    }
}

// ---- per-tile segment sort -----------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t block_exclusive_scan_u32_256(uint32_t v, uint32_t *tmp) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t incl = v;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, incl, off);
        if (lane >= off) incl += t;
    }
    if (lane == 31) tmp[warp] = incl;
    __syncthreads();
    uint32_t base = 0;
#pragma unroll
    for (int w = 0; w < kSegWarps; ++w)
        if (w < warp) base += tmp[w];
    __syncthreads();
    return base + incl - v;
}

struct SegRange {
    int32_t start, n;
    uint64_t hi;  // (camera << tile_n_bits | tile) << 32
};

__device__ __forceinline__ SegRange segment_of(uint32_t slot, uint32_t n_slots, uint32_t n_tiles, uint32_t tile_n_bits,
                                               const int32_t *offsets, int64_t n_total) {
    SegRange s;
    s.start = offsets[slot];
    int64_t end = slot + 1 < n_slots ? (int64_t)offsets[slot + 1] : n_total;
    if (end > n_total) end = n_total;
    s.n = (int32_t)(end - s.start);
    const uint64_t cam = slot / n_tiles, tile = slot - (uint32_t)cam * n_tiles;
    s.hi = ((cam << tile_n_bits) | tile) << 32;
    return s;
}

// Stable in-warp ranking of one digit per lane against the warp's running digit counters (shared memory):
// returns the number of earlier elements of this warp with the same digit.  The peer set comes from eight ballots
// (22 cycles per warp and SM on B200; MATCH.ANY measures 60).
__device__ __forceinline__ uint32_t warp_rank_digit(uint32_t d, bool valid, uint32_t *warp_cnt, uint32_t lane) {
    uint32_t peers = __ballot_sync(0xffffffffu, valid);
#pragma unroll
    for (int b = 0; b < 8; ++b) {
        const bool bit = (d >> b) & 1u;
        const uint32_t m = __ballot_sync(0xffffffffu, bit);
        peers &= bit ? m : ~m;
    }
    const int leader = __ffs(peers) - 1;
    uint32_t prev = 0;
    if ((int)lane == leader && valid) {
        prev = warp_cnt[d];
        warp_cnt[d] = prev + __popc(peers);
    }
    prev = __shfl_sync(0xffffffffu, prev, leader);
    __syncwarp();
    return prev + __popc(peers & ((1u << lane) - 1u));
}

__global__ void __launch_bounds__(kSegThreads)
segment_sort_kernel(uint32_t n_slots, uint32_t n_tiles, uint32_t tile_n_bits, const int32_t *__restrict__ offsets,
                    const int64_t *__restrict__ n_isects_dev, int64_t capacity, const uint64_t *__restrict__ keyval,
                    int64_t *__restrict__ isect_ids, int32_t *__restrict__ flatten_ids) {
    __shared__ uint32_t s_key[2][kSegMax];
    __shared__ uint32_t s_id[2][kSegMax];
    __shared__ uint32_t s_cnt[kSegWarps * 256];
    __shared__ uint32_t s_base[256];
    __shared__ uint32_t s_tmp[kSegWarps];
    __shared__ uint32_t s_min[kSegWarps], s_max[kSegWarps];

    const int64_t n_total = min(*n_isects_dev, capacity);
    const SegRange seg = segment_of(blockIdx.x, n_slots, n_tiles, tile_n_bits, offsets, n_total);
    const int32_t n = seg.n;
    if (n <= 0 || n > kSegMax) return;  // long segments: segment_sort_big_kernel
    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint64_t *src = keyval + seg.start;

    // ---- load, depth range of the segment ---------------------------------------------------------------------
    uint32_t kmin = 0xffffffffu, kmax = 0u;
    for (int32_t i = (int32_t)tid; i < n; i += kSegThreads) {
        const uint64_t kv = src[i];
        const uint32_t k = (uint32_t)(kv >> 32);
        s_key[0][i] = k;
        s_id[0][i] = (uint32_t)kv;
        kmin = min(kmin, k);
        kmax = max(kmax, k);
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        kmin = min(kmin, __shfl_xor_sync(0xffffffffu, kmin, off));
        kmax = max(kmax, __shfl_xor_sync(0xffffffffu, kmax, off));
    }
    if (lane == 0) s_min[warp] = kmin, s_max[warp] = kmax;
    __syncthreads();
#pragma unroll
    for (int w = 0; w < kSegWarps; ++w) {
        kmin = min(kmin, s_min[w]);
        kmax = max(kmax, s_max[w]);
    }
    const uint32_t lo = kmin;
    const int nbits = 32 - __clz(kmax - kmin);  // 0 when every depth is identical
    const int passes = (nbits + 7) >> 3;

    // each warp owns a contiguous chunk (a multiple of 32 elements); order inside = (item, lane)
    const int32_t chunk = ((n + kSegThreads - 1) / kSegThreads) * 32;
    const int32_t wbase = (int32_t)warp * chunk;
    uint32_t *my_cnt = s_cnt + warp * 256;
    int cur = 0;

    for (int p = 0; p < passes; ++p) {
        const int shift = 8 * p;
#pragma unroll
        for (int k = 0; k < kSegWarps; ++k) s_cnt[k * 256 + tid] = 0;
        __syncthreads();
        uint32_t key[kSegItems], rank[kSegItems];
#pragma unroll
        for (int it = 0; it < kSegItems; ++it) {
            if (it * 32 < chunk) {
                const int32_t e = wbase + it * 32 + (int32_t)lane;
                const bool valid = e < n;
                key[it] = valid ? s_key[cur][e] : 0u;
                const uint32_t d = valid ? (((key[it] - lo) >> shift) & 255u) : 0u;
                rank[it] = warp_rank_digit(d, valid, my_cnt, lane);
            }
        }
        __syncthreads();
        // digit `tid`: counts of the warps -> exclusive offsets of the warps, segment-wide count
        uint32_t cnt_d = 0;
#pragma unroll
        for (int w = 0; w < kSegWarps; ++w) {
            const uint32_t c = s_cnt[w * 256 + tid];
            s_cnt[w * 256 + tid] = cnt_d;
            cnt_d += c;
        }
        s_base[tid] = block_exclusive_scan_u32_256(cnt_d, s_tmp);
        __syncthreads();
#pragma unroll
        for (int it = 0; it < kSegItems; ++it) {
            if (it * 32 < chunk) {
                const int32_t e = wbase + it * 32 + (int32_t)lane;
                if (e < n) {
                    const uint32_t d = ((key[it] - lo) >> shift) & 255u;
                    const uint32_t pos = s_base[d] + my_cnt[d] + rank[it];
                    s_key[cur ^ 1][pos] = key[it];
                    s_id[cur ^ 1][pos] = s_id[cur][e];
                }
            }
        }
        __syncthreads();
        cur ^= 1;
    }

    // ---- write out; bit-identical depths are ordered by flatten id (each element ranks itself in its run) ------
    const uint32_t *K = s_key[cur], *I = s_id[cur];
    int64_t *out_keys = isect_ids + seg.start;
    int32_t *out_vals = flatten_ids + seg.start;
    for (int32_t i = (int32_t)tid; i < n; i += kSegThreads) {
        const uint32_t k = K[i], id = I[i];
        int32_t a = i, less = 0;
        while (a > 0 && K[a - 1] == k) {
            --a;
            less += I[a] < id;
        }
        for (int32_t j = i + 1; j < n && K[j] == k; ++j) less += I[j] < id;
        out_keys[a + less] = (int64_t)(seg.hi | (uint64_t)k);
        out_vals[a + less] = (int32_t)id;
    }
}

// Long segments (n > kSegMax): LSD radix sort of the full 64-bit words (depth_bits << 32 | id: unique, so no tie
// pass) in global memory, ping-ponging between the segment's span of `keyval` and of `alt`.  One CTA per segment,
// found by striding over the slots; correctness for any length matters here, speed does not.
__global__ void __launch_bounds__(kSegThreads)
segment_sort_big_kernel(uint32_t n_slots, uint32_t n_tiles, uint32_t tile_n_bits, const int32_t *__restrict__ offsets,
                        const int64_t *__restrict__ n_isects_dev, int64_t capacity, uint64_t *keyval, uint64_t *alt,
                        int64_t *__restrict__ isect_ids, int32_t *__restrict__ flatten_ids) {
    __shared__ uint32_t s_cnt[kSegWarps * 256];
    __shared__ uint32_t s_base[256];
    __shared__ uint32_t s_tmp[kSegWarps];
    __shared__ unsigned long long s_or[kSegWarps];
    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int64_t n_total = min(*n_isects_dev, capacity);
    for (uint32_t slot = blockIdx.x; slot < n_slots; slot += gridDim.x) {
        const SegRange seg = segment_of(slot, n_slots, n_tiles, tile_n_bits, offsets, n_total);
        const int32_t n = seg.n;
        if (n <= kSegMax) continue;
        uint64_t *src = keyval + seg.start, *dst = alt + seg.start;
        // bits that differ anywhere in the segment
        const uint64_t first = src[0];
        unsigned long long diff = 0ull;
        for (int32_t i = (int32_t)tid; i < n; i += kSegThreads) diff |= src[i] ^ first;
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) diff |= __shfl_xor_sync(0xffffffffu, diff, off);
        __syncthreads();
        if (lane == 0) s_or[warp] = diff;
        __syncthreads();
#pragma unroll
        for (int w = 0; w < kSegWarps; ++w) diff |= s_or[w];

        const int32_t chunk = (int32_t)(((int64_t)n + kSegThreads - 1) / kSegThreads) * 32;
        const int32_t wbase = (int32_t)warp * chunk;
        uint32_t *my_cnt = s_cnt + warp * 256;
        for (int shift = 0; shift < 64; shift += 8) {
            if (((diff >> shift) & 0xffull) == 0ull) continue;  // this digit is constant: the pass is the identity
#pragma unroll
            for (int k = 0; k < kSegWarps; ++k) s_cnt[k * 256 + tid] = 0;
            __syncthreads();
            for (int32_t e0 = wbase; e0 < wbase + chunk && e0 < n; e0 += 32) {
                const int32_t e = e0 + (int32_t)lane;
                const bool valid = e < n;
                const uint32_t d = valid ? (uint32_t)((src[e] >> shift) & 0xffull) : 0u;
                (void)warp_rank_digit(d, valid, my_cnt, lane);
            }
            __syncthreads();
            uint32_t cnt_d = 0;
#pragma unroll
            for (int w = 0; w < kSegWarps; ++w) {
                const uint32_t c = s_cnt[w * 256 + tid];
                s_cnt[w * 256 + tid] = cnt_d;
                cnt_d += c;
            }
            s_base[tid] = block_exclusive_scan_u32_256(cnt_d, s_tmp);
            __syncthreads();
            // second walk: the counters now run from each warp's exclusive offset
            for (int32_t e0 = wbase; e0 < wbase + chunk && e0 < n; e0 += 32) {
                const int32_t e = e0 + (int32_t)lane;
                const bool valid = e < n;
                const uint64_t kv = valid ? src[e] : 0ull;
                const uint32_t d = valid ? (uint32_t)((kv >> shift) & 0xffull) : 0u;
                const uint32_t r = warp_rank_digit(d, valid, my_cnt, lane);
                if (valid) dst[s_base[d] + r] = kv;
            }
            __syncthreads();
            uint64_t *t = src;
            src = dst;
            dst = t;
        }
        for (int32_t i = (int32_t)tid; i < n; i += kSegThreads) {
            const uint64_t kv = src[i];
            isect_ids[seg.start + i] = (int64_t)(seg.hi | (kv >> 32));
            flatten_ids[seg.start + i] = (int32_t)(uint32_t)kv;
        }
        __syncthreads();
    }
}

}  // namespace

size_t bin_delta_bytes(int C, int tile_width, int tile_height) { return delta_bytes(C, tile_width, tile_height); }

}  // namespace ubs

extern "C" size_t ubs_isect_bin_workspace_bytes(int C, int tile_width, int tile_height, int64_t capacity) {
    using namespace ubs;
    if (C < 0 || tile_width <= 0 || tile_height <= 0) return 0;
    return bin_ws_bytes(C, tile_width, tile_height, capacity < 0 ? 0 : capacity);
}

extern "C" int ubs_isect_bin_sort(int C, int64_t N, const float *means2d, const int32_t *radii, const float *depths,
                                  int tile_size, int tile_width, int tile_height, int deltas_ready,
                                  int32_t *tiles_per_gauss, int64_t capacity, int64_t *n_isects, int64_t *isect_ids,
                                  int32_t *flatten_ids, int32_t *offsets, int32_t *status, void *workspace,
                                  size_t workspace_bytes, void *stream) {
    using namespace ubs;
    UBS_CHECK_ARG(C >= 0 && N >= 0 && tile_size > 0 && tile_width > 0 && tile_height > 0 && capacity >= 0,
                  "isect_bin_sort: bad sizes");
    UBS_CHECK_ARG(n_isects != nullptr, "isect_bin_sort: n_isects is null");
    cudaStream_t s = (cudaStream_t)stream;
    const int64_t CN = (int64_t)C * N;
    const uint32_t n_tiles = (uint32_t)tile_width * (uint32_t)tile_height;
    const int64_t n_slots64 = (int64_t)C * n_tiles;
    if (C == 0) {
        UBS_CUDA_TRY(cudaMemsetAsync(n_isects, 0, sizeof(int64_t), s));
        return UBS_OK;
    }
    UBS_CHECK_ARG(offsets != nullptr && workspace != nullptr, "isect_bin_sort: null offsets / workspace");
    UBS_CHECK_ARG(n_slots64 < ((int64_t)1 << 31) && CN < ((int64_t)1 << 31) && capacity < ((int64_t)1 << 31),
                  "isect_bin_sort: C*tiles, C*N and capacity must fit int32");
    const int tile_n_bits = id_bits(n_tiles), cam_n_bits = id_bits((uint32_t)C);
    UBS_CHECK_ARG(tile_n_bits + cam_n_bits <= 32, "isect_bin_sort: tile+camera ids need more than 32 bits");
    UBS_CHECK_ARG(((uintptr_t)workspace & 7) == 0, "isect_bin_sort: workspace must be 8-byte aligned");
    if (workspace_bytes < bin_ws_bytes(C, tile_width, tile_height, capacity)) {
        set_error("isect_bin_sort: workspace %zu < %zu", workspace_bytes,
                  bin_ws_bytes(C, tile_width, tile_height, capacity));
        return UBS_ENOSPC;
    }
    const BinWorkspace w = bin_carve(workspace, C, tile_width, tile_height, capacity);
    const uint32_t n_slots = (uint32_t)n_slots64;
    if (CN > 0) UBS_CHECK_ARG(means2d && radii && depths, "isect_bin_sort: null primitive arrays");
    if (capacity > 0) UBS_CHECK_ARG(isect_ids && flatten_ids, "isect_bin_sort: null pair arrays");

    if (!deltas_ready) {
        UBS_CUDA_TRY(cudaMemsetAsync(w.delta, 0, delta_bytes(C, tile_width, tile_height), s));
        if (CN > 0) {
            bin_count_kernel<<<(unsigned)ceil_div(CN, kIsectThreads), kIsectThreads, 0, s>>>(
                CN, N, means2d, radii, (uint32_t)tile_size, (uint32_t)tile_width, (uint32_t)tile_height,
                tiles_per_gauss, w.delta);
            UBS_LAUNCH_CHECK("bin_count_kernel");
        }
    }
    bin_scan_kernel<<<(unsigned)C, kScanThreads, 0, s>>>((uint32_t)C, (uint32_t)tile_width, (uint32_t)tile_height,
                                                         capacity, w.delta, offsets, w.cursor, w.cam_total, n_isects,
                                                         status, 1);
    UBS_LAUNCH_CHECK("bin_scan_kernel");
    if (C > 1) {
        bin_scan_kernel<<<(unsigned)C, kScanThreads, 0, s>>>((uint32_t)C, (uint32_t)tile_width, (uint32_t)tile_height,
                                                             capacity, w.delta, offsets, w.cursor, w.cam_total,
                                                             n_isects, status, 2);
        UBS_LAUNCH_CHECK("bin_scan_kernel");
    }
    if (CN == 0 || capacity == 0) return UBS_OK;
    bin_emit_kernel<<<(unsigned)ceil_div(CN, kIsectThreads), kIsectThreads, 0, s>>>(
        CN, N, means2d, radii, depths, (uint32_t)tile_size, (uint32_t)tile_width, (uint32_t)tile_height, capacity,
        w.cursor, w.keyval);
    UBS_LAUNCH_CHECK("bin_emit_kernel");
    segment_sort_kernel<<<n_slots, kSegThreads, 0, s>>>(n_slots, n_tiles, (uint32_t)tile_n_bits, offsets, n_isects,
                                                        capacity, w.keyval, isect_ids, flatten_ids);
    UBS_LAUNCH_CHECK("segment_sort_kernel");
    int sm = ubs_device_sm_count();
    if (sm <= 0) sm = 148;
    const unsigned big_grid = n_slots < (unsigned)(2 * sm) ? n_slots : (unsigned)(2 * sm);
    segment_sort_big_kernel<<<big_grid, kSegThreads, 0, s>>>(n_slots, n_tiles, (uint32_t)tile_n_bits, offsets, n_isects,
                                                             capacity, w.keyval, w.alt, isect_ids, flatten_ids);
    UBS_LAUNCH_CHECK("segment_sort_big_kernel");
    return UBS_OK;
}
