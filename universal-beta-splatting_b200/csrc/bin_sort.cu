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
// Segments longer than kSegMax (2048) are sorted by segment_sort_long_kernel: in shared memory up to 8192 pairs,
// in global memory beyond (any length).
//
// Precondition: depths of visible primitives are >= +0 (near_plane > 0), so the reference's sign extension of the
// depth bits (isect_tiles.cu:91) is the identity; callers with near_plane <= 0 use ubs_isect_emit_sort.
#include "common.cuh"
#include "isect.cuh"

namespace ubs {
namespace {

constexpr int kSegThreads = 256;
constexpr int kSegWarps = kSegThreads / 32;
constexpr int kSegItems = 8;                       // elements per thread in the shared-memory sort (short segments)
constexpr int kSegMax = kSegThreads * kSegItems;   // longest segment of the first shared-memory instantiation
constexpr int kSegItemsMid = 16, kSegItemsLong = 32;                 // 2049..4096 and 4097..8192 pairs per tile
constexpr int kSegMaxShared = kSegThreads * kSegItemsLong;          // beyond this: sort_segment_global
constexpr int kTopBits = 11;    // MSD shortcut of the segment sort: bucket = top 11 varying depth bits
constexpr int kBucketMax = 64;  // MSD shortcut of the segment sort: largest bucket ranked by comparisons
constexpr int kScanThreads = 1024;
// Atomic targets are spread to one per 32-byte sector: with 4-byte spacing the whole cursor array sits in a few
// L2 slices and one slice's atomic unit saturates (lts__d_atomic_input_cycles_active: max 65 %, mean 8.5 %).
constexpr int kCursorStride = 8;

struct BinWorkspace {
    int32_t *delta;      // [C][(th+1)*(tw+1)] corner deltas (zeroed per frame)
    int32_t *cursor;     // [C*n_tiles*kCursorStride] next free slot of every tile segment
    int64_t *cam_total;  // [C] pairs per camera
    uint64_t *keyval;    // [capacity] depth_bits << 32 | flatten_id, grouped by tile
    uint64_t *alt;       // [capacity] ping-pong buffer of the long-segment sort
};

size_t delta_bytes(int C, int tw, int th) {
    return align_up(sizeof(int32_t) * (size_t)C * (th + 1) * (tw + 1) * kDeltaStride, 256);
}

size_t bin_ws_bytes(int C, int tw, int th, int64_t capacity) {
    size_t b = delta_bytes(C, tw, th);
    b += align_up(sizeof(int32_t) * (size_t)C * tw * th * kCursorStride, 256);
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
    p += align_up(sizeof(int32_t) * (size_t)C * tw * th * kCursorStride, 256);
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
        if (cnt > 0) add_tile_deltas(delta + (idx / N) * (int64_t)((tile_height + 1) * (tile_width + 1)) * kDeltaStride, t,
                                     tile_width);
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

// Phase 1 (one CTA per camera): delta grid -> per-tile counts (written to `counts`), camera total.  `g` is the
// grid the 2-D prefix sum runs in: a compact copy in shared memory when it fits (every dependent step is then a
// 29-cycle shared access instead of an L2 round trip), else the strided global delta grid itself.
__device__ void bin_counts_of_camera(int32_t *g, uint32_t gw, uint32_t gs, uint32_t tw, uint32_t th, int32_t *counts,
                                     int64_t *cam_total, int64_t *s_warp) {
    // rows: inclusive scan along x, one warp per row
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (uint32_t y = warp; y < th; y += kScanThreads / 32) {
        int32_t carry = 0;
        for (uint32_t x0 = 0; x0 < tw; x0 += 32) {
            const uint32_t x = x0 + lane;
            int32_t v = x < tw ? g[(size_t)(y * gw + x) * gs] : 0;
#pragma unroll
            for (int off = 1; off < 32; off <<= 1) {
                const int32_t t = __shfl_up_sync(0xffffffffu, v, off);
                if (lane >= off) v += t;
            }
            v += carry;
            if (x < tw) g[(size_t)(y * gw + x) * gs] = v;
            carry = __shfl_sync(0xffffffffu, v, 31);
        }
    }
    __syncthreads();
    // columns: running sum down y, one thread per column; the result is the number of pairs of tile (y, x)
    int64_t mine = 0;
    for (uint32_t x = threadIdx.x; x < tw; x += kScanThreads) {
        int32_t acc = 0;
        for (uint32_t y0 = 0; y0 < th; y0 += 8) {
            int32_t d[8];  // independent loads first: the running sum must not serialise eight round trips
#pragma unroll
            for (int k = 0; k < 8; ++k) d[k] = y0 + k < th ? g[(size_t)((y0 + k) * gw + x) * gs] : 0;
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                if (y0 + k < th) {
                    acc += d[k];
                    counts[(y0 + k) * tw + x] = acc;
                    g[(size_t)((y0 + k) * gw + x) * gs] = acc;  // kept for the offsets phase (shared-memory path)
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
            cursor[(size_t)t * kCursorStride] = o;
        }
        base += tot;
    }
    if (cam == C - 1 && threadIdx.x == 0) {
        *n_isects = base;
        report_truncation(status, base > capacity);
    }
}

// Phase 2, single camera with the counts still in shared memory: every thread owns a run of consecutive tiles, so the
// whole exclusive scan needs ONE block scan and no global reads.
__device__ void bin_offsets_from_smem(const int32_t *cnt, uint32_t n_tiles, int64_t capacity, int32_t *offsets,
                                      int32_t *cursor, int64_t *n_isects, int32_t *status, int64_t *s_warp) {
    const uint32_t per = (n_tiles + kScanThreads - 1) / kScanThreads;
    const uint32_t t0 = threadIdx.x * per, t1 = min(n_tiles, t0 + per);
    int64_t mine = 0, total;
    for (uint32_t t = t0; t < t1; ++t) mine += cnt[t];
    int64_t run = block_inclusive_scan_i64(mine, s_warp, &total) - mine;
    for (uint32_t t = t0; t < t1; ++t) {
        const int32_t o = (int32_t)(run < capacity ? run : capacity);
        offsets[t] = o;
        cursor[(size_t)t * kCursorStride] = o;
        run += cnt[t];
    }
    if (threadIdx.x == 0) {
        *n_isects = total;
        report_truncation(status, total > capacity);
    }
}

__global__ void __launch_bounds__(kScanThreads)
bin_scan_kernel(uint32_t C, uint32_t tw, uint32_t th, int64_t capacity, int32_t *__restrict__ delta,
                int32_t *__restrict__ offsets, int32_t *__restrict__ cursor, int64_t *__restrict__ cam_total,
                int64_t *__restrict__ n_isects, int32_t *__restrict__ status, int phase, int grid_in_smem) {
    extern __shared__ int32_t s_grid[];  // th * tw ints when grid_in_smem
    __shared__ int64_t s_warp[kScanThreads / 32];
    const uint32_t cam = blockIdx.x, n_tiles = tw * th;
    if (phase != 2) {
        int32_t *d = delta + (size_t)cam * (th + 1) * (tw + 1) * kDeltaStride;
        if (grid_in_smem) {
            for (uint32_t i = threadIdx.x; i < n_tiles; i += kScanThreads) {
                const uint32_t y = i / tw, x = i - y * tw;
                s_grid[i] = d[(size_t)(y * (tw + 1) + x) * kDeltaStride];
            }
            __syncthreads();
            bin_counts_of_camera(s_grid, tw, 1, tw, th, offsets + (size_t)cam * n_tiles, cam_total + cam, s_warp);
        } else {
            bin_counts_of_camera(d, tw + 1, kDeltaStride, tw, th, offsets + (size_t)cam * n_tiles, cam_total + cam,
                                 s_warp);
        }
    }
    // a single camera needs no second launch: its total is already visible to this CTA
    if (phase != 2 && C == 1 && grid_in_smem) {
        bin_offsets_from_smem(s_grid, n_tiles, capacity, offsets, cursor, n_isects, status, s_warp);
        return;
    }
    if (phase == 2 || C == 1)
        bin_offsets_of_camera(cam, C, n_tiles, cam_total, capacity, offsets + (size_t)cam * n_tiles,
                              cursor + (size_t)cam * n_tiles * kCursorStride, n_isects, status, s_warp);
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
    int32_t *cur = cursor + (idx / N) * (int64_t)(tile_width * tile_height) * kCursorStride;
    // four slot reservations in flight before the first dependent store: the returning atomics are what this
    // kernel waits on (long-scoreboard stalls, 13 % issue utilisation when issued one at a time)
    const uint32_t cnt = (t.x1 - t.x0) * (t.y1 - t.y0);
    uint32_t tx = t.x0, ty = t.y0;
    for (uint32_t k0 = 0; k0 < cnt; k0 += 4) {
        uint32_t pos[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            pos[q] = 0xffffffffu;
            if (k0 + q < cnt) {
                pos[q] = (uint32_t)atomicAdd(cur + (size_t)(ty * tile_width + tx) * kCursorStride, 1);
                if (++tx == t.x1) {
                    tx = t.x0;
                    ++ty;
                }
            }
        }
#pragma unroll
        for (int q = 0; q < 4; ++q)
            if ((int64_t)pos[q] < capacity) keyval[pos[q]] = kv;
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

// One stable LSD pass over BITS key bits of a segment held in shared memory as 64-bit words key << 32 | id.
//   * peer sets from BITS ballots (4 instructions per bit: sign mask, predicate, VOTE, LOP3);
//   * per-warp digit counters -> (thread per digit, or per digit pair when BITS == 9) exclusive offsets over warps,
//     one block scan, segment-wide digit bases folded back into the per-warp counters, so the scatter position is
//     one shared load + the in-warp rank.
template <int ITEMS>
struct SegSmem {
    unsigned long long kv[2][kSegThreads * ITEMS];  // 32 / 64 / 128 KB
    __align__(16) uint32_t cnt[kSegWarps * 512];  // 16 KB
    uint32_t tmp[kSegWarps];
    uint32_t wmin[kSegWarps], wmax[kSegWarps];
};

template <int BITS, int ITEMS>
__device__ __noinline__ void segment_radix_pass(SegSmem<ITEMS> &sm, int cur, int shift, int32_t n, int32_t chunk) {
    constexpr int NDIG = 1 << BITS;
    constexpr int PER = NDIG > kSegThreads ? NDIG / kSegThreads : 1;  // digits per thread in the prefix step
    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int i = (int)tid; i < kSegWarps * NDIG / 4; i += kSegThreads)
        reinterpret_cast<uint4 *>(sm.cnt)[i] = make_uint4(0u, 0u, 0u, 0u);
    __syncthreads();
    uint32_t *my_cnt = sm.cnt + warp * NDIG;
    const int32_t wbase = (int32_t)warp * chunk;
    uint32_t rank2[ITEMS / 2];  // two 16-bit in-warp ranks per register (a rank is < kSegMax)
#pragma unroll
    for (int it = 0; it < ITEMS / 2; ++it) rank2[it] = 0;
#pragma unroll
    for (int it = 0; it < ITEMS; ++it) {
        if (it * 32 < chunk) {
            const int32_t e = wbase + it * 32 + (int32_t)lane;
            const bool valid = e < n;
            const uint32_t key = valid ? (uint32_t)(sm.kv[cur][e] >> 32) : 0u;
            const uint32_t d = (key >> shift) & (uint32_t)(NDIG - 1);
            uint32_t peers = __ballot_sync(0xffffffffu, valid);
#pragma unroll
            for (int b = 0; b < BITS; ++b) {
                const uint32_t sgn = (uint32_t)((int32_t)(d << (31 - b)) >> 31);  // all ones iff bit b of d is set
                const uint32_t m = __ballot_sync(0xffffffffu, sgn != 0u);
                peers &= ~(m ^ sgn);                                               // bit ? m : ~m
            }
            const int leader = __ffs(peers) - 1;
            uint32_t prev = 0;
            if ((int)lane == leader && valid) {
                prev = my_cnt[d];
                my_cnt[d] = prev + __popc(peers);
            }
            prev = __shfl_sync(0xffffffffu, prev, leader);
            __syncwarp();
            rank2[it >> 1] |= (prev + __popc(peers & ((1u << lane) - 1u))) << (16 * (it & 1));
        }
    }
    __syncthreads();
    // digits of this thread: counts of the warps -> exclusive offsets over warps; segment-wide digit totals
    uint32_t tot[PER], sum = 0;
    if (tid * PER < NDIG) {
#pragma unroll
        for (int q = 0; q < PER; ++q) {
            const int d = (int)tid * PER + q;
            uint32_t acc = 0;
#pragma unroll
            for (int w = 0; w < kSegWarps; ++w) {
                const uint32_t c = sm.cnt[w * NDIG + d];
                sm.cnt[w * NDIG + d] = acc;
                acc += c;
            }
            tot[q] = acc;
            sum += acc;
        }
    }
    uint32_t base = block_exclusive_scan_u32_256(sum, sm.tmp);  // contains two barriers
    if (tid * PER < NDIG) {
#pragma unroll
        for (int q = 0; q < PER; ++q) {
            const int d = (int)tid * PER + q;
#pragma unroll
            for (int w = 0; w < kSegWarps; ++w) sm.cnt[w * NDIG + d] += base;
            base += tot[q];
        }
    }
    __syncthreads();
#pragma unroll
    for (int it = 0; it < ITEMS; ++it) {
        if (it * 32 < chunk) {
            const int32_t e = wbase + it * 32 + (int32_t)lane;
            if (e < n) {
                const unsigned long long kv = sm.kv[cur][e];
                const uint32_t d = ((uint32_t)(kv >> 32) >> shift) & (uint32_t)(NDIG - 1);
                sm.kv[cur ^ 1][my_cnt[d] + ((rank2[it >> 1] >> (16 * (it & 1))) & 0xffffu)] = kv;
            }
        }
    }
    __syncthreads();
}

// Sort of one tile's segment in shared memory (all threads of the CTA; ends with the sorted pairs written out).
// ITEMS pairs per thread: 8 (tiles of up to 2048 pairs), 16 (up to 4096), 32 (up to 8192).
template <int ITEMS>
__device__ __forceinline__ void sort_segment_shared(SegSmem<ITEMS> &sm, const SegRange seg,
                                                    const uint64_t *__restrict__ keyval,
                                                    int64_t *__restrict__ isect_ids, int32_t *__restrict__ flatten_ids) {
    const int32_t n = seg.n;
    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint64_t *src = keyval + seg.start;

    // bucket tables of the MSD shortcut below (they live in the counter array of the LSD passes)
    constexpr int kDig = 1 << kTopBits, kPer = kDig / kSegThreads;  // digits, digits per thread in the scan
    static_assert(2 * kDig <= kSegWarps * 512 && kPer % 4 == 0, "bucket tables live in the counter array");
    uint32_t *start = sm.cnt, *hist = sm.cnt + kDig;  // first position / number of members of every bucket
#pragma unroll
    for (int q = 0; q < kPer / 4; ++q) reinterpret_cast<uint4 *>(hist)[tid * (kPer / 4) + q] = make_uint4(0u, 0u, 0u, 0u);

    // ---- depth range of the segment; keys are taken relative to the smallest one -------------------------------
    unsigned long long mine[ITEMS];
    uint32_t kmin = 0xffffffffu, kmax = 0u;
#pragma unroll
    for (int it = 0; it < ITEMS; ++it) {
        const int32_t i = it * kSegThreads + (int32_t)tid;
        mine[it] = 0ull;
        if (i < n) {
            mine[it] = src[i];
            const uint32_t k = (uint32_t)(mine[it] >> 32);
            kmin = min(kmin, k);
            kmax = max(kmax, k);
        }
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        kmin = min(kmin, __shfl_xor_sync(0xffffffffu, kmin, off));
        kmax = max(kmax, __shfl_xor_sync(0xffffffffu, kmax, off));
    }
    if (lane == 0) sm.wmin[warp] = kmin, sm.wmax[warp] = kmax;
    __syncthreads();
#pragma unroll
    for (int w = 0; w < kSegWarps; ++w) {
        kmin = min(kmin, sm.wmin[w]);
        kmax = max(kmax, sm.wmax[w]);
    }
    const uint32_t lo = kmin;
#pragma unroll
    for (int it = 0; it < ITEMS; ++it) mine[it] -= (unsigned long long)lo << 32;  // (slots past n are never used)
    const int nbits = 32 - __clz(kmax - kmin);  // 0 when every depth is identical
    const int passes = (nbits + 8) / 9;         // digits of at most 9 bits
    const int width = passes > 0 ? (nbits + passes - 1) / passes : 0;

    // ---- MSD shortcut: ONE counting pass over the top (at most kTopBits) varying depth bits, then every element
    // ranks itself inside its bucket by comparing the full 64-bit words depth << 32 | id (unique, so this is the final
    // order and the tie rule at once).  A tile holds a few hundred pairs spread over up to 2048 buckets, so a bucket
    // has a handful of members and the ranking costs less than the LSD passes it replaces.  The bucket pass need not
    // be stable, so it is a counting sort with ONE shared-memory atomic per element, issued from the registers the
    // element was loaded into: the returned count is the element's slot inside its bucket (spread-address ATOMS cost
    // 2 cycles per lane -- the pass is bound by them).  Segments whose depths cluster (a bucket above kBucketMax
    // members) take the LSD passes below instead.
    if (nbits > 0) {
        const int cap_w = n >= 768 ? kTopBits : kTopBits - 2;  // short segments: fewer, fuller buckets measure faster
        const int top_w = nbits < cap_w ? nbits : cap_w, top_shift = nbits - top_w;
        const int ndig = 1 << top_w;
        uint32_t slot[ITEMS];
#pragma unroll
        for (int it = 0; it < ITEMS; ++it) {
            slot[it] = 0u;
            if (it * kSegThreads + (int32_t)tid < n) slot[it] = atomicAdd(&hist[(uint32_t)(mine[it] >> 32) >> top_shift], 1u);
        }
        __syncthreads();
        uint32_t c[kPer], sum = 0;
        bool big = false;
#pragma unroll
        for (int q = 0; q < kPer / 4; ++q) {
            const uint4 v = reinterpret_cast<const uint4 *>(hist)[tid * (kPer / 4) + q];
            c[4 * q] = v.x, c[4 * q + 1] = v.y, c[4 * q + 2] = v.z, c[4 * q + 3] = v.w;
        }
#pragma unroll
        for (int q = 0; q < kPer; ++q) {
            sum += c[q];
            big |= c[q] > (uint32_t)kBucketMax;
        }
        uint32_t run = block_exclusive_scan_u32_256(sum, sm.tmp);  // contains two barriers
#pragma unroll
        for (int q = 0; q < kPer; ++q) {
            const uint32_t t = c[q];
            c[q] = run;
            run += t;
        }
#pragma unroll
        for (int q = 0; q < kPer / 4; ++q)
            reinterpret_cast<uint4 *>(start)[tid * (kPer / 4) + q] = make_uint4(c[4 * q], c[4 * q + 1], c[4 * q + 2], c[4 * q + 3]);
        if (!__syncthreads_or(big)) {
#pragma unroll
            for (int it = 0; it < ITEMS; ++it)
                if (it * kSegThreads + (int32_t)tid < n)
                    sm.kv[1][start[(uint32_t)(mine[it] >> 32) >> top_shift] + slot[it]] = mine[it];
            __syncthreads();
            const unsigned long long *B = sm.kv[1];
            int64_t *keys_out = isect_ids + seg.start;
            int32_t *vals_out = flatten_ids + seg.start;
            for (int32_t i = (int32_t)tid; i < n; i += kSegThreads) {
                const unsigned long long e = B[i];
                const uint32_t k = (uint32_t)(e >> 32);
                const int d = (int)(k >> top_shift);
                const uint32_t b0 = start[d], b1 = d + 1 < ndig ? start[d + 1] : (uint32_t)n;
                uint32_t pos = b0;
#pragma unroll 4
                for (uint32_t j = b0; j < b1; ++j) pos += B[j] < e;
                keys_out[pos] = (int64_t)(seg.hi | (uint64_t)(k + lo));
                vals_out[pos] = (int32_t)(uint32_t)e;
            }
            return;
        }
    }

    // ---- LSD passes (clustered depths, or all depths identical) ------------------------------------------------
#pragma unroll
    for (int it = 0; it < ITEMS; ++it) {
        const int32_t i = it * kSegThreads + (int32_t)tid;
        if (i < n) sm.kv[0][i] = mine[it];
    }
    __syncthreads();
    // each warp owns a contiguous chunk (a multiple of 32 elements); order inside = (item, lane)
    const int32_t chunk = ((n + kSegThreads - 1) / kSegThreads) * 32;
    int cur = 0;
    for (int p = 0; p < passes; ++p) {
        const int shift = p * width;
        switch (width) {  // uniform across the CTA
            case 1: segment_radix_pass<1, ITEMS>(sm, cur, shift, n, chunk); break;
            case 2: segment_radix_pass<2, ITEMS>(sm, cur, shift, n, chunk); break;
            case 3: segment_radix_pass<3, ITEMS>(sm, cur, shift, n, chunk); break;
            case 4: segment_radix_pass<4, ITEMS>(sm, cur, shift, n, chunk); break;
            case 5: segment_radix_pass<5, ITEMS>(sm, cur, shift, n, chunk); break;
            case 6: segment_radix_pass<6, ITEMS>(sm, cur, shift, n, chunk); break;
            case 7: segment_radix_pass<7, ITEMS>(sm, cur, shift, n, chunk); break;
            case 8: segment_radix_pass<8, ITEMS>(sm, cur, shift, n, chunk); break;
            default: segment_radix_pass<9, ITEMS>(sm, cur, shift, n, chunk); break;
        }
        cur ^= 1;
    }

    // ---- write out; bit-identical depths are ordered by flatten id (each element ranks itself in its run) ------
    const unsigned long long *S = sm.kv[cur];
    int64_t *out_keys = isect_ids + seg.start;
    int32_t *out_vals = flatten_ids + seg.start;
    for (int32_t i = (int32_t)tid; i < n; i += kSegThreads) {
        const unsigned long long kv = S[i];
        const uint32_t k = (uint32_t)(kv >> 32), id = (uint32_t)kv;
        int32_t a = i, less = 0;
        while (a > 0 && (uint32_t)(S[a - 1] >> 32) == k) {
            --a;
            less += (uint32_t)S[a] < id;
        }
        for (int32_t j = i + 1; j < n && (uint32_t)(S[j] >> 32) == k; ++j) less += (uint32_t)S[j] < id;
        out_keys[a + less] = (int64_t)(seg.hi | (uint64_t)(k + lo));
        out_vals[a + less] = (int32_t)id;
    }
}

// Short segments (the common case): one CTA per tile, 4 CTAs per SM.
__global__ void __launch_bounds__(kSegThreads, 4)
segment_sort_kernel(uint32_t n_slots, uint32_t n_tiles, uint32_t tile_n_bits, const int32_t *__restrict__ offsets,
                    const int64_t *__restrict__ n_isects_dev, int64_t capacity, const uint64_t *__restrict__ keyval,
                    int64_t *__restrict__ isect_ids, int32_t *__restrict__ flatten_ids) {
    extern __shared__ __align__(16) unsigned char seg_smem_raw[];
    SegSmem<kSegItems> &sm = *reinterpret_cast<SegSmem<kSegItems> *>(seg_smem_raw);
    const int64_t n_total = min(*n_isects_dev, capacity);
    const SegRange seg = segment_of(blockIdx.x, n_slots, n_tiles, tile_n_bits, offsets, n_total);
    if (seg.n <= 0 || seg.n > kSegMax) return;  // longer segments: segment_sort_long_kernel
    sort_segment_shared<kSegItems>(sm, seg, keyval, isect_ids, flatten_ids);
}

// Very long segments (n > kSegMaxShared = 8192): LSD radix sort of the full 64-bit words (depth_bits << 32 | id: unique,
// so no tie pass) in global memory, ping-ponging between the segment's span of `keyval` and of `alt`.  Correct for
// any length; speed matters little here.
__device__ __noinline__ void sort_segment_global(const SegRange seg, uint64_t *keyval, uint64_t *alt,
                                                 int64_t *__restrict__ isect_ids, int32_t *__restrict__ flatten_ids) {
    __shared__ uint32_t s_cnt[kSegWarps * 256];
    __shared__ uint32_t s_base[256];
    __shared__ uint32_t s_tmp[kSegWarps];
    __shared__ unsigned long long s_or[kSegWarps];
    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int32_t n = seg.n;
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

// Tiles with more than 2048 pairs -- dense regions, coarse tile grids -- in ONE launch: a persistent grid in which
// every CTA owns a contiguous range of slots, finds its long ones with all threads in parallel and sorts them one
// after the other: up to 4096 / 8192 pairs with the shared-memory sort at 16 / 32 pairs per thread, beyond that in
// global memory.  (One CTA per slot would spend its time launching CTAs that exit: few tiles are this long.)
__global__ void __launch_bounds__(kSegThreads, 1)
segment_sort_long_kernel(uint32_t n_slots, uint32_t n_tiles, uint32_t tile_n_bits, const int32_t *__restrict__ offsets,
                         const int64_t *__restrict__ n_isects_dev, int64_t capacity, uint64_t *keyval, uint64_t *alt,
                         int64_t *__restrict__ isect_ids, int32_t *__restrict__ flatten_ids) {
    extern __shared__ __align__(16) unsigned char seg_smem_raw[];
    __shared__ uint32_t s_long[kSegThreads];
    __shared__ uint32_t s_n_long;
    const uint32_t tid = threadIdx.x;
    const int64_t n_total = min(*n_isects_dev, capacity);
    const uint32_t per_cta = (n_slots + gridDim.x - 1) / gridDim.x;
    const uint32_t slot_begin = blockIdx.x * per_cta, slot_end = min(n_slots, slot_begin + per_cta);
    for (uint32_t s0 = slot_begin; s0 < slot_end; s0 += kSegThreads) {
        if (tid == 0) s_n_long = 0;
        __syncthreads();
        if (s0 + tid < slot_end && segment_of(s0 + tid, n_slots, n_tiles, tile_n_bits, offsets, n_total).n > kSegMax)
            s_long[atomicAdd(&s_n_long, 1u)] = s0 + tid;
        __syncthreads();
        const uint32_t n_long = s_n_long;
        for (uint32_t li = 0; li < n_long; ++li) {
            const SegRange seg = segment_of(s_long[li], n_slots, n_tiles, tile_n_bits, offsets, n_total);
            if (seg.n <= kSegThreads * kSegItemsMid)
                sort_segment_shared<kSegItemsMid>(*reinterpret_cast<SegSmem<kSegItemsMid> *>(seg_smem_raw), seg, keyval,
                                                  isect_ids, flatten_ids);
            else if (seg.n <= kSegMaxShared)
                sort_segment_shared<kSegItemsLong>(*reinterpret_cast<SegSmem<kSegItemsLong> *>(seg_smem_raw), seg,
                                                   keyval, isect_ids, flatten_ids);
            else
                sort_segment_global(seg, keyval, alt, isect_ids, flatten_ids);
            __syncthreads();  // the next segment reuses the shared buffers
        }
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
    // the camera's tile grid is staged in shared memory when it fits (up to 160 KB = 40960 tiles)
    const size_t grid_smem = sizeof(int32_t) * (size_t)n_tiles;
    const int in_smem = grid_smem <= 160 * 1024 ? 1 : 0;
    static size_t scan_smem_set = 0;
    if (in_smem && grid_smem > 48 * 1024 && grid_smem > scan_smem_set) {
        UBS_CUDA_TRY(cudaFuncSetAttribute(bin_scan_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
        scan_smem_set = 160 * 1024;
    }
    bin_scan_kernel<<<(unsigned)C, kScanThreads, in_smem ? grid_smem : 0, s>>>(
        (uint32_t)C, (uint32_t)tile_width, (uint32_t)tile_height, capacity, w.delta, offsets, w.cursor, w.cam_total,
        n_isects, status, 1, in_smem);
    UBS_LAUNCH_CHECK("bin_scan_kernel");
    if (C > 1) {
        bin_scan_kernel<<<(unsigned)C, kScanThreads, 0, s>>>((uint32_t)C, (uint32_t)tile_width, (uint32_t)tile_height,
                                                             capacity, w.delta, offsets, w.cursor, w.cam_total,
                                                             n_isects, status, 2, 0);
        UBS_LAUNCH_CHECK("bin_scan_kernel");
    }
    if (CN == 0 || capacity == 0) return UBS_OK;
    bin_emit_kernel<<<(unsigned)ceil_div(CN, kIsectThreads), kIsectThreads, 0, s>>>(
        CN, N, means2d, radii, depths, (uint32_t)tile_size, (uint32_t)tile_width, (uint32_t)tile_height, capacity,
        w.cursor, w.keyval);
    UBS_LAUNCH_CHECK("bin_emit_kernel");
    int sm = ubs_device_sm_count();
    if (sm <= 0) sm = 148;
    static bool seg_attr_set = false;
    if (!seg_attr_set) {
        UBS_CUDA_TRY(cudaFuncSetAttribute(segment_sort_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          (int)sizeof(SegSmem<kSegItems>)));
        UBS_CUDA_TRY(cudaFuncSetAttribute(segment_sort_long_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          (int)sizeof(SegSmem<kSegItemsLong>)));
        seg_attr_set = true;
    }
    segment_sort_kernel<<<n_slots, kSegThreads, sizeof(SegSmem<kSegItems>), s>>>(
        n_slots, n_tiles, (uint32_t)tile_n_bits, offsets, n_isects, capacity, w.keyval, isect_ids, flatten_ids);
    UBS_LAUNCH_CHECK("segment_sort_kernel");
    if (capacity > kSegMax) {  // a tile of more than 2048 pairs is possible at all
        const unsigned g_long = n_slots < (unsigned)sm ? n_slots : (unsigned)sm;
        segment_sort_long_kernel<<<g_long, kSegThreads, sizeof(SegSmem<kSegItemsLong>), s>>>(
            n_slots, n_tiles, (uint32_t)tile_n_bits, offsets, n_isects, capacity, w.keyval, w.alt, isect_ids,
            flatten_ids);
        UBS_LAUNCH_CHECK("segment_sort_long_kernel");
    }
    return UBS_OK;
}
