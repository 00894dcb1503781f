// Onesweep LSD radix sort (see radix_sort.cuh).  HBM-bound integer work: per pass 12 B read + 12 B written per
// pair, plus one 8 B/pair histogram read for the whole sort.
#include "radix_sort.cuh"

namespace ubs {

namespace {

constexpr uint32_t kFlagMask = 3u << 30;
constexpr uint32_t kFlagAggregate = 1u << 30;  // tile-local digit count published
constexpr uint32_t kFlagPrefix = 2u << 30;     // inclusive prefix over tiles [0, tile] published
constexpr uint32_t kValueMask = ~kFlagMask;

__device__ __forceinline__ uint32_t digit_of(uint64_t key, int shift, uint32_t mask) {
    return (uint32_t)(key >> shift) & mask;
}

// ---- global histograms for every pass: one read of the keys ------------------------------------------------
__global__ void __launch_bounds__(256)
radix_histogram_kernel(const uint64_t *__restrict__ keys, const int64_t *__restrict__ n_dev, int64_t capacity,
                       int begin_bit, int end_bit, int passes, uint32_t *__restrict__ hist) {
    __shared__ uint32_t s_hist[kSortMaxPasses * kRadix];
    for (int i = threadIdx.x; i < passes * kRadix; i += blockDim.x) s_hist[i] = 0;
    __syncthreads();
    int64_t n = *n_dev;
    if (n > capacity) n = capacity;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const uint64_t k = keys[i];
        for (int p = 0; p < passes; ++p) {
            const int shift = begin_bit + p * kRadixBits;
            const int bits = min(kRadixBits, end_bit - shift);
            atomicAdd(&s_hist[p * kRadix + digit_of(k, shift, (1u << bits) - 1u)], 1u);
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < passes * kRadix; i += blockDim.x) {
        const uint32_t v = s_hist[i];
        if (v) atomicAdd(&hist[i], v);
    }
}

// ---- one onesweep pass ---------------------------------------------------------------------------------------
struct PassSmem {
    uint64_t keys[kSortTile];
    int32_t vals[kSortTile];
    uint32_t warp_cnt[(kSortThreads / 32) * kRadix];  // per-warp digit counts -> per-warp digit offsets
    uint32_t digit_excl[kRadix];                      // exclusive scan of this tile's digit counts
    uint32_t gdst[kRadix];                            // global base of digit d minus digit_excl[d]
    uint32_t scan_tmp[kSortThreads / 32];
    uint32_t tile;
};

__device__ __forceinline__ uint32_t block_exclusive_scan_256(uint32_t v, uint32_t *tmp, uint32_t *total) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t incl = v;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, incl, off);
        if (lane >= off) incl += t;
    }
    if (lane == 31) tmp[warp] = incl;
    __syncthreads();
    uint32_t base = 0, tot = 0;
#pragma unroll
    for (int w = 0; w < kSortThreads / 32; ++w) {
        const uint32_t t = tmp[w];
        if (w < warp) base += t;
        tot += t;
    }
    if (total) *total = tot;
    __syncthreads();
    return base + incl - v;
}

__global__ void __launch_bounds__(kSortThreads)
onesweep_pass_kernel(const uint64_t *__restrict__ keys_in, const int32_t *__restrict__ vals_in,
                     uint64_t *__restrict__ keys_out, int32_t *__restrict__ vals_out,
                     const int64_t *__restrict__ n_dev, int64_t capacity, const uint32_t *__restrict__ hist,
                     uint32_t *__restrict__ tile_counter, volatile uint32_t *__restrict__ lookback, int shift,
                     int bits) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    PassSmem &sm = *reinterpret_cast<PassSmem *>(smem_raw);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t dmask = (1u << bits) - 1u;

    int64_t n64 = *n_dev;
    if (n64 > capacity) n64 = capacity;
    const uint32_t n = (uint32_t)n64;
    const uint32_t n_tiles = (n + kSortTile - 1) / kSortTile;

    // Dynamic tile assignment: a tile id is only handed out after all lower ids were, so every tile the
    // look-back waits on belongs to a CTA that is already running (forward progress without co-residency).
    if (tid == 0) sm.tile = atomicAdd(tile_counter, 1u);
    for (int i = tid; i < (kSortThreads / 32) * kRadix; i += kSortThreads) sm.warp_cnt[i] = 0;
    __syncthreads();
    const uint32_t tile = sm.tile;
    if (tile >= n_tiles) return;
    const uint32_t tile_base = tile * kSortTile;
    const uint32_t tile_n = min((uint32_t)kSortTile, n - tile_base);

    // ---- load (warp-striped: item i of lane l in warp w is element w*512 + i*32 + l) and rank ---------------
    uint64_t key[kSortItems];
    uint32_t rank[kSortItems];
    const uint32_t warp_base = warp * (32 * kSortItems);
#pragma unroll
    for (int i = 0; i < kSortItems; ++i) {
        const uint32_t e = warp_base + i * 32 + lane;
        key[i] = e < tile_n ? keys_in[tile_base + e] : ~0ull;
    }
    uint32_t *my_cnt = sm.warp_cnt + warp * kRadix;
#pragma unroll
    for (int i = 0; i < kSortItems; ++i) {
        const uint32_t e = warp_base + i * 32 + lane;
        const bool valid = e < tile_n;
        // MATCH.ANY costs ~60 cycles per warp and SM on B200 (eight ballots: 22), but the ballot form triples the
        // instruction count of this loop and measured slower here (81 vs 62 us per pass at 6.2M pairs)
        const uint32_t d = valid ? digit_of(key[i], shift, dmask) : (uint32_t)kRadix;  // invalid lanes match only each other
        const uint32_t peers = __match_any_sync(0xffffffffu, d);
        const int leader = __ffs(peers) - 1;
        uint32_t prev = 0;
        if (lane == leader && valid) {
            prev = my_cnt[d];
            my_cnt[d] = prev + __popc(peers);
        }
        prev = __shfl_sync(0xffffffffu, prev, leader);
        rank[i] = prev + __popc(peers & ((1u << lane) - 1u));
        __syncwarp();
    }
    __syncthreads();

    // ---- per digit (thread d): warp counts -> warp offsets, tile count ---------------------------------------
    uint32_t tile_cnt = 0;
    {
        const int d = tid;  // kSortThreads == kRadix
#pragma unroll
        for (int w = 0; w < kSortThreads / 32; ++w) {
            const uint32_t c = sm.warp_cnt[w * kRadix + d];
            sm.warp_cnt[w * kRadix + d] = tile_cnt;
            tile_cnt += c;
        }
    }
    // publish this tile's aggregate as early as possible
    volatile uint32_t *my_state = lookback + (size_t)tile * kRadix + tid;
    if (tile == 0) *my_state = kFlagPrefix | tile_cnt;
    else *my_state = kFlagAggregate | tile_cnt;

    // exclusive scans: of the global histogram (digit bases) and of this tile's digit counts (smem layout)
    const uint32_t hist_excl = block_exclusive_scan_256(hist[tid], sm.scan_tmp, nullptr);
    const uint32_t local_excl = block_exclusive_scan_256(tile_cnt, sm.scan_tmp, nullptr);

    // ---- decoupled look-back: exclusive prefix of digit `tid` over tiles [0, tile) ----------------------------
    uint32_t excl = 0;
    if (tile > 0) {
        int64_t t = (int64_t)tile - 1;
        while (true) {
            const uint32_t s = lookback[(size_t)t * kRadix + tid];
            const uint32_t flag = s & kFlagMask;
            if (flag == 0) continue;  // not published yet: spin
            excl += s & kValueMask;
            if (flag == kFlagPrefix) break;
            --t;
        }
        *my_state = kFlagPrefix | (excl + tile_cnt);
    }
    sm.digit_excl[tid] = local_excl;
    sm.gdst[tid] = hist_excl + excl - local_excl;  // modular arithmetic; adding the smem slot gives the address
    __syncthreads();

    // ---- scatter into shared memory in digit order (stable: warps in order, items in order) ------------------
#pragma unroll
    for (int i = 0; i < kSortItems; ++i) {
        const uint32_t e = warp_base + i * 32 + lane;
        if (e < tile_n) {
            const uint32_t d = digit_of(key[i], shift, dmask);
            const uint32_t pos = sm.digit_excl[d] + sm.warp_cnt[warp * kRadix + d] + rank[i];
            sm.keys[pos] = key[i];
            sm.vals[pos] = vals_in[tile_base + e];
        }
    }
    __syncthreads();

    // ---- coalesced write-out: consecutive smem slots of one digit go to consecutive global addresses ---------
    for (uint32_t j = tid; j < tile_n; j += kSortThreads) {
        const uint64_t k = sm.keys[j];
        const uint32_t dst = sm.gdst[digit_of(k, shift, dmask)] + j;
        keys_out[dst] = k;
        vals_out[dst] = sm.vals[j];
    }
}

}  // namespace

size_t sort_workspace_bytes(int64_t capacity) {
    const int64_t n_tiles = ceil_div(capacity > 0 ? capacity : 1, kSortTile);
    size_t b = 0;
    b += align_up(sizeof(uint32_t) * kSortMaxPasses * kRadix, 256);
    b += 256;  // tile counters
    b += align_up(sizeof(uint32_t) * kSortMaxPasses * n_tiles * kRadix, 256);
    b += align_up(sizeof(int64_t) * (size_t)capacity, 256);
    b += align_up(sizeof(int32_t) * (size_t)capacity, 256);
    return b;
}

SortWorkspace sort_workspace_carve(void *base, int64_t capacity) {
    SortWorkspace ws;
    const int64_t n_tiles = ceil_div(capacity > 0 ? capacity : 1, kSortTile);
    unsigned char *p = (unsigned char *)base;
    ws.hist = (uint32_t *)p;
    p += align_up(sizeof(uint32_t) * kSortMaxPasses * kRadix, 256);
    ws.tile_counter = (uint32_t *)p;
    p += 256;
    ws.lookback = (uint32_t *)p;
    p += align_up(sizeof(uint32_t) * kSortMaxPasses * n_tiles * kRadix, 256);
    ws.zero_bytes = (size_t)(p - (unsigned char *)base);
    ws.alt_keys = (int64_t *)p;
    p += align_up(sizeof(int64_t) * (size_t)capacity, 256);
    ws.alt_vals = (int32_t *)p;
    ws.n_tiles_cap = n_tiles;
    return ws;
}

int radix_sort_pairs_pingpong(const int64_t *n_dev, int64_t capacity, int64_t *src_keys, int32_t *src_vals,
                              int64_t *dst_keys, int32_t *dst_vals, int begin_bit, int end_bit,
                              const SortWorkspace &ws, cudaStream_t stream) {
    UBS_CHECK_ARG(begin_bit >= 0 && end_bit <= 64 && begin_bit < end_bit, "radix sort: bad bit range [%d,%d)",
                  begin_bit, end_bit);
    UBS_CHECK_ARG(capacity >= 0 && capacity <= kSortMaxN, "radix sort: capacity %lld out of range",
                  (long long)capacity);
    const int passes = sort_num_passes(begin_bit, end_bit);
    if (capacity == 0) return passes;

    static bool smem_attr_set = false;
    if (!smem_attr_set) {
        UBS_CUDA_TRY(cudaFuncSetAttribute(onesweep_pass_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          (int)sizeof(PassSmem)));
        smem_attr_set = true;
    }
    // only the first `passes` look-back slabs are used
    const size_t zero = (size_t)((unsigned char *)ws.lookback - (unsigned char *)ws.hist) +
                        sizeof(uint32_t) * (size_t)passes * ws.n_tiles_cap * kRadix;
    UBS_CUDA_TRY(cudaMemsetAsync(ws.hist, 0, zero, stream));

    int sm_count = 148;
    {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, dev);
    }
    const int64_t want = ceil_div(capacity, 256 * 8);
    const int hist_blocks = (int)(want < (int64_t)sm_count * 8 ? (want < 1 ? 1 : want) : (int64_t)sm_count * 8);
    radix_histogram_kernel<<<hist_blocks, 256, 0, stream>>>((const uint64_t *)src_keys, n_dev, capacity, begin_bit,
                                                            end_bit, passes, ws.hist);
    UBS_LAUNCH_CHECK("radix_histogram_kernel");

    const unsigned grid = (unsigned)ceil_div(capacity, kSortTile);
    int64_t *ka = src_keys, *kb = dst_keys;
    int32_t *va = src_vals, *vb = dst_vals;
    for (int p = 0; p < passes; ++p) {
        const int shift = begin_bit + p * kRadixBits;
        const int bits = (end_bit - shift) < kRadixBits ? (end_bit - shift) : kRadixBits;
        onesweep_pass_kernel<<<grid, kSortThreads, sizeof(PassSmem), stream>>>(
            (const uint64_t *)ka, va, (uint64_t *)kb, vb, n_dev, capacity, ws.hist + p * kRadix, ws.tile_counter + p,
            ws.lookback + (size_t)p * ws.n_tiles_cap * kRadix, shift, bits);
        UBS_LAUNCH_CHECK("onesweep_pass_kernel");
        int64_t *tk = ka;
        ka = kb;
        kb = tk;
        int32_t *tv = va;
        va = vb;
        vb = tv;
    }
    return passes;
}

}  // namespace ubs

extern "C" size_t ubs_radix_sort_workspace_bytes(int64_t capacity) {
    return ubs::sort_workspace_bytes(capacity < 0 ? 0 : capacity);
}

extern "C" int ubs_radix_sort_pairs(const int64_t *n_dev, int64_t capacity, int64_t *keys_in, int32_t *vals_in,
                                    int64_t *keys_out, int32_t *vals_out, int begin_bit, int end_bit, void *workspace,
                                    size_t workspace_bytes, void *stream) {
    using namespace ubs;
    if (capacity == 0) return UBS_OK;
    UBS_CHECK_ARG(n_dev && keys_in && vals_in && keys_out && vals_out && workspace, "radix_sort_pairs: null pointer");
    if (workspace_bytes < sort_workspace_bytes(capacity)) {
        set_error("radix_sort_pairs: workspace %zu < %zu", workspace_bytes, sort_workspace_bytes(capacity));
        return UBS_ENOSPC;
    }
    cudaStream_t s = (cudaStream_t)stream;
    const SortWorkspace ws = sort_workspace_carve(workspace, capacity);
    const int passes = sort_num_passes(begin_bit, end_bit);
    int rc;
    if (passes % 2 == 1) {
        rc = radix_sort_pairs_pingpong(n_dev, capacity, keys_in, vals_in, keys_out, vals_out, begin_bit, end_bit, ws, s);
    } else {
        // even pass count: start from the output arrays so the ping-pong ends there
        UBS_CUDA_TRY(cudaMemcpyAsync(keys_out, keys_in, sizeof(int64_t) * (size_t)capacity, cudaMemcpyDeviceToDevice, s));
        UBS_CUDA_TRY(cudaMemcpyAsync(vals_out, vals_in, sizeof(int32_t) * (size_t)capacity, cudaMemcpyDeviceToDevice, s));
        rc = radix_sort_pairs_pingpong(n_dev, capacity, keys_out, vals_out, keys_in, vals_in, begin_bit, end_bit, ws, s);
    }
    return rc < 0 ? rc : UBS_OK;
}
