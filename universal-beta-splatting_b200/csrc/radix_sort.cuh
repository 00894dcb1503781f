// Hand-written onesweep LSD radix sort for (int64 key, int32 value) pairs, sm_100a.
//
// Replaces the reference's cub::DeviceRadixSort::SortPairs call (isect_tiles.cu:230-278): stable, ascending,
// over key bits [begin_bit, end_bit), 8-bit digits.  One global-histogram kernel reads the keys once for all
// passes; each pass is ONE kernel that ranks a 4096-key tile in shared memory, obtains its global digit bases
// through decoupled look-back over per-tile digit counts, and scatters keys and values through shared memory
// so global writes are coalesced per digit run.  The element count lives in DEVICE memory so the sort composes
// with a capacity-bounded pair list without any host synchronisation.
#pragma once
#include "common.cuh"

namespace ubs {

constexpr int kRadixBits = 8;
constexpr int kRadix = 1 << kRadixBits;
constexpr int kSortThreads = 256;
constexpr int kSortItems = 16;
constexpr int kSortTile = kSortThreads * kSortItems;  // 4096 keys per CTA
constexpr int kSortMaxPasses = 8;
constexpr int64_t kSortMaxN = (int64_t)1 << 30;  // look-back words carry 30-bit counts

struct SortWorkspace {
    uint32_t *hist;          // [kSortMaxPasses][kRadix]
    uint32_t *tile_counter;  // [kSortMaxPasses]
    uint32_t *lookback;      // [passes][n_tiles_cap][kRadix]
    int64_t *alt_keys;       // [capacity]
    int32_t *alt_vals;       // [capacity]
    size_t zero_bytes;       // bytes from `hist` that must be zeroed before a sort
    int64_t n_tiles_cap;
};

size_t sort_workspace_bytes(int64_t capacity);
SortWorkspace sort_workspace_carve(void *base, int64_t capacity);

// Sorts `src` into `dst` when the pass count is odd, and leaves the result in `src` when it is even
// (ping-pong).  Returns the number of passes, or a negative error code.
int radix_sort_pairs_pingpong(const int64_t *n_dev, int64_t capacity, int64_t *src_keys, int32_t *src_vals,
                              int64_t *dst_keys, int32_t *dst_vals, int begin_bit, int end_bit,
                              const SortWorkspace &ws, cudaStream_t stream);

static inline int sort_num_passes(int begin_bit, int end_bit) { return (end_bit - begin_bit + kRadixBits - 1) / kRadixBits; }

}  // namespace ubs
