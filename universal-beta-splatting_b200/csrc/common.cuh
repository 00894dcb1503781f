// Shared host/device helpers for libubs_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "ubs_b200.h"

namespace ubs {

void set_error(const char *fmt, ...);
void count_launch();  // bumps the process-wide kernel-launch counter (ubs_launch_count)

#define UBS_CHECK_ARG(cond, ...)                                                                                       \
    do {                                                                                                               \
        if (!(cond)) {                                                                                                 \
            ::ubs::set_error(__VA_ARGS__);                                                                             \
            return UBS_EINVAL;                                                                                         \
        }                                                                                                              \
    } while (0)

#define UBS_CUDA_TRY(expr)                                                                                             \
    do {                                                                                                               \
        cudaError_t _e = (expr);                                                                                       \
        if (_e != cudaSuccess) {                                                                                       \
            ::ubs::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__);              \
            return UBS_ECUDA;                                                                                          \
        }                                                                                                              \
    } while (0)

#define UBS_LAUNCH_CHECK(name)                                                                                         \
    do {                                                                                                               \
        ::ubs::count_launch();                                                                                         \
        cudaError_t _e = cudaGetLastError();                                                                           \
        if (_e != cudaSuccess) {                                                                                       \
            ::ubs::set_error("launch of %s failed: %s", name, cudaGetErrorString(_e));                                 \
            return UBS_ECUDA;                                                                                          \
        }                                                                                                              \
    } while (0)

static inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }
static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// floor(log2(x)) + 1 for x >= 1: number of bits the reference reserves for tile / camera ids
// (reference: isect_tiles.cu:137-138 uses (uint32_t)floor(log2(x)) + 1).
static inline int id_bits(uint32_t x) {
    int b = 0;
    while (x) {
        ++b;
        x >>= 1;
    }
    return b < 1 ? 1 : b;
}

}  // namespace ubs
