// Library-level entry points: error channel, version, device query.
#include <stdarg.h>
#include <string.h>

#include <atomic>

#include "common.cuh"

namespace ubs {
static thread_local char g_err[512] = "";
void set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
static std::atomic<unsigned long long> g_launches{0};
void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }
}  // namespace ubs

extern "C" unsigned long long ubs_launch_count(void) { return ubs::g_launches.load(std::memory_order_relaxed); }
extern "C" const char *ubs_last_error(void) { return ubs::g_err; }
extern "C" int ubs_version(void) { return 5; }
extern "C" int ubs_device_sm_count(void) {
    int dev = 0, n = 0;
    UBS_CUDA_TRY(cudaGetDevice(&dev));
    UBS_CUDA_TRY(cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev));
    return n;
}
extern "C" int ubs_record_stride(int D) {
    if (D < 4 || D > 8) return UBS_EINVAL;
    return UBS_RECORD_STRIDE(D);
}
