// Entry points declared in ubs_b200.h whose kernels are not built yet return UBS_EUNSUPPORTED (loudly, never a
// silent fallback).  Each stub disappears when its kernel file lands.
#include "common.cuh"

#define UBS_STUB(name, ...)                                                                                            \
    extern "C" int name(__VA_ARGS__) {                                                                                 \
        ubs::set_error(#name ": not implemented in this build");                                                       \
        return UBS_EUNSUPPORTED;                                                                                       \
    }

UBS_STUB(ubs_rasterize_bwd, int, int64_t, const int64_t *, int64_t, const float *, const float *, const float *, const float *,
         const float *, const float *, const uint8_t *, int, int, int, int, const int32_t *, const int32_t *,
         const float *, const int32_t *, const float *, const float *, float *, float *, float *, float *, float *,
         void *)
UBS_STUB(ubs_fused_project_bwd, int, int64_t, int, const float *, const float *, const float *, const float *,
         const float *, int, int, float, int, const int32_t *, const float *, const float *, const float *,
         const float *, const float *, const float *, const float *, float *, void *)
