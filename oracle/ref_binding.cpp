// TEST INFRASTRUCTURE ONLY -- not part of the product path.
//
// pybind11 module exposing the *reference's own* CUDA kernels for the hot path
// (compiled in place from /root/reference/submodules/gsplat/cuda/csrc/*.cu by
// oracle/build_ref.py; no reference source is copied into this repository).
// It replaces the reference's csrc/ext.cpp:3-58 for the 13 functions on the
// path and leaves out the dead legacy ops (proj_*, world_to_cam_*,
// quat_scale_to_covar_preci_*), so only 12 translation units need building.
// Declarations come from the reference header csrc/bindings.h:34-273, found on
// the include path at build time.
#include "bindings.h"

PYBIND11_MODULE(TORCH_EXTENSION_NAME, m) {
    m.def("cond_mean_convariance_opacity_fwd", &gsplat::cond_mean_convariance_opacity_fwd_tensor);
    m.def("cond_mean_convariance_opacity_bwd", &gsplat::cond_mean_convariance_opacity_bwd_tensor);
    m.def("rot_scale_l_triangle_to_covar_fwd", &gsplat::rot_scale_l_triangle_to_covar_fwd_tensor);
    m.def("rot_scale_l_triangle_to_covar_bwd", &gsplat::rot_scale_l_triangle_to_covar_bwd_tensor);
    m.def("l_triangle_to_rotmat_fwd", &gsplat::l_triangle_to_rotmat_fwd_tensor);
    m.def("l_triangle_to_rotmat_bwd", &gsplat::l_triangle_to_rotmat_bwd_tensor);
    m.def("fully_fused_projection_fwd", &gsplat::fully_fused_projection_fwd_tensor);
    m.def("fully_fused_projection_bwd", &gsplat::fully_fused_projection_bwd_tensor);
    m.def("isect_tiles", &gsplat::isect_tiles_tensor);
    m.def("isect_offset_encode", &gsplat::isect_offset_encode_tensor);
    m.def("rasterize_to_pixels_fwd", &gsplat::rasterize_to_pixels_fwd_tensor);
    m.def("rasterize_to_pixels_bwd", &gsplat::rasterize_to_pixels_bwd_tensor);
    m.def("rasterize_to_indices_in_range", &gsplat::rasterize_to_indices_in_range_tensor);
}
