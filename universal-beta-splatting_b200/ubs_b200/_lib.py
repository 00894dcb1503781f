"""ctypes loader for libubs_b200.so (the C-ABI declared in include/ubs_b200.h).

There is NO fallback: if the library is missing or a symbol is absent, importing / calling fails loudly.
Replaces the reference's JIT/pybind backend loader (submodules/gsplat/cuda/_backend.py:79-142).
"""
import ctypes
import os
from ctypes import c_char_p, c_double, c_float, c_int, c_int64, c_size_t, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libubs_b200.so")

_P = c_void_p  # every device pointer crosses the ABI as an opaque address

# name -> (restype, argtypes); must list every symbol of include/ubs_b200.h
SIGNATURES = {
    "ubs_last_error": (c_char_p, []),
    "ubs_launch_count": (ctypes.c_ulonglong, []),
    "ubs_version": (c_int, []),
    "ubs_device_sm_count": (c_int, []),
    "ubs_record_stride": (c_int, [c_int]),
    "ubs_l_triangle_to_rotmat_fwd": (c_int, [c_int64, _P, _P, _P]),
    "ubs_l_triangle_to_rotmat_bwd": (c_int, [c_int64, _P, _P, _P]),
    "ubs_rot_scale_l_triangle_to_covar_fwd": (c_int, [c_int64, c_int, c_int, _P, _P, _P, _P, _P]),
    "ubs_rot_scale_l_triangle_to_covar_bwd": (c_int, [c_int64, c_int, c_int, _P, _P, _P, _P, _P, _P, _P, _P]),
    "ubs_cond_mean_covar_opacity_fwd": (c_int, [c_int64, c_int, _P, _P, _P, _P, _P, _P, _P, _P, _P]),
    "ubs_cond_mean_covar_opacity_bwd": (c_int, [c_int64, c_int] + [_P] * 13),
    "ubs_projection_fwd": (c_int, [c_int, c_int64, _P, _P, _P, _P, c_int, c_int, c_float, c_float, c_float, c_float,
                                   _P, _P, _P, _P, _P, _P]),
    "ubs_projection_bwd": (c_int, [c_int, c_int64, _P, _P, _P, _P, c_int, c_int, c_float] + [_P] * 11),
    "ubs_isect_workspace_bytes": (c_size_t, [c_int64, c_int64]),
    "ubs_isect_count": (c_int, [c_int, c_int64, _P, _P, c_int, c_int, c_int, _P, _P, _P, c_size_t, _P]),
    "ubs_isect_emit_sort": (c_int, [c_int, c_int64, _P, _P, _P, c_int, c_int, c_int, c_int, _P, _P, c_int64, _P, _P,
                                    _P, _P, _P, c_size_t, _P]),
    "ubs_isect_offset_encode": (c_int, [c_int64, _P, c_int, c_int, c_int, _P, _P]),
    "ubs_isect_bin_workspace_bytes": (c_size_t, [c_int, c_int, c_int, c_int64]),
    "ubs_isect_bin_sort": (c_int, [c_int, c_int64, _P, _P, _P, c_int, c_int, c_int, c_int, _P, c_int64, _P, _P, _P, _P,
                                   _P, _P, c_size_t, _P]),
    "ubs_radix_sort_workspace_bytes": (c_size_t, [c_int64]),
    "ubs_radix_sort_pairs": (c_int, [_P, c_int64, _P, _P, _P, _P, c_int, c_int, _P, c_size_t, _P]),
    "ubs_rasterize_fwd": (c_int, [c_int, c_int64, _P, c_int64, _P, _P, _P, _P, _P, _P, _P, c_int, c_int, c_int, c_int, _P, _P,
                                  _P, _P, _P, _P]),
    "ubs_rasterize_fwd_splats": (c_int, [c_int, c_int64, _P, c_int64, _P, _P, _P, _P, c_int, c_int, c_int, c_int, _P, _P, _P,
                                         _P, _P, _P]),
    "ubs_rasterize_bwd_splats": (c_int, [c_int, c_int64, _P, c_int64, _P, _P, _P, _P, c_int, c_int, c_int, c_int] +
                                 [_P] * 13),
    "ubs_rasterize_bwd_rows": (c_int, [c_int, c_int64, _P, c_int64, _P, _P, _P, c_int, c_int, c_int] + [_P] * 9),
    "ubs_fused_project_bwd_adam_pull": (c_int, [c_int64, c_int, c_int, c_int, c_int64, _P, _P, _P, _P, _P, _P, c_int, c_int,
                                                c_float, c_int, _P, _P, _P, c_double, c_double, c_double, c_int64,
                                                c_double, c_double, _P]),
    "ubs_rasterize_bwd": (c_int, [c_int, c_int64, _P, c_int64, _P, _P, _P, _P, _P, _P, _P, c_int, c_int, c_int, c_int] +
                          [_P] * 12),
    "ubs_rasterize_count": (c_int, [c_int, _P, c_int64, _P, _P, _P, _P, c_int, c_int, c_int, _P, _P, _P, _P]),
    "ubs_fused_project_fwd": (c_int, [c_int, c_int64, c_int, _P, _P, _P, _P, _P, _P, c_int, c_int, c_float, c_float,
                                      c_float, c_float, c_int, c_int, c_int, c_int] + [_P] * 12 +
                              [c_size_t, c_int, _P, _P]),
    "ubs_fused_project_bwd": (c_int, [c_int, c_int64, c_int, _P, _P, _P, _P, _P, c_int, c_int, c_float, c_int] +
                              [_P, _P, _P, c_int, _P, _P] + [c_int, _P, _P, _P]),
    "ubs_pack_gradient_rows": (c_int, [c_int64] + [_P] * 8),
    "ubs_fused_project_bwd_unpacked": (c_int, [c_int, c_int64, c_int, _P, _P, _P, _P, _P, c_int, c_int, c_float, c_int] +
                                       [_P, _P, _P, c_int] + [_P] * 8 + [c_int, _P, _P, _P]),
    "ubs_pack_records": (c_int, [c_int64, c_int] + [_P] * 9),
    "ubs_unpack_records": (c_int, [c_int64, c_int] + [_P] * 9),
    "ubs_l1_ssim_workspace_bytes": (c_size_t, [c_int, c_int, c_int, c_int]),
    "ubs_l1_ssim_loss": (c_int, [c_int, c_int, c_int, c_int, _P, c_int64, c_int64, c_int64, c_int64, _P, c_int64,
                                 c_int64, c_int64, c_int64, c_float, c_float, _P, _P, _P, c_size_t, _P]),
    "ubs_adam_step": (c_int, [c_int64, c_int, c_int64, c_int64, _P, _P, _P, _P, _P, c_double, c_double, c_double, c_int64, c_double,
                              c_double, _P]),
    "ubs_fused_project_bwd_adam": (c_int, [c_int, c_int64, c_int, _P, _P, _P, _P, _P, c_int, c_int, c_float, c_int] +
                                   [_P, _P, _P, c_int, _P, _P, _P] +
                                   [c_double, c_double, c_double, c_int64, c_double, c_double, _P, _P]),
    "ubs_fused_project_bwd_scatter": (c_int, [c_int64, c_int, _P, _P, _P, _P, _P, c_int, c_int, c_float, c_int] +
                                      [_P, _P, _P, c_int] + [c_int, c_int, c_int64, _P, _P, _P]),
    "ubs_reduce_adam_gather": (c_int, [c_int64, c_int, c_int, c_int, c_int64, _P, _P, _P, _P, _P, _P, c_double, c_double,
                                       c_double, c_int64, c_double, c_double, _P]),
    "ubs_mcmc_relocate": (c_int, [c_int64, c_int, _P, _P, _P, c_int64, c_int64, c_int64, _P, _P, _P, _P]),
    "ubs_sgld_noise": (c_int, [c_int64, c_int, _P, _P, c_double, c_double, _P]),
}

_lib = None


class UbsError(RuntimeError):
    pass


def load():
    """Loads the shared library (once). Raises if it has not been built: there is no CPU or torch fallback."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise UbsError(
            "libubs_b200.so not found at %s -- build it with `python universal-beta-splatting_b200/build.py` "
            "(the B200 CUDA library is the only implementation; there is no fallback path)" % LIB_PATH
        )
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the library does not export a declared symbol
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc, what):
    if rc != 0:
        msg = load().ubs_last_error()
        raise UbsError("%s failed (code %d): %s" % (what, rc, msg.decode() if msg else "?"))


def ptr(t):
    """Device address of a tensor (None -> NULL)."""
    return None if t is None else t.data_ptr()
