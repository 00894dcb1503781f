"""TEST INFRASTRUCTURE ONLY -- builds the reference's own CUDA kernels as a parity oracle.

Compiles the hot-path translation units of the reference extension *where they lie* under
/root/reference/submodules/gsplat/cuda/csrc (nothing is copied into this repository) together with
oracle/ref_binding.cpp into oracle/_ref/ubs_ref_cuda.so, for sm_100a, with the reference's own flags
(`-O3 --use_fast_math`, reference: submodules/gsplat/cuda/_backend.py:93-99).  The result is git-ignored but
travels to the GPU box with gpurun.  /root/reference does not exist on the GPU box, so this script is a no-op
there (the prebuilt .so is used).

Usage: python oracle/build_ref.py [--force]
"""
import glob
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF_CSRC = "/root/reference/submodules/gsplat/cuda/csrc"
OUT_DIR = os.path.join(HERE, "_ref")
NAME = "ubs_ref_cuda"

HOT_PATH_TUS = [
    "cond_mean_convariance_opacity_fwd.cu",
    "cond_mean_convariance_opacity_bwd.cu",
    "rot_scale_l_triangle_to_covar_fwd.cu",
    "rot_scale_l_triangle_to_covar_bwd.cu",
    "l_triagnle_to_rotmat_fwd.cu",
    "l_triagnle_to_rotmat_bwd.cu",
    "fully_fused_projection_fwd.cu",
    "fully_fused_projection_bwd.cu",
    "isect_tiles.cu",
    "rasterize_to_pixels_fwd.cu",
    "rasterize_to_pixels_bwd.cu",
    "rasterize_to_indices_in_range.cu",
]


def so_path():
    return os.path.join(OUT_DIR, NAME + ".so")


def build(force=False, verbose=True):
    if os.path.exists(so_path()) and not force:
        return so_path()
    if not os.path.isdir(REF_CSRC):
        return None  # GPU box: only the prebuilt file can be used
    os.makedirs(OUT_DIR, exist_ok=True)
    os.environ.setdefault("TORCH_CUDA_ARCH_LIST", "10.0a")
    os.environ.setdefault("MAX_JOBS", str(os.cpu_count() or 8))
    from torch.utils.cpp_extension import load

    sources = [os.path.join(REF_CSRC, f) for f in HOT_PATH_TUS] + [os.path.join(HERE, "ref_binding.cpp")]
    load(
        name=NAME,
        sources=sources,
        extra_cflags=["-O3"],
        extra_cuda_cflags=["-O3", "--use_fast_math"],
        extra_include_paths=[REF_CSRC, os.path.join(REF_CSRC, "third_party", "glm")],
        build_directory=OUT_DIR,
        verbose=verbose,
        is_python_module=False,
    )
    return so_path() if os.path.exists(so_path()) else None


if __name__ == "__main__":
    p = build(force="--force" in sys.argv)
    print("reference oracle:", p)
