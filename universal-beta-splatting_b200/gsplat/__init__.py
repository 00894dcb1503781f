"""Drop-in shim: put `universal-beta-splatting_b200/` on sys.path ahead of the reference's `submodules/` and the
imports of scene/beta_model.py:12-22 (`gsplat.rendering`, `gsplat.cuda._wrapper`, `gsplat.cuda._torch_impl`)
resolve to the B200 library unchanged.  See INTEGRATION.md."""
from ubs_b200 import rasterization  # noqa: F401
from ubs_b200.ops import (  # noqa: F401
    cond_mean_convariance_opacity,
    fully_fused_projection,
    isect_offset_encode,
    isect_tiles,
    l_triangle_to_rotmat,
    rasterize_to_pixels,
    rot_scale_l_triangle_to_covar,
)

__version__ = "ubs_b200"
