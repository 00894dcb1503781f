"""`from gsplat.cuda._wrapper import ...` (scene/beta_model.py:18-22) -> the ubs_b200 operator set."""
from ubs_b200.ops import (  # noqa: F401
    cond_mean_convariance_opacity,
    fully_fused_projection,
    isect_offset_encode,
    isect_tiles,
    l_triangle_to_rotmat,
    rasterize_to_pixels,
    rot_scale_l_triangle_to_covar,
)
