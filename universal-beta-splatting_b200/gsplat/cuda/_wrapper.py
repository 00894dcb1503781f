"""`from gsplat.cuda._wrapper import ...` (scene/beta_model.py:18-22) -> the ubs_b200 operator set.  The three
companion operators of the conditioning chain come in their deferred form (ubs_b200/dropin.py): called the way
BetaModel.render calls them, they cost nothing until rasterization() runs the whole chain as fused kernels; used any
other way they compute through the stand-alone kernels of ubs_b200.ops."""
from ubs_b200.dropin import (  # noqa: F401
    cond_mean_convariance_opacity,
    l_triangle_to_rotmat,
    rot_scale_l_triangle_to_covar,
)
from ubs_b200.ops import (  # noqa: F401
    fully_fused_projection,
    isect_offset_encode,
    isect_tiles,
    rasterize_to_pixels,
)
