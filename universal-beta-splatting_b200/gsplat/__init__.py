"""Drop-in shim: put `universal-beta-splatting_b200/` on sys.path ahead of the reference's `submodules/` and the
imports of scene/beta_model.py:12-22 (`gsplat.rendering`, `gsplat.cuda._wrapper`, `gsplat.cuda._torch_impl`)
resolve to the B200 library unchanged.  See INTEGRATION.md."""
from ubs_b200 import rasterization  # noqa: F401
from ubs_b200.dropin import (  # noqa: F401  (deferred forms: rasterization() fuses them, see ubs_b200/dropin.py)
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

__version__ = "ubs_b200"
