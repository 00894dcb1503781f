"""`from gsplat.rendering import rasterization` (scene/beta_model.py:12) -> ubs_b200.rendering.rasterization."""
from ubs_b200.rendering import depth_to_normal, rasterization  # noqa: F401
