"""ubs_b200 -- B200 (sm_100a) rasteriser for Universal Beta Splatting behind the reference's operator API."""
from ._lib import LIB_PATH, UbsError, load  # noqa: F401
from .ops import (  # noqa: F401
    fully_fused_projection,
    isect_offset_encode,
    isect_tiles,
    radix_sort_pairs,
    rasterize_to_pixels,
)
from .rendering import rasterization  # noqa: F401
