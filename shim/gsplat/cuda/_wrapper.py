"""``gsplat.cuda._wrapper`` names imported by gaussian_renderer/render.py:14."""
from horizongs_b200.cuda._wrapper import (  # noqa: F401
    fully_fused_projection,
    fully_fused_projection_2dgs,
    isect_offset_encode,
    isect_tiles,
    rasterize_to_pixels,
    rasterize_to_pixels_2dgs,
    spherical_harmonics,
)
