"""Drop-in ``gsplat`` package: put this directory's parent (``<repo>/shim``) on sys.path *before* any
real gsplat, and Horizon-GS's unmodified ``gaussian_renderer/render.py`` (``import gsplat`` at :13,
``from gsplat.cuda._wrapper import ...`` at :14) runs on the B200 kernels.  See INTEGRATION.md."""
from horizongs_b200 import *  # noqa: F401,F403
from horizongs_b200 import (  # noqa: F401
    rasterization,
    rasterization_2dgs,
    fully_fused_projection,
    fully_fused_projection_2dgs,
    isect_tiles,
    isect_offset_encode,
    rasterize_to_pixels,
    rasterize_to_pixels_2dgs,
    spherical_harmonics,
)
from . import cuda  # noqa: F401
