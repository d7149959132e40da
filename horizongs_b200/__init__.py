"""horizongs_b200 -- B200 (sm_100a) implementation of the gsplat rasterization hot path that
Horizon-GS reaches from gaussian_renderer/render.py.

Public surface = the gsplat names the reference uses:
    rasterization, rasterization_2dgs                      (render.py:40,62)
    cuda._wrapper.fully_fused_projection[_2dgs]            (render.py:14,149,171)
plus the stage operators (isect_tiles, isect_offset_encode, rasterize_to_pixels, spherical_harmonics).
Add ``shim/`` to sys.path to make ``import gsplat`` resolve to this package.

Next to the path, as submodules:
    distributed   view-sharded multi-GPU steps: FusedBackwardExchange / PeerGradientExchange / NCCL helpers
    losses        fused training loss (1 - l) * L1 + l * (1 - SSIM)          (train.py:158-160)
    decode        fused LOD filter + anchor -> neural-Gaussian decode          (basic_model.py:297-371)
    ply_io        explicit-Gaussian and anchor PLY layouts                     (lod_model.py:374-465,681-832)
    scenes        seeded synthetic scenes of BASELINE.json's configs
"""
from .rendering import rasterization, rasterization_2dgs, depth_to_normal  # noqa: F401
from .cuda._wrapper import (  # noqa: F401
    fully_fused_projection,
    fully_fused_projection_2dgs,
    isect_offset_encode,
    isect_tiles,
    rasterize_to_pixels,
    rasterize_to_pixels_2dgs,
    spherical_harmonics,
)

__version__ = "0.1.0"
