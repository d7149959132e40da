"""CPU oracle for the rasterization hot path.  TEST INFRASTRUCTURE ONLY.

Nothing under ``horizongs_b200/`` may import this package.  The only callers
are ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline``
/ ``--impl reference`` legs, and there only as the checker / the CPU arm.

PARITY UNPINNED: the arithmetic of
this path lives in the third-party package ``gsplat`` (unpinned in the
reference's environment.yml:28, fork named without commit at README.md:29),
which is neither vendored in /root/reference nor installed here, and the
reference ships no tests, golden vectors or fixtures.  This oracle restates the
published gsplat ~v1.4 algorithm; the only pieces pinned against reference code
are the conventions the reference states in-tree (SH basis, wxyz quaternion ->
rotation, world->view construction), see tests/golden/make_golden.py.
"""
from . import constants  # noqa: F401
from .gsplat_oracle import (  # noqa: F401
    fully_fused_projection,
    fully_fused_projection_2dgs,
    spherical_harmonics,
    isect_tiles,
    isect_offset_encode,
    rasterize_to_pixels,
    rasterize_to_pixels_2dgs,
    rasterization,
    rasterization_2dgs,
    depth_to_normal,
)
