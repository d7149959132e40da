"""Numeric constants of the rasterization path, one place for the oracle.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  The CUDA side keeps the
same values in horizongs_b200/csrc/hgs_constants.cuh; tests/test_oracle_cpu.py (test_constants_agree_between_oracle_and_cuda_header)
checks the two files agree, so a single edit re-aligns both once a real gsplat
is available to compare against (SURVEY.md section 7 "hard parts").

Source of every value: the published algorithm of gsplat ~v1.4 (the package the
reference calls at gaussian_renderer/render.py:40,62,149,171; unpinned in
environment.yml:28, therefore PARITY UNPINNED).
"""

TILE_SIZE = 16                 # rasterization(tile_size=16) default; render.py:40-54 does not override
ALPHA_MAX = 0.999              # alpha clamp in rasterize_to_pixels
ALPHA_MIN = 1.0 / 255.0        # skip threshold
T_EPS = 1e-4                   # stop when T * (1 - alpha) <= T_EPS
RADIUS_SIGMA = 3.0             # radius = ceil(3 * sqrt(lambda_max))
EIG_FLOOR = 0.01               # max(0.01, b*b - det) under the sqrt of the eigenvalue
FOV_MARGIN = 0.3               # tan-fov clamp margin: lim = (W - cx)/fx + 0.3 * tan_fovx
EPS2D_DEFAULT = 0.3            # render.py:158
NEAR_DEFAULT = 0.01            # render.py:160
FAR_DEFAULT = 1e10             # render.py:161
ED_ALPHA_FLOOR = 1e-10         # expected depth = acc_depth / alpha.clamp(min=1e-10)
SH_OFFSET = 0.5                # colors = clamp_min(sh_eval + 0.5, 0)
FILTER_INV_SQUARE_2DGS = 2.0   # 2DGS screen-space low-pass weight
RADIUS_FLOOR_2DGS = 1e-4       # max(1e-4, extent) under the sqrt (2DGS)
MEDIAN_T_2DGS = 0.5            # median depth = depth of last Gaussian blended while T > 0.5
