// Numeric constants of the rasterization path (CUDA side).
// Mirror of oracle/constants.py; tests/test_oracle_cpu.py::test_constants_agree_between_oracle_and_cuda_header checks the two agree.
// Values restate the published gsplat ~v1.4 algorithm (PARITY UNPINNED).
#pragma once

#define HGS_TILE_SIZE 16
#define HGS_ALPHA_MAX 0.999f
#define HGS_ALPHA_MIN (1.0f / 255.0f)
#define HGS_T_EPS 1e-4f
#define HGS_RADIUS_SIGMA 3.0f
#define HGS_EIG_FLOOR 0.01f
#define HGS_FOV_MARGIN 0.3f
#define HGS_ED_ALPHA_FLOOR 1e-10f
#define HGS_SH_OFFSET 0.5f
#define HGS_FILTER_INV_SQUARE_2DGS 2.0f
#define HGS_RADIUS_FLOOR_2DGS 1e-4f
#define HGS_MEDIAN_T_2DGS 0.5f
