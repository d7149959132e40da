#!/usr/bin/env python
"""Culling statistics of the blend kernels on bench.py's 6M scene (instrumented replay kernel hgs_blend3d_stats):
(warp, Gaussian) pairs that survive the 8x4 sub-tile cull, and the iteration counts two independent half-warps
would need with 4x4 / 8x2 half rectangles (measured before deciding against half-warp culling)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import horizongs_b200 as hgs  # noqa: E402
from horizongs_b200 import scenes  # noqa: E402
from horizongs_b200.cuda import _wrapper as Wr  # noqa: E402

sc, views, Ks, W, H = scenes.config4()
sc = sc.to("cuda")
views, Ks = views.cuda(), Ks.cuda()
with torch.no_grad():
    for v in range(4):
        rc, ra, meta = hgs.rasterization(sc.means, sc.quats, sc.scales, sc.opacities, sc.colors, views[v:v + 1], Ks[v:v + 1],
                                         W, H, sh_degree=2, render_mode="RGB+ED")
        pe, pb = Wr.blend3d_pair_stats(meta["means2d"], meta["conics"], meta["opacities"].contiguous(), meta["radii"], W, H,
                                       16, meta["isect_offsets"], meta["flatten_ids"])
        print(v, "I", meta["flatten_ids"].numel(), "P_eval", pe, "P_blend", pb, Wr.blend3d_pair_stats.last_cull)
