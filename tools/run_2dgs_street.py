"""Development aid: the configs[2] street view alone (fwd + bwd), for ncu captures of the 2DGS kernels."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import horizongs_b200 as hgs
from horizongs_b200 import scenes

view = sys.argv[1] if len(sys.argv) > 1 else "street"
sc, V, Ks, W, H = scenes.config1(n=1_000_000, view=view)
sc = sc.to("cuda")
params = [t.requires_grad_() for t in (sc.means, sc.quats, sc.scales, sc.opacities, sc.colors)]
for it in range(3):
    (rc, ra, rn, rnd, rd, rm), meta = hgs.rasterization_2dgs(*params, V.cuda(), Ks.cuda(), W, H, render_mode="RGB+ED")
    (rc.mean() + ra.mean() + 0.05 * (rn * rnd).sum(-1).mean()).backward()
    for p in params:
        p.grad = None
torch.cuda.synchronize()
off = meta["isect_offsets"].flatten()
d = torch.diff(torch.cat([off, off.new_tensor([meta["flatten_ids"].numel()])]))
print("I", meta["flatten_ids"].numel(), "n_vis", int((meta["radii"] > 0).sum()), "max depth", int(d.max()), "mean", float(d.float().mean()))
# ---- how deep does the forward go?  (alpha saturation and termination depth per tile)
with torch.no_grad():
    (rc, ra, rn, rnd, rd, rm), meta = hgs.rasterization_2dgs(*params, V.cuda(), Ks.cuda(), W, H, render_mode="RGB+ED")
    a = ra[0, ..., 0]
    print("alpha: mean", float(a.mean()), "frac > 0.9999", float((a > 0.9999).float().mean()), "frac < 0.5", float((a < 0.5).float().mean()))
    # per 16x16 tile: does every pixel saturate?
    th, tw = (H + 15) // 16, (W + 15) // 16
    pad = torch.zeros(th * 16, tw * 16, device="cuda")
    pad[:H, :W] = (a > 0.9999).float()
    pad[H:, :] = 1.0
    tiles_sat = pad.reshape(th, 16, tw, 16).permute(0, 2, 1, 3).reshape(th, tw, 256).min(-1).values
    print("tiles where every pixel saturates:", float(tiles_sat.mean()))
    for r in range(0, th, 8):
        print("row", r, "alpha mean", [round(float(a[r * 16:(r + 1) * 16, c * 16:(c + 1) * 16].mean()), 3) for c in range(0, tw, 20)])
