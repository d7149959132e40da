"""Development aid: where do the gradient deviations of configs[0] (100k Gaussians, 256x256) come from?"""
import math
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import horizongs_b200 as hgs
from horizongs_b200 import scenes
from oracle import gsplat_oracle as O

n = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000
sc, V, Ks, W, H = scenes.config0(n=n)
torch.set_num_threads(os.cpu_count())
g = torch.Generator().manual_seed(5)
w_rc, w_ra = torch.rand(1, H, W, 4, generator=g), torch.rand(1, H, W, 1, generator=g)


def rel(a, b):
    return float((a - b).abs().max() / (b.abs().max() + 1e-20))


# ---- blend stage alone, oracle intermediates as inputs (float32 and float64 references)
radii, m2, d, con, _ = O.fully_fused_projection(sc.means, None, sc.quats, sc.scales, V, Ks, W, H)
tw, th = math.ceil(W / 16), math.ceil(H / 16)
_, ids, flat = O.isect_tiles(m2, radii, d, 16, tw, th)
off = O.isect_offset_encode(ids, 1, tw, th)
cols = torch.cat([sc.colors[None], d[..., None]], -1).contiguous()
op = sc.opacities[None].contiguous()
res = {}
for dt in (torch.float32, torch.float64):
    ins = [t.clone().to(dt).requires_grad_() for t in (m2, con, cols, op)]
    rc, ra = O.rasterize_to_pixels(*ins, W, H, 16, off, flat)
    loss = (rc * w_rc.to(dt)).sum() + (ra * w_ra.to(dt)).sum()
    res[dt] = (rc.detach(), torch.autograd.grad(loss, ins))
cins = [t.cuda().requires_grad_() for t in (m2, con, cols, op)]
crc, cra = hgs.rasterize_to_pixels(*cins, W, H, 16, off.cuda(), flat.cuda())
got = torch.autograd.grad((crc * w_rc.cuda()).sum() + (cra * w_ra.cuda()).sum(), cins)
print("blend stage: image err vs f32", float((crc.detach().cpu() - res[torch.float32][0]).abs().max()),
      "vs f64", float((crc.detach().cpu().double() - res[torch.float64][0]).abs().max()),
      " f32 vs f64", float((res[torch.float32][0].double() - res[torch.float64][0]).abs().max()))
for i, name in enumerate(("v_means2d", "v_conics", "v_colors", "v_opacities")):
    r32, r64, gg = res[torch.float32][1][i], res[torch.float64][1][i], got[i].cpu()
    print(f"  {name:12s} cuda-vs-f32 {rel(gg, r32):.2e}  cuda-vs-f64 {rel(gg.double(), r64):.2e}  f32-vs-f64 {rel(r32.double(), r64):.2e}")
    if name == "v_means2d":
        e = (gg.double() - r64).abs().amax(-1)[0]
        k = int(e.argmax())
        print("    worst Gaussian", k, "cuda", gg[0, k].tolist(), "f32", r32[0, k].tolist(), "f64", r64[0, k].tolist(),
              "radius", int(radii[0, k]), "opacity", float(op[0, k]), "conic", con[0, k].tolist())
        print("    Gaussians with rel err > 1e-3 (vs f64):", int((e > 1e-3 * float(r64.abs().max())).sum()), "of",
              int((radii > 0).sum()))
