"""Development aid: configs[2] (1M surfels) window -- forward deviation of the fast 2DGS blend: cull or arithmetic?"""
import math
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import horizongs_b200 as hgs
from horizongs_b200 import scenes
from oracle import gsplat_oracle as O

torch.set_num_threads(os.cpu_count())
view = sys.argv[1] if len(sys.argv) > 1 else "aerial"
sc, V, Ks, Wd, Hd = scenes.config1(n=1_000_000, view=view)
w, h = 384, 256
K2 = Ks.clone()
K2[0, 0, 2] -= Wd // 2 - w // 2
K2[0, 1, 2] -= Hd // 2 - h // 2
radii, m2, d, rt, nrm = O.fully_fused_projection_2dgs(sc.means, sc.quats, sc.scales, V, None, K2, w, h)
got = hgs.fully_fused_projection_2dgs(sc.means.cuda(), sc.quats.cuda(), sc.scales.cuda(), V.cuda(), None, K2.cuda(), w, h)
bad = torch.nonzero(got[0].cpu() != radii)
print("radii mismatches:", bad.shape[0], "of visible", int((radii > 0).sum()))
for b in bad[:5]:
    k = int(b[1])
    print("   id", k, "cuda", int(got[0][0, k]), "oracle", int(radii[0, k]), "depth", float(d[0, k]), "scales", sc.scales[k].tolist())
for name, g_, r_ in zip(("means2d", "depths", "ray_transforms", "normals"), got[1:], (m2, d, rt, nrm)):
    print("   ", name, "max abs diff", float((g_.cpu() - r_).abs().max()), "n differing", int((g_.cpu() != r_).sum()))
tw, th = math.ceil(w / 16), math.ceil(h / 16)
_, ids, flat = O.isect_tiles(m2, radii, d, 16, tw, th)
off = O.isect_offset_encode(ids, 1, tw, th)
op = sc.opacities[None].contiguous()
cols3 = sc.colors[None].contiguous()
ref = O.rasterize_to_pixels_2dgs(m2, rt, cols3, op, nrm, w, h, 16, off, flat)
fast = hgs.rasterize_to_pixels_2dgs(m2.cuda(), rt.cuda(), cols3.cuda(), op.cuda(), nrm.cuda(), None, w, h, 16, off.cuda(), flat.cuda())
plain = hgs.rasterize_to_pixels_2dgs(m2.cuda(), rt.cuda(), cols3[..., :2].contiguous().cuda(), op.cuda(), nrm.cuda(), None, w, h, 16,
                                     off.cuda(), flat.cuda())
ref64 = O.rasterize_to_pixels_2dgs(m2.double(), rt.double(), cols3.double(), op.double(), nrm.double(), w, h, 16, off, flat)
for name, i in (("colors", 0), ("alphas", 1), ("normals", 2)):
    f, p, r, r64 = fast[i].cpu(), plain[i].cpu(), ref[i], ref64[i]
    if name == "colors":
        p, r2, r642 = p, r[..., :2], r64[..., :2]
        print(name, "fast-vs-f32", float((f - r).abs().max()), "plain-vs-f32", float((p - r2).abs().max()),
              "f32-vs-f64", float((r.double() - r64).abs().max()), "fast-vs-f64", float((f.double() - r64).abs().max()))
    else:
        print(name, "fast-vs-f32", float((f - r).abs().max()), "plain-vs-f32", float((p - r).abs().max()),
              "f32-vs-f64", float((r.double() - r64).abs().max()), "fast-vs-f64", float((f.double() - r64).abs().max()),
              "plain-vs-f64", float((p.double() - r64).abs().max()))
e = (fast[1].cpu() - ref[1]).abs()[0, ..., 0]
k = int(e.argmax())
print("worst alpha pixel", divmod(k, w), "err", float(e.flatten()[k]), "n pixels with err > 1e-4:", int((e > 1e-4).sum()))
