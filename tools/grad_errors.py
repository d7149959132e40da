"""Development aid: gradient errors of the 2DGS kernels against the float32 AND the float64 oracle (same integer
stage inputs), per tensor -- tells whether a deviation above 1e-3 is the kernel's or the float32 oracle's own."""
import math
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import horizongs_b200 as hgs
from oracle import gsplat_oracle as O
from tests.helpers import rel_err, small_scene


def _rand_like(t, seed):
    return torch.rand(t.shape, generator=torch.Generator().manual_seed(seed))


def blend2d_stage(D, distloss, n=2500, scale=0.15, seed=0):
    sc, V, Ks, Wd, H = small_scene(n=n, width=128, height=96, scale=scale, seed=seed)
    radii, m2, d, rt, nrm = O.fully_fused_projection_2dgs(sc.means, sc.quats, sc.scales, V, None, Ks, Wd, H)
    tw, th = math.ceil(Wd / 16), math.ceil(H / 16)
    _, ids, flat = O.isect_tiles(m2, radii, d, 16, tw, th)
    off = O.isect_offset_encode(ids, 1, tw, th)
    cols = torch.cat([sc.colors[None], d[..., None]], -1)[..., 4 - D:].contiguous()
    op = sc.opacities[None].contiguous()
    names = ("v_means2d", "v_ray_transforms", "v_colors", "v_opacities", "v_normals")
    res = {}
    ws = None
    for dt in (torch.float32, torch.float64):
        ins = [t.clone().to(dt).requires_grad_() for t in (m2, rt, cols, op, nrm)]
        outs = O.rasterize_to_pixels_2dgs(ins[0], ins[1], ins[2], ins[3], ins[4], Wd, H, 16, off, flat, distloss=distloss)
        if ws is None:
            ws = [_rand_like(o, 30 + i) for i, o in enumerate(outs)]
        loss = sum((o * w.to(dt)).sum() for o, w in zip(outs, ws))
        res[dt] = torch.autograd.grad(loss, ins, allow_unused=True)
    cins = [t.cuda().requires_grad_() for t in (m2, rt, cols, op, nrm)]
    couts = hgs.rasterize_to_pixels_2dgs(cins[0], cins[1], cins[2], cins[3], cins[4], None, Wd, H, 16, off.cuda(),
                                         flat.cuda(), distloss=distloss)
    got = torch.autograd.grad(sum((o * w.cuda()).sum() for o, w in zip(couts, ws)), cins, allow_unused=True)
    print(f"blend2d stage D={D} distloss={distloss} n={n} seed={seed}")
    for i, name in enumerate(names):
        r32, r64, g = res[torch.float32][i], res[torch.float64][i], got[i]
        if r32 is None:
            continue
        print(f"  {name:18s} cuda-vs-f32 {rel_err(g.cpu(), r32):.2e}   cuda-vs-f64 {rel_err(g.cpu().double(), r64):.2e}"
              f"   f32-vs-f64 {rel_err(r32.double(), r64):.2e}")


def pipeline2d(distloss, seed=0):
    sc, V, Ks, Wd, H = small_scene(n=3000, width=144, height=96, scale=0.15, seed=seed)
    ins = [t.clone().requires_grad_() for t in (sc.means, sc.quats, sc.scales, sc.opacities, sc.colors)]
    (rc, ra, rn, rnd, rd, rm), meta = O.rasterization_2dgs(*ins, V, Ks, Wd, H, render_mode="RGB+ED", distloss=distloss)
    outs = [rc, ra, rn, rnd, rd, rm]
    ws = [_rand_like(o, 20 + i) for i, o in enumerate(outs)]
    names = ("colors", "alphas", "normals", "normals_from_depth", "distort", "median")
    print(f"2dgs pipeline distloss={distloss}: gradient of each output term separately (cuda vs f32 oracle)")
    cins = [t.cuda().requires_grad_() for t in (sc.means, sc.quats, sc.scales, sc.opacities, sc.colors)]
    (crc, cra, crn, crnd, crd, crm), cmeta = hgs.rasterization_2dgs(*cins, V.cuda(), Ks.cuda(), Wd, H,
                                                                    render_mode="RGB+ED", distloss=distloss)
    couts = [crc, cra, crn, crnd, crd, crm]
    for k, name in enumerate(names):
        if not outs[k].requires_grad:
            continue
        ref = torch.autograd.grad((outs[k] * ws[k]).sum(), ins, retain_graph=True, allow_unused=True)
        got = torch.autograd.grad((couts[k] * ws[k].cuda()).sum(), cins, retain_graph=True, allow_unused=True)
        row = []
        for pn, g, r in zip(("means", "quats", "scales", "opac", "colors"), got, ref):
            if r is None or g is None:
                row.append(f"{pn} -")
            else:
                row.append(f"{pn} {rel_err(g.cpu(), r):.1e}")
        print(f"  {name:20s} " + "  ".join(row))
    ref = torch.autograd.grad(sum((o * w).sum() for o, w in zip(outs, ws)), ins, allow_unused=True)
    got = torch.autograd.grad(sum((o * w.cuda()).sum() for o, w in zip(couts, ws)), cins, allow_unused=True)
    print("  ALL                  " + "  ".join(f"{pn} {rel_err(g.cpu(), r):.1e}" for pn, g, r in
                                                 zip(("means", "quats", "scales", "opac", "colors"), got, ref)))


if __name__ == "__main__":
    for D, dl in ((3, False), (4, True), (2, True)):
        blend2d_stage(D, dl)
    blend2d_stage(4, False, n=6000, scale=0.1, seed=2)
    for dl in (False, True):
        pipeline2d(dl)
