"""Per-stage CUDA-event timings of the hot path on the named synthetic configs (development aid)."""
import argparse
import json
import math
import sys
import os

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import horizongs_b200 as hgs
from horizongs_b200 import scenes
from horizongs_b200.cuda import _wrapper as W
from horizongs_b200 import rendering as R


def timed(fn, iters=10, warm=3):
    for _ in range(warm):
        out = fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return out, ts[len(ts) // 2]


def run(name, sc, V, Ks, Wd, H):
    dev = "cuda"
    sc = sc.to(dev)
    V, Ks = V.to(dev), Ks.to(dev)
    C, N = V.shape[0], sc.n
    tw, th = math.ceil(Wd / 16), math.ceil(H / 16)
    res = {"config": name, "N": N, "C": C, "W": Wd, "H": H}
    proj, t = timed(lambda: W._project3d(sc.means, sc.quats, sc.scales, V, Ks, Wd, H, 0.3, 0.01, 1e10, 0.0, False, 16))
    res["project_fwd_ms"] = t
    radii, m2, d, con, _, tiles = proj
    res["n_visible"] = int((radii > 0).sum())
    (ids, flat, off, vis_ids), t = timed(lambda: W._isect_sorted_from_counts(m2, radii, d, tiles, C, N, 16, tw, th))
    res["isect_sort_ms"] = t
    res["I"] = int(ids.numel())
    res["max_tile_depth"] = int((torch.diff(torch.cat([off.flatten(), off.new_tensor([ids.numel()])]))).max())
    if sc.sh_degree is not None:
        campos = R._camera_positions(V)
        cols, t = timed(lambda: W._sh_view_colors(sc.sh_degree, sc.means, campos, sc.colors, radii))
        res["sh_fwd_ms"] = t
    else:
        cols = sc.colors[None]
    op = sc.opacities[None]
    (rc, ra), t = timed(lambda: W._blend3d(m2, con, cols, d, op, None, Wd, H, 16, off, flat))
    res["blend_fwd_ms"] = t
    # backward of the blend stage alone
    ins = [x.detach().clone().requires_grad_() for x in (m2, con, cols.contiguous(), d, op)]
    rc, ra = W._blend3d(ins[0], ins[1], ins[2], ins[3], ins[4], None, Wd, H, 16, off, flat)
    g1, g2 = torch.rand_like(rc), torch.rand_like(ra)
    _, t = timed(lambda: torch.autograd.grad((rc, ra), ins, (g1, g2), retain_graph=True))
    res["blend_bwd_ms"] = t
    # whole pipeline
    params = [x.detach().clone().requires_grad_() for x in (sc.means, sc.quats, sc.scales, sc.opacities, sc.colors)]

    def fwd():
        return hgs.rasterization(*params, V, Ks, Wd, H, sh_degree=sc.sh_degree, render_mode="RGB+ED")

    with torch.no_grad():
        _, t = timed(fwd)
    res["pipeline_fwd_ms"] = t

    def fwdbwd():
        rc, ra, meta = fwd()
        (rc * g1).sum().backward()
        for p in params:
            p.grad = None

    _, t = timed(fwdbwd)
    res["pipeline_fwd_bwd_ms"] = t
    print(json.dumps(res), flush=True)
    return res


def run2d(name, sc, V, Ks, Wd, H, distloss):
    """config 2: 2DGS pipeline (rasterization_2dgs), forward and forward+backward."""
    dev = "cuda"
    sc = sc.to(dev)
    V, Ks = V.to(dev), Ks.to(dev)
    params = [x.detach().clone().requires_grad_() for x in (sc.means, sc.quats, sc.scales, sc.opacities, sc.colors)]
    res = {"config": name, "N": sc.n, "W": Wd, "H": H, "distloss": distloss}

    def fwd():
        return hgs.rasterization_2dgs(*params, V, Ks, Wd, H, sh_degree=sc.sh_degree, render_mode="RGB+ED",
                                      distloss=distloss)

    with torch.no_grad():
        (outs, meta), t = timed(fwd)
    res["pipeline_fwd_ms"] = t
    res["I"] = int(meta["flatten_ids"].numel())
    res["n_visible"] = int((meta["radii"] > 0).sum())
    g = [torch.rand_like(o) for o in outs[:3]]

    def fwdbwd():
        (rc, ra, rn, rnd, rd, rm), meta = fwd()
        loss = (rc * g[0]).sum() + (ra * g[1]).sum() + 0.05 * (rn * g[2]).sum() + 0.05 * (1 - (rn * rnd).sum(-1)).mean()
        if distloss:
            loss = loss + 0.01 * rd.mean()
        loss.backward()
        for p in params:
            p.grad = None

    _, t = timed(fwdbwd)
    res["pipeline_fwd_bwd_ms"] = t
    print(json.dumps(res), flush=True)


def run_lod(n_anchors=500_000, fused_decode=False, standard_dist=26.686):
    """config 3: LOD anchor model through the reference adapter's control flow (tests/lod_harness.py): anchor
    mask + prefilter (fully_fused_projection) + MLP decode (PyTorch, as in the reference) + rasterization."""
    from tests import lod_harness as LH
    dev = "cuda"
    Wd, H = 1920, 1080
    model = LH.TinyAnchorModel(n_anchors=n_anchors, levels=4, extent=25.0, voxel0=0.12, standard_dist=standard_dist).to(dev)
    model.level = model.level.to(dev)
    V = scenes.aerial_camera(12.0, 45.0, 30.0, (2.0, -3.0)).to(dev)
    Km = scenes.intrinsics(Wd, H).to(dev)
    bg = torch.zeros(3, device=dev)
    res = {"config": "config3-lod-aerial", "anchors": n_anchors, "W": Wd, "H": H, "fused_decode": fused_decode,
           "standard_dist": standard_dist}
    marks = []
    W.set_stage_hook(lambda name, ph: marks.append((name, ph, _ev())))

    def _ev():
        e = torch.cuda.Event(enable_timing=True)
        e.record()
        return e

    def fwd():
        return LH.render(model, V, Km, Wd, H, bg, hgs, fused_decode=fused_decode)

    with torch.no_grad():
        o, t = timed(fwd)
    res["render_fwd_ms"] = t
    res["n_gaussians"] = int(o["n_gaussians"])
    res["n_visible_anchors"] = int(o["visible_mask"].sum())

    def fwdbwd():
        o = fwd()
        (o["render"].mean() + 0.1 * o["render_depth"].mean()).backward()
        model.zero_grad(set_to_none=True)

    marks.clear()
    _, t = timed(fwdbwd, iters=5, warm=2)
    res["render_fwd_bwd_ms"] = t
    W.set_stage_hook(None)
    torch.cuda.synchronize()
    open_, agg = {}, {}
    for name, ph, ev in marks:
        if ph == 0:
            open_[name] = ev
        elif name in open_:
            agg.setdefault(name, []).append(open_.pop(name).elapsed_time(ev))
    res["rasterizer_stage_ms"] = {k: round(sum(v) / len(v), 4) for k, v in agg.items()}
    res["rasterizer_total_ms"] = round(sum(res["rasterizer_stage_ms"].values()), 4)
    print(json.dumps(res), flush=True)


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--configs", default="1a,1s,4")
    a = ap.parse_args()
    for c in a.configs.split(","):
        if c == "0":
            sc, V, Ks, Wd, H = scenes.config0()
            run("config0", sc, V, Ks, Wd, H)
        elif c == "1a":
            run("config1-aerial", *scenes.config1(view="aerial"))
        elif c == "1s":
            run("config1-street", *scenes.config1(view="street"))
        elif c == "2":
            for view in ("aerial", "street"):
                for dl in (False, True):
                    run2d(f"config2-2dgs-{view}", *scenes.config1(view=view), dl)
        elif c == "3":
            for sd in (26.686, 200.0):          # the reference's standard_dist, and one that keeps most anchors
                run_lod(fused_decode=False, standard_dist=sd)
                run_lod(fused_decode=True, standard_dist=sd)
        elif c == "4":
            sc, V, Ks, Wd, H = scenes.config4(n_views=2)
            run("config4-view0-aerial", sc, V[:1], Ks[:1], Wd, H)
            run("config4-view1-street", sc, V[1:2], Ks[1:2], Wd, H)
