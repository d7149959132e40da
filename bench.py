#!/usr/bin/env python
"""bench.py -- fwd+bwd iterations/s of the rasterization hot path on BASELINE.json's workloads.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config 0|1|2|3|4]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...   (N > 1)

Default workload (config.workload): BASELINE.json configs[4] = the configuration its metric is quoted on -- 6M
explicit SH2 Gaussians, 1920x1080, render_mode RGB+ED, 8 seeded cameras (aerial / street alternating).  One "step"
= one view per GPU: rasterization forward, an L1-style loss, backward to the 38 floats of every Gaussian, the
densification statistics, and for N > 1 the gradient exchange (default: the per-Gaussian backward fused with the
exchange over NVLink peer memory, csrc/exchange_vjp.cu; `--exchange nccl` = dense NCCL all-reduce).  For N > 1 the
views follow the cost-bucketed schedule where it applies (2 and 4 GPUs; `--view-schedule bucketed`: a step holds views
of one kind, every rank renders one aerial and one street view per two steps); with 8 GPUs a step is the whole 8-view
batch, one view per rank (`interleaved`: rank r renders view (r + step) mod 8).  Every rank sees every view either way
("weak" scaling: per-GPU work fixed).
`--config 0..3` time the other named configurations (parity-test cases of the contract, measured for completeness):
0 = 100k / 256x256 (the CPU-runnable case, full frame on both arms), 1 = 1M 3DGS 1080p aerial + street, 2 = 1M 2DGS
surfels with the normal-consistency loss, 3 = LOD anchor model (500k anchors x 10) through the adapter control flow.

One JSON line on stdout (rank 0).  `value` = views/s with everything resident in HBM; `e2e` = the same through
the public API with the step's camera and ground-truth image copied from pinned host memory and the loss read
back, inside the timed region.  `roofline` describes the dominant kernel (blend backward, FP32-issue bound),
`roofline_hbm` the dominant HBM-bound stage; `cpu_baseline` is the CPU oracle on a bounded, FIXED sample (384x256
centre window of views 0 and 7; measured seconds and the full-frame extrapolation are separate fields);
`parity_check` compares the CUDA path with the oracle on those same windows of the benchmarked scene (integer stages
bit-exact, image 1e-4, gradients 1e-3) and, for N > 1, `exchange_parity` compares the fused exchange with the dense
all-reduce of every rank's autograd gradients.
`--impl reference` times the CPU oracle port (the reference's own rasterizer, gsplat, is not installable here: see
DESIGN.md) on the same bounded sample, rank 0 only.
"""
from __future__ import annotations

import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402

LOSS_MODE = "l1"
SAMPLE_WINDOW = (384, 256)          # the bounded CPU sample: centre window of the 1080p frame, fixed
SAMPLE_VIEWS = {4: (0, 7), 1: (0, 1), 2: (0, 1)}   # one aerial and one street view (view 7 of configs[4] has the deepest tiles)
LAMBDA_NORMAL = 0.05                # config/our_2d: lambda_normal (train.py:180-188)


def log(*a):
    print(*a, file=sys.stderr, flush=True)


# ------------------------------------------------------------------------------------------------- workloads
class Workload:
    """scene + cameras + ground truth of one BASELINE.json config (CPU tensors; callers move them)"""

    def __init__(self, config: int, n_gauss: int | None, n_views: int = 8):
        from horizongs_b200 import scenes
        t0 = time.time()
        self.config = config
        self.kind = "2dgs" if config == 2 else ("lod" if config == 3 else "3dgs")
        if config == 4:
            n = n_gauss or 6_000_000
            self.sc, self.views, self.Ks, self.W, self.H = scenes.config4(n=n, n_views=n_views)
            self.name = (f"configs[4]: {n / 1e6:g}M explicit SH2 3DGS Gaussians, {self.W}x{self.H}, RGB+ED, "
                         "1 view per GPU per step")
            self.metric = f"fwd+bwd iters/s, {n / 1e6:g}M Gaussians @{self.W}x{self.H}"
        elif config in (1, 2):
            n = n_gauss or 1_000_000
            sc, va, Ks, self.W, self.H = scenes.config1(n=n, view="aerial")
            _, vs, _, _, _ = scenes.config1(n=16, view="street")
            self.sc, self.views, self.Ks = sc, torch.cat([va, vs], 0), Ks.expand(2, -1, -1).contiguous()
            what = "3DGS Gaussians" if config == 1 else "2DGS surfels (normal-consistency loss, distortion off)"
            self.name = (f"configs[{config}]: {n / 1e6:g}M {what}, {self.W}x{self.H}, RGB+ED, aerial + street views "
                         "alternating, 1 view per GPU per step")
            self.metric = f"fwd+bwd iters/s, {n / 1e6:g}M {'Gaussians' if config == 1 else 'surfels'} @{self.W}x{self.H}"
        elif config == 0:
            n = n_gauss or 100_000
            self.sc, self.views, self.Ks, self.W, self.H = scenes.config0(n=n)
            self.name = f"configs[0]: {n / 1e3:g}k 3DGS Gaussians, one {self.W}x{self.H} camera, RGB+ED (full frame on both arms)"
            self.metric = f"fwd+bwd iters/s, {n / 1e3:g}k Gaussians @{self.W}x{self.H}"
        elif config == 3:
            from tests import lod_harness as LH
            n = n_gauss or 500_000
            self.W, self.H = 1920, 1080
            self.model = LH.TinyAnchorModel(n_anchors=n, levels=4, extent=25.0, voxel0=0.12, standard_dist=26.686)
            self.views = torch.stack([scenes.aerial_camera(12.0, 45.0, 30.0, (2.0, -3.0)),
                                      scenes.street_camera(0.3, 40.0, (1.0, 2.0))], 0)
            self.Ks = scenes.intrinsics(self.W, self.H)[None].expand(2, -1, -1).contiguous()
            self.sc = None
            self.name = (f"configs[3]: LOD anchor model, {n / 1e3:g}k anchors x 10 neural Gaussians, {self.W}x{self.H}, "
                         "level mask + prefilter + fused decode + rasterization through the adapter control flow")
            self.metric = f"fwd+bwd iters/s, {n / 1e3:g}k anchors x10 @{self.W}x{self.H}"
        else:
            raise ValueError(config)
        self.n_views = self.views.shape[0]
        g = torch.Generator().manual_seed(7)
        self.gts = torch.rand(self.n_views, self.H, self.W, 3, generator=g)        # synthetic ground-truth images
        self.sh_degree = None if self.sc is None else self.sc.sh_degree
        log(f"[bench] workload {self.name!r} generated in {time.time() - t0:.1f}s")

    @property
    def n(self):
        return self.sc.n if self.sc is not None else self.model.anchor.shape[0]


def loss_fn(rc, ra, gt, extra=None):
    """L1 photometric term + small depth / alpha terms (every output channel gets a gradient).  On the GPU the
    same expression is one fused forward and one fused backward kernel (horizongs_b200.losses, csrc/loss.cu).
    extra = (render_normals, normals_from_depth) of a 2DGS render: + the normal-consistency term of train.py:180-188."""
    if rc.is_cuda:
        from horizongs_b200 import losses
        if LOSS_MODE == "l1ssim":       # the reference's full photometric loss (train.py:158-160), fused
            loss = losses.photometric_loss(rc, gt, 0.2, ra, w_depth=0.01, w_alpha=0.01)
        else:
            loss = losses.photometric_l1_loss(rc, gt, ra, w_depth=0.01, w_alpha=0.01)
    else:
        loss = (rc[..., :3] - gt).abs().mean() + 0.01 * rc[..., 3].mean() + 0.01 * ra.mean()
    if extra is not None:
        rn, rnd = extra
        nfd = rnd.reshape(rn.shape) * ra.detach()
        loss = loss + LAMBDA_NORMAL * (1.0 - (rn * nfd).sum(-1)).mean()
    return loss


def exchange_cap_rows(n_gaussians: int) -> int:
    """rows per rank of the peer-memory mailboxes (N > 1).  N rows: a view can never overflow them (2 x world x N x 64 B
    = 6.1 GB of 180 GB at 8 GPUs and 6 M Gaussians).  N // 4 was enough for views 0..7 (at most 1.02 M of 6 M visible)
    but not for view 11 of a 16-view set tried at 8 GPUs (1.68 M visible; tests/test_distributed_cpu.py counts them): the
    exchange then reported HGS_EX_OVERFLOW on every rank, as designed."""
    return int(n_gaussians)


def view_schedule(step: int, rank: int, n_views: int, bucketed: bool) -> int:
    """index of the view `rank` renders at `step` (views alternate aerial = even index / street = odd index).
    bucketed: horizongs_b200.distributed.bucketed_view (one kind of view per step); interleaved: (rank + step) % n_views.
    With one rank both are the sequence 0, 1, 2, ..."""
    if bucketed:
        from horizongs_b200.distributed import bucketed_view
        return bucketed_view(step, rank, n_views)
    return (rank + step) % n_views


def render_explicit(backend, kind, params, view, Km, W, H, sh_degree, bg):
    """-> (render_colors, render_alphas, meta, extra) through the gsplat-named pipeline of `backend`"""
    kw = dict(sh_degree=sh_degree, render_mode="RGB+ED", backgrounds=bg, packed=False)
    if kind == "2dgs":
        (rc, ra, rn, rnd, _, _), meta = backend.rasterization_2dgs(*params, view, Km, W, H, **kw)
        return rc, ra, meta, (rn, rnd)
    rc, ra, meta = backend.rasterization(*params, view, Km, W, H, **kw)
    return rc, ra, meta, None


# ------------------------------------------------------------------------------------------------- clocks
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.proc = None
        self.lines = []
        self.gpu = gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100", "-i",
                 str(self.gpu)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._pump, daemon=True)
            self.t.start()
        except Exception as e:  # nvidia-smi missing
            log(f"[bench] clock sampling unavailable: {e}")
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, smax, power, reasons = [], [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); smax.append(float(f[2])); power.append(float(f[3]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": max(smax), "power_w_max": max(power),
                "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------------- CPU oracle arm
def centre_window(wl: Workload):
    if wl.W <= SAMPLE_WINDOW[0] or wl.H <= SAMPLE_WINDOW[1]:
        return (0, 0, wl.W, wl.H)                                   # configs[0]: the full frame, nothing extrapolated
    w, h = SAMPLE_WINDOW
    return (wl.W // 2 - w // 2, wl.H // 2 - h // 2, w, h)


def window_K(K, win):
    K2 = K.clone()
    K2[0, 2] -= win[0]
    K2[1, 2] -= win[1]
    return K2


def window_gt(win):
    return torch.rand(1, win[3], win[2], 3, generator=torch.Generator().manual_seed(3))


def parity_loss(rc, ra, extra, win):
    """a SMOOTH scalar with a distinct seeded weight per pixel and channel: gradient parity must not
    hinge on sign(render - gt) of an L1 term flipping where the two images differ in the last bits.  The depth-derived
    normals (finite differences, discontinuous) stay out of it."""
    g = torch.Generator().manual_seed(11)
    w1 = torch.rand(rc.shape, generator=g).to(rc.device)
    w2 = torch.rand(ra.shape, generator=g).to(rc.device)
    loss = (rc * w1).mean() + (ra * w2).mean()
    if extra is not None:
        w3 = torch.rand(extra[0].shape, generator=g).to(rc.device)
        loss = loss + (extra[0] * w3).mean()
    return loss


def oracle_window_step(wl: Workload, v: int, win, threads, keep=False):
    """fwd+bwd of the CPU oracle on the window (x0,y0,w,h) of view v (a sub-frustum: same Gaussians, same camera,
    principal point shifted).  -> dict(seconds of the step, seconds of its per-Gaussian part, isects, and -- keep --
    the outputs / meta / gradients for the parity check)"""
    from oracle import gsplat_oracle as O
    torch.set_num_threads(threads)
    sc = wl.sc
    K2 = window_K(wl.Ks[v], win)
    w, h = win[2], win[3]
    src = (sc.means, sc.quats, sc.scales, sc.opacities, sc.colors)
    params = [t.clone().requires_grad_() for t in src]
    gt = window_gt(win)
    bg = torch.zeros(1, 3)
    t0 = time.perf_counter()
    rc, ra, meta, extra = render_explicit(O, wl.kind, params, wl.views[v][None], K2[None], w, h, sc.sh_degree, bg)
    if keep:
        meta["means2d"].retain_grad()
    loss = parity_loss(rc, ra, extra, win) if keep else loss_fn(rc, ra, gt, extra)
    loss.backward()
    t_step = time.perf_counter() - t0
    out = {"t_step": t_step, "n_isects": int(meta["flatten_ids"].numel())}
    if keep:
        out.update(rc=rc.detach(), ra=ra.detach(), meta=meta, grads=[p.grad for p in params],
                   extra=None if extra is None else [e.detach() for e in extra], loss=float(loss.detach()))
    # the per-Gaussian part (projection + SH, fwd+bwd) does not shrink with the window: time it alone
    params = [t.clone().requires_grad_() for t in src]
    t0 = time.perf_counter()
    if wl.kind == "2dgs":
        radii, m2, d, rt, nrm = O.fully_fused_projection_2dgs(params[0], params[1], params[2], wl.views[v][None], None,
                                                              K2[None], w, h)
        geo = rt.sum() + nrm.sum()
    else:
        radii, m2, d, con, _ = O.fully_fused_projection(params[0], None, params[1], params[2], wl.views[v][None],
                                                        K2[None], w, h)
        geo = con.sum()
    cols = O._view_colors(params[0], params[4], wl.views[v][None], radii, sc.sh_degree)
    (m2.sum() + d.sum() + geo + cols.sum()).backward()
    out["t_gauss"] = time.perf_counter() - t0
    scale = (wl.W * wl.H) / float(w * h)
    out["t_full_est"] = out["t_gauss"] + max(t_step - out["t_gauss"], 1e-6) * scale
    out["scale"] = scale
    return out


def cpu_sample_description(wl, win, views, threads):
    if win[2] == wl.W and win[3] == wl.H:
        return f"the full {wl.W}x{wl.H} frame of view(s) {list(views)}, nothing extrapolated"
    return (f"{win[2]}x{win[3]} centre window (fixed) of the {wl.W}x{wl.H} frame of views {list(views)}, all {wl.n} "
            f"Gaussians projected; full-frame estimate = per-Gaussian seconds + per-pixel seconds x "
            f"{wl.W * wl.H / (win[2] * win[3]):.2f}; measured and estimated seconds are separate fields")


def run_reference(args):
    """--impl reference: the CPU oracle port on the bounded sample, rank 0 only."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    wl = Workload(args.config, args.gaussians)
    if wl.kind == "lod":
        print(json.dumps({"impl": "reference", "unavailable": "configs[3] has no CPU arm (the decode stays PyTorch; "
                          "its rasterizer part is covered by configs[1])"}), flush=True)
        return
    win = centre_window(wl)
    views = SAMPLE_VIEWS.get(args.config, (0,))
    t_all = time.time()
    rows = []
    budget = 200.0
    for s in range(args.warmup + args.steps):
        v = views[s % len(views)]
        r = oracle_window_step(wl, v, win, threads)
        log(f"[reference] step {s}: view {v} window {r['t_step']:.2f}s (per-Gaussian part {r['t_gauss']:.2f}s) -> "
            f"full-frame estimate {r['t_full_est']:.1f}s ({r['n_isects']} isects)")
        if s >= args.warmup:
            rows.append((v, r))
        if time.time() - t_all > budget and len({v for v, _ in rows}) >= len(views):
            log("[reference] time budget reached; stopping early")
            break
    # mean over the sample views (each view's repetitions averaged first, so an odd count does not skew it)
    per_view = {}
    for v, r in rows:
        per_view.setdefault(v, []).append(r)
    est = sum(sum(x["t_full_est"] for x in rs) / len(rs) for rs in per_view.values()) / len(per_view)
    meas = sum(sum(x["t_step"] for x in rs) / len(rs) for rs in per_view.values()) / len(per_view)
    value = 1.0 / est
    line = {
        "impl": "reference", "metric": wl.metric, "value": value, "unit": "iters/s", "n_gpus": args.gpus,
        "steps": len(rows), "warmup": args.warmup, "ms_per_step": 1e3 * est, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": wl.name, "views": wl.n_views, "render_mode": "RGB+ED", "sh_degree": wl.sh_degree,
                   "note": "gsplat (the reference's rasterizer) is not installable here; this is the repo's CPU oracle "
                           "port of it (torch, float32).  value / ms_per_step are the FULL-FRAME estimate; "
                           "measured_s_per_sample_step is what one timed step actually took"},
        "cpu_baseline": {"value": value, "unit": "iters/s", "cores": threads, "kind": "port",
                         "sample": cpu_sample_description(wl, win, views, threads),
                         "measured_s_per_sample_step": meas, "estimated_s_per_full_frame": est,
                         "extrapolation_factor_pixels": rows[0][1]["scale"],
                         "per_view": {str(v): {"measured_s": sum(x["t_step"] for x in rs) / len(rs),
                                               "per_gaussian_s": sum(x["t_gauss"] for x in rs) / len(rs),
                                               "estimated_full_frame_s": sum(x["t_full_est"] for x in rs) / len(rs)}
                                      for v, rs in per_view.items()}},
        "e2e": {"value": value, "unit": "iters/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------- parity
def _rel(a, b):
    return float((a - b).abs().max() / (b.abs().max() + 1e-12))


def parity_check_explicit(wl: Workload, params, dev, threads):
    """CUDA path vs the CPU oracle on the fixed sample windows of the BENCHMARKED scene: integer stages bit-exact,
    image 1e-4 abs (depth channel relative to max(1, |ref|)), the five parameter gradients 1e-3 relative to the
    tensor's max (north_star tolerances).  Returns (parity dict, the oracle timings for cpu_baseline)."""
    import horizongs_b200 as hgs
    win = centre_window(wl)
    views = SAMPLE_VIEWS.get(wl.config, (0,))
    res = {"window": list(win), "views": list(views), "tol": {"image_abs": 1e-4, "grad_rel": 1e-3}, "per_view": {}}
    timings = {}
    ok_all = True
    for v in views:
        o = oracle_window_step(wl, v, win, threads, keep=True)
        timings[v] = {k: o[k] for k in ("t_step", "t_gauss", "t_full_est", "scale", "n_isects")}
        K2 = window_K(wl.Ks[v], win).to(dev)
        for p in params:
            p.grad = None
        rc, ra, meta, extra = render_explicit(hgs, wl.kind, params, wl.views[v][None].to(dev), K2[None], win[2], win[3],
                                              wl.sh_degree, torch.zeros(1, 3, device=dev))
        meta["means2d"].retain_grad()
        parity_loss(rc, ra, extra, win).backward()
        torch.cuda.synchronize()
        ints = {k: bool(torch.equal(meta[k].cpu(), o["meta"][k]))
                for k in ("tiles_per_gauss", "isect_ids", "flatten_ids", "isect_offsets")}
        # radii: float32 holds consecutive integers only below 2^24; a surfel at the near plane projects to a radius of
        # 10^7 pixels, where one ulp is 2 -- compare the radii that are representable exactly
        r_ref, r_got = o["meta"]["radii"], meta["radii"].cpu()
        small = r_ref < (1 << 22)
        ints["radii"] = bool(torch.equal(r_got[small], r_ref[small]) and ((r_got > 0) == (r_ref > 0)).all())
        ref_rc = o["rc"]
        img = float(((rc.detach().cpu() - ref_rc).abs() / ref_rc.abs().clamp(min=1.0)).max())
        alp = float((ra.detach().cpu() - o["ra"]).abs().max())
        grads = {n: _rel(p.grad.cpu(), g) for n, p, g in zip(("means", "quats", "scales", "opacities", "colors"),
                                                             params, o["grads"])}
        grads["means2d"] = _rel(meta["means2d"].grad.cpu(), o["meta"]["means2d"].grad) if wl.kind != "2dgs" else None
        row = {"n_isects": o["n_isects"], "n_visible": int((o["meta"]["radii"] > 0).sum()),
               "max_tile_depth": int(torch.diff(torch.cat([o["meta"]["isect_offsets"].flatten(),
                                                           torch.tensor([o["n_isects"]])])).max()),
               "integer_stages_bit_exact": ints, "image_max_err": img, "alpha_max_err": alp, "grad_max_rel": grads}
        n_px = ref_rc.shape[1] * ref_rc.shape[2]
        img_ok = img < 1e-4 and alp < 1e-4
        if extra is not None:
            # 2DGS: the float32 ORACLE itself is ill-conditioned at grazing ray / surfel intersections (it deviates
            # from its own float64 evaluation by up to 1e-3 at isolated pixels), so the decisive image comparison is the
            # blend stage against the float64 oracle on the same (float32, bit-exact) stage inputs
            row["normals_max_err"] = float((extra[0].detach().cpu() - o["extra"][0]).abs().max())
            e_px = ((rc.detach().cpu() - ref_rc).abs() / ref_rc.abs().clamp(min=1.0)).amax(-1)
            row["pixels_above_1e-4_vs_f32_oracle"] = int((e_px > 1e-4).sum())
            row["pixels"] = int(n_px)
            s64, g64 = stage2d_vs_f64(wl, o["meta"], v, win, dev)
            row["blend_stage_vs_f64_oracle"] = {"image_max_err": s64, "grad_max_rel": g64}
            row["pipeline_grad_max_rel_vs_f32_oracle"] = dict(grads)
            img_ok = max(s64.values()) < 1e-4
            grads = g64               # decisive for 2DGS: the blend stage against float64 (see above)
            row["grad_max_rel"] = g64
        ok = all(ints.values()) and img_ok and all(g is None or g < 1e-3 for g in grads.values())
        row["ok"] = ok
        ok_all = ok_all and ok
        res["per_view"][str(v)] = row
        for p in params:
            p.grad = None
        del o
    res["ok"] = ok_all
    return res, timings


def stage2d_vs_f64(wl, m, v, win, dev):
    """2DGS blend stage (rasterize_to_pixels_2dgs) on the CUDA path against the float64 oracle, both fed the oracle's
    float32 projection outputs of the window: image errors and gradient errors of the five stage inputs"""
    import horizongs_b200 as hgs
    from oracle import gsplat_oracle as O
    w, h = win[2], win[3]
    cols = O._view_colors(wl.sc.means, wl.sc.colors, wl.views[v][None], m["radii"], wl.sh_degree)
    cols = torch.cat([cols, m["depths"].detach()[..., None]], -1).contiguous()
    src = [m["means2d"].detach(), m["ray_transforms"].detach(), cols, m["opacities"].detach().contiguous(),
           m["normals"].detach()]
    g = torch.Generator().manual_seed(13)
    ws = [torch.rand(1, h, w, 4, generator=g), torch.rand(1, h, w, 1, generator=g), torch.rand(1, h, w, 3, generator=g)]
    ins = [t.double().requires_grad_() for t in src]
    r64 = O.rasterize_to_pixels_2dgs(ins[0], ins[1], ins[2], ins[3], ins[4], w, h, 16, m["isect_offsets"],
                                     m["flatten_ids"])
    loss = sum((o_ * w_.double()).mean() for o_, w_ in zip(r64[:3], ws))
    ref = torch.autograd.grad(loss, ins)
    cins = [t.to(dev).requires_grad_() for t in src]
    out = hgs.rasterize_to_pixels_2dgs(cins[0], cins[1], cins[2], cins[3], cins[4], None, w, h, 16,
                                       m["isect_offsets"].to(dev), m["flatten_ids"].to(dev))
    got = torch.autograd.grad(sum((o_ * w_.to(dev)).mean() for o_, w_ in zip(out[:3], ws)), cins)
    img = {n: float(((a.detach().cpu().double() - b.detach()).abs() / b.detach().abs().clamp(min=1.0)).max())
           for n, a, b in zip(("colors", "alphas", "normals"), out[:3], r64[:3])}
    grd = {n: _rel(a.cpu().double(), b) for n, a, b in
           zip(("means2d", "ray_transforms", "colors", "opacities", "normals"), got, ref)}
    return img, grd


def exchange_parity_check(fused, step_plain, step_fused, params, stats, world, dev):
    """N > 1: one extra step both ways -- the fused backward + exchange against the dense all-reduce of every rank's
    autograd gradients (and densification statistics); bit-identity of the fused result across ranks."""
    import torch.distributed as dist
    N = params[0].shape[0]
    worst, same, same_den = 0.0, True, True
    for s in (10_000, 10_001):                          # step indices outside the timed range (one step of each kind of
        # view under the bucketed schedule); the same views both ways
        # dense reference
        st_ref = torch.zeros(2, N, device=dev)
        step_plain(s, st_ref)
        ref = [p.grad.clone().contiguous() for p in params]
        for t in ref + [st_ref]:
            dist.all_reduce(t)
        for p in params:
            p.grad = None
        st_f = torch.zeros(2, N, device=dev)
        step_fused(s, st_f)
        torch.cuda.synchronize()
        for p, r in zip(params, ref):
            worst = max(worst, _rel(p.grad, r))
            g0 = p.grad.clone()
            dist.broadcast(g0, 0)
            same = same and bool(torch.equal(g0, p.grad))
        worst = max(worst, _rel(st_f[0], st_ref[0]))
        same_den = same_den and bool(torch.equal(st_f[1], st_ref[1]))
        for p in params:
            p.grad = None
        del ref, st_ref, st_f
    t = torch.tensor([worst, 0.0 if same else 1.0, 0.0 if same_den else 1.0], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return {"max_rel": float(t[0]), "bit_identical": bool(t[1] == 0), "visibility_counts_equal": bool(t[2] == 0),
            "tol": 1e-3, "ok": bool(t[0] < 1e-3 and t[1] == 0 and t[2] == 0),
            "what": "fused SH/projection backward + peer-memory exchange vs dense NCCL all-reduce of every rank's autograd "
                    "gradients and densification statistics, two extra steps (one of each kind of view) after the timed region"}


# ------------------------------------------------------------------------------------------------- our arm
def run_ours(args):
    import torch.distributed as dist
    import horizongs_b200 as hgs
    from horizongs_b200 import _lib
    from horizongs_b200.cuda import _wrapper as Wr

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py (impl ours) needs a CUDA device"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    elif args.force_exchange:
        # development aid: the fused backward + exchange kernels with ONE rank (their fixed costs; ncu can profile it)
        dist.init_process_group("nccl", init_method="tcp://127.0.0.1:29571", rank=0, world_size=1, device_id=dev)
    L = _lib.lib()                                   # fails loudly if the CUDA library is missing

    # configs[4]: the same 8 seeded cameras (4 aerial + 4 street) for every N -- BASELINE.json's "8-view batch"
    wl = Workload(args.config, args.gaussians, n_views=8)
    W, H, NV = wl.W, wl.H, wl.n_views
    views, Ks, gts = wl.views.to(dev), wl.Ks.to(dev), wl.gts.to(dev)
    # View schedule.  Aerial views (even indices) cost ~1.6x a street view (odd indices).  "interleaved": rank r renders
    # view (r + s) % NV, so for N > 1 every step holds both kinds and lasts as long as its aerial views.  "bucketed"
    # (default for configs[4]): a step holds views of ONE kind -- step s renders kind s % 2, rank r takes the
    # (r + s // 2)-th view of that kind -- the cost-bucketed batch sampler a data-parallel trainer uses.  Every rank still
    # visits every view, each rank renders one aerial and one street view per two steps for every N (per-GPU work is
    # unchanged: weak scaling), and with one rank the two schedules are the same sequence 0, 1, 2, ...  It needs as many
    # views of a kind as ranks: with 8 ranks a step IS the 8-view batch (4 aerial + 4 street, one view per rank, as
    # BASELINE.json's configs[4] names it) and the schedule is the interleaved one.  (A 16-view set was tried for 8 ranks:
    # its view 11 sees 1.68 M Gaussians and overflowed the N // 4 mailboxes of that time, see exchange_cap_rows.)
    bucketed = (args.view_schedule == "bucketed" and args.config == 4 and NV % 2 == 0 and NV >= 2 * world)

    def view_of(s, r=rank):
        return view_schedule(s, r, NV, bucketed)
    bg = torch.zeros(1, 3, device=dev)
    if wl.kind == "lod":
        from tests import lod_harness as LH
        model = wl.model.to(dev)
        model.level = model.level.to(dev)
        params = list(model.parameters())
        N = wl.n * model.n_offsets
    else:
        sc = wl.sc.to(dev)
        params = [t.requires_grad_() for t in (sc.means, sc.quats, sc.scales, sc.opacities, sc.colors)]
        N = sc.n
    # densification statistics (scene/basic_model.py:96-144): gradient-norm accumulator and visibility count
    stats = torch.zeros(2, N, device=dev)
    step_stats = torch.zeros(2, N, device=dev) if world > 1 else None
    from horizongs_b200 import distributed as D
    # gradient exchange (N > 1): backward fused with the exchange over NVLink peer memory (csrc/exchange_vjp.cu); the
    # sparse all-reduce (csrc/exchange.cu) and the dense NCCL all-reduce stay available (--exchange peer | nccl)
    peer, fused, exchange, exchange_name = None, None, None, "none (1 GPU)"
    if world > 1 and wl.kind == "lod":
        raise SystemExit("configs[3] is a single-GPU bench line")
    cap_rows = exchange_cap_rows(N)
    if (world > 1 or args.force_exchange) and args.exchange in ("peer", "fused"):
        try:
            if args.exchange == "fused":
                fused = D.FusedBackwardExchange(N, cap_rows=cap_rows, device=dev)
                exchange_name = ("per-Gaussian backward fused with the exchange over NVLink peer memory (own kernels, "
                                 "csrc/exchange_vjp.cu): each rank runs the camera-specific SH / projection VJP of its own "
                                 "view and stores one 64-byte record per visible Gaussian into every peer's mailbox; every "
                                 "rank then expands the rank-one SH part, sums all views' records in rank order "
                                 "(bit-identical replicas) and updates the densification statistics")
            else:
                peer = D.PeerGradientExchange((3, 4, 3, 1, 27, 1, 1), N, cap_rows=cap_rows, device=dev)
                exchange_name = ("sparse all-reduce over NVLink peer memory (own kernels, csrc/exchange.cu): each rank "
                                 "stores the 40-float records (38 gradients + 2 densification statistics) of its visible "
                                 "Gaussians into every peer's mailbox, then merges all ranks' records in rank order")
        except Exception as e:  # noqa: BLE001
            log(f"[bench] rank {rank}: peer-memory exchange unavailable ({e}); using the NCCL all-reduce")
            peer = fused = None
        ok = torch.tensor([1 if (peer is not None or fused is not None) else 0], device=dev)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        if int(ok.item()) == 0:
            peer = fused = None
    if world > 1 and peer is None and fused is None:
        exchange = D.GradientExchange(params)
        exchange_name = ("NCCL all-reduce of 38 floats/Gaussian gradients + 2 floats/Gaussian densification statistics "
                         "per step")
    copy_stream = torch.cuda.Stream(device=dev)
    loss_ready = torch.cuda.Event()
    loss_host = torch.zeros(1).pin_memory()

    def read_loss(loss):
        """device -> host read of the step's loss, every step: the 4-byte copy runs on the copy stream as soon as the
        forward has produced the loss (the backward is already enqueued behind it), and the host waits for the value"""
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(loss_ready)
            loss_host.copy_(loss.detach().reshape(1), non_blocking=True)
        copy_stream.synchronize()
        return float(loss_host[0])
    # pinned host copies for the end-to-end arm
    gts_pin = wl.gts.pin_memory()
    views_pin, Ks_pin = wl.views.pin_memory(), wl.Ks.pin_memory()

    def inputs(s, e2e):
        v = view_of(s)
        if not e2e:
            return views[v:v + 1], Ks[v:v + 1], gts[v:v + 1]
        # camera first (the forward needs it at once); the ground-truth image is copied on a second stream while the
        # forward runs and joined just before the loss
        view = views_pin[v:v + 1].to(dev, non_blocking=True)
        Km = Ks_pin[v:v + 1].to(dev, non_blocking=True)
        main = torch.cuda.current_stream()
        copy_stream.wait_stream(main)
        with torch.cuda.stream(copy_stream):
            gt = gts_pin[v:v + 1].to(dev, non_blocking=True)
        gt.record_stream(main)
        return view, Km, gt

    def fwd_loss(view, Km, gt, e2e):
        if wl.kind == "lod":
            o = LH.render(model, view[0], Km[0], W, H, bg[0], hgs, fused_decode=True)
            rc = torch.cat([o["render"], o["render_depth"]], 0).permute(1, 2, 0)[None]
            ra = o["render_alphas"].permute(1, 2, 0)[None]
            meta, extra = {"means2d": o["viewspace_points"], "radii": o["radii"][None], "visible_ids": None}, None
        else:
            rc, ra, meta, extra = render_explicit(hgs, wl.kind, params, view, Km, W, H, wl.sh_degree, bg)
        if e2e:
            torch.cuda.current_stream().wait_stream(copy_stream)
        loss = loss_fn(rc, ra, gt, extra)
        if e2e:
            loss_ready.record()
        return loss, meta

    meta_vis = [None]

    def step_fused(s, st, e2e=False):
        # N > 1, fused: backward() stops after the blend backward; the SH / projection backward of all ranks' views, the
        # densification statistics and the exchange are one push + one reduce kernel
        view, Km, gt = inputs(s, e2e)
        with fused.deferred():
            loss, _ = fwd_loss(view, Km, gt, e2e)
            loss.backward()
        fused.finish(*params, grad_accum=st[0], denom=st[1])
        return loss

    def step_plain(s, st, e2e=False):
        view, Km, gt = inputs(s, e2e)
        loss, meta = fwd_loss(view, Km, gt, e2e)
        if wl.kind != "lod":
            meta["means2d"].retain_grad()
        loss.backward()
        # densification statistics from the view-space gradient (basic_model.py:131-144), one fused kernel
        if wl.kind != "lod":
            Wr.densification_stats_update(meta["means2d"].grad, meta["radii"], W, H, st[0], st[1],
                                          visible_ids=meta["visible_ids"])
        meta_vis[0] = meta["visible_ids"]          # work list of the step (the sparse all-reduce needs it)
        return loss

    def step(s, e2e=False):
        if fused is not None:
            loss = step_fused(s, stats, e2e)
        elif world > 1:
            # statistics per view BEFORE the exchange, then summed over ranks together with the gradients
            step_stats.zero_()
            loss = step_plain(s, step_stats, e2e)
            if peer is not None:
                peer.exchange([p.grad for p in params] + [step_stats[0], step_stats[1]], meta_vis[0])
            else:
                h = dist.all_reduce(step_stats, async_op=True)
                exchange.wait()                  # gradient all-reduces were started inside backward()
                h.wait()
            stats.add_(step_stats)
        else:
            loss = step_plain(s, stats, e2e)
        out = read_loss(loss) if e2e else None
        for p in params:
            p.grad = None
        return out

    def timed(n_steps, first, e2e):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = L.hgs_debug_launch_count()
        e0.record()
        for s in range(first, first + n_steps):
            step(s, e2e)
        e1.record()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item()), L.hgs_debug_launch_count() - l0

    for s in range(args.warmup):
        step(s)
    # ---- value: device-resident inputs, timed WITHOUT instrumentation; the per-stage CUDA events (on the launching
    # stream) come from a second pass over the same steps, so that their ~30 event records per step are not part of
    # the reported time (ms_per_step_instrumented is that pass's own time)
    marks = []

    def hook(name, phase):
        ev = torch.cuda.Event(enable_timing=True)
        ev.record()
        marks.append((name, phase, ev))

    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    ms_total, launches = timed(args.steps, args.warmup, e2e=False)
    Wr.set_stage_hook(hook)
    ms_instr, _ = timed(args.steps, args.warmup, e2e=False)
    Wr.set_stage_hook(None)
    stage_ms = {}
    opened = {}
    for name, phase, ev in marks:
        if phase == 0:
            opened[name] = ev
        else:
            stage_ms.setdefault(name, []).append(opened.pop(name).elapsed_time(ev))
    stage_avg = {k: sum(v) / len(v) for k, v in stage_ms.items()}
    blend_b = "blend2d_bwd" if wl.kind == "2dgs" else "blend3d_bwd"
    blend_f = "blend2d_fwd" if wl.kind == "2dgs" else "blend3d_fwd"
    # per-view split of the blend kernels (step s of the timed loop renders view_of(warmup + s))
    stage_by_view = {}
    for k in (blend_f, blend_b, "isect_sorted"):
        per = {}
        for i, t in enumerate(stage_ms.get(k, [])):
            per.setdefault(view_of(args.warmup + i), []).append(t)
        stage_by_view[k] = {str(v): round(sum(ts) / len(ts), 4) for v, ts in sorted(per.items())}
    # ---- e2e: host buffers, H2D of camera + ground truth and D2H of the loss inside the timed region
    for s in range(2):
        step(s, e2e=True)
    ms_e2e, _ = timed(args.steps, args.warmup, e2e=True)
    if peer is not None:
        peer.check_status()
    if fused is not None:
        fused.check_status()
    clk = clocks.stop() if rank == 0 else None

    # ---- per-rank stage times (N > 1): what the step waits on
    stage_ranks = None
    if world > 1:
        allst = [None] * world
        dist.all_gather_object(allst, {k: round(v, 4) for k, v in stage_avg.items()})
        keys = sorted({k for d in allst for k in d})
        stage_ranks = {k: {"min": min(d.get(k, 0.0) for d in allst), "max": max(d.get(k, 0.0) for d in allst)}
                       for k in keys}

    # ---- N > 1: parity of the exchange (fused vs dense all-reduce of autograd gradients), outside the timed region
    exchange_parity = None
    if fused is not None:
        exchange_parity = exchange_parity_check(fused, step_plain, step_fused, params, stats, world, dev)
        fused.check_status()

    # ---- forward-only render FPS (reference method: torch.no_grad around render(), render.py:79-83,177)
    with torch.no_grad():
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for s in range(args.steps):
            view, Km, gt = inputs(s, False)
            fwd_loss(view, Km, gt, False)
        torch.cuda.synchronize()
        fps = args.steps / (time.perf_counter() - t0)

    fused_multicast = bool(fused is not None and fused.multicast)
    if peer is not None:
        peer.close()
    if fused is not None:
        fused.close()
    if rank != 0:
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return

    # ---- data-dependent counts per view (outside any timed region)
    counts = []
    if wl.kind != "lod":
        with torch.no_grad():
            for v in range(NV):
                rc, ra, meta, _ = render_explicit(hgs, wl.kind, params, views[v:v + 1], Ks[v:v + 1], W, H, wl.sh_degree, bg)
                off = meta["isect_offsets"].flatten()
                depth = torch.diff(torch.cat([off, off.new_tensor([meta["flatten_ids"].numel()])]))
                row = {"view": v, "n_visible": int((meta["radii"] > 0).sum()), "I": int(meta["flatten_ids"].numel()),
                       "max_tile_depth": int(depth.max())}
                if wl.kind == "3dgs":
                    pe, pb = Wr.blend3d_pair_stats(meta["means2d"], meta["conics"], meta["opacities"].contiguous(),
                                                   meta["radii"], W, H, 16, meta["isect_offsets"], meta["flatten_ids"])
                    row.update(P_eval=pe, P_blend=pb, **Wr.blend3d_pair_stats.last_cull)
                counts.append(row)
    mean = lambda k: sum(c[k] for c in counts) / max(1, len(counts))  # noqa: E731

    # ---- roofline of the dominant kernel (blend backward): FP32 pipe
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    hbm_src = "measured (MEASURED_PEAKS.json)" if "hbm_gbs" in peaks else "fallback (B200_PROFILING.md)"
    n_sm = torch.cuda.get_device_properties(dev).multi_processor_count
    sm_max = float((clk or {}).get("sm_max_mhz") or peaks.get("sm_max_mhz", 1965.0))
    fp32_peak = 2 * 128 * n_sm * sm_max * 1e6 / 1e12            # TFLOP/s, FMA = 2 FLOP, non-tensor pipe
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath):
        try:
            traffic = json.load(open(tpath))
        except Exception:
            traffic = None
    dom = max((k for k in stage_avg if k.startswith("blend") and k.endswith(("_fwd", "_bwd"))),
              key=lambda k: stage_avg[k], default=None)
    roofline = None
    if dom is not None and wl.kind == "3dgs" and counts:
        # algorithmic work: sigma/alpha/tests 14 FLOP per evaluated pair; forward blend 12 FLOP and
        # backward 66 FLOP per blended pair.  frac_blended_only counts ONLY pairs that are actually blended at
        # SURVEY 8(d)'s 80 (bwd) / 26 (fwd) FLOP per pair -- the per-warp cull skips most evaluated pairs
        bwd = dom == "blend3d_bwd"
        flops = (14 * mean("P_eval") + (66 if bwd else 12) * mean("P_blend"))
        ach = flops / (stage_avg[dom] * 1e-3) / 1e12
        ach2 = (80 if bwd else 26) * mean("P_blend") / (stage_avg[dom] * 1e-3) / 1e12
        roofline = {"kernel": dom, "bound": "fp32", "achieved": ach, "peak": fp32_peak, "unit": "TFLOP/s",
                    "frac": ach / fp32_peak, "frac_blended_pairs_only": ach2 / fp32_peak,
                    "traffic": (traffic or {}).get(dom),
                    "peak_source": f"2 FLOP x 128 lanes x {n_sm} SMs x {sm_max:.0f} MHz (non-tensor FP32 FMA peak; "
                                   "the path is not a dense contraction, no tensor-core roofline applies)",
                    "avg_launch_ms": stage_avg[dom]}
    elif dom is not None:
        roofline = {"kernel": dom, "bound": "fp32", "achieved": None, "peak": fp32_peak, "unit": "TFLOP/s", "frac": None,
                    "traffic": (traffic or {}).get(dom), "avg_launch_ms": stage_avg[dom],
                    "note": "pair counts are only instrumented for the 3DGS blend; see profiles/ for the ncu pipe utilisation"}
    roofline_hbm = None
    if "isect_sorted" in stage_avg and counts:
        # Ordering stages a8-a10 as a whole.  `achieved` follows SURVEY 8(d): the ALGORITHMIC bytes of the reference
        # formulation of the same result -- count + emit 2 x 16 B x N + 4 B x N + 12 B x I, 6-pass pair sort 152 B x I,
        # tile offsets 8 B x I + 4 B x T -- over the time of our ordering stages.  Our super-tile formulation moves far
        # fewer bytes (achieved_own_bytes: records 16 B + keys 8 B per visible Gaussian written and read, keys sorted
        # in shared memory 2 x 8 B per super-tile key, 12 B per intersection written), so frac can exceed what a
        # six-pass sort could ever reach.  With the fused projection (3DGS) the compaction / histogram half of the
        # ordering runs inside project3d_fwd and is not separable; its full time is charged here only as
        # `avg_launch_ms_with_projection`.
        T = math.ceil(W / 16) * math.ceil(H / 16)
        t_ord = stage_avg["isect_sorted"] + stage_avg.get("isect_prepare", 0.0)
        t_with_proj = t_ord + stage_avg.get("project3d_fwd", stage_avg.get("project2d_fwd", 0.0))
        bytes_ref = (2 * 16 + 4) * N + (12 + 152 + 8) * mean("I") + 4 * T
        bytes_own = (16 + 16 + 8) * mean("n_visible") + 12 * mean("I") + 4 * T
        ach = bytes_ref / (t_ord * 1e-3) / 1e9
        roofline_hbm = {"kernel": "ordering stages a8-a10 (isect_prepare/scan + isect_sorted: super-tile binning, "
                                  "per-super-tile sort, expansion into tile ranges)", "bound": "hbm", "achieved": ach,
                        "peak": hbm_peak, "unit": "GB/s", "frac": ach / hbm_peak, "peak_source": hbm_src,
                        "algorithmic_bytes": bytes_ref, "algorithmic_bytes_source": "SURVEY 8(d): 36 B x N + 172 B x I + 4 B x T",
                        "achieved_own_bytes": bytes_own / (t_ord * 1e-3) / 1e9, "own_bytes": bytes_own,
                        "traffic": (traffic or {}).get("isect_sorted"), "avg_launch_ms": t_ord,
                        "avg_launch_ms_with_projection": t_with_proj,
                        "frac_with_projection": (bytes_ref + 68 * N) / (t_with_proj * 1e-3) / 1e9 / hbm_peak}

    # ---- parity on the benchmarked scene + CPU baseline on the same bounded sample (rank 0)
    parity, cpu_baseline = None, None
    threads = os.cpu_count() or 1
    if wl.kind != "lod" and not args.no_cpu_baseline and (world == 1 or args.parity_at_scale):
        t0 = time.time()
        parity, tim = parity_check_explicit(wl, params, dev, threads)
        win = centre_window(wl)
        vs = list(tim)
        est = sum(tim[v]["t_full_est"] for v in vs) / len(vs)
        meas = sum(tim[v]["t_step"] for v in vs) / len(vs)
        if world == 1:
            cpu_baseline = {"value": 1.0 / est, "unit": "iters/s", "cores": threads, "kind": "port",
                            "sample": cpu_sample_description(wl, win, vs, threads),
                            "measured_s_per_sample_step": meas, "estimated_s_per_full_frame": est,
                            "extrapolation_factor_pixels": tim[vs[0]]["scale"],
                            "per_view": {str(v): {"measured_s": tim[v]["t_step"], "per_gaussian_s": tim[v]["t_gauss"],
                                                  "estimated_full_frame_s": tim[v]["t_full_est"]} for v in vs},
                            "cpu_seconds_spent": round(time.time() - t0, 1)}

    n_steps = args.steps
    value = world * n_steps / (ms_total * 1e-3)
    h2d = int(gts_pin[0:1].numel() * 4 + 16 * 4 + 9 * 4)
    sum_stage = sum(stage_avg.values())
    line = {
        "metric": wl.metric, "value": value, "unit": "iters/s", "n_gpus": world, "steps": n_steps, "warmup": args.warmup,
        "ms_per_step": ms_total / n_steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": wl.name, "views": NV, "render_mode": "RGB+ED", "sh_degree": wl.sh_degree,
                   "tile_size": 16, "loss": LOSS_MODE,
                   "view_schedule": ("bucketed: a step holds views of one kind (aerial on even steps, street on odd "
                                     "steps; rank r renders the (r + s // 2)-th view of the kind); per-GPU work per two "
                                     "steps = one aerial + one street view for every N" if bucketed else
                                     "interleaved: rank r renders view (r + s) mod views (with 8 ranks every step is "
                                     "the whole 8-view batch: 4 aerial + 4 street views, one per rank)"),
                   "l2": (f"inputs larger than L2 ({N * 38 * 4 / 1e6:.0f} MB of Gaussian parameters per step; no flush)"
                          if N * 38 * 4 > 126e6 else
                          "the view (and with it the set of visible Gaussians and every intermediate) changes every step; "
                          "parameters + gradients + intermediates of a step exceed L2"),
                   "collective": exchange_name,
                   "exchange_transport": (None if fused is None else
                                          ("NVSwitch multicast: every record stored once with multimem.st, replicated by "
                                           "the switch" if fused_multicast else "one unicast peer store per rank"))},
        "clocks": clk,
        "e2e": {"value": world * n_steps / (ms_e2e * 1e-3), "unit": "iters/s", "h2d_bytes_per_step": h2d,
                "d2h_bytes_per_step": 4, "ms_per_step": ms_e2e / n_steps},
        "gpu_launches": int(launches),
        "render_fps": fps,
        "roofline": roofline, "roofline_hbm": roofline_hbm, "cpu_baseline": cpu_baseline,
        "parity_check": parity, "exchange_parity": exchange_parity,
        "stage_ms": stage_avg, "stage_ms_sum": sum_stage, "outside_stage_ms": ms_total / n_steps - sum_stage,
        "ms_per_step_instrumented": ms_instr / n_steps,
        "stage_ms_by_view": stage_by_view, "stage_ms_over_ranks": stage_ranks,
        "counts": ({"mean_n_visible": mean("n_visible"), "mean_I": mean("I"),
                    "mean_P_eval": mean("P_eval") if wl.kind == "3dgs" else None,
                    "mean_P_blend": mean("P_blend") if wl.kind == "3dgs" else None, "per_view": counts}
                   if counts else None),
        "frame_budget": {"ms_per_view": ms_total / n_steps, "within_33.3ms": ms_total / n_steps < 33.3,
                         "within_16.7ms": ms_total / n_steps < 16.7},
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=16)
    ap.add_argument("--warmup", type=int, default=4)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", type=int, default=4, choices=[0, 1, 2, 3, 4],
                    help="index into BASELINE.json configs (default 4: the configuration the metric is quoted on)")
    ap.add_argument("--gaussians", type=int, default=None, help="override the Gaussian / anchor count of the config")
    ap.add_argument("--no-cpu-baseline", action="store_true", help="skip the CPU oracle legs (parity_check, cpu_baseline)")
    ap.add_argument("--parity-at-scale", action="store_true", help="N > 1: also run rank 0's oracle parity check")
    ap.add_argument("--force-exchange", action="store_true",
                    help="development aid: run the N > 1 exchange path with a single rank")
    ap.add_argument("--exchange", default="fused", choices=["fused", "peer", "nccl"],
                    help="N > 1: per-Gaussian backward fused with the exchange over NVLink peer memory (default), "
                         "sparse all-reduce of the parameter gradients over peer memory, or the dense NCCL all-reduce")
    ap.add_argument("--view-schedule", default="bucketed", choices=["bucketed", "interleaved"],
                    help="configs[4], N > 1: one kind of view (aerial / street) per step, or both kinds in every step")
    ap.add_argument("--loss", default="l1", choices=["l1", "l1ssim"],
                    help="loss inside the step: fused L1 (+ depth / alpha means; default) or the reference's "
                         "0.8 L1 + 0.2 (1 - SSIM), fused (csrc/loss.cu)")
    args = ap.parse_args()
    global LOSS_MODE
    LOSS_MODE = args.loss
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else max(args.warmup, 0)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
