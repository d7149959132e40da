#!/usr/bin/env python
"""bench.py -- fwd+bwd iterations/s of the rasterization hot path on BASELINE.json's headline workload.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...   (N > 1)

Workload (config.workload): BASELINE.json configs[4] = the configuration its metric is quoted on -- 6M explicit
SH2 Gaussians, 1920x1080, render_mode RGB+ED, 8 seeded cameras (aerial / street alternating).  One "step" =
one view per GPU: rasterization forward, an L1-style loss, backward to the 38 floats of every Gaussian, and
for N > 1 the NCCL all-reduce of those gradients and of the densification statistics.  Rank r renders view
(r + step) mod 8, so every rank sees every view ("weak" scaling: per-GPU work fixed).

One JSON line on stdout (rank 0).  `value` = views/s with everything resident in HBM; `e2e` = the same through
the public API with the step's camera and ground-truth image copied from pinned host memory and the loss read
back, inside the timed region.  `roofline` describes the dominant kernel (blend backward, FP32-pipe bound),
`roofline_hbm` the dominant HBM-bound stage; `cpu_baseline` is the CPU oracle on a bounded sample.
`--impl reference` times the CPU oracle port (the reference's own rasterizer, gsplat, is not installable
here: see DESIGN.md) on a bounded sample of the same workload, rank 0 only.
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

METRIC = "fwd+bwd iters/s, 6M Gaussians @1920x1080"
LOSS_MODE = "l1"
N_VIEWS = 8
WIDTH, HEIGHT = 1920, 1080


def log(*a):
    print(*a, file=sys.stderr, flush=True)


# ------------------------------------------------------------------------------------------------- workload
def make_workload(n_gauss: int):
    from horizongs_b200 import scenes
    t0 = time.time()
    sc, views, Ks, W, H = scenes.config4(n=n_gauss, n_views=N_VIEWS, width=WIDTH, height=HEIGHT)
    g = torch.Generator().manual_seed(7)
    gts = torch.rand(N_VIEWS, H, W, 3, generator=g)             # synthetic ground-truth images
    log(f"[bench] scene {n_gauss} Gaussians generated in {time.time() - t0:.1f}s")
    return sc, views, Ks, W, H, gts


def loss_fn(rc, ra, gt):
    """L1 photometric term + small depth / alpha terms (every output channel gets a gradient).  On the GPU the
    same expression is one fused forward and one fused backward kernel (horizongs_b200.losses, csrc/loss.cu)."""
    if rc.is_cuda:
        from horizongs_b200 import losses
        if LOSS_MODE == "l1ssim":       # the reference's full photometric loss (train.py:158-160), fused
            return losses.photometric_loss(rc, gt, 0.2, ra, w_depth=0.01, w_alpha=0.01)
        return losses.photometric_l1_loss(rc, gt, ra, w_depth=0.01, w_alpha=0.01)
    return (rc[..., :3] - gt).abs().mean() + 0.01 * rc[..., 3].mean() + 0.01 * ra.mean()


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
def oracle_window_step(sc, view, K, win, threads):
    """fwd+bwd of the CPU oracle on a window (x0,y0,w,h) of the full frame (a sub-frustum: same Gaussians, same
    camera, principal point shifted).  Returns (seconds for the window step, seconds of its projection/SH part)."""
    from oracle import gsplat_oracle as O
    torch.set_num_threads(threads)
    x0, y0, w, h = win
    K2 = K.clone()
    K2[0, 2] -= x0
    K2[1, 2] -= y0
    params = [t.clone().requires_grad_() for t in (sc.means, sc.quats, sc.scales, sc.opacities, sc.colors)]
    gt = torch.rand(1, h, w, 3, generator=torch.Generator().manual_seed(3))
    t0 = time.perf_counter()
    rc, ra, meta = O.rasterization(*params, view[None], K2[None], w, h, sh_degree=sc.sh_degree, render_mode="RGB+ED",
                                   backgrounds=torch.zeros(1, 3))
    loss_fn(rc, ra, gt).backward()
    t_step = time.perf_counter() - t0
    # the per-Gaussian part (projection + SH, fwd+bwd) does not shrink with the window: time it alone
    params = [t.clone().requires_grad_() for t in (sc.means, sc.quats, sc.scales, sc.opacities, sc.colors)]
    t0 = time.perf_counter()
    radii, m2, d, con, _ = O.fully_fused_projection(params[0], None, params[1], params[2], view[None], K2[None], w, h)
    cols = O._view_colors(params[0], params[4], view[None], radii, sc.sh_degree)
    (m2.sum() + d.sum() + con.sum() + cols.sum()).backward()
    t_gauss = time.perf_counter() - t0
    return t_step, t_gauss, int(meta["flatten_ids"].numel())


def oracle_full_frame_estimate(sc, view, K, win, threads):
    t_step, t_gauss, n_isect = oracle_window_step(sc, view, K, win, threads)
    scale = (WIDTH * HEIGHT) / float(win[2] * win[3])
    t_pix = max(t_step - t_gauss, 1e-6)
    return t_gauss + t_pix * scale, t_step, n_isect


def centre_window(w, h):
    return (WIDTH // 2 - w // 2, HEIGHT // 2 - h // 2, w, h)


def pick_window(sc, view, K, threads):
    """the bounded sample: a centre window sized so that one oracle step is roughly 10-30 s of CPU work on this host
    (192x128 first; a fast host re-runs with 4x / 16x the pixels).  -> (estimate, seconds, isects, window)"""
    win = centre_window(192, 128)
    est, t_step, n_isect = oracle_full_frame_estimate(sc, view, K, win, threads)
    for w, h in ((384, 256), (768, 512)):
        if t_step >= 5.0:
            break
        win = centre_window(w, h)
        est, t_step, n_isect = oracle_full_frame_estimate(sc, view, K, win, threads)
    return est, t_step, n_isect, win


def run_reference(args):
    """--impl reference: the CPU oracle port on a bounded sample, rank 0 only."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    steps, warm = args.steps, args.warmup
    sc, views, Ks, W, H, _ = make_workload(args.gaussians)
    # window sized so that one sample is 10-30 s of CPU work and (steps + warmup) samples end within a few minutes
    t_all = time.time()
    _, _, _, win = pick_window(sc, views[0], Ks[0], threads)
    ests = []
    for s in range(warm + steps):
        v = s % N_VIEWS
        est, t_step, n_isect = oracle_full_frame_estimate(sc, views[v], Ks[v], win, threads)
        if s >= warm:
            ests.append(est)
        log(f"[reference] step {s}: window {t_step:.2f}s -> full-frame estimate {est:.1f}s ({n_isect} isects)")
        if time.time() - t_all > 240 and len(ests) >= 1:
            log("[reference] time budget reached; stopping early")
            break
    ms = 1e3 * sum(ests) / len(ests)
    value = 1e3 / ms
    sample = (f"{win[2]}x{win[3]} centre window of the 1920x1080 frame, all {sc.n} Gaussians projected; per-pixel cost "
              f"scaled by {WIDTH * HEIGHT / (win[2] * win[3]):.0f}x, per-Gaussian cost unscaled; mean of {len(ests)} views")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "iters/s", "n_gpus": args.gpus,
        "steps": len(ests), "warmup": warm, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(args.gaussians), "views": N_VIEWS, "render_mode": "RGB+ED",
                   "sh_degree": 2, "note": "gsplat (the reference's rasterizer) is not installable here; this is the "
                   "repo's CPU oracle port of it (torch, float32)"},
        "cpu_baseline": {"value": value, "unit": "iters/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "iters/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def workload_name(n):
    return f"configs[4]: {n / 1e6:g}M explicit SH2 3DGS Gaussians, {WIDTH}x{HEIGHT}, RGB+ED, 1 view per GPU per step"


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
    L = _lib.lib()                                   # fails loudly if the CUDA library is missing

    sc_cpu, views_cpu, Ks_cpu, W, H, gts_cpu = make_workload(args.gaussians)
    sc = sc_cpu.to(dev)
    views, Ks = views_cpu.to(dev), Ks_cpu.to(dev)
    gts = gts_cpu.to(dev)
    bg = torch.zeros(1, 3, device=dev)
    params = [t.requires_grad_() for t in (sc.means, sc.quats, sc.scales, sc.opacities, sc.colors)]
    N = sc.n
    # densification statistics (scene/basic_model.py:96-144): gradient-norm accumulator and visibility count
    stats = torch.zeros(2, N, device=dev)
    step_stats = torch.zeros(2, N, device=dev) if world > 1 else None
    from horizongs_b200 import distributed as D
    # gradient exchange (N > 1): sparse all-reduce over NVLink peer memory (csrc/exchange.cu); the dense NCCL
    # all-reduce stays available (--exchange nccl) and is the fallback if peer memory cannot be mapped
    peer, fused, exchange, exchange_name = None, None, None, "none (1 GPU)"
    if world > 1 and args.exchange in ("peer", "fused"):
        try:
            if args.exchange == "fused":
                fused = D.FusedBackwardExchange(N, cap_rows=N // 4, device=dev)
                exchange_name = ("SH / projection backward fused with the exchange over NVLink peer memory (own kernels, "
                                 "csrc/exchange_vjp.cu): each rank stores the 12-float blend-gradient rows of its visible "
                                 "Gaussians into every peer's mailbox; every rank then runs the per-Gaussian backward of "
                                 "all views' rows, summing in rank order, and updates the densification statistics")
            else:
                peer = D.PeerGradientExchange((3, 4, 3, 1, 27, 1, 1), N, cap_rows=N // 4, device=dev)
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
    gts_pin = gts_cpu.pin_memory()
    views_pin, Ks_pin = views_cpu.pin_memory(), Ks_cpu.pin_memory()

    def step(s, e2e=False):
        v = (rank + s) % N_VIEWS
        if e2e:
            # camera first (the forward needs it at once); the 25 MB ground-truth image is copied on a second
            # stream while the forward runs and joined just before the loss
            view = views_pin[v:v + 1].to(dev, non_blocking=True)
            Km = Ks_pin[v:v + 1].to(dev, non_blocking=True)
            main = torch.cuda.current_stream()
            copy_stream.wait_stream(main)
            with torch.cuda.stream(copy_stream):
                gt = gts_pin[v:v + 1].to(dev, non_blocking=True)
            gt.record_stream(main)
        else:
            view, Km, gt = views[v:v + 1], Ks[v:v + 1], gts[v:v + 1]
        if fused is not None:
            # N > 1, fused: backward() stops after the blend backward; the SH / projection backward of all ranks'
            # views, the densification statistics and the exchange are one push + one reduce kernel
            with fused.deferred():
                rc, ra, meta = hgs.rasterization(params[0], params[1], params[2], params[3], params[4], view, Km, W, H,
                                                 sh_degree=2, render_mode="RGB+ED", backgrounds=bg, packed=False)
                if e2e:
                    torch.cuda.current_stream().wait_stream(copy_stream)
                loss = loss_fn(rc, ra, gt)
                if e2e:
                    loss_ready.record()
                loss.backward()
            fused.finish(*params, grad_accum=stats[0], denom=stats[1])
            out = read_loss(loss) if e2e else None
            for p in params:
                p.grad = None
            return out
        rc, ra, meta = hgs.rasterization(params[0], params[1], params[2], params[3], params[4], view, Km, W, H,
                                         sh_degree=2, render_mode="RGB+ED", backgrounds=bg, packed=False)
        meta["means2d"].retain_grad()
        if e2e:
            torch.cuda.current_stream().wait_stream(copy_stream)
        loss = loss_fn(rc, ra, gt)
        if e2e:
            loss_ready.record()
        loss.backward()
        # densification statistics from the view-space gradient (basic_model.py:131-144), one fused kernel;
        # computed per view BEFORE the exchange, then summed over ranks together with the gradients
        if world > 1:
            step_stats.zero_()
            Wr.densification_stats_update(meta["means2d"].grad, meta["radii"], W, H, step_stats[0], step_stats[1],
                                          visible_ids=meta["visible_ids"])
            if peer is not None:
                peer.exchange([p.grad for p in params] + [step_stats[0], step_stats[1]], meta["visible_ids"])
            else:
                h = dist.all_reduce(step_stats, async_op=True)
                exchange.wait()                  # gradient all-reduces were started inside backward()
                h.wait()
            stats.add_(step_stats)
        else:
            Wr.densification_stats_update(meta["means2d"].grad, meta["radii"], W, H, stats[0], stats[1],
                                          visible_ids=meta["visible_ids"])
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
    # ---- value: device-resident inputs, with per-stage CUDA events on the launching stream
    marks = []

    def hook(name, phase):
        ev = torch.cuda.Event(enable_timing=True)
        ev.record()
        marks.append((name, phase, ev))

    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    Wr.set_stage_hook(hook)
    ms_total, launches = timed(args.steps, args.warmup, e2e=False)
    Wr.set_stage_hook(None)
    stage_ms = {}
    opened = {}
    for name, phase, ev in marks:
        if phase == 0:
            opened[name] = ev
        else:
            stage_ms.setdefault(name, []).append(opened.pop(name).elapsed_time(ev))
    stage_avg = {k: sum(v) / len(v) for k, v in stage_ms.items()}
    # per-view split of the two blend kernels (step s of the timed loop renders view (rank + warmup + s) % 8)
    stage_by_view = {}
    for k in ("blend3d_fwd", "blend3d_bwd", "isect_sorted"):
        per = {}
        for i, t in enumerate(stage_ms.get(k, [])):
            per.setdefault((rank + args.warmup + i) % N_VIEWS, []).append(t)
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

    # ---- forward-only render FPS (reference method: torch.no_grad around render(), render.py:79-83,177)
    with torch.no_grad():
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for s in range(args.steps):
            v = (rank + s) % N_VIEWS
            hgs.rasterization(params[0], params[1], params[2], params[3], params[4], views[v:v + 1], Ks[v:v + 1], W, H,
                              sh_degree=2, render_mode="RGB+ED", backgrounds=bg)
        torch.cuda.synchronize()
        fps = args.steps / (time.perf_counter() - t0)

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
    with torch.no_grad():
        for v in range(N_VIEWS):
            rc, ra, meta = hgs.rasterization(params[0], params[1], params[2], params[3], params[4], views[v:v + 1],
                                             Ks[v:v + 1], W, H, sh_degree=2, render_mode="RGB+ED")
            pe, pb = Wr.blend3d_pair_stats(meta["means2d"], meta["conics"], meta["opacities"].contiguous(),
                                           meta["radii"], W, H, 16, meta["isect_offsets"], meta["flatten_ids"])
            off = meta["isect_offsets"].flatten()
            depth = torch.diff(torch.cat([off, off.new_tensor([meta["flatten_ids"].numel()])]))
            counts.append({"view": v, "n_visible": int((meta["radii"] > 0).sum()), "I": int(meta["flatten_ids"].numel()),
                           "P_eval": pe, "P_blend": pb, "max_tile_depth": int(depth.max())})
    mean = lambda k: sum(c[k] for c in counts) / len(counts)  # noqa: E731

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
    dom = max((k for k in stage_avg if k.startswith("blend3d")), key=lambda k: stage_avg[k], default=None)
    roofline = None
    if dom is not None:
        # algorithmic work: sigma/alpha/tests 14 FLOP per evaluated pair; forward blend 12 FLOP and
        # backward 66 FLOP per blended pair
        flops = (14 * mean("P_eval") + (66 if dom == "blend3d_bwd" else 12) * mean("P_blend"))
        ach = flops / (stage_avg[dom] * 1e-3) / 1e12
        roofline = {"kernel": dom, "bound": "fp32", "achieved": ach, "peak": fp32_peak, "unit": "TFLOP/s",
                    "frac": ach / fp32_peak, "traffic": (traffic or {}).get(dom),
                    "peak_source": f"2 FLOP x 128 lanes x {n_sm} SMs x {sm_max:.0f} MHz (non-tensor FP32 FMA peak; "
                                   "the path is not a dense contraction, no tensor-core roofline applies)",
                    "avg_launch_ms": stage_avg[dom]}
    roofline_hbm = None
    if "isect_sorted" in stage_avg:
        # emit 8 B + 2 tile-bit passes x (4 + 8 + 8) B + finalize (8 + 8 + 4) B per intersection
        bytes_i = (8 + 2 * 20 + 20) * mean("I")
        ach = bytes_i / (stage_avg["isect_sorted"] * 1e-3) / 1e9
        roofline_hbm = {"kernel": "isect_sorted (emit + tile partition + finalize)", "bound": "hbm", "achieved": ach,
                        "peak": hbm_peak, "unit": "GB/s", "frac": ach / hbm_peak, "peak_source": hbm_src,
                        "traffic": (traffic or {}).get("isect_sorted"), "avg_launch_ms": stage_avg["isect_sorted"]}

    # ---- CPU baseline on a bounded sample (rank 0, N == 1 only)
    cpu_baseline = None
    if world == 1 and not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        t0 = time.time()
        est, t_step, _, win = pick_window(sc_cpu, views_cpu[0], Ks_cpu[0], threads)
        cpu_baseline = {"value": 1.0 / est, "unit": "iters/s", "cores": threads, "kind": "port",
                        "sample": f"view 0, {win[2]}x{win[3]} centre window of the 1920x1080 frame with all {N} Gaussians "
                                  f"projected ({t_step:.1f}s measured); per-pixel cost scaled to the full frame, "
                                  f"per-Gaussian cost unscaled; CPU oracle (torch float32), {time.time() - t0:.0f}s of CPU work"}

    n_steps = args.steps
    value = world * n_steps / (ms_total * 1e-3)
    h2d = int(gts_pin[0:1].numel() * 4 + 16 * 4 + 9 * 4)
    line = {
        "metric": METRIC, "value": value, "unit": "iters/s", "n_gpus": world, "steps": n_steps, "warmup": args.warmup,
        "ms_per_step": ms_total / n_steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(N), "views": N_VIEWS, "render_mode": "RGB+ED", "sh_degree": 2,
                   "tile_size": 16, "l2": "inputs larger than L2 (912 MB of Gaussian parameters per step; no flush)",
                   "collective": exchange_name},
        "clocks": clk,
        "e2e": {"value": world * n_steps / (ms_e2e * 1e-3), "unit": "iters/s", "h2d_bytes_per_step": h2d,
                "d2h_bytes_per_step": 4, "ms_per_step": ms_e2e / n_steps},
        "gpu_launches": int(launches),
        "render_fps": fps,
        "roofline": roofline, "roofline_hbm": roofline_hbm, "cpu_baseline": cpu_baseline,
        "stage_ms": stage_avg, "stage_ms_by_view": stage_by_view,
        "counts": {"mean_n_visible": mean("n_visible"), "mean_I": mean("I"), "mean_P_eval": mean("P_eval"),
                   "mean_P_blend": mean("P_blend"), "per_view": counts},
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
    ap.add_argument("--gaussians", type=int, default=6_000_000)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--exchange", default="fused", choices=["fused", "peer", "nccl"],
                    help="N > 1: per-Gaussian backward fused with the exchange over NVLink peer memory (default), "
                         "sparse all-reduce of the parameter gradients over peer memory, or the dense NCCL all-reduce")
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
