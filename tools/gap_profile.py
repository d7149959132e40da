#!/usr/bin/env python
"""GPU idle gaps of one bench step (torch.profiler / CUPTI): kernel time vs wall span, largest gaps and what follows them."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import horizongs_b200 as hgs
from horizongs_b200 import scenes, losses
from horizongs_b200.cuda import _wrapper as Wr
from torch.profiler import profile, ProfilerActivity

sc, views, Ks, W, H = scenes.config4()
sc = sc.to('cuda'); views = views.cuda(); Ks = Ks.cuda()
gts = torch.rand(8, H, W, 3, device='cuda')
params = [t.requires_grad_() for t in (sc.means, sc.quats, sc.scales, sc.opacities, sc.colors)]
stats = torch.zeros(2, sc.n, device='cuda')
bg = torch.zeros(1, 3, device='cuda')

def step(s):
    v = s % 8
    rc, ra, meta = hgs.rasterization(*params, views[v:v+1], Ks[v:v+1], W, H, sh_degree=2, render_mode="RGB+ED", backgrounds=bg)
    meta["means2d"].retain_grad()
    loss = losses.photometric_l1_loss(rc, gts[v:v+1], ra, 0.01, 0.01)
    loss.backward()
    Wr.densification_stats_update(meta["means2d"].grad, meta["radii"], W, H, stats[0], stats[1], visible_ids=meta["visible_ids"])
    for p in params:
        p.grad = None

for s in range(6):
    step(s)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for s in range(6, 14):
        step(s)
    torch.cuda.synchronize()
evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
ks = sorted([(e.time_range.start, e.time_range.end, e.name) for e in evs])
span = ks[-1][1] - ks[0][0]
busy = 0; gaps = []
cur_end = ks[0][0]
for a, b, n in ks:
    if a > cur_end:
        gaps.append((a - cur_end, n))
    busy += max(0, b - max(a, cur_end))
    cur_end = max(cur_end, b)
print(json.dumps({"steps": 8, "span_us": span, "busy_us": busy, "idle_us": span - busy, "idle_per_step_us": (span - busy) / 8,
                  "n_kernels": len(ks)}))
agg = {}
for g, n in gaps:
    agg.setdefault(n[:60], [0, 0]); agg[n[:60]][0] += g; agg[n[:60]][1] += 1
for n, (g, c) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:15]:
    print(f"{g/8:8.1f} us/step idle before  x{c/8:.1f}  {n}")
