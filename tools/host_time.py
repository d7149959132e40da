"""Development aid: host enqueue time vs GPU time of one fwd+bwd step of bench.py's workload (is the step host-bound?)."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import horizongs_b200 as hgs
from horizongs_b200 import losses, scenes

n = int(sys.argv[1]) if len(sys.argv) > 1 else 6_000_000
sc, views, Ks, W, H = scenes.config4(n=n, n_views=8)
dev = torch.device("cuda")
sc = sc.to(dev)
views, Ks = views.to(dev), Ks.to(dev)
gts = torch.rand(8, H, W, 3, device=dev)
bg = torch.zeros(1, 3, device=dev)
params = [t.requires_grad_() for t in (sc.means, sc.quats, sc.scales, sc.opacities, sc.colors)]


def step(s):
    v = s % 8
    rc, ra, meta = hgs.rasterization(*params, views[v:v + 1], Ks[v:v + 1], W, H, sh_degree=2, render_mode="RGB+ED",
                                     backgrounds=bg)
    meta["means2d"].retain_grad()
    loss = losses.photometric_l1_loss(rc, gts[v:v + 1], ra, w_depth=0.01, w_alpha=0.01)
    loss.backward()
    for p in params:
        p.grad = None


for s in range(5):
    step(s)
torch.cuda.synchronize()
K = 32
t0 = time.perf_counter()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for s in range(K):
    step(s)
e1.record()
t1 = time.perf_counter()
torch.cuda.synchronize()
t2 = time.perf_counter()
print(f"host enqueue {1e3 * (t1 - t0) / K:.3f} ms/step, wall {1e3 * (t2 - t0) / K:.3f} ms/step, gpu {e0.elapsed_time(e1) / K:.3f} ms/step")
# forward only
with torch.no_grad():
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for s in range(K):
        v = s % 8
        hgs.rasterization(*params, views[v:v + 1], Ks[v:v + 1], W, H, sh_degree=2, render_mode="RGB+ED", backgrounds=bg)
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
print(f"forward only: host enqueue {1e3 * (t1 - t0) / K:.3f} ms/view, wall {1e3 * (t2 - t0) / K:.3f} ms/view")
