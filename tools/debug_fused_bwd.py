import sys, torch
sys.path.insert(0, "/root/repo")
import horizongs_b200 as hgs
from horizongs_b200.cuda import _wrapper as Wr
from tests.helpers import small_scene
sc, V, Ks, Wd, H = small_scene(n=4000, C=1, sh_degree=2, width=160, height=96, scale=0.1)
for retain in (False, True):
    ins = [t.clone().cuda().requires_grad_() for t in (sc.means, sc.quats, sc.scales, sc.opacities, sc.colors)]
    rc, ra, meta = hgs.rasterization(*ins, V.cuda(), Ks.cuda(), Wd, H, sh_degree=2, render_mode="RGB+ED")
    if retain:
        meta["means2d"].retain_grad()
    (rc.sum() + ra.sum()).backward()
    print(retain, dict(Wr.FUSED_BWD_COUNTS))
