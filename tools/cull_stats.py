import sys, torch
sys.path.insert(0, '/root/repo')
import horizongs_b200 as hgs
from horizongs_b200 import scenes
from horizongs_b200.cuda import _wrapper as Wr
sc, views, Ks, W, H = scenes.config4()
sc = sc.to('cuda'); views = views.cuda(); Ks = Ks.cuda()
with torch.no_grad():
    for v in range(4):
        rc, ra, meta = hgs.rasterization(sc.means, sc.quats, sc.scales, sc.opacities, sc.colors, views[v:v+1], Ks[v:v+1], W, H, sh_degree=2, render_mode="RGB+ED")
        pe, pb = Wr.blend3d_pair_stats(meta["means2d"], meta["conics"], meta["opacities"].contiguous(), meta["radii"], W, H, 16, meta["isect_offsets"], meta["flatten_ids"])
        print(v, 'I', meta["flatten_ids"].numel(), 'P_eval', pe, 'P_blend', pb, Wr.blend3d_pair_stats.last_cull)
