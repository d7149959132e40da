"""Generate tests/golden/reference_conventions.npz from the REFERENCE's own Python helpers.

Run in the build container (needs /root/reference, which does not exist on the GPU box):
    python tests/golden/make_golden.py
The hot path's arithmetic lives in gsplat (absent), so the only reference code that can pin the oracle
is the in-tree statement of conventions:
    utils/sh_utils.py:57-112      eval_sh        -> SH basis constants / signs / coefficient order
    utils/general_utils.py:113-145 build_rotation, build_scaling_rotation -> wxyz quaternion, L = R S
    utils/graphics_utils.py:38-49 getWorld2View2 -> world->view matrix (scene/cameras.py:91 stores its transpose)
Two import shims are needed and nothing else is altered: utils/general_utils.py imports matplotlib
(absent here) at module scope, and build_rotation allocates with device='cuda' (no GPU here).
"""
import os
import sys
import types

import numpy as np
import torch

REF = "/root/reference"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_conventions.npz")


def main():
    sys.path.insert(0, REF)
    # shim 1: matplotlib.cm is imported but irrelevant to the functions we call
    mpl = types.ModuleType("matplotlib")
    mpl.cm = types.ModuleType("matplotlib.cm")
    sys.modules.setdefault("matplotlib", mpl)
    sys.modules.setdefault("matplotlib.cm", mpl.cm)
    import utils.general_utils as gu
    import utils.graphics_utils as gr
    import utils.sh_utils as sh

    # shim 2: build_rotation / build_scaling_rotation hard-code device="cuda"
    real_zeros = torch.zeros

    def cpu_zeros(*a, **k):
        k.pop("device", None)
        return real_zeros(*a, **k)

    g = torch.Generator().manual_seed(20261018)
    n = 64
    dirs = torch.nn.functional.normalize(torch.randn(n, 3, generator=g, dtype=torch.float64), dim=-1)
    coeffs_nk3 = torch.randn(n, 25, 3, generator=g, dtype=torch.float64)       # our layout [N,K,3]
    sh_out = {}
    for deg in range(5):
        # reference layout is [..., C, K]
        sh_out[f"sh_deg{deg}"] = sh.eval_sh(deg, coeffs_nk3.permute(0, 2, 1), dirs).numpy()

    quats = torch.randn(n, 4, generator=g)
    scales = torch.rand(n, 3, generator=g) + 0.1
    torch.zeros = cpu_zeros
    try:
        R = gu.build_rotation(quats).numpy()
        L = gu.build_scaling_rotation(scales, quats).numpy()
    finally:
        torch.zeros = real_zeros

    # a camera: R is camera-to-world rotation as stored by the readers, T the world->camera translation
    ang = 0.7
    Rc2w = np.array([[np.cos(ang), 0, np.sin(ang)], [0, 1, 0], [-np.sin(ang), 0, np.cos(ang)]]) @ \
        np.array([[1, 0, 0], [0, np.cos(0.3), -np.sin(0.3)], [0, np.sin(0.3), np.cos(0.3)]])
    T = np.array([0.3, -1.2, 4.0])
    w2v = gr.getWorld2View2(Rc2w, T)

    np.savez(OUT, dirs=dirs.numpy(), coeffs=coeffs_nk3.numpy(), quats=quats.numpy(), scales=scales.numpy(),
             rot=R, scaling_rot=L, cam_R=Rc2w, cam_T=T, world2view=w2v, **sh_out)
    print("wrote", OUT, os.path.getsize(OUT), "bytes")


if __name__ == "__main__":
    main()
