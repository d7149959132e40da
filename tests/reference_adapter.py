"""Test infrastructure: load the REFERENCE's own adapter file, gaussian_renderer/render.py, UNMODIFIED, and run
its render() / prefilter_voxel() against a chosen implementation of the four gsplat names it imports
(render.py:13-14 ``import gsplat``; ``from gsplat.cuda._wrapper import fully_fused_projection,
fully_fused_projection_2dgs``).

* backend "oracle": a module object named ``gsplat`` whose four callables are the CPU oracle's (checker side).
* backend "shim":   ``<repo>/shim`` first on sys.path, so ``import gsplat`` resolves to the drop-in package and
                    the calls land in libhgs_raster.so (product side; needs a CUDA device).

The file is located at /root/reference/gaussian_renderer/render.py (build container) or at
baseline/_ref/gaussian_renderer/render.py (the copy __graft_entry__.build() installs, git-ignored, so that the
GPU box -- where /root/reference does not exist -- can run the same test).

render.py hard-codes ``device="cuda"`` for the intrinsics (:32-36, :131-135) and the 2DGS ``densifications``
(:167-169).  For the oracle backend (CPU tensors) the module's global ``torch`` is replaced by a proxy that maps
that one keyword to the CPU; the source text is not touched.

The mock ``pc`` / ``viewpoint_camera`` objects expose exactly the attributes render.py reads (scene/lod_model.py,
scene/basic_model.py:297-383, scene/cameras.py:91-99) and nothing else.
"""
import importlib
import importlib.util
import math
import os
import sys
import types

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CANDIDATES = ("/root/reference/gaussian_renderer/render.py",
              os.path.join(ROOT, "baseline", "_ref", "gaussian_renderer", "render.py"))


def reference_render_path():
    for p in CANDIDATES:
        if os.path.exists(p):
            return p
    return None


class _TorchOnCpu:
    """``torch`` with device="cuda" mapped to the CPU in the two factory calls render.py makes"""

    def __getattr__(self, name):
        return getattr(torch, name)

    @staticmethod
    def _fix(kw):
        if str(kw.get("device", "")).startswith("cuda"):
            kw["device"] = "cpu"
        return kw

    def tensor(self, *a, **kw):
        return torch.tensor(*a, **self._fix(kw))

    def zeros(self, *a, **kw):
        return torch.zeros(*a, **self._fix(kw))


def _oracle_gsplat_module():
    from oracle import gsplat_oracle as O
    g = types.ModuleType("gsplat")
    g.rasterization, g.rasterization_2dgs = O.rasterization, O.rasterization_2dgs
    g.cuda = types.ModuleType("gsplat.cuda")
    g.cuda._wrapper = types.ModuleType("gsplat.cuda._wrapper")
    g.cuda._wrapper.fully_fused_projection = O.fully_fused_projection
    g.cuda._wrapper.fully_fused_projection_2dgs = O.fully_fused_projection_2dgs
    return {"gsplat": g, "gsplat.cuda": g.cuda, "gsplat.cuda._wrapper": g.cuda._wrapper}


def load_reference_render(backend: str):
    """-> the module object of the reference's render.py, bound to `backend` ("oracle" | "shim")"""
    path = reference_render_path()
    assert path is not None, "reference render.py not found"
    names = ("gsplat", "gsplat.cuda", "gsplat.cuda._wrapper")
    saved = {k: sys.modules.pop(k, None) for k in names}
    shim_dir = os.path.join(ROOT, "shim")
    try:
        if backend == "oracle":
            sys.modules.update(_oracle_gsplat_module())
        else:
            sys.path.insert(0, shim_dir)
            g = importlib.import_module("gsplat")
            assert os.path.abspath(g.__file__).startswith(shim_dir), f"gsplat resolved to {g.__file__}, not the shim"
        spec = importlib.util.spec_from_file_location(f"_reference_render_{backend}", path)
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
    finally:
        if backend != "oracle" and shim_dir in sys.path:
            sys.path.remove(shim_dir)
        for k in names:
            sys.modules.pop(k, None)
            if saved[k] is not None:
                sys.modules[k] = saved[k]
    if backend == "oracle":
        mod.torch = _TorchOnCpu()          # the oracle runs on the CPU, also on the GPU box
    return mod


class Camera:
    """the attributes of scene/cameras.py's Camera that render.py reads"""

    def __init__(self, viewmat, K, width, height, resolution_scale=1.0):
        self.world_view_transform = viewmat.transpose(0, 1).contiguous()       # stored transposed (cameras.py:91)
        self.camera_center = torch.linalg.inv(viewmat)[:3, 3]                  # cameras.py:94
        self.fx, self.fy = float(K[0, 0]), float(K[1, 1])
        self.cx, self.cy = float(K[0, 2]), float(K[1, 2])
        self.image_width, self.image_height = int(width), int(height)
        self.resolution_scale = resolution_scale
        self.uid = 0


class Pipe:
    def __init__(self, add_prefilter=True):
        self.add_prefilter = add_prefilter


class AnchorModel:
    """anchor (neural-Gaussian) branch of render(): what scene/lod_model.py + basic_model.py expose to it"""
    explicit_gs = False

    def __init__(self, tiny, gs_attr="3D", render_mode="RGB+ED"):
        self.m = tiny
        self.gs_attr, self.render_mode = gs_attr, render_mode
        self._anchor_mask = None

    @property
    def get_anchor(self):
        return self.m.anchor

    @property
    def get_scaling(self):
        return torch.exp(self.m.scaling)                                      # basic_model.py scaling_activation

    @property
    def get_rotation(self):
        return self.m.rotation.to(self.m.anchor.device)

    def set_anchor_mask(self, cam_center, resolution_scale):                   # lod_model.py:286-290, floor mode
        m = self.m
        dist = torch.sqrt(torch.sum((m.anchor.detach() - cam_center) ** 2, dim=1)) * resolution_scale
        pred = torch.log2(m.standard_dist / dist) / math.log2(m.fork)
        int_level = torch.clamp(torch.floor(pred).int(), min=0, max=m.levels - 1)
        self._anchor_mask = m.level.to(dist.device) <= int_level

    def generate_neural_gaussians(self, viewpoint_camera, visible_mask=None):  # basic_model.py:297-371
        m = self.m
        xyz, color, opacity, scales, quats = m.decode(viewpoint_camera.camera_center, visible_mask)
        n_all = int(visible_mask.sum()) * m.n_offsets
        sel = torch.ones(n_all, dtype=torch.bool, device=xyz.device)          # mask of kept offsets: unused by render()
        return xyz, None, color, opacity, scales, quats, None, sel


class ExplicitModel:
    """explicit-Gaussian branch of render() (basic_model.py:373-383, lod_model.py:292-296)"""
    explicit_gs = True

    def __init__(self, scene, gs_attr="3D", render_mode="RGB+ED", standard_dist=9.0, fork=2, levels=4, seed=0):
        self.sc = scene
        self.gs_attr, self.render_mode = gs_attr, render_mode
        self.standard_dist, self.fork, self.levels = standard_dist, fork, levels
        g = torch.Generator().manual_seed(seed)
        self._level = torch.randint(0, levels, (scene.n,), generator=g).to(scene.means.device)
        self.params = [t.clone().requires_grad_() for t in (scene.means, scene.colors, scene.opacities[:, None],
                                                            scene.scales, scene.quats)]
        self._gs_mask = None

    def set_gs_mask(self, cam_center, resolution_scale):
        dist = torch.sqrt(torch.sum((self.params[0].detach() - cam_center) ** 2, dim=1)) * resolution_scale
        pred = torch.log2(self.standard_dist / dist) / math.log2(self.fork)
        int_level = torch.clamp(torch.floor(pred).int(), min=0, max=self.levels - 1)
        self._gs_mask = self._level <= int_level

    def generate_explicit_gaussians(self, visible_mask=None):
        xyz, color, opacity, scaling, rot = (p[visible_mask] for p in self.params)
        mask = torch.ones(self.params[0].shape[0], dtype=torch.bool, device=xyz.device)
        return xyz, color, opacity, scaling, rot, self.sc.sh_degree, mask
