"""Seeded synthetic scenes and cameras of BASELINE.json's configs.

Generated on the CPU with torch.manual_seed-style generators so the oracle (CPU)
and the CUDA path see identical bits; callers move the tensors where they need
them.  Camera conventions follow the reference: world->camera matrices with
OpenCV axes (x right, y down, z forward; scene/dataset_readers.py:358-359,
utils/graphics_utils.py:38-49), wxyz quaternions (utils/general_utils.py:113-134),
post-activation scales/opacities (scene/basic_model.py:328-361).
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Optional

import torch


@dataclass
class Scene:
    means: torch.Tensor      # [N,3]
    quats: torch.Tensor      # [N,4] wxyz
    scales: torch.Tensor     # [N,3] post-activation
    opacities: torch.Tensor  # [N]   post-activation
    colors: torch.Tensor     # [N,3] RGB  or [N,K,3] SH coefficients
    sh_degree: Optional[int]

    def to(self, device):
        return Scene(self.means.to(device), self.quats.to(device), self.scales.to(device),
                     self.opacities.to(device), self.colors.to(device), self.sh_degree)

    @property
    def n(self):
        return self.means.shape[0]


def make_scene(n: int, extent_xy: float, height_z: float, log_scale_mean: float, log_scale_std: float,
               sh_degree: Optional[int] = None, seed: int = 0) -> Scene:
    """means ~ U([-e,e]^2 x [0,h]); log-scales ~ N(log m, s^2); quats ~ normalise(N(0,I));
    opacities = clamp(sigmoid(N(0,1.5^2)), .01, .99); colors ~ U(0,1) or SH coeffs."""
    g = torch.Generator().manual_seed(seed)
    means = torch.rand(n, 3, generator=g)
    means[:, :2] = (means[:, :2] * 2 - 1) * extent_xy
    means[:, 2] = means[:, 2] * height_z
    scales = torch.exp(math.log(log_scale_mean) + log_scale_std * torch.randn(n, 3, generator=g))
    quats = torch.nn.functional.normalize(torch.randn(n, 4, generator=g), dim=-1)
    opacities = torch.sigmoid(1.5 * torch.randn(n, generator=g)).clamp(0.01, 0.99)
    if sh_degree is None:
        colors = torch.rand(n, 3, generator=g)
    else:
        k = (sh_degree + 1) ** 2
        colors = torch.randn(n, k, 3, generator=g) * 0.3
        colors[:, 0, :] = (torch.rand(n, 3, generator=g) - 0.5) / 0.28209479177387814
    return Scene(means, quats, scales, opacities, colors, sh_degree)


def look_at(eye, target, up=(0.0, 0.0, 1.0)) -> torch.Tensor:
    """world->camera [4,4], OpenCV axes (z forward, y down); world z is up."""
    eye = torch.tensor(eye, dtype=torch.float64)
    fwd = torch.tensor(target, dtype=torch.float64) - eye
    fwd = fwd / fwd.norm()
    upv = torch.tensor(up, dtype=torch.float64)
    right = torch.linalg.cross(fwd, upv)
    if right.norm() < 1e-9:                       # looking straight down/up
        right = torch.tensor([1.0, 0.0, 0.0], dtype=torch.float64)
    right = right / right.norm()
    down = torch.linalg.cross(fwd, right)
    R = torch.stack([right, down, fwd], 0)        # rows = camera axes in world
    V = torch.eye(4, dtype=torch.float64)
    V[:3, :3] = R
    V[:3, 3] = -R @ eye
    return V.to(torch.float32)


def intrinsics(width: int, height: int, fov_x_deg: float = 60.0) -> torch.Tensor:
    f = 0.5 * width / math.tan(math.radians(fov_x_deg) / 2)
    return torch.tensor([[f, 0, width / 2], [0, f, height / 2], [0, 0, 1]], dtype=torch.float32)


def aerial_camera(height: float = 12.0, pitch_deg: float = 45.0, yaw_deg: float = 0.0, centre=(0.0, 0.0)):
    """camera at `height` looking down at `pitch_deg` below the horizon toward the scene centre."""
    back = height / math.tan(math.radians(pitch_deg))
    yaw = math.radians(yaw_deg)
    eye = (centre[0] - back * math.cos(yaw), centre[1] - back * math.sin(yaw), height)
    return look_at(eye, (centre[0], centre[1], 0.0))


def street_camera(height: float = 0.3, yaw_deg: float = 0.0, pos=(0.0, 0.0)):
    yaw = math.radians(yaw_deg)
    eye = (pos[0], pos[1], height)
    return look_at(eye, (pos[0] + math.cos(yaw), pos[1] + math.sin(yaw), height))


# ---- the named configurations (indices into BASELINE.json.configs) --------------------
def config0(seed=0, n=100_000):
    """100k 3DGS, one 256x256 camera at (0,0,6) tilted 30 degrees."""
    sc = make_scene(n, 4.0, 1.0, 0.03, 0.5, None, seed)
    t = math.radians(30.0)
    view = look_at((0.0, -6.0 * math.sin(t), 6.0 * math.cos(t)), (0.0, 0.0, 0.0), up=(0.0, 1.0, 0.0))
    return sc, view[None], intrinsics(256, 256)[None], 256, 256


def config1(seed=0, n=1_000_000, view="aerial", width=1920, height=1080, sh_degree=None):
    """Block_small-shaped slab: 1M Gaussians, aerial (h=12, 45 deg) or street (h=0.3) 1080p view."""
    sc = make_scene(n, 10.0, 2.0, 0.02, 0.6, sh_degree, seed)
    v = aerial_camera() if view == "aerial" else street_camera()
    return sc, v[None], intrinsics(width, height)[None], width, height


def config4(seed=0, n=6_000_000, n_views=8, width=1920, height=1080):
    """Block_A-scale: 6M explicit SH2 Gaussians; n_views seeded cameras, alternating aerial/street."""
    sc = make_scene(n, 25.0, 2.0, 0.02, 0.6, 2, seed)
    g = torch.Generator().manual_seed(seed + 1)
    views = []
    for i in range(n_views):
        r = torch.rand(3, generator=g).tolist()
        cx, cy, yaw = (r[0] * 2 - 1) * 12.0, (r[1] * 2 - 1) * 12.0, r[2] * 360.0
        if i % 2 == 0:
            views.append(aerial_camera(12.0, 45.0, yaw, (cx, cy)))
        else:
            views.append(street_camera(0.3, yaw, (cx, cy)))
    Ks = intrinsics(width, height)[None].expand(n_views, -1, -1).contiguous()
    return sc, torch.stack(views, 0), Ks, width, height


def fixed_weight_image(shape, seed=123) -> torch.Tensor:
    """seeded random per-pixel weights so every pixel has a distinct upstream gradient."""
    g = torch.Generator().manual_seed(seed)
    return torch.rand(*shape, generator=g)
