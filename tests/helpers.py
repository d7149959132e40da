"""Shared test helpers: small seeded scenes, comparison utilities."""
import math

import torch

from horizongs_b200 import scenes


def small_scene(n=2000, seed=0, sh_degree=None, width=160, height=120, C=1, extent=3.0, scale=0.08):
    sc = scenes.make_scene(n, extent, 1.0, scale, 0.5, sh_degree, seed)
    views = []
    for c in range(C):
        ang = 0.5 * c
        eye = (6.0 * math.sin(ang), -5.0 * math.cos(ang), 4.0 + 0.5 * c)
        views.append(scenes.look_at(eye, (0.0, 0.0, 0.3)))
    K = scenes.intrinsics(width, height, 70.0)
    Ks = K[None].expand(C, -1, -1).contiguous()
    return sc, torch.stack(views, 0), Ks, width, height


def rel_err(a, b, floor=1e-8):
    """max |a-b| / (max|b| + floor) -- relative to the tensor's scale (atomic-order tolerant)"""
    return float((a - b).abs().max() / (b.abs().max() + floor))


def to_cuda(*ts):
    return [None if t is None else t.cuda() for t in ts]
