"""Fused anchor -> neural-Gaussian decode of the LOD model (csrc/decode.cu).

``generate_neural_gaussians`` mirrors scene/basic_model.py:297-371 for the configurations Horizon-GS ships
(feat_dim 32, appearance_dim 0, n_offsets <= 16; view_dim 3 with RGB colours, or view_dim 0 with SH colours of degree
<= 3 -- color_attr 'RGB' / 'SH<d>', lod_model.py:58-61): three MLPs on cat(anchor_feat, unit view direction) (or on
anchor_feat alone), the opacity > 0 mask, the compaction and the post-processing are three kernels (count / forward /
backward) with the MLP weights in shared memory, and the kept Gaussians are written straight into the tensors
``rasterization()`` consumes.  Differentiable w.r.t. anchor, anchor_feat, offset, the (post-activation) grid scaling
and all MLP parameters.  CUDA only (no fallback)."""
from __future__ import annotations

import ctypes as C
from typing import Tuple

import torch
from torch import Tensor, nn

from . import _lib
from ._lib import check, ptr


def _mlp_tensors(mlp: nn.Sequential):
    lin = [m for m in mlp.modules() if isinstance(m, nn.Linear)]
    if len(lin) != 2:
        raise NotImplementedError("expected Linear -> ReLU -> Linear [-> activation] (scene/lod_model.py:67-84)")
    return [lin[0].weight, lin[0].bias, lin[1].weight, lin[1].bias]


class _AnchorDecode(torch.autograd.Function):
    @staticmethod
    def forward(ctx, anchor, feat, offset, scaling, cam_center, vis, color_sigmoid, view_dim, color_dim, *mlp):
        L = _lib.lib()
        dev = anchor.device
        V, k, F = int(vis.numel()), int(offset.shape[1]), int(feat.shape[1])
        st = torch.cuda.current_stream().cuda_stream
        mlp = [t.contiguous() for t in mlp]
        mp = (C.c_void_p * 12)(*[t.data_ptr() for t in mlp])
        opac_all = torch.empty((V, k), dtype=torch.float32, device=dev)
        bits = torch.empty(V, dtype=torch.int32, device=dev)
        cnt = torch.empty(V, dtype=torch.int32, device=dev)
        check(L.hgs_decode_count(mp, ptr(anchor), ptr(feat), ptr(cam_center), ptr(vis), V, F, k, view_dim, color_dim,
                                 ptr(opac_all), ptr(bits), ptr(cnt), st), "hgs_decode_count")
        incl = torch.cumsum(cnt, 0, dtype=torch.int64)
        row0 = (incl - cnt).contiguous()
        M = int(incl[-1].item()) if V > 0 else 0            # the one host read (the reference's boolean gather has one too)
        xyz = torch.empty((M, 3), dtype=torch.float32, device=dev)
        color = torch.empty((M, color_dim), dtype=torch.float32, device=dev)
        opacity = torch.empty((M,), dtype=torch.float32, device=dev)
        scales = torch.empty((M, 3), dtype=torch.float32, device=dev)
        quats = torch.empty((M, 4), dtype=torch.float32, device=dev)
        check(L.hgs_decode_fwd(mp, ptr(anchor), ptr(feat), ptr(offset), ptr(scaling), ptr(cam_center), ptr(vis), V, F, k,
                               view_dim, color_dim, int(color_sigmoid), ptr(opac_all), ptr(bits), ptr(row0), ptr(xyz), ptr(color), ptr(opacity),
                               ptr(scales), ptr(quats), st), "hgs_decode_fwd")
        mask = ((bits[:, None] >> torch.arange(k, device=dev, dtype=torch.int32)[None]) & 1).bool().reshape(-1)
        ctx.save_for_backward(anchor, feat, offset, scaling, cam_center, vis, opac_all, bits, row0, *mlp)
        ctx.cfg = (V, k, F, int(color_sigmoid), view_dim, color_dim)
        ctx.mark_non_differentiable(mask)
        return xyz, color, opacity, scales, quats, mask

    @staticmethod
    def backward(ctx, v_xyz, v_color, v_opacity, v_scales, v_quats, _v_mask):
        anchor, feat, offset, scaling, cam_center, vis, opac_all, bits, row0, *mlp = ctx.saved_tensors
        V, k, F, color_sigmoid, view_dim, color_dim = ctx.cfg
        L = _lib.lib()
        dev = anchor.device
        rows = next((t.shape[0] for t in (v_xyz, v_color, v_opacity, v_scales, v_quats) if t is not None), 0)

        def dense(t, shape):
            return torch.zeros(shape, dtype=torch.float32, device=dev) if t is None else t.contiguous()
        v_xyz, v_color = dense(v_xyz, (rows, 3)), dense(v_color, (rows, color_dim))
        v_opacity, v_scales, v_quats = dense(v_opacity, (rows,)), dense(v_scales, (rows, 3)), dense(v_quats, (rows, 4))
        g_anchor, g_feat = torch.zeros_like(anchor), torch.zeros_like(feat)
        g_offset, g_scaling = torch.zeros_like(offset), torch.zeros_like(scaling)
        g_mlp = [torch.zeros_like(t) for t in mlp]
        mp = (C.c_void_p * 12)(*[t.data_ptr() for t in mlp])
        gp = (C.c_void_p * 12)(*[t.data_ptr() for t in g_mlp])
        check(L.hgs_decode_bwd(mp, gp, ptr(anchor), ptr(feat), ptr(offset), ptr(scaling), ptr(cam_center), ptr(vis), V, F, k,
                               view_dim, color_dim, color_sigmoid, ptr(opac_all), ptr(bits), ptr(row0), ptr(v_xyz), ptr(v_color), ptr(v_opacity),
                               ptr(v_scales), ptr(v_quats), ptr(g_anchor), ptr(g_feat), ptr(g_offset), ptr(g_scaling),
                               torch.cuda.current_stream().cuda_stream), "hgs_decode_bwd")
        return (g_anchor, g_feat, g_offset, g_scaling, None, None, None, None, None, *g_mlp)


def generate_neural_gaussians(anchor: Tensor, anchor_feat: Tensor, offset: Tensor, scaling: Tensor, cam_center: Tensor,
                              visible_mask: Tensor, mlp_opacity: nn.Sequential, mlp_cov: nn.Sequential,
                              mlp_color: nn.Sequential, dist2level: str = "floor"
                              ) -> Tuple[Tensor, Tensor, Tensor, Tensor, Tensor, Tensor]:
    """-> (xyz [M,3], color [M,3] (RGB) or [M,K,3] (SH), opacity [M,1], scaling [M,3], rot [M,4], mask [V*k] bool), the tensors
    scene/basic_model.py:297-371 returns (without the pass-through `offsets` / `active_sh_degree`).

    anchor [A,3], anchor_feat [A,32], offset [A,k,3], scaling [A,6] POST-activation (get_scaling = exp(_scaling),
    lod_model.py:182-183), cam_center [3], visible_mask [A] bool; the MLPs are the nn.Sequential modules of
    scene/lod_model.py:67-84 (a trailing nn.Sigmoid on the colour MLP is honoured).  dist2level: the model's level
    mode; 'progressive' (which scales the opacities by _prog_ratio, lod_model.py:215-222) raises."""
    if not anchor.is_cuda:
        raise ValueError("generate_neural_gaussians runs on CUDA tensors only (no CPU fallback)")
    if dist2level == "progressive":
        # scene/lod_model.py:215-222 multiplies the opacities by the per-anchor _prog_ratio in that mode
        # (smooth_complement); the fused decode implements smooth_complement == 1 (basic_model.py:43-44)
        raise NotImplementedError("dist2level='progressive' (opacity x _prog_ratio, lod_model.py:215-222) is not "
                                  "supported by the fused decode; use the PyTorch decode of the reference")
    A, k = anchor.shape[0], offset.shape[1]
    assert anchor_feat.shape == (A, 32), "feat_dim 32 (one lane per hidden unit)"
    assert offset.shape == (A, k, 3) and 1 <= k <= 16 and scaling.shape == (A, 6) and visible_mask.shape == (A,)
    last = lambda seq: [m for m in seq.modules() if not isinstance(m, nn.Sequential)][-1]  # noqa: E731
    if not isinstance(last(mlp_opacity), nn.Tanh):
        raise NotImplementedError("the opacity MLP must end in Tanh (scene/lod_model.py:67-72)")
    mlp = _mlp_tensors(mlp_opacity) + _mlp_tensors(mlp_cov) + _mlp_tensors(mlp_color)
    view_dim = mlp[0].shape[1] - 32
    color_dim = mlp[10].shape[0] // k
    if (view_dim not in (0, 3) or any(mlp[i].shape != (32, 32 + view_dim) for i in (0, 4, 8)) or mlp[2].shape != (k, 32)
            or mlp[6].shape != (7 * k, 32) or mlp[10].shape != (color_dim * k, 32) or color_dim % 3 or not 3 <= color_dim <= 48):
        raise NotImplementedError("supported: feat_dim 32, view_dim 3 or 0, appearance_dim 0, colour_dim 3 (RGB) or "
                                  "3 (d + 1)^2 (SH degree d <= 3)")
    color_sigmoid = isinstance(last(mlp_color), nn.Sigmoid)
    vis = torch.nonzero(visible_mask).flatten().contiguous()                      # int64 work list (one host read)
    xyz, color, opacity, scales, quats, mask = _AnchorDecode.apply(
        anchor.contiguous(), anchor_feat.contiguous(), offset.contiguous(), scaling.contiguous(),
        cam_center.detach().to(torch.float32).contiguous(), vis, color_sigmoid, int(view_dim), int(color_dim), *mlp)
    if color_dim != 3:
        color = color.reshape(color.shape[0], color_dim // 3, 3)      # SH coefficients [M,K,3], basic_model.py:368-369
    return xyz, color, opacity[:, None], scales, quats, mask


@torch.no_grad()
def anchor_visibility(anchor: Tensor, scaling: Tensor, rotation: Tensor, viewmat: Tensor, K: Tensor, width: int,
                      height: int, level: Tensor = None, extra_level: Tensor = None, cam_center: Tensor = None,
                      standard_dist: float = 1.0, fork: float = 2.0, max_level: int = 0, resolution_scale: float = 1.0,
                      dist2level: str = "floor", eps2d: float = 0.3, near_plane: float = 0.01, far_plane: float = 1e10,
                      radius_clip: float = 0.0) -> Tensor:
    """bool [A]: the anchors that pass the LOD level test (scene/lod_model.py:286-290, basic_model.py:192-203;
    skipped when `level` is None) AND the prefilter of gaussian_renderer/render.py:120-197 (projected as Gaussians
    with scales = scaling[:, :3], quats = rotation, radii > 0) -- one kernel (hgs_anchor_filter) instead of the
    elementwise chain, the boolean gather, the projection call and the boolean scatter.
    anchor [A,3], scaling [A,>=3] post-activation, rotation [A,4] (wxyz), viewmat [4,4], K [3,3]."""
    if not anchor.is_cuda:
        raise ValueError("anchor_visibility runs on CUDA tensors only (no CPU fallback)")
    L = _lib.lib()
    A = anchor.shape[0]
    mode = {"floor": 0, "round": 1, "ceil": 2}.get(dist2level)
    if mode is None:
        raise NotImplementedError(f"dist2level '{dist2level}' is not supported (floor / round / ceil)")
    dev = anchor.device
    f32 = lambda t: None if t is None else t.detach().to(device=dev, dtype=torch.float32).contiguous()  # noqa: E731
    anchor, rotation, viewmat, K = f32(anchor), f32(rotation), f32(viewmat), f32(K)
    scaling = scaling.detach()
    if scaling.stride(-1) != 1 or scaling.dtype != torch.float32 or scaling.device != dev:
        scaling = scaling.to(device=dev, dtype=torch.float32).contiguous()
    lvl = None if level is None else level.reshape(A).to(device=dev, dtype=torch.int32).contiguous()
    if lvl is not None and cam_center is None:
        cam_center = torch.linalg.inv(viewmat)[:3, 3]
    out = torch.empty(A, dtype=torch.uint8, device=anchor.device)
    check(L.hgs_anchor_filter(ptr(anchor), ptr(lvl), ptr(f32(extra_level)), ptr(scaling), int(scaling.stride(0)),
                              ptr(rotation), ptr(f32(cam_center)), float(resolution_scale), float(standard_dist),
                              float(fork), int(max_level), mode, ptr(viewmat), ptr(K), A, int(width), int(height),
                              float(eps2d), float(near_plane), float(far_plane), float(radius_clip), ptr(out),
                              torch.cuda.current_stream().cuda_stream), "hgs_anchor_filter")
    return out.bool()
