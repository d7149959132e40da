"""The two pipelines Horizon-GS calls: ``rasterization`` (gaussian_renderer/render.py:40-54) and
``rasterization_2dgs`` (render.py:56-76), with gsplat's Python signatures and outputs.

Every stage is a hand-written sm_100a kernel behind the C ABI (include/hgs_raster.h); this file
only sequences them and owns the tensors.  There is no CPU or PyTorch fallback.
"""
from __future__ import annotations

import math
import weakref
from typing import Dict, Optional, Tuple

import torch
from torch import Tensor

from .cuda import _wrapper as W

_RENDER_MODES = ("RGB", "D", "ED", "RGB+D", "RGB+ED")


def _camera_positions(viewmats: Tensor) -> Tensor:
    """camera centres [C,3] = -R^T t of world->camera matrices (closed form of inverse(viewmats)[:, :3, 3])."""
    R = viewmats[:, :3, :3]
    t = viewmats[:, :3, 3]
    return -(R.transpose(1, 2) @ t[..., None])[..., 0]


def _validate(means, quats, scales, opacities, colors, viewmats, Ks, render_mode, sh_degree, backgrounds):
    N = means.shape[0]
    C = viewmats.shape[0]
    assert means.shape == (N, 3), means.shape
    assert quats.shape == (N, 4), quats.shape
    assert scales.shape == (N, 3), scales.shape
    assert opacities.shape == (N,), opacities.shape
    assert viewmats.shape == (C, 4, 4), viewmats.shape
    assert Ks.shape == (C, 3, 3), Ks.shape
    assert render_mode in _RENDER_MODES, render_mode
    if sh_degree is None:
        # post-activation colours [N,D] or [C,N,D]
        assert (colors.dim() == 2 and colors.shape[0] == N) or (
            colors.dim() == 3 and colors.shape[:2] == (C, N)), colors.shape
    else:
        # SH coefficients [N,K,3]
        assert colors.dim() == 3 and colors.shape[0] == N and colors.shape[2] == 3, colors.shape
        assert (sh_degree + 1) ** 2 <= colors.shape[1], colors.shape
    if backgrounds is not None:
        assert backgrounds.dim() == 2 and backgrounds.shape[0] == C, backgrounds.shape
    if not means.is_cuda:
        raise ValueError("horizongs_b200 runs on CUDA tensors only (no CPU fallback)")
    return C, N


def _per_view_features(means, colors, viewmats, radii, sh_degree, C, vis_ids=None, defer=None, n_vis_dev=None,
                       holder=None, campos=None):
    """-> [C,N,CH] colour features for blending (CH = 3 for SH).  holder["sh_precomputed"]: the colours of the visible
    rows, already evaluated by the projection kernel (the SH node then only carries the backward)."""
    if sh_degree is None:
        if colors.dim() == 2:
            return colors[None] if C == 1 else colors[None].expand(C, -1, -1)
        return colors
    return W._sh_view_colors(sh_degree, means, _camera_positions(viewmats) if campos is None else campos, colors, radii,
                             vis_ids, defer, n_vis_dev, holder)


def _mode_features(feats, depths, backgrounds, render_mode):
    """split into (colour features or None, depth channel or None, backgrounds padded for the depth channel)"""
    if render_mode in ("RGB+D", "RGB+ED"):
        if backgrounds is not None:
            backgrounds = torch.cat([backgrounds, backgrounds.new_zeros(backgrounds.shape[0], 1)], -1)
        return feats, depths, backgrounds
    if render_mode in ("D", "ED"):
        if backgrounds is not None:
            backgrounds = backgrounds.new_zeros(backgrounds.shape[0], 1)
        return depths[..., None], None, backgrounds
    return feats, None, backgrounds


def rasterization(
    means: Tensor, quats: Tensor, scales: Tensor, opacities: Tensor, colors: Tensor, viewmats: Tensor, Ks: Tensor,
    width: int, height: int, near_plane: float = 0.01, far_plane: float = 1e10, radius_clip: float = 0.0,
    eps2d: float = 0.3, sh_degree: Optional[int] = None, packed: bool = False, tile_size: int = 16,
    backgrounds: Optional[Tensor] = None, render_mode: str = "RGB", sparse_grad: bool = False,
    absgrad: bool = False, rasterize_mode: str = "classic", channel_chunk: int = 32, distributed: bool = False,
    camera_model: str = "pinhole", covars: Optional[Tensor] = None,
) -> Tuple[Tensor, Tensor, Dict]:
    """3DGS rasterization with gsplat.rasterization's signature (call site render.py:40-54).

    -> render_colors [C,H,W,3|4|1], render_alphas [C,H,W,1], meta.  meta["means2d"] is a non-leaf
    tensor that accepts retain_grad() and receives pixel-unit gradients (render.py:91,101);
    meta["radii"] is [C,N] int32 (render.py:89).
    """
    if packed or sparse_grad or distributed or covars is not None or camera_model != "pinhole":
        raise NotImplementedError("packed / sparse_grad / distributed / covars / non-pinhole are not supported")
    assert rasterize_mode in ("classic", "antialiased"), rasterize_mode
    if rasterize_mode == "antialiased" and torch.is_grad_enabled() and any(
            t.requires_grad for t in (means, quats, scales)):
        # gsplat back-propagates v_compensations into the covariance; that VJP is not implemented here, so training with
        # it would silently drop a gradient term (the reference only uses 'classic', render.py:40-54)
        raise NotImplementedError("rasterize_mode='antialiased' is forward-only here: the gradient of the opacity "
                                  "compensation w.r.t. means / quats / scales is not implemented (use torch.no_grad(), "
                                  "or rasterize_mode='classic' as the reference does)")
    C, N = _validate(means, quats, scales, opacities, colors, viewmats, Ks, render_mode, sh_degree, backgrounds)
    width, height = int(width), int(height)
    tile_width = math.ceil(width / float(tile_size))
    tile_height = math.ceil(height / float(tile_size))

    holder: Dict = {"park_means_grad": sh_degree is not None, "fuse_bin": True}
    # shading fused into the projection kernel: SH colours (or given [N,3] colours) + the blend records
    campos = None
    if (rasterize_mode == "classic" and not absgrad and render_mode in ("RGB", "RGB+D", "RGB+ED") and colors.is_cuda
            and colors.dtype == torch.float32 and opacities.is_cuda and opacities.dtype == torch.float32
            and ((sh_degree is not None and colors.dim() == 3 and 0 <= sh_degree <= 4)
                 or (sh_degree is None and colors.dim() == 2 and colors.shape[-1] == 3))):
        with torch.no_grad():
            campos = _camera_positions(viewmats.detach()).contiguous() if sh_degree is not None else None
            holder["shade"] = {"mode": -1 if sh_degree is None else int(sh_degree),
                               "feats": colors.detach().contiguous(), "opacities": opacities.detach().contiguous(),
                               "campos": campos, "depth_channel": render_mode != "RGB"}
    radii, means2d, depths, conics, comps, tiles_per_gauss = W._project3d(
        means, quats, scales, viewmats, Ks, width, height, eps2d, near_plane, far_plane, radius_clip,
        rasterize_mode == "antialiased", tile_size, holder)
    opac = opacities[None] if C == 1 else opacities[None].expand(C, -1)
    if comps is not None:
        opac = opac * comps

    # phase 1 of the ordering: everything up to the one host read of (n_visible, n_isects) -- the copy is
    # asynchronous, and the stages that only need the device-side count (SH colours, record packing) are enqueued
    # BEFORE the host waits for it, so the GPU is busy while the host reads the two numbers
    with torch.no_grad():
        if "bin" in holder:         # the compaction / histogram ran inside the projection kernel
            prep = W._isect_scan_async(holder.pop("bin"), C, N, tile_size, tile_width, tile_height)
        else:
            prep = W._isect_prepare_async(means2d, radii, depths, tiles_per_gauss, C, N, tile_size, tile_width,
                                          tile_height)
    vis_full, n_vis_dev = prep["vis_full"], prep["counts"]

    # multi-GPU: the SH / projection backward is deferred and runs fused with the gradient exchange
    defer = W.current_deferred_sink() if torch.is_grad_enabled() else None
    if defer is not None:
        if C != 1 or render_mode not in ("RGB", "RGB+D", "RGB+ED") or absgrad or comps is not None:
            raise NotImplementedError("deferred backward needs one camera per rank, an RGB(+depth) render mode, "
                                      "no absgrad and rasterize_mode='classic'")
        if sh_degree is None and (colors.dim() != 2 or colors.shape[-1] != 3):
            raise NotImplementedError("deferred backward needs [N,3] colours or SH coefficients")
        holder["defer"] = defer
        defer.clear()
        defer.update(viewmats=viewmats.detach().contiguous(), Ks=Ks.detach().contiguous(),
                     campos=_camera_positions(viewmats.detach()).contiguous(), width=width, height=height, eps2d=eps2d,
                     near_plane=near_plane, far_plane=far_plane, sh_degree=sh_degree, n=N)

    shaded = holder.pop("shade_out", None)
    holder.pop("shade", None)
    if shaded is not None and shaded["colors"] is not None:
        holder["sh_precomputed"] = shaded["colors"]
    feats = _per_view_features(means, colors, viewmats, radii, sh_degree, C, vis_full, defer, n_vis_dev, holder, campos)
    feats, depth_ch, bgs = _mode_features(feats, depths, backgrounds, render_mode)
    n_ch = feats.shape[-1] + (1 if depth_ch is not None else 0)
    fuse_norm = render_mode in ("ED", "RGB+ED") and n_ch <= 4 and not absgrad
    records = None if shaded is None else shaded["records"]
    if records is None and n_ch <= 4 and not absgrad:
        records = W._pack3d(W._f32c(means2d, "means2d"), W._f32c(conics, "conics"), W._f32c(feats, "colors"),
                            W._f32c(depth_ch, "depths"), W._f32c(opac, "opacities"), radii, vis_full, n_vis_dev)

    # phase 2: the host read, then emit + tile partition
    with torch.no_grad():
        isect_ids, flatten_ids, isect_offsets, vis_ids = W._isect_finish(
            prep, means2d, radii, depths, C, N, tile_size, tile_width, tile_height)
    holder["vis_ids"] = vis_ids          # work list for the backward of the projection and of the SH stage
    if defer is not None:
        defer["vis_ids"] = vis_ids

    # dense gradients the backward of this call will have to zero-fill (done on a second stream, see W._prezero_async)
    prezero = None
    if C == 1 and torch.is_grad_enabled() and defer is None:
        shapes = {}
        if means.requires_grad:
            shapes["v_means"] = tuple(means.shape)
        if quats.requires_grad and scales.requires_grad and means.requires_grad:
            shapes.update(v_quats=tuple(quats.shape), v_scales=tuple(scales.shape))
        if sh_degree is not None and colors.requires_grad:
            shapes["v_coeffs"] = tuple(colors.shape)
        prezero = {"holder": holder, "shapes": shapes}
    render_colors, render_alphas = W._blend3d(means2d, conics, feats, depth_ch, opac, bgs, width, height, tile_size,
                                              isect_offsets, flatten_ids, absgrad, radii=radii,
                                              normalize_depth=fuse_norm, vis_ids=vis_ids, defer=defer, records=records,
                                              aux=(None if sh_degree is None else
                                                   [weakref.ref(t) for t in (conics, feats, depth_ch) if t is not None]),
                                              prezero=prezero)
    if render_mode in ("ED", "RGB+ED") and not fuse_norm:
        render_colors = torch.cat(
            [render_colors[..., :-1], render_colors[..., -1:] / render_alphas.clamp(min=1e-10)], dim=-1)

    meta = {
        "camera_ids": None, "gaussian_ids": None, "radii": radii, "means2d": means2d, "depths": depths,
        "conics": conics, "opacities": opac, "tile_width": tile_width, "tile_height": tile_height,
        "tiles_per_gauss": tiles_per_gauss, "isect_ids": isect_ids, "flatten_ids": flatten_ids,
        "isect_offsets": isect_offsets, "width": width, "height": height, "tile_size": tile_size, "n_cameras": C,
        "visible_ids": vis_ids,
    }
    return render_colors, render_alphas, meta


class _DensifyProbe(torch.autograd.Function):
    """Identity on means2d; together with _DensifyInject it makes ``meta["means2d"].grad`` carry the
    densification gradient without polluting the gradient that reaches the projection.

    Horizon-GS reads ``info["means2d"].grad`` for densification in BOTH modes (render.py:91,101;
    scene/basic_model.py:131-134).  For surfels the true d(loss)/d(means2d) is almost always zero (means2d
    only enters the rarely-taken screen-space low-pass branch), so -- like the 2DGS reference rasterizer --
    the positional gradient of the ray transform is reported there instead.

        projection -> means2d --Probe--> means2d_info (returned in meta) --Inject--> blend
    backward:  blend stores the pseudo-gradient in ``box``; Inject adds it (so means2d_info.grad = true +
    pseudo); Probe subtracts it again (so the projection receives the true gradient only).
    """

    @staticmethod
    def forward(ctx, means2d, box):
        ctx.box = box
        return means2d.view_as(means2d)

    @staticmethod
    def backward(ctx, v):
        extra = ctx.box.pop("densify", None)
        return (v if extra is None else v - extra), None


class _DensifyInject(torch.autograd.Function):
    @staticmethod
    def forward(ctx, means2d_info, box):
        ctx.box = box
        return means2d_info.view_as(means2d_info)

    @staticmethod
    def backward(ctx, v):
        extra = ctx.box.get("densify", None)
        return (v if extra is None else v + extra), None


def rasterization_2dgs(
    means: Tensor, quats: Tensor, scales: Tensor, opacities: Tensor, colors: Tensor, viewmats: Tensor, Ks: Tensor,
    width: int, height: int, near_plane: float = 0.01, far_plane: float = 1e10, radius_clip: float = 0.0,
    eps2d: float = 0.3, sh_degree: Optional[int] = None, packed: bool = False, tile_size: int = 16,
    backgrounds: Optional[Tensor] = None, render_mode: str = "RGB", sparse_grad: bool = False,
    absgrad: bool = False, distloss: bool = False, depth_mode: str = "expected",
):
    """2DGS rasterization with the signature and NESTED return the reference unpacks at render.py:56-76:
    ((render_colors, render_alphas, render_normals, render_normals_from_depth, render_distort,
      render_median), meta).

    render_normals [C,H,W,3] are world-space; render_normals_from_depth is [H,W,3] for C == 1 (squeezed,
    train.py:184-185 handles both); render_distort is zeros unless distloss=True (lambda_dist is 0 in every
    shipped config).  meta["means2d"].grad receives the densification gradient (see _DensifyProbe);
    meta["gradient_2dgs"] carries the same quantity upstream-gsplat style.
    """
    if packed or sparse_grad or absgrad:
        raise NotImplementedError("packed / sparse_grad / absgrad are not supported")
    assert depth_mode in ("expected", "median"), depth_mode
    C, N = _validate(means, quats, scales, opacities, colors, viewmats, Ks, render_mode, sh_degree, backgrounds)
    width, height = int(width), int(height)
    tile_width = math.ceil(width / float(tile_size))
    tile_height = math.ceil(height / float(tile_size))

    holder: Dict = {"park_means_grad": sh_degree is not None}
    radii, means2d, depths, ray_transforms, normals, tiles_per_gauss = W._project2d(
        means, quats, scales, viewmats, Ks, width, height, near_plane, far_plane, radius_clip, tile_size, holder)
    opac = opacities[None] if C == 1 else opacities[None].expand(C, -1)

    with torch.no_grad():
        isect_ids, flatten_ids, isect_offsets, vis_ids = W._isect_sorted_from_counts(
            means2d, radii, depths, tiles_per_gauss, C, N, tile_size, tile_width, tile_height)
    holder["vis_ids"] = vis_ids

    # multi-GPU: the SH / surfel-projection backward is deferred and runs fused with the gradient exchange
    defer = W.current_deferred_sink() if torch.is_grad_enabled() else None
    if defer is not None:
        if C != 1 or render_mode not in ("RGB", "RGB+D", "RGB+ED"):
            raise NotImplementedError("deferred backward needs one camera per rank and an RGB(+depth) render mode")
        if sh_degree is None and (colors.dim() != 2 or colors.shape[-1] != 3):
            raise NotImplementedError("deferred backward needs [N,3] colours or SH coefficients")
        holder["defer"] = defer
        defer.clear()
        defer.update(vis_ids=vis_ids, viewmats=viewmats.detach().contiguous(), Ks=Ks.detach().contiguous(),
                     campos=_camera_positions(viewmats.detach()).contiguous(), width=width, height=height, eps2d=eps2d,
                     near_plane=near_plane, far_plane=far_plane, sh_degree=sh_degree, n=N)

    feats = _per_view_features(means, colors, viewmats, radii, sh_degree, C, vis_ids, defer, holder=holder)
    feats, depth_ch, bgs = _mode_features(feats, depths, backgrounds, render_mode)
    n_ch = feats.shape[-1] + (1 if depth_ch is not None else 0)
    fuse_norm = render_mode in ("ED", "RGB+ED") and n_ch in (1, 3, 4)

    grad_on = torch.is_grad_enabled() and means2d.requires_grad
    box: Optional[Dict] = {} if grad_on else None
    densify = torch.zeros_like(means2d, requires_grad=True) if grad_on else None
    means2d_info = _DensifyProbe.apply(means2d, box) if grad_on else means2d
    means2d_in = _DensifyInject.apply(means2d_info, box) if grad_on else means2d

    render_colors, render_alphas, render_normals, render_distort, render_median = W._blend2d(
        means2d_in, ray_transforms, feats, depth_ch, normals, opac, densify, bgs, width, height, tile_size,
        isect_offsets, flatten_ids, distloss, box, radii=radii, normalize_depth=fuse_norm, vis_ids=vis_ids, defer=defer)

    render_normals_from_depth = None
    if render_mode in ("ED", "RGB+ED") and not fuse_norm:
        render_colors = torch.cat(
            [render_colors[..., :-1], render_colors[..., -1:] / render_alphas.clamp(min=1e-10)], dim=-1)
    # post-ops (a13): normals to the world frame + normals from the depth map, one kernel each way (csrc/normals.cu)
    if render_mode in ("RGB+D", "RGB+ED"):
        depth_src = render_colors if depth_mode == "expected" else render_median
        render_normals, render_normals_from_depth = W.normals_post(render_normals, depth_src, viewmats, Ks, True)
        render_normals_from_depth = render_normals_from_depth.squeeze(0)
    else:
        render_normals, _ = W.normals_post(render_normals, None, viewmats, Ks, False)

    meta = {
        "camera_ids": None, "gaussian_ids": None, "radii": radii, "means2d": means2d_info, "depths": depths,
        "ray_transforms": ray_transforms, "normals": normals, "opacities": opac, "tile_width": tile_width,
        "tile_height": tile_height, "tiles_per_gauss": tiles_per_gauss, "isect_ids": isect_ids,
        "flatten_ids": flatten_ids, "isect_offsets": isect_offsets, "width": width, "height": height,
        "tile_size": tile_size, "n_cameras": C, "render_distort": render_distort, "gradient_2dgs": densify,
        "visible_ids": vis_ids,
    }
    return (render_colors, render_alphas, render_normals, render_normals_from_depth, render_distort,
            render_median), meta


def depth_to_normal(depths: Tensor, camtoworlds: Tensor, Ks: Tensor) -> Tensor:
    """gsplat.utils.depth_to_normal: z-depth maps [C,H,W,1] -> world-space finite-difference normals
    [C,H,W,3] (one-pixel zero border)."""
    C, H, Wd, _ = depths.shape
    dev, dt = depths.device, depths.dtype
    x, y = torch.meshgrid(torch.arange(Wd, device=dev, dtype=dt), torch.arange(H, device=dev, dtype=dt), indexing="xy")
    fx, fy = Ks[:, 0, 0][:, None, None], Ks[:, 1, 1][:, None, None]
    cx, cy = Ks[:, 0, 2][:, None, None], Ks[:, 1, 2][:, None, None]
    dirs_c = torch.stack([(x[None] - cx + 0.5) / fx, (y[None] - cy + 0.5) / fy, torch.ones_like(x)[None].expand(C, -1, -1)], -1)
    dirs_w = torch.einsum("cij,chwj->chwi", camtoworlds[:, :3, :3], dirs_c)
    pts = camtoworlds[:, None, None, :3, 3] + depths * dirs_w
    dx = pts[:, 2:, 1:-1] - pts[:, :-2, 1:-1]
    dy = pts[:, 1:-1, 2:] - pts[:, 1:-1, :-2]
    n = torch.nn.functional.normalize(torch.cross(dx, dy, dim=-1), dim=-1)
    return torch.nn.functional.pad(n, (0, 0, 1, 1, 1, 1), value=0.0)
