"""Operator layer: the names and call signatures of ``gsplat.cuda._wrapper`` that Horizon-GS
imports (gaussian_renderer/render.py:14 ``from gsplat.cuda._wrapper import fully_fused_projection,
fully_fused_projection_2dgs``) plus the other stage operators, each a torch.autograd.Function over
the C-ABI library (include/hgs_raster.h).  PyTorch is plumbing only: it owns the tensors, the
autograd graph and the stream; all arithmetic is in libhgs_raster.so.

Only the un-packed layout is implemented (the reference passes packed=False at render.py:50,72,158,180).
"""
from __future__ import annotations

import math
from typing import Optional, Tuple

import os

import torch
from torch import Tensor

from .. import _lib
from .._lib import check, ptr

_TILE_SIZES = (16,)


def _stream():
    return torch.cuda.current_stream().cuda_stream


# Optional stage hook for measurement (bench.py): called as hook(stage_name, 0) before and hook(stage_name, 1)
# after the kernels of a stage are enqueued.  None (the default) costs one attribute test per stage.
_STAGE_HOOK = None


def set_stage_hook(fn):
    global _STAGE_HOOK
    _STAGE_HOOK = fn


def _mark(name: str, phase: int):
    if _STAGE_HOOK is not None:
        _STAGE_HOOK(name, phase)


# Deferred per-Gaussian backward (multi-GPU, horizongs_b200.distributed.FusedBackwardExchange): inside
# `with deferred_backward(sink):` a rasterization() call records what the SH / projection backward would need in
# `sink`, and loss.backward() stops after the blend backward -- the SH / projection backward of ALL ranks' views
# then runs fused with the gradient exchange (csrc/exchange_vjp.cu).
_DEFER_SINK = None


class deferred_backward:
    def __init__(self, sink: dict):
        self.sink = sink

    def __enter__(self):
        global _DEFER_SINK
        self._prev, _DEFER_SINK = _DEFER_SINK, self.sink
        return self.sink

    def __exit__(self, *exc):
        global _DEFER_SINK
        _DEFER_SINK = self._prev
        return False


def current_deferred_sink():
    return _DEFER_SINK


_SIDE_STREAMS = {}


def _prezero_async(holder, shapes, dev):
    """Start zero-filling the dense gradient tensors of a rasterization call's backward on a SECOND stream, so that
    the ~1 GB of memsets (6 M Gaussians) runs in the shadow of the compute-bound blend backward instead of in front of
    the SH / projection backward kernels.  holder["zeros"] = {name: tensor}, holder["zeros_event"] = completion."""
    main = torch.cuda.current_stream(dev)
    side = _SIDE_STREAMS.get(dev)
    if side is None:
        side = _SIDE_STREAMS[dev] = torch.cuda.Stream(device=dev)
    bufs = {k: torch.empty(shp, dtype=torch.float32, device=dev) for k, shp in shapes.items()}
    side.wait_stream(main)              # the allocator may hand out memory whose last use is still queued on `main`
    with torch.cuda.stream(side):
        for t in bufs.values():
            t.zero_()
            t.record_stream(side)
    holder["zeros"] = bufs
    holder["zeros_event"] = side.record_event()


def _take_zeros(holder, name, shape):
    """the pre-zeroed tensor `name` of this call (after making the current stream wait for the fill), or None"""
    if holder is None:
        return None
    bufs = holder.get("zeros")
    if not bufs or name not in bufs:
        return None
    t = bufs.pop(name)
    if tuple(t.shape) != tuple(shape):
        return None
    torch.cuda.current_stream(t.device).wait_event(holder["zeros_event"])
    return t


_FUSED_BWD = os.environ.get("HGS_FUSED_BWD", "1") != "0"     # development switch: 0 = three separate kernels
FUSED_BWD_COUNTS = {"proj_direct": 0, "proj_delta": 0, "sh_direct": 0, "sh_delta": 0}   # how the nodes were served


def _same_tensor(a, b) -> bool:
    """is `a` the very gradient tensor `b` we handed to autograd (no other contribution was added to it)?"""
    if a is b:
        return True
    if a is None or b is None:
        return False
    return a.data_ptr() == b.data_ptr() and a.shape == b.shape and a.stride() == b.stride() and a.dtype == b.dtype


def _f32c(t: Optional[Tensor], name: str) -> Optional[Tensor]:
    if t is None:
        return None
    if not t.is_cuda:
        raise ValueError(f"{name} must be a CUDA tensor (this path has no CPU fallback)")
    if t.dtype != torch.float32:
        raise ValueError(f"{name} must be float32, got {t.dtype}")
    return t.contiguous()


def _rows(t: Optional[Tensor], width: int):
    """(tensor to take the pointer of, row stride in floats) for a [..., width] gradient.  Slices of the
    packed blend-gradient buffer ([C,N,12] rows) are passed as they are; anything else is made dense."""
    if t is None:
        return None, width
    if t.dtype == torch.float32 and t.is_cuda and t.dim() >= 2 and t.shape[-1] == width and t.numel() > 0:
        st = t.stride()
        ld = st[-2] if width > 1 or t.dim() >= 2 else 1
        ok = st[-1] == 1 and ld >= width
        # leading dims must be a plain row-major walk over rows of stride ld
        expect = ld
        for size, stride in zip(reversed(t.shape[:-1]), reversed(st[:-1])):
            if size != 1 and stride != expect:
                ok = False
                break
            expect *= size
        if ok:
            return t, int(ld)
    return t.contiguous(), width


def _n_bits(n: int) -> int:
    return int(n).bit_length()  # floor(log2(n)) + 1 for n >= 1


# =====================================================================================
# a3: fully_fused_projection
# =====================================================================================
class _Project3D(torch.autograd.Function):
    @staticmethod
    def forward(ctx, means, quats, scales, viewmats, Ks, width, height, eps2d, near_plane, far_plane, radius_clip,
                calc_compensations, tile_size, holder):
        L = _lib.lib()
        ctx.holder = holder       # rendering.py drops the visible-Gaussian work list here once it is known
        ctx.set_materialize_grads(False)      # no zero tensors for the gradients of the integer outputs
        C, N = viewmats.shape[0], means.shape[0]
        dev = means.device
        radii = torch.empty((C, N), dtype=torch.int32, device=dev)
        means2d = torch.empty((C, N, 2), dtype=torch.float32, device=dev)
        depths = torch.empty((C, N), dtype=torch.float32, device=dev)
        conics = torch.empty((C, N, 3), dtype=torch.float32, device=dev)
        comps = torch.empty((C, N), dtype=torch.float32, device=dev) if calc_compensations else None
        tiles = torch.empty((C, N), dtype=torch.int32, device=dev) if tile_size > 0 else None
        _mark("project3d_fwd", 0)
        if holder is not None and holder.get("fuse_bin") and tile_size > 0:
            # projection + ordered compaction of the visible Gaussians + super-tile histogram in one launch
            tw, th = -(-width // tile_size), -(-height // tile_size)
            vis_ids = torch.empty(C * N, dtype=torch.int32, device=dev)
            counts = torch.empty(3, dtype=torch.int64, device=dev)
            tb = L.hgs_isect_bin_temp_bytes(C * N, C, tw, th)
            temp = torch.empty(tb, dtype=torch.uint8, device=dev)
            # shading fused as well (rendering.py decides): colour evaluation + the 64-byte blend records
            sh = holder.get("shade") if comps is None else None
            mode, feats, sh_k, opac, campos, colors_out, records = -2, None, 0, None, None, None, None
            if sh is not None:
                mode, feats, opac, campos = sh["mode"], sh["feats"], sh["opacities"], sh["campos"]
                sh_k = feats.shape[1] if mode >= 0 else 0
                records = torch.empty(L.hgs_blend3d_pack_bytes(C * N), dtype=torch.uint8, device=dev)
                colors_out = torch.empty((C, N, 3), dtype=torch.float32, device=dev) if mode >= 0 else None
                holder["shade_out"] = {"records": records, "colors": colors_out}
            check(L.hgs_project3d_fwd_bin(ptr(means), ptr(quats), ptr(scales), ptr(viewmats), ptr(Ks), C, N, width, height,
                                          eps2d, near_plane, far_plane, radius_clip, tile_size, ptr(radii), ptr(means2d),
                                          ptr(depths), ptr(conics), ptr(comps), ptr(tiles), ptr(vis_ids), ptr(counts),
                                          ptr(temp), tb, mode, ptr(feats), sh_k, ptr(opac), ptr(campos),
                                          int(bool(sh and sh["depth_channel"])), ptr(colors_out), ptr(records),
                                          _stream()), "hgs_project3d_fwd_bin")
            holder["bin"] = {"vis_full": vis_ids, "counts": counts, "temp": temp}
        else:
            check(L.hgs_project3d_fwd(ptr(means), ptr(quats), ptr(scales), ptr(viewmats), ptr(Ks), C, N, width, height,
                                      eps2d, near_plane, far_plane, radius_clip, max(tile_size, 1), ptr(radii),
                                      ptr(means2d), ptr(depths), ptr(conics), ptr(comps), ptr(tiles), _stream()),
                  "hgs_project3d_fwd")
        _mark("project3d_fwd", 1)
        ctx.save_for_backward(means, quats, scales, viewmats, Ks, radii)
        ctx.cfg = (width, height, eps2d, near_plane, far_plane)
        if holder is not None and C == 1 and comps is None:
            # what the fused per-Gaussian backward (hgs_gauss_bwd_fused, launched by the blend backward) needs
            holder["proj_ctx"] = (means, quats, scales, viewmats, Ks, ctx.cfg)
        ctx.mark_non_differentiable(radii)
        if tiles is not None:
            ctx.mark_non_differentiable(tiles)
        if comps is not None:
            ctx.mark_non_differentiable(comps)  # TODO(antialiased): compensation gradient not implemented
        return radii, means2d, depths, conics, comps, tiles

    @staticmethod
    def backward(ctx, _v_radii, v_means2d, v_depths, v_conics, _v_comps, _v_tiles):
        if ctx.holder is not None and ctx.holder.get("defer") is not None:
            return (None,) * 14           # runs later, fused with the exchange (FusedBackwardExchange.finish)
        means, quats, scales, viewmats, Ks, radii = ctx.saved_tensors
        width, height, eps2d, near_plane, far_plane = ctx.cfg
        L = _lib.lib()
        C, N = radii.shape
        fb = None if ctx.holder is None else ctx.holder.get("fused_bwd")
        if fb is not None:
            # the blend backward already ran this node's kernel (fused with the SH backward) on the gradients it
            # produced.  If autograd added nothing to them, that is the answer; otherwise the map is linear: add the
            # backward of the difference.
            ctx.holder["proj_bwd_done"] = True
            extra = ctx.holder.pop("v_means_sh", None)
            # pop: autograd accumulates a leaf gradient without a copy only if nobody else references the tensor
            v_means, v_quats, v_scales = fb.pop("v_means"), fb.pop("v_quats"), fb.pop("v_scales")
            direct = (_same_tensor(v_means2d, fb["v_means2d"]) and _same_tensor(v_conics, fb["v_conics"])
                      and _same_tensor(v_depths, fb["v_depths"]))
            FUSED_BWD_COUNTS["proj_direct" if direct else "proj_delta"] += 1
            if not direct:
                def diff(a, b, shape):
                    a = torch.zeros(shape, dtype=torch.float32, device=means.device) if a is None else a
                    return (a - b).contiguous() if b is not None else a.contiguous()
                d_m2 = diff(v_means2d, fb["v_means2d"], (C, N, 2))
                d_c = diff(v_conics, fb["v_conics"], (C, N, 3))
                d_d = diff(v_depths, fb["v_depths"], (C, N))
                g_m, g_q, g_s = torch.empty_like(means), torch.empty_like(quats), torch.empty_like(scales)
                check(L.hgs_project3d_bwd(ptr(means), ptr(quats), ptr(scales), ptr(viewmats), ptr(Ks), C, N, width,
                                          height, eps2d, near_plane, far_plane, ptr(radii), ptr(d_m2), 2, ptr(d_d), 1,
                                          ptr(d_c), 3, None, 0, ptr(g_m), ptr(g_q), ptr(g_s), 0, _stream()),
                      "hgs_project3d_bwd")
                v_means, v_quats, v_scales = v_means + g_m, v_quats + g_q, v_scales + g_s
            if extra is not None:
                v_means = v_means + extra
            return (v_means, v_quats, v_scales) + (None,) * 11
        # the SH stage ran its backward first and left its direction gradient for the means here: add to it
        # instead of letting autograd sum two dense [N,3] tensors
        v_means = _take_sh_means_grad(ctx.holder, means)
        acc = v_means is not None
        vis = None if ctx.holder is None else ctx.holder.get("vis_ids")
        zq, zs = _take_zeros(ctx.holder, "v_quats", quats.shape), _take_zeros(ctx.holder, "v_scales", scales.shape)
        zm = None if acc else _take_zeros(ctx.holder, "v_means", means.shape)
        zeroed = vis is not None and zq is not None and zs is not None and (acc or zm is not None)
        if not acc:
            v_means = zm if zeroed else torch.empty_like(means)
        v_quats = zq if zeroed else torch.empty_like(quats)
        v_scales = zs if zeroed else torch.empty_like(scales)
        v_means2d, ld_m = _rows(means.new_zeros((C, N, 2)) if v_means2d is None else v_means2d, 2)
        v_conics, ld_c = _rows(means.new_zeros((C, N, 3)) if v_conics is None else v_conics, 3)
        v_depths, ld_d = _rows(None if v_depths is None else v_depths.unsqueeze(-1), 1)
        _mark("project3d_bwd", 0)
        check(L.hgs_project3d_bwd(ptr(means), ptr(quats), ptr(scales), ptr(viewmats), ptr(Ks), C, N, width, height,
                                  eps2d, near_plane, far_plane, ptr(radii), ptr(v_means2d), ld_m, ptr(v_depths), ld_d,
                                  ptr(v_conics), ld_c, ptr(vis), 0 if vis is None else vis.numel(), ptr(v_means),
                                  ptr(v_quats), ptr(v_scales), int(acc) | (2 if zeroed else 0), _stream()),
              "hgs_project3d_bwd")
        _mark("project3d_bwd", 1)
        return (v_means, v_quats, v_scales) + (None,) * 11


def _take_sh_means_grad(holder, means):
    """the [N,3] direction gradient the SH backward of the same rasterization call parked in `holder` (or None)"""
    if holder is None:
        return None
    holder["proj_bwd_done"] = True
    g = holder.pop("v_means_sh", None)
    if g is not None and (g.shape != means.shape or not g.is_contiguous()):
        raise _lib.HgsError("internal: parked SH gradient has the wrong layout")
    return g


def _check_proj_inputs(means, quats, scales, viewmats, Ks):
    N = means.shape[0]
    C = viewmats.shape[0]
    assert means.shape == (N, 3), means.shape
    assert quats is not None and quats.shape == (N, 4), None if quats is None else quats.shape
    assert scales is not None and scales.shape == (N, 3), None if scales is None else scales.shape
    assert viewmats.shape == (C, 4, 4), viewmats.shape
    assert Ks.shape == (C, 3, 3), Ks.shape
    return C, N


def _project3d(means, quats, scales, viewmats, Ks, width, height, eps2d, near_plane, far_plane, radius_clip,
               calc_compensations, tile_size, holder=None):
    _check_proj_inputs(means, quats, scales, viewmats, Ks)
    means, quats, scales = _f32c(means, "means"), _f32c(quats, "quats"), _f32c(scales, "scales")
    viewmats, Ks = _f32c(viewmats.detach(), "viewmats"), _f32c(Ks.detach(), "Ks")
    return _Project3D.apply(means, quats, scales, viewmats, Ks, int(width), int(height), float(eps2d),
                            float(near_plane), float(far_plane), float(radius_clip), bool(calc_compensations),
                            int(tile_size), holder)


def fully_fused_projection(
    means: Tensor, covars: Optional[Tensor], quats: Optional[Tensor], scales: Optional[Tensor], viewmats: Tensor,
    Ks: Tensor, width: int, height: int, eps2d: float = 0.3, near_plane: float = 0.01, far_plane: float = 1e10,
    radius_clip: float = 0.0, packed: bool = False, sparse_grad: bool = False, calc_compensations: bool = False,
) -> Tuple[Tensor, Tensor, Tensor, Tensor, Optional[Tensor]]:
    """Same call as gsplat.cuda._wrapper.fully_fused_projection at render.py:149-165.

    -> (radii[C,N] int32, means2d[C,N,2], depths[C,N], conics[C,N,3], compensations[C,N] | None).
    radii == 0 marks a culled Gaussian (its other outputs are zeros).
    """
    if covars is not None:
        raise NotImplementedError("covars input is not supported; pass quats and scales (as render.py:151 does)")
    if packed or sparse_grad:
        raise NotImplementedError("packed / sparse_grad are not supported (the reference passes False)")
    radii, means2d, depths, conics, comps, _ = _project3d(
        means, quats, scales, viewmats, Ks, width, height, eps2d, near_plane, far_plane, radius_clip,
        calc_compensations, 0)
    return radii, means2d, depths, conics, comps


# =====================================================================================
# a7: spherical harmonics
# =====================================================================================
class _SphericalHarmonics(torch.autograd.Function):
    """colors[C,N,3] from coeffs[N,K,3]; direction either dirs[C,N,3] or means[N,3]-campos[C,3]."""

    @staticmethod
    def forward(ctx, degree, dirs, means, campos, coeffs, radii, post, vis_ids=None, defer=None, n_vis_dev=None,
                holder=None):
        ctx.defer = defer
        ctx.holder = holder     # rendering.py drops the exact work list here once the host knows its length
        L = _lib.lib()
        N, K = coeffs.shape[0], coeffs.shape[1]
        C = dirs.shape[0] if dirs is not None else campos.shape[0]
        precomputed = None if holder is None else holder.pop("sh_precomputed", None)
        if precomputed is not None:
            # the projection kernel evaluated the colours of the visible rows (hgs_project3d_fwd_bin, shade >= 0);
            # this node only carries the backward
            colors = precomputed
        else:
            colors = torch.empty((C, N, 3), dtype=torch.float32, device=coeffs.device)
            _mark("sh_fwd", 0)
            check(L.hgs_sh_fwd(degree, K, ptr(dirs), ptr(means), ptr(campos), ptr(coeffs), ptr(radii), ptr(vis_ids),
                               0 if vis_ids is None else vis_ids.numel(), ptr(n_vis_dev), C, N, int(post), ptr(colors),
                               _stream()), "hgs_sh_fwd")
            _mark("sh_fwd", 1)
        ctx.vis_ids = vis_ids
        ctx.save_for_backward(dirs, means, campos, coeffs, radii, colors if post else None)
        ctx.cfg = (degree, K, C, N, int(post))
        if holder is not None and post and dirs is None and C == 1:
            # colors.detach(): the output object itself gets this node as grad_fn -- holding it here would close a
            # reference cycle (node -> ctx -> holder -> output -> node) and keep the whole step's tensors alive
            holder["sh_ctx"] = (degree, K, coeffs, means, campos, colors.detach())
        return colors

    @staticmethod
    def backward(ctx, v_colors):
        dirs, means, campos, coeffs, radii, colors = ctx.saved_tensors
        degree, K, C, N, post = ctx.cfg
        if ctx.defer is not None:
            ctx.defer["colors_fwd"] = colors      # the clamp mask of `post`; the gradient itself comes from vpack
            return (None,) * 11
        L = _lib.lib()
        fb = None if ctx.holder is None else ctx.holder.get("fused_bwd")
        delta_of = None
        if fb is not None:
            # computed by the blend backward's fused per-Gaussian kernel from the gradient it produced; the direction
            # part of the mean gradient is already inside the projection node's result
            if _same_tensor(v_colors, fb["v_colors"]):
                FUSED_BWD_COUNTS["sh_direct"] += 1
                return None, None, None, None, fb.pop("v_coeffs"), None, None, None, None, None, None
            FUSED_BWD_COUNTS["sh_delta"] += 1
            delta_of = {"v_coeffs": fb.pop("v_coeffs")}                       # something else was added: backward of the difference (linear map)
            v_colors = (v_colors - fb["v_colors"]).contiguous()
        v_colors, ld_vc = _rows(v_colors, 3)
        need_dirs = dirs is not None and ctx.needs_input_grad[1]
        need_means = means is not None and ctx.needs_input_grad[2]
        v_dirs = torch.empty_like(dirs) if need_dirs else None
        vis = ctx.vis_ids if ctx.holder is None else ctx.holder.get("vis_ids", ctx.vis_ids)
        zc = _take_zeros(ctx.holder, "v_coeffs", coeffs.shape) if (vis is not None and C == 1 and dirs is None) else None
        zm = _take_zeros(ctx.holder, "v_means", means.shape) if (zc is not None and need_means) else None
        zeroed = zc is not None and (zm is not None or not need_means)
        v_coeffs = zc if zeroed else torch.empty_like(coeffs)
        v_means = (zm if zeroed else torch.empty_like(means)) if need_means else None
        _mark("sh_bwd", 0)
        check(L.hgs_sh_bwd(degree, K, ptr(dirs), ptr(means), ptr(campos), ptr(coeffs), ptr(radii), ptr(vis),
                           0 if vis is None else vis.numel(), ptr(colors), ptr(v_colors), ld_vc, C, N, post,
                           ptr(v_coeffs), ptr(v_dirs), ptr(v_means), int(zeroed), _stream()),
              "hgs_sh_bwd")
        _mark("sh_bwd", 1)
        if delta_of is not None:
            v_coeffs = v_coeffs + delta_of["v_coeffs"]
        if v_means is not None and ctx.holder is not None and ctx.holder.get("park_means_grad") and \
                not ctx.holder.get("proj_bwd_done"):
            # the projection backward of this call has not run yet: it will ADD its gradient into this tensor
            # (hgs_project3d_bwd accumulate_means) -- saves autograd's dense [N,3] + [N,3] addition
            ctx.holder["v_means_sh"] = v_means
            v_means = None
        return None, v_dirs, v_means, None, v_coeffs, None, None, None, None, None, None


def spherical_harmonics(degrees_to_use: int, dirs: Tensor, coeffs: Tensor, masks: Optional[Tensor] = None) -> Tensor:
    """gsplat.spherical_harmonics: dirs[..., 3] (un-normalised), coeffs[..., K, 3] -> colors[..., 3].

    Supported shapes: dirs [N,3] or [C,N,3] with coeffs [N,K,3] (shared over C).  masks: bool [..] like dirs[...,0].
    """
    assert (degrees_to_use + 1) ** 2 <= coeffs.shape[-2], coeffs.shape
    assert dirs.shape[-1] == 3 and coeffs.shape[-1] == 3
    if coeffs.dim() != 3:
        raise NotImplementedError("coeffs must be [N,K,3]")
    squeeze = dirs.dim() == 2
    d = _f32c(dirs, "dirs")
    d = d[None] if squeeze else d
    assert d.dim() == 3 and d.shape[1] == coeffs.shape[0], (dirs.shape, coeffs.shape)
    radii = None
    if masks is not None:
        radii = masks.reshape(d.shape[:2]).to(torch.int32).contiguous()
    out = _SphericalHarmonics.apply(int(degrees_to_use), d, None, None, _f32c(coeffs, "coeffs"), radii, False)
    return out[0] if squeeze else out


def _sh_view_colors(sh_degree: int, means: Tensor, campos: Tensor, coeffs: Tensor, radii: Tensor,
                    vis_ids: Optional[Tensor] = None, defer: Optional[dict] = None, n_vis_dev: Optional[Tensor] = None,
                    holder: Optional[dict] = None) -> Tensor:
    """fused path used by rasterization*: clamp_min(SH(means - campos) + 0.5, 0), masked by radii > 0
    (vis_ids = the work list of visible flat indices, equivalent to the mask but without idle threads)."""
    return _SphericalHarmonics.apply(int(sh_degree), None, _f32c(means, "means"), _f32c(campos, "campos"),
                                     _f32c(coeffs, "colors"), radii, True, vis_ids, defer, n_vis_dev, holder)


# =====================================================================================
# a8-a10: tile intersection / sort / offsets
# =====================================================================================
_PINNED_RING = {}          # device -> (list of pinned [2] int64 buffers, next slot): one buffer per call in flight
_PINNED_SLOTS = 16


def _pinned_counts(dev):
    """a pinned 16-byte host buffer for this call's (n_visible, n_isects); rotating ring per device, so that several
    rasterizations in flight (other streams / threads) never share one"""
    ring = _PINNED_RING.get(dev)
    if ring is None:
        ring = _PINNED_RING[dev] = [[torch.empty(3, dtype=torch.int64).pin_memory() for _ in range(_PINNED_SLOTS)], 0]
    buf = ring[0][ring[1] % _PINNED_SLOTS]
    ring[1] += 1
    return buf


def _isect_prepare_async(means2d, radii, depths, tiles_per_gauss, C, N, tile_size, tile_width, tile_height):
    """phase 1 of the sorted path: ordered compaction of the visible Gaussians + (camera, tile) histogram + its scan
    (= the per-tile ranges) on the device, and an ASYNCHRONOUS copy of (n_visible, n_isects) into pinned host memory.
    Work that only needs the device-side count (SH colours, record packing) can be enqueued before _isect_finish()
    makes the host wait for the two numbers."""
    L = _lib.lib()
    dev = means2d.device
    CN = C * N
    st = _stream()
    vis_ids = torch.empty(CN, dtype=torch.int32, device=dev)
    counts = torch.empty(3, dtype=torch.int64, device=dev)
    _mark("isect_prepare", 0)
    tb = L.hgs_isect_bin_temp_bytes(CN, C, tile_width, tile_height)
    temp = torch.empty(tb, dtype=torch.uint8, device=dev)
    check(L.hgs_isect_bin_prepare(ptr(means2d), ptr(radii), ptr(depths), ptr(tiles_per_gauss), C, N, tile_size, tile_width,
                                  tile_height, ptr(vis_ids), ptr(counts), ptr(temp), tb, st),
          "hgs_isect_bin_prepare")
    _mark("isect_prepare", 1)
    host = _pinned_counts(dev)
    host.copy_(counts, non_blocking=True)
    ev = torch.cuda.Event()
    ev.record()
    return {"vis_full": vis_ids, "counts": counts, "host": host, "event": ev, "temp": temp}


def _isect_scan_async(prep, C, N, tile_size, tile_width, tile_height):
    """phase 1 when its first launch ran fused in the projection (hgs_project3d_fwd_bin): the scan of the super-tile
    histogram, then the asynchronous copy of the counts as in _isect_prepare_async"""
    L = _lib.lib()
    temp, counts = prep["temp"], prep["counts"]
    _mark("isect_prepare", 0)
    check(L.hgs_isect_bin_scan(C, N, tile_size, tile_width, tile_height, ptr(counts), ptr(temp), temp.numel(), _stream()),
          "hgs_isect_bin_scan")
    _mark("isect_prepare", 1)
    host = _pinned_counts(counts.device)
    host.copy_(counts, non_blocking=True)
    ev = torch.cuda.Event()
    ev.record()
    prep.update(host=host, event=ev)
    return prep


_ISECT_CAPS = {}            # (device, C, N, tile_width, tile_height) -> [capacity of I, capacity of super-tile keys]
_CAP_SLACK = 1.5


def _isect_finish(prep, means2d, radii, depths, C, N, tile_size, tile_width, tile_height):
    """phase 2: scatter into the super-tile ranges + per-super-tile sort + expansion into tile ranges, and the one host
    read of (n_visible, n_isects, n_super) that sizes what the caller sees.

    The first call of a (device, C, N, tile grid) reads the counts, then allocates exactly and launches.  Later calls
    launch BEFORE the read, into buffers of 1.5 x the largest counts seen so far -- the device then has the whole
    ordering stage queued while the host waits for the three numbers -- and return exact-length views.  If the counts
    turn out larger than the guess the kernels did nothing (hgs_isect_bin_sorted's capacity rule) and the call is
    repeated with exact sizes."""
    L = _lib.lib()
    dev = means2d.device
    st = _stream()
    temp, counts = prep["temp"], prep["counts"]
    key = (dev, C, N, tile_width, tile_height)
    offsets = torch.empty((C, tile_height, tile_width), dtype=torch.int32, device=dev)

    def launch(cap_i, cap_s):
        ids = torch.empty(cap_i, dtype=torch.int64, device=dev)
        flat = torch.empty(cap_i, dtype=torch.int32, device=dev)
        bb = L.hgs_isect_bin_bucket_bytes(cap_s)
        bucket = torch.empty(bb, dtype=torch.uint8, device=dev)
        check(L.hgs_isect_bin_sorted(ptr(counts), C, N, C * N, cap_i, cap_s, tile_size, tile_width, tile_height,
                                     ptr(offsets), ptr(ids), ptr(flat), ptr(temp), temp.numel(), ptr(bucket), bb, st),
              "hgs_isect_bin_sorted")
        return ids, flat

    caps = _ISECT_CAPS.get(key)
    ids = flat = None
    if caps is not None:
        _mark("isect_sorted", 0)
        ids, flat = launch(caps[0], caps[1])
        _mark("isect_sorted", 1)
    prep["event"].synchronize()
    n_visible, n_isects, n_super = prep["host"].tolist()
    if n_isects >= 2 ** 31:
        raise _lib.HgsError(f"{n_isects} tile intersections exceed the 32-bit index range")
    if caps is None or n_isects > caps[0] or n_super > caps[1]:
        _mark("isect_sorted", 0)
        ids, flat = launch(n_isects, n_super)
        _mark("isect_sorted", 1)
    want = (int(n_isects * _CAP_SLACK) + 1, int(n_super * _CAP_SLACK) + 1)
    if caps is None:
        _ISECT_CAPS[key] = [want[0], want[1]]
    else:
        caps[0], caps[1] = max(caps[0], want[0]), max(caps[1], want[1])
    return ids[:n_isects], flat[:n_isects], offsets, prep["vis_full"][:n_visible]


def _isect_sorted_from_counts(means2d, radii, depths, tiles_per_gauss, C, N, tile_size, tile_width, tile_height):
    """compaction + tile histogram + ranges, (one D2H read of I), scatter + per-tile sort.  All int work in
    libhgs_raster."""
    prep = _isect_prepare_async(means2d, radii, depths, tiles_per_gauss, C, N, tile_size, tile_width, tile_height)
    return _isect_finish(prep, means2d, radii, depths, C, N, tile_size, tile_width, tile_height)


@torch.no_grad()
def isect_tiles(means2d: Tensor, radii: Tensor, depths: Tensor, tile_size: int, tile_width: int, tile_height: int,
                sort: bool = True, packed: bool = False, n_cameras: Optional[int] = None,
                camera_ids: Optional[Tensor] = None, gaussian_ids: Optional[Tensor] = None,
                _with_offsets: bool = False):
    """gsplat.isect_tiles -> (tiles_per_gauss[C,N] i32, isect_ids[I] i64, flatten_ids[I] i32)."""
    if packed:
        raise NotImplementedError("packed layout is not supported")
    C, N = radii.shape
    assert means2d.shape == (C, N, 2) and depths.shape == (C, N)
    L = _lib.lib()
    dev = means2d.device
    means2d, depths = _f32c(means2d, "means2d"), _f32c(depths, "depths")
    radii = radii.to(torch.int32).contiguous()
    tiles = torch.empty((C, N), dtype=torch.int32, device=dev)
    st = _stream()
    check(L.hgs_isect_count(ptr(means2d), ptr(radii), C * N, tile_size, tile_width, tile_height, ptr(tiles), st),
          "hgs_isect_count")
    if sort:
        isect_ids, flatten_ids, offsets, _ = _isect_sorted_from_counts(
            means2d, radii, depths, tiles, C, N, tile_size, tile_width, tile_height)
        if _with_offsets:
            return tiles, isect_ids, flatten_ids, offsets
        return tiles, isect_ids, flatten_ids
    cum = torch.empty(C * N, dtype=torch.int32, device=dev)
    total = torch.empty(1, dtype=torch.int64, device=dev)
    tb = L.hgs_scan_temp_bytes(C * N)
    temp = torch.empty(tb, dtype=torch.uint8, device=dev)
    check(L.hgs_exclusive_scan_i32(ptr(tiles), ptr(cum), ptr(total), C * N, ptr(temp), tb, st), "hgs_exclusive_scan")
    n_isects = int(total.item())
    isect_ids = torch.empty(n_isects, dtype=torch.int64, device=dev)
    flatten_ids = torch.empty(n_isects, dtype=torch.int32, device=dev)
    check(L.hgs_isect_emit(ptr(means2d), ptr(radii), ptr(depths), ptr(cum), C, N, tile_size, tile_width, tile_height,
                           ptr(isect_ids), ptr(flatten_ids), st), "hgs_isect_emit")
    return tiles, isect_ids, flatten_ids


@torch.no_grad()
def isect_offset_encode(isect_ids: Tensor, n_cameras: int, tile_width: int, tile_height: int) -> Tensor:
    """gsplat.isect_offset_encode -> offsets[C, tile_height, tile_width] int32."""
    L = _lib.lib()
    isect_ids = isect_ids.contiguous()
    assert isect_ids.dtype == torch.int64 and isect_ids.is_cuda
    offsets = torch.empty((n_cameras, tile_height, tile_width), dtype=torch.int32, device=isect_ids.device)
    check(L.hgs_isect_offset_encode(ptr(isect_ids), isect_ids.numel(), n_cameras, tile_width, tile_height,
                                    ptr(offsets), _stream()), "hgs_isect_offset_encode")
    return offsets


# =====================================================================================
# a11: rasterize_to_pixels (3DGS)
# =====================================================================================
class _Blend3D(torch.autograd.Function):
    """rasterize_to_pixels.  <= 4 channels: packed-record fast kernels (TMA-staged, per-warp culling,
    fused expected-depth normalisation); 5..8 channels: plain kernels."""

    @staticmethod
    def forward(ctx, means2d, conics, colors, depths, opacities, backgrounds, width, height, tile_size,
                isect_offsets, flatten_ids, absgrad, radii, normalize_depth, vis_ids=None, defer=None, records=None,
                aux=None, prezero=None):
        L = _lib.lib()
        ctx.defer = defer
        ctx.vis_ids = vis_ids
        ctx.prezero = prezero   # {"holder": dict shared with the projection / SH stages, "shapes": their dense gradients}
        ctx.aux = aux           # weak references to the per-Gaussian inputs (does anybody retain their gradient?)
        C, N = opacities.shape
        CH = colors.shape[-1]
        D = CH + (1 if depths is not None else 0)
        dev = means2d.device
        render_colors = torch.empty((C, height, width, D), dtype=torch.float32, device=dev)
        render_alphas = torch.empty((C, height, width, 1), dtype=torch.float32, device=dev)
        last_ids = torch.empty((C, height, width), dtype=torch.int32, device=dev)
        fast = D <= 4 and not absgrad
        ctx.fast = fast
        ctx.cfg = (width, height, tile_size, absgrad, bool(normalize_depth), CH, D)
        st = _stream()
        if fast:
            if records is None:
                records = _pack3d(means2d, conics, colors, depths, opacities, radii, vis_ids, None)
            _mark("blend3d_fwd", 0)
            check(L.hgs_blend3d_fwd_packed(ptr(records), ptr(backgrounds), C, D, int(normalize_depth), width, height,
                                           tile_size, ptr(isect_offsets), ptr(flatten_ids), flatten_ids.numel(),
                                           ptr(render_colors), ptr(render_alphas), ptr(last_ids), st),
                  "hgs_blend3d_fwd_packed")
            _mark("blend3d_fwd", 1)
            ctx.save_for_backward(records, backgrounds, isect_offsets, flatten_ids, render_colors, render_alphas,
                                  last_ids)
            ctx.shapes = (means2d.shape, depths is not None)
        else:
            check(L.hgs_blend3d_fwd(ptr(means2d), ptr(conics), ptr(colors), ptr(depths), ptr(opacities),
                                    ptr(backgrounds), C, N, CH, width, height, tile_size, ptr(isect_offsets),
                                    ptr(flatten_ids), flatten_ids.numel(), ptr(render_colors), ptr(render_alphas),
                                    ptr(last_ids), st), "hgs_blend3d_fwd")
            if normalize_depth:
                raise NotImplementedError("fused depth normalisation needs <= 4 channels")
            ctx.save_for_backward(means2d, conics, colors, depths, opacities, backgrounds, isect_offsets, flatten_ids,
                                  render_alphas, last_ids)
        return render_colors, render_alphas

    @staticmethod
    def backward(ctx, v_render_colors, v_render_alphas):
        width, height, tile_size, absgrad, normalize_depth, CH, D = ctx.cfg
        L = _lib.lib()
        v_render_colors = v_render_colors.contiguous()
        v_render_alphas = v_render_alphas.contiguous()
        tail = (None,) * 13
        if ctx.fast:
            records, backgrounds, isect_offsets, flatten_ids, render_colors, render_alphas, last_ids = ctx.saved_tensors
            (C, N, _), has_depth = ctx.shapes
            vpack = _vpack_alloc(C, N, 12, ctx.vis_ids, ctx.aux, records.device)
            pz = ctx.prezero
            if pz is not None and ctx.vis_ids is not None and ctx.defer is None:
                # dense zero fills of this backward pass: on a second stream, in the shadow of the blend backward
                shapes = dict(pz["shapes"])
                shapes.update(v_means2d=(C, N, 2), v_opacities=(C, N))
                _prezero_async(pz["holder"], shapes, records.device)
            _mark("blend3d_bwd", 0)
            check(L.hgs_blend3d_bwd_packed(ptr(records), ptr(backgrounds), C, D, int(normalize_depth), width, height,
                                           tile_size, ptr(isect_offsets), ptr(flatten_ids), flatten_ids.numel(),
                                           ptr(render_colors), ptr(render_alphas), ptr(last_ids),
                                           ptr(v_render_colors), ptr(v_render_alphas), ptr(vpack), _stream()),
                  "hgs_blend3d_bwd_packed")
            _mark("blend3d_bwd", 1)
            if ctx.defer is not None:
                ctx.defer["vpack"] = vpack
            v_means2d, v_conics, v_opacities = vpack[..., 0:2], vpack[..., 2:5], vpack[..., 5]
            v_colors = vpack[..., 8:8 + CH]
            v_depths = vpack[..., 8 + CH] if has_depth else None
            hold = None if ctx.prezero is None else ctx.prezero["holder"]
            fused = None
            if (_FUSED_BWD and ctx.vis_ids is not None and ctx.defer is None and hold is not None and C == 1 and CH == 3
                    and "sh_ctx" in hold and "proj_ctx" in hold
                    and all(k in hold.get("zeros", {}) for k in ("v_means2d", "v_opacities", "v_coeffs", "v_means",
                                                                 "v_quats", "v_scales"))):
                # one pass over the visible Gaussians' rows does the dense unpack, the SH backward and the projection
                # backward (hgs_gauss_bwd_fused); the SH / projection nodes then only hand out what is computed here
                degree, K, coeffs, means, campos, colors_fwd = hold["sh_ctx"]
                p_means, quats, scales, viewmats, Ks, (pw, ph, eps2d, near_plane, far_plane) = hold["proj_ctx"]
                if p_means.data_ptr() == means.data_ptr():
                    _mark("zeros_wait", 0)
                    z = {k: _take_zeros(hold, k, shp) for k, shp in (
                        ("v_means2d", (C, N, 2)), ("v_opacities", (C, N)), ("v_coeffs", coeffs.shape),
                        ("v_means", means.shape), ("v_quats", quats.shape), ("v_scales", scales.shape))}
                    _mark("zeros_wait", 1)
                    if all(t is not None for t in z.values()):
                        _mark("gauss_bwd", 0)
                        check(L.hgs_gauss_bwd_fused(ptr(vpack), int(has_depth), ptr(ctx.vis_ids), ctx.vis_ids.numel(), N,
                                                    ptr(means), ptr(quats), ptr(scales), ptr(viewmats), ptr(Ks), pw, ph,
                                                    eps2d, near_plane, far_plane, degree, K, ptr(campos), ptr(coeffs),
                                                    ptr(colors_fwd), ptr(z["v_means2d"]), ptr(z["v_opacities"]),
                                                    ptr(z["v_coeffs"]), ptr(z["v_means"]), ptr(z["v_quats"]),
                                                    ptr(z["v_scales"]), _stream()), "hgs_gauss_bwd_fused")
                        _mark("gauss_bwd", 1)
                        v_means2d, v_opacities = z["v_means2d"], z["v_opacities"]
                        fused = dict(z, v_colors=v_colors, v_conics=v_conics, v_depths=v_depths)
                        del fused["v_opacities"], z
                        hold["fused_bwd"] = fused
            if fused is None and ctx.vis_ids is not None and ctx.defer is None:
                # autograd consumes these two as dense tensors (retain_grad clone, leaf accumulation): copy the visible
                # rows into dense zero-filled tensors instead of handing out strided views of the 48-byte rows
                _mark("zeros_wait", 0)
                z2, zo = _take_zeros(hold, "v_means2d", (C, N, 2)), _take_zeros(hold, "v_opacities", (C, N))
                _mark("zeros_wait", 1)
                zeroed = z2 is not None and zo is not None
                v_means2d = z2 if zeroed else torch.empty((C, N, 2), dtype=torch.float32, device=records.device)
                v_opacities = zo if zeroed else torch.empty((C, N), dtype=torch.float32, device=records.device)
                check(L.hgs_blend3d_unpack(ptr(vpack), ptr(ctx.vis_ids), ctx.vis_ids.numel(), C * N, ptr(v_means2d),
                                           ptr(v_opacities), int(zeroed), _stream()), "hgs_blend3d_unpack")
            v_bg = None
            if backgrounds is not None and ctx.needs_input_grad[5]:
                vrc = v_render_colors
                if normalize_depth:
                    vrc = torch.cat([vrc[..., :-1], vrc[..., -1:] / render_alphas.clamp(min=1e-10)], -1)
                v_bg = (vrc * (1.0 - render_alphas)).sum(dim=(1, 2))
            return (v_means2d, v_conics, v_colors, v_depths, v_opacities, v_bg) + tail
        (means2d, conics, colors, depths, opacities, backgrounds, isect_offsets, flatten_ids, render_alphas,
         last_ids) = ctx.saved_tensors
        C, N = opacities.shape
        v_means2d = torch.zeros_like(means2d)
        v_conics = torch.zeros_like(conics)
        v_colors = torch.zeros_like(colors)
        v_depths = torch.zeros_like(depths) if depths is not None else None
        v_opacities = torch.zeros_like(opacities)
        v_abs = torch.zeros_like(means2d) if absgrad else None
        check(L.hgs_blend3d_bwd(ptr(means2d), ptr(conics), ptr(colors), ptr(depths), ptr(opacities), ptr(backgrounds),
                                C, N, CH, width, height, tile_size, ptr(isect_offsets), ptr(flatten_ids),
                                flatten_ids.numel(), ptr(render_alphas), ptr(last_ids), ptr(v_render_colors),
                                ptr(v_render_alphas), ptr(v_means2d), ptr(v_abs), ptr(v_conics), ptr(v_colors),
                                ptr(v_depths), ptr(v_opacities), _stream()), "hgs_blend3d_bwd")
        if absgrad:
            means2d.absgrad = v_abs
        v_bg = None
        if backgrounds is not None and ctx.needs_input_grad[5]:
            v_bg = (v_render_colors * (1.0 - render_alphas)).sum(dim=(1, 2))
        return (v_means2d, v_conics, v_colors, v_depths, v_opacities, v_bg) + tail


@torch.no_grad()
def _pack3d(means2d, conics, colors, depths, opacities, radii, vis_ids, n_vis_dev):
    """64-byte blend records of the visible Gaussians (hgs_blend3d_pack).  With n_vis_dev (device-side length of the
    work list) the call can be enqueued before the host knows how many Gaussians are visible."""
    L = _lib.lib()
    C, N = opacities.shape
    CH = colors.shape[-1]
    records = torch.empty(L.hgs_blend3d_pack_bytes(C * N), dtype=torch.uint8, device=means2d.device)
    _mark("blend3d_pack", 0)
    check(L.hgs_blend3d_pack(ptr(means2d), ptr(conics), ptr(colors), ptr(depths), ptr(opacities), ptr(radii),
                             ptr(vis_ids), 0 if vis_ids is None else vis_ids.numel(), ptr(n_vis_dev), C * N, CH,
                             ptr(records), _stream()), "hgs_blend3d_pack")
    _mark("blend3d_pack", 1)
    return records


def _vpack_alloc(C, N, width, vis_ids, aux, dev):
    """the packed gradient accumulator of a blend backward.  When every consumer of its rows (projection / SH
    backward, the dense unpack, the fused exchange) goes through the work list of visible Gaussians, only those rows
    are zeroed (hgs_zero_rows).  The caller says so by passing `aux`, the weak references of the per-Gaussian blend
    inputs that stay reachable (meta["conics"], meta["depths"]): aux is None when the colours are the caller's own
    tensor (their gradient is a dense view of the rows), and a live input with retain_grad() could be inspected at
    the rows of culled Gaussians -- in both cases the whole buffer is zero-filled."""
    dense = vis_ids is None or aux is None or any(r() is not None and r().retains_grad for r in aux)
    if dense:
        return torch.zeros((C, N, width), dtype=torch.float32, device=dev)
    vpack = torch.empty((C, N, width), dtype=torch.float32, device=dev)
    check(_lib.lib().hgs_zero_rows(ptr(vpack), width, ptr(vis_ids), vis_ids.numel(), _stream()), "hgs_zero_rows")
    return vpack


def _blend3d(means2d, conics, colors, depths, opacities, backgrounds, width, height, tile_size, isect_offsets,
             flatten_ids, absgrad=False, radii=None, normalize_depth=False, vis_ids=None, defer=None, records=None,
             aux=None, prezero=None):
    if tile_size not in _TILE_SIZES:
        raise NotImplementedError(f"tile_size {tile_size} is not supported (supported: {_TILE_SIZES})")
    D = colors.shape[-1] + (1 if depths is not None else 0)
    if D > 8:
        raise NotImplementedError(f"{D} render channels requested; at most 8 are supported")
    return _Blend3D.apply(_f32c(means2d, "means2d"), _f32c(conics, "conics"), _f32c(colors, "colors"),
                          _f32c(depths, "depths"), _f32c(opacities, "opacities"), _f32c(backgrounds, "backgrounds"),
                          int(width), int(height), int(tile_size), isect_offsets.contiguous(),
                          flatten_ids.contiguous(), bool(absgrad), radii, bool(normalize_depth), vis_ids, defer, records,
                          aux, prezero)


def rasterize_to_pixels(means2d: Tensor, conics: Tensor, colors: Tensor, opacities: Tensor, image_width: int,
                        image_height: int, tile_size: int, isect_offsets: Tensor, flatten_ids: Tensor,
                        backgrounds: Optional[Tensor] = None, masks: Optional[Tensor] = None, packed: bool = False,
                        absgrad: bool = False) -> Tuple[Tensor, Tensor]:
    """gsplat.rasterize_to_pixels: means2d[C,N,2], conics[C,N,3], colors[C,N,D], opacities[C,N]
    -> (render_colors[C,H,W,D], render_alphas[C,H,W,1])."""
    if packed or masks is not None:
        raise NotImplementedError("packed layout / tile masks are not supported")
    C, N = opacities.shape
    assert means2d.shape == (C, N, 2) and conics.shape == (C, N, 3) and colors.shape[:2] == (C, N)
    return _blend3d(means2d, conics, colors, None, opacities, backgrounds, image_width, image_height, tile_size,
                    isect_offsets, flatten_ids, absgrad)


# =====================================================================================
# a4: fully_fused_projection_2dgs
# =====================================================================================
class _Project2D(torch.autograd.Function):
    @staticmethod
    def forward(ctx, means, quats, scales, viewmats, Ks, width, height, near_plane, far_plane, radius_clip,
                tile_size, holder):
        L = _lib.lib()
        ctx.holder = holder
        ctx.set_materialize_grads(False)
        C, N = viewmats.shape[0], means.shape[0]
        dev = means.device
        radii = torch.empty((C, N), dtype=torch.int32, device=dev)
        means2d = torch.empty((C, N, 2), dtype=torch.float32, device=dev)
        depths = torch.empty((C, N), dtype=torch.float32, device=dev)
        ray_transforms = torch.empty((C, N, 3, 3), dtype=torch.float32, device=dev)
        normals = torch.empty((C, N, 3), dtype=torch.float32, device=dev)
        tiles = torch.empty((C, N), dtype=torch.int32, device=dev) if tile_size > 0 else None
        check(L.hgs_project2d_fwd(ptr(means), ptr(quats), ptr(scales), ptr(viewmats), ptr(Ks), C, N, width, height,
                                  near_plane, far_plane, radius_clip, max(tile_size, 1), ptr(radii), ptr(means2d),
                                  ptr(depths), ptr(ray_transforms), ptr(normals), ptr(tiles), _stream()),
              "hgs_project2d_fwd")
        ctx.save_for_backward(means, quats, scales, viewmats, Ks, radii)
        ctx.cfg = (width, height, near_plane, far_plane)
        ctx.mark_non_differentiable(radii)
        if tiles is not None:
            ctx.mark_non_differentiable(tiles)
        return radii, means2d, depths, ray_transforms, normals, tiles

    @staticmethod
    def backward(ctx, _v_radii, v_means2d, v_depths, v_ray_transforms, v_normals, _v_tiles):
        if ctx.holder is not None and ctx.holder.get("defer") is not None:
            return (None,) * 12           # runs later, fused with the exchange (FusedBackwardExchange.finish)
        means, quats, scales, viewmats, Ks, radii = ctx.saved_tensors
        width, height, near_plane, far_plane = ctx.cfg
        L = _lib.lib()
        C, N = radii.shape
        v_means = _take_sh_means_grad(ctx.holder, means)
        acc = v_means is not None
        if not acc:
            v_means = torch.empty_like(means)
        v_quats = torch.empty_like(quats)
        v_scales = torch.empty_like(scales)
        v_means2d, ld_m = _rows(v_means2d, 2)
        v_depths, ld_d = _rows(None if v_depths is None else v_depths.unsqueeze(-1), 1)
        v_rt, ld_rt = _rows(None if v_ray_transforms is None else v_ray_transforms.flatten(-2), 9)
        v_normals, ld_n = _rows(v_normals, 3)
        vis = None if ctx.holder is None else ctx.holder.get("vis_ids")
        _mark("project2d_bwd", 0)
        check(L.hgs_project2d_bwd(ptr(means), ptr(quats), ptr(scales), ptr(viewmats), ptr(Ks), C, N, width, height,
                                  near_plane, far_plane, ptr(radii), ptr(v_means2d), ld_m, ptr(v_depths), ld_d,
                                  ptr(v_rt), ld_rt, ptr(v_normals), ld_n, ptr(vis), 0 if vis is None else vis.numel(),
                                  ptr(v_means), ptr(v_quats), ptr(v_scales), int(acc), _stream()), "hgs_project2d_bwd")
        _mark("project2d_bwd", 1)
        return (v_means, v_quats, v_scales) + (None,) * 9


def _project2d(means, quats, scales, viewmats, Ks, width, height, near_plane, far_plane, radius_clip, tile_size,
               holder=None):
    _check_proj_inputs(means, quats, scales, viewmats, Ks)
    means, quats, scales = _f32c(means, "means"), _f32c(quats, "quats"), _f32c(scales, "scales")
    viewmats, Ks = _f32c(viewmats.detach(), "viewmats"), _f32c(Ks.detach(), "Ks")
    return _Project2D.apply(means, quats, scales, viewmats, Ks, int(width), int(height), float(near_plane),
                            float(far_plane), float(radius_clip), int(tile_size), holder)


def fully_fused_projection_2dgs(
    means: Tensor, quats: Tensor, scales: Tensor, viewmats: Tensor, densifications: Optional[Tensor], Ks: Tensor,
    width: int, height: int, eps2d: float = 0.3, near_plane: float = 0.01, far_plane: float = 1e10,
    radius_clip: float = 0.0, packed: bool = False, sparse_grad: bool = False,
) -> Tuple[Tensor, Tensor, Tensor, Tensor, Tensor]:
    """Same call as the gsplat 2DGS fork's fully_fused_projection_2dgs at render.py:171-186
    (``densifications`` [C,N,2] is the fork's 5th positional: a gradient slot, never read; eps2d is unused
    by surfel projection).

    -> (radii[C,N] int32, means2d[C,N,2], depths[C,N], ray_transforms[C,N,3,3], normals[C,N,3]).
    """
    if packed or sparse_grad:
        raise NotImplementedError("packed / sparse_grad are not supported (the reference passes False)")
    radii, means2d, depths, ray_transforms, normals, _ = _project2d(
        means, quats, scales, viewmats, Ks, width, height, near_plane, far_plane, radius_clip, 0)
    return radii, means2d, depths, ray_transforms, normals


# =====================================================================================
# a12: rasterize_to_pixels_2dgs
# =====================================================================================
class _Blend2D(torch.autograd.Function):
    """rasterize_to_pixels_2dgs.  1 / 3 / 4 channels: packed-record fast kernels (TMA-staged, per-warp cull box,
    fused expected-depth normalisation); otherwise the plain kernels."""

    @staticmethod
    def forward(ctx, means2d, ray_transforms, colors, depths, normals, opacities, densify, backgrounds, width,
                height, tile_size, isect_offsets, flatten_ids, distloss, box, radii, normalize_depth, vis_ids,
                defer=None):
        L = _lib.lib()
        ctx.defer = defer
        C, N = opacities.shape
        CH = colors.shape[-1]
        D = CH + (1 if depths is not None else 0)
        dev = means2d.device
        render_colors = torch.empty((C, height, width, D), dtype=torch.float32, device=dev)
        render_alphas = torch.empty((C, height, width, 1), dtype=torch.float32, device=dev)
        render_normals = torch.empty((C, height, width, 3), dtype=torch.float32, device=dev)
        render_distort = torch.empty((C, height, width, 1), dtype=torch.float32, device=dev) if distloss else None
        render_median = torch.empty((C, height, width, 1), dtype=torch.float32, device=dev)
        last_ids = torch.empty((C, height, width), dtype=torch.int32, device=dev)
        median_ids = torch.empty((C, height, width), dtype=torch.int32, device=dev)
        fast = D in (1, 3, 4)
        ctx.fast = fast
        ctx.cfg = (width, height, tile_size, distloss, bool(normalize_depth), CH, D)
        ctx.box = box
        st = _stream()
        if fast:
            records = torch.empty(L.hgs_blend2d_pack_bytes(C * N), dtype=torch.uint8, device=dev)
            _mark("blend2d_pack", 0)
            check(L.hgs_blend2d_pack(ptr(means2d), ptr(ray_transforms), ptr(colors), ptr(depths), ptr(normals),
                                     ptr(opacities), ptr(radii), ptr(vis_ids),
                                     0 if vis_ids is None else vis_ids.numel(), C * N, CH, ptr(records), st),
                  "hgs_blend2d_pack")
            _mark("blend2d_pack", 1)
            _mark("blend2d_fwd", 0)
            check(L.hgs_blend2d_fwd_packed(ptr(records), ptr(backgrounds), C, D, int(normalize_depth), width, height,
                                           tile_size, ptr(isect_offsets), ptr(flatten_ids), flatten_ids.numel(),
                                           ptr(render_colors), ptr(render_alphas), ptr(render_normals),
                                           ptr(render_distort), ptr(render_median), ptr(last_ids), ptr(median_ids),
                                           st), "hgs_blend2d_fwd_packed")
            _mark("blend2d_fwd", 1)
            ctx.save_for_backward(records, backgrounds, isect_offsets, flatten_ids, render_colors, render_alphas,
                                  last_ids, median_ids)
            ctx.shapes = (means2d.shape, depths is not None)
        else:
            if normalize_depth:
                raise NotImplementedError("fused depth normalisation needs 1, 3 or 4 channels")
            check(L.hgs_blend2d_fwd(ptr(means2d), ptr(ray_transforms), ptr(colors), ptr(depths), ptr(normals),
                                    ptr(opacities), ptr(backgrounds), C, N, CH, width, height, tile_size,
                                    ptr(isect_offsets), ptr(flatten_ids), flatten_ids.numel(), ptr(render_colors),
                                    ptr(render_alphas), ptr(render_normals), ptr(render_distort), ptr(render_median),
                                    ptr(last_ids), ptr(median_ids), st), "hgs_blend2d_fwd")
            ctx.save_for_backward(means2d, ray_transforms, colors, depths, normals, opacities, backgrounds,
                                  isect_offsets, flatten_ids, render_colors, render_alphas, last_ids, median_ids)
        if not distloss:
            render_distort = torch.zeros((C, height, width, 1), dtype=torch.float32, device=dev)
            ctx.mark_non_differentiable(render_distort)
        return render_colors, render_alphas, render_normals, render_distort, render_median

    @staticmethod
    def backward(ctx, v_render_colors, v_render_alphas, v_render_normals, v_render_distort, v_render_median):
        width, height, tile_size, distloss, normalize_depth, CH, D = ctx.cfg
        L = _lib.lib()
        cg = lambda t: None if t is None else t.contiguous()  # noqa: E731
        tail = (None,) * 11
        if ctx.fast:
            (records, backgrounds, isect_offsets, flatten_ids, render_colors, render_alphas, last_ids,
             median_ids) = ctx.saved_tensors
            (C, N, _), has_depth = ctx.shapes
        else:
            (means2d, ray_transforms, colors, depths, normals, opacities, backgrounds, isect_offsets, flatten_ids,
             render_colors, render_alphas, last_ids, median_ids) = ctx.saved_tensors
            C, N = opacities.shape
        v_render_colors = cg(v_render_colors) if v_render_colors is not None else torch.zeros_like(render_colors)
        v_render_alphas = cg(v_render_alphas) if v_render_alphas is not None else torch.zeros_like(render_alphas)
        v_render_normals, v_render_median = cg(v_render_normals), cg(v_render_median)
        v_render_distort = cg(v_render_distort) if distloss else None
        if ctx.fast:
            vpack = torch.zeros((C, N, 24), dtype=torch.float32, device=records.device)
            _mark("blend2d_bwd", 0)
            check(L.hgs_blend2d_bwd_packed(ptr(records), ptr(backgrounds), C, D, int(normalize_depth), width, height,
                                           tile_size, ptr(isect_offsets), ptr(flatten_ids), flatten_ids.numel(),
                                           ptr(render_colors), ptr(render_alphas), ptr(last_ids), ptr(median_ids),
                                           ptr(v_render_colors), ptr(v_render_alphas), ptr(v_render_normals),
                                           ptr(v_render_distort), ptr(v_render_median), ptr(vpack), _stream()),
                  "hgs_blend2d_bwd_packed")
            _mark("blend2d_bwd", 1)
            if ctx.defer is not None:
                ctx.defer.update(vpack=vpack, surfel=True, has_depth=bool(has_depth))
            v_means2d = vpack[..., 0:2]
            v_rt = vpack[..., 2:11].unflatten(-1, (3, 3))
            v_normals = vpack[..., 11:14]
            v_opacities = vpack[..., 14]
            v_colors = vpack[..., 16:16 + CH]
            v_depths = vpack[..., 16 + CH] if has_depth else None
            v_densify = vpack[..., 20:22]
        else:
            v_means2d = torch.zeros_like(means2d)
            v_rt = torch.zeros_like(ray_transforms)
            v_colors = torch.zeros_like(colors)
            v_depths = torch.zeros_like(depths) if depths is not None else None
            v_normals = torch.zeros_like(normals)
            v_opacities = torch.zeros_like(opacities)
            v_densify = torch.zeros_like(means2d) if (ctx.needs_input_grad[6] or ctx.box is not None) else None
            check(L.hgs_blend2d_bwd(ptr(means2d), ptr(ray_transforms), ptr(colors), ptr(depths), ptr(normals),
                                    ptr(opacities), ptr(backgrounds), C, N, CH, width, height, tile_size,
                                    ptr(isect_offsets), ptr(flatten_ids), flatten_ids.numel(), ptr(render_colors),
                                    ptr(render_alphas), ptr(last_ids), ptr(median_ids), ptr(v_render_colors),
                                    ptr(v_render_alphas), ptr(v_render_normals), ptr(v_render_distort),
                                    ptr(v_render_median), ptr(v_means2d), ptr(v_rt), ptr(v_colors), ptr(v_depths),
                                    ptr(v_normals), ptr(v_opacities), ptr(v_densify), _stream()), "hgs_blend2d_bwd")
        if ctx.box is not None:
            ctx.box["densify"] = v_densify  # picked up by rendering._DensifyInject / _DensifyProbe
        v_bg = None
        if backgrounds is not None and ctx.needs_input_grad[7]:
            vrc = v_render_colors
            if normalize_depth:
                vrc = torch.cat([vrc[..., :-1], vrc[..., -1:] / render_alphas.clamp(min=1e-10)], -1)
            v_bg = (vrc * (1.0 - render_alphas)).sum(dim=(1, 2))
        return (v_means2d, v_rt, v_colors, v_depths, v_normals, v_opacities,
                v_densify if ctx.needs_input_grad[6] else None, v_bg) + tail


def _blend2d(means2d, ray_transforms, colors, depths, normals, opacities, densify, backgrounds, width, height,
             tile_size, isect_offsets, flatten_ids, distloss=False, box=None, radii=None, normalize_depth=False,
             vis_ids=None, defer=None):
    if tile_size not in _TILE_SIZES:
        raise NotImplementedError(f"tile_size {tile_size} is not supported (supported: {_TILE_SIZES})")
    D = colors.shape[-1] + (1 if depths is not None else 0)
    if D > 8:
        raise NotImplementedError(f"{D} render channels requested; at most 8 are supported")
    return _Blend2D.apply(_f32c(means2d, "means2d"), _f32c(ray_transforms, "ray_transforms"),
                          _f32c(colors, "colors"), _f32c(depths, "depths"), _f32c(normals, "normals"),
                          _f32c(opacities, "opacities"), densify, _f32c(backgrounds, "backgrounds"), int(width),
                          int(height), int(tile_size), isect_offsets.contiguous(), flatten_ids.contiguous(),
                          bool(distloss), box, radii, bool(normalize_depth), vis_ids, defer)


def rasterize_to_pixels_2dgs(means2d: Tensor, ray_transforms: Tensor, colors: Tensor, opacities: Tensor,
                             normals: Tensor, densify: Optional[Tensor], image_width: int, image_height: int,
                             tile_size: int, isect_offsets: Tensor, flatten_ids: Tensor,
                             backgrounds: Optional[Tensor] = None, masks: Optional[Tensor] = None,
                             packed: bool = False, absgrad: bool = False, distloss: bool = False):
    """gsplat.rasterize_to_pixels_2dgs -> (render_colors, render_alphas, render_normals, render_distort,
    render_median).  ``densify`` [C,N,2] (zeros, requires_grad) receives the screen-space positional
    gradient used for densification in its .grad."""
    if packed or masks is not None or absgrad:
        raise NotImplementedError("packed layout / tile masks / absgrad are not supported")
    return _blend2d(means2d, ray_transforms, colors, None, normals, opacities, densify, backgrounds, image_width,
                    image_height, tile_size, isect_offsets, flatten_ids, distloss)


# =====================================================================================
# a13: 2DGS post-ops (normals to world frame, normals from depth)
# =====================================================================================
class _NormalsPost(torch.autograd.Function):
    """(render_normals camera frame [C,H,W,3], depth_src [C,H,W,Dd] whose LAST channel is the z-depth map or None)
    -> (render_normals world frame, normals_from_depth [C,H,W,3] or None)"""

    @staticmethod
    def forward(ctx, normals_cam, depth_src, viewmats, Ks, want_nfd):
        L = _lib.lib()
        C, H, Wd, _ = normals_cam.shape
        ctx.set_materialize_grads(False)
        normals_world = torch.empty_like(normals_cam)
        nfd = torch.empty_like(normals_cam) if want_nfd else None
        ld = depth_src.shape[-1] if depth_src is not None else 1
        dptr = None if depth_src is None else _lib.C.c_void_p(depth_src.data_ptr() + 4 * (ld - 1))
        _mark("normals_post_fwd", 0)
        check(L.hgs_normals_post_fwd(ptr(normals_cam), dptr, ld, ptr(viewmats), ptr(Ks), C, H, Wd, ptr(normals_world),
                                     ptr(nfd), _stream()), "hgs_normals_post_fwd")
        _mark("normals_post_fwd", 1)
        ctx.save_for_backward(depth_src, viewmats, Ks)
        ctx.dims = (C, H, Wd, ld)
        return normals_world, nfd

    @staticmethod
    def backward(ctx, v_normals_world, v_nfd):
        depth_src, viewmats, Ks = ctx.saved_tensors
        C, H, Wd, ld = ctx.dims
        L = _lib.lib()
        dev = viewmats.device
        v_normals_world = None if v_normals_world is None else v_normals_world.contiguous()
        v_nfd = None if v_nfd is None else v_nfd.contiguous()
        v_normals_cam = torch.empty((C, H, Wd, 3), dtype=torch.float32, device=dev) if ctx.needs_input_grad[0] else None
        v_src, vptr = None, None
        if depth_src is not None and ctx.needs_input_grad[1] and v_nfd is not None:
            v_src = torch.zeros_like(depth_src) if ld > 1 else torch.empty_like(depth_src)
            vptr = _lib.C.c_void_p(v_src.data_ptr() + 4 * (ld - 1))
        dptr = None if depth_src is None else _lib.C.c_void_p(depth_src.data_ptr() + 4 * (ld - 1))
        _mark("normals_post_bwd", 0)
        check(L.hgs_normals_post_bwd(dptr, ld, ptr(viewmats), ptr(Ks), C, H, Wd, ptr(v_normals_world), ptr(v_nfd),
                                     ptr(v_normals_cam), vptr, ld, _stream()), "hgs_normals_post_bwd")
        _mark("normals_post_bwd", 1)
        return v_normals_cam, v_src, None, None, None


def normals_post(render_normals: Tensor, depth_src: Optional[Tensor], viewmats: Tensor, Ks: Tensor, want_nfd: bool):
    """gsplat.rasterization_2dgs' post-ops: render_normals to the world frame and (want_nfd) depth_to_normal of the
    z-depth map in the last channel of depth_src [C,H,W,Dd]."""
    return _NormalsPost.apply(_f32c(render_normals, "render_normals"),
                              None if depth_src is None else _f32c(depth_src, "depth map"),
                              _f32c(viewmats.detach(), "viewmats"), _f32c(Ks.detach(), "Ks"), bool(want_nfd))


@torch.no_grad()
def blend3d_pair_stats(means2d, conics, opacities, radii, width, height, tile_size, isect_offsets, flatten_ids):
    """(P_eval, P_blend) of a view -- measurement aid for bench.py's roofline figures (not on the product path)."""
    L = _lib.lib()
    C, N = opacities.shape
    dev = means2d.device
    records = torch.empty(L.hgs_blend3d_pack_bytes(C * N), dtype=torch.uint8, device=dev)
    dummy = torch.zeros((C, N, 1), dtype=torch.float32, device=dev)
    st = _stream()
    check(L.hgs_blend3d_pack(ptr(means2d.contiguous()), ptr(conics.contiguous()), ptr(dummy), None,
                             ptr(opacities.contiguous()), ptr(radii), None, 0, None, C * N, 1, ptr(records), st),
          "hgs_blend3d_pack")
    counters = torch.zeros(8, dtype=torch.int64, device=dev)
    check(L.hgs_blend3d_stats(ptr(records), C, int(width), int(height), int(tile_size), ptr(isect_offsets),
                              ptr(flatten_ids), flatten_ids.numel(), ptr(counters), st), "hgs_blend3d_stats")
    vals = counters.tolist()
    blend3d_pair_stats.last_cull = {"warp_pairs_8x4": int(vals[2]), "half_iters_4x4": int(vals[3]),
                                    "half_iters_8x2": int(vals[4])}
    return int(vals[0]), int(vals[1])


@torch.no_grad()
def densification_stats_update(means2d_grad: Tensor, radii: Tensor, width: int, height: int, grad_accum: Tensor,
                               denom: Tensor, max_radii: Optional[Tensor] = None, mode: str = "mean",
                               visible_ids: Optional[Tensor] = None) -> None:
    """In-place update of the densification accumulators from one rendered batch of views
    (scene/basic_model.py:96-144): grad_accum[N] (+= or max= the scaled view-space gradient norm),
    denom[N] (+= views that saw the Gaussian), max_radii[N] (optional).  visible_ids = meta["visible_ids"] of the
    rasterization call (optional work list; same result, no idle threads)."""
    assert mode in ("mean", "max")
    L = _lib.lib()
    C, N = radii.shape
    g, ld = _rows(means2d_grad, 2)
    assert grad_accum.is_contiguous() and denom.is_contiguous() and grad_accum.numel() == N and denom.numel() == N
    check(L.hgs_densify_stats(ptr(g), ld, ptr(radii.contiguous()), ptr(visible_ids),
                              0 if visible_ids is None else visible_ids.numel(), C, N, int(width), int(height),
                              1 if mode == "max" else 0, ptr(grad_accum), ptr(denom), ptr(max_radii), _stream()),
          "hgs_densify_stats")
