"""Fused loss terms that sit right after the rasterization path.

``photometric_l1_loss`` is the reference's ``l1_loss(image, gt_image)`` (utils/loss_utils.py:17-18, train.py:158)
-- optionally with the means of the depth channel and of the alpha map as regularisers -- as ONE forward and ONE
backward kernel (csrc/loss.cu) instead of the dozen elementwise / reduction launches autograd makes of it.  The
gradient images come out contiguous, in the layout the blend backward reads.  CUDA only (no fallback)."""
from __future__ import annotations

from typing import Optional

import torch
from torch import Tensor

from . import _lib
from ._lib import check, ptr


class _PhotometricL1(torch.autograd.Function):
    @staticmethod
    def forward(ctx, render_colors, render_alphas, gt, w_depth, w_alpha):
        L = _lib.lib()
        D = render_colors.shape[-1]
        P = render_colors.numel() // D
        dev = render_colors.device
        partials = torch.empty(L.hgs_l1_loss_partials(), dtype=torch.float32, device=dev)
        loss = torch.empty((), dtype=torch.float32, device=dev)
        check(L.hgs_l1_loss_fwd(ptr(render_colors), ptr(render_alphas), ptr(gt), P, D, w_depth, w_alpha, ptr(partials),
                                ptr(loss), torch.cuda.current_stream().cuda_stream), "hgs_l1_loss_fwd")
        ctx.save_for_backward(render_colors, gt)
        ctx.cfg = (P, D, w_depth, w_alpha, render_alphas is not None, None if render_alphas is None else render_alphas.shape)
        return loss

    @staticmethod
    def backward(ctx, v_loss):
        render_colors, gt = ctx.saved_tensors
        P, D, w_depth, w_alpha, has_alpha, a_shape = ctx.cfg
        L = _lib.lib()
        v_rc = torch.empty_like(render_colors)
        v_ra = torch.empty(a_shape, dtype=torch.float32, device=render_colors.device) if has_alpha else None
        check(L.hgs_l1_loss_bwd(ptr(render_colors), ptr(gt), ptr(v_loss.contiguous()), P, D, w_depth, w_alpha, ptr(v_rc),
                                ptr(v_ra), torch.cuda.current_stream().cuda_stream), "hgs_l1_loss_bwd")
        return v_rc, v_ra, None, None, None


def photometric_l1_loss(render_colors: Tensor, gt: Tensor, render_alphas: Optional[Tensor] = None,
                        w_depth: float = 0.0, w_alpha: float = 0.0) -> Tensor:
    """mean |render_colors[..., :3] - gt| + w_depth * mean(render_colors[..., 3]) + w_alpha * mean(render_alphas).

    render_colors [C,H,W,3|4] (channels-last, as the rasterizer returns it), gt [C,H,W,3], render_alphas [C,H,W,1]."""
    if not render_colors.is_cuda:
        raise ValueError("photometric_l1_loss runs on CUDA tensors only (no CPU fallback)")
    D = render_colors.shape[-1]
    assert D in (3, 4) and gt.shape[-1] == 3 and gt.shape[:-1] == render_colors.shape[:-1], (render_colors.shape, gt.shape)
    assert render_colors.dtype == torch.float32 and gt.dtype == torch.float32
    if render_alphas is not None:
        assert render_alphas.numel() == render_colors.numel() // D
        render_alphas = render_alphas.contiguous()
    return _PhotometricL1.apply(render_colors.contiguous(), render_alphas, gt.contiguous(), float(w_depth), float(w_alpha))


class _PhotometricL1SSIM(torch.autograd.Function):
    """(1 - lam) * L1 + lam * (1 - SSIM) + w_depth * mean(depth) + w_alpha * mean(alpha); train.py:158-160"""

    @staticmethod
    def forward(ctx, render_colors, render_alphas, gt, lam, w_depth, w_alpha):
        L = _lib.lib()
        C, H, W, D = render_colors.shape
        P = C * H * W
        dev = render_colors.device
        st = torch.cuda.current_stream().cuda_stream
        partials = torch.empty(max(L.hgs_l1_loss_partials(), L.hgs_ssim_partials(C, H, W)), dtype=torch.float32, device=dev)
        out = torch.empty(2, dtype=torch.float32, device=dev)          # [l1 (+ regularisers / (1 - lam)), mean ssim]
        k = 1.0 / (1.0 - lam) if lam < 1.0 else 0.0
        check(L.hgs_l1_loss_fwd(ptr(render_colors), ptr(render_alphas), ptr(gt), P, D, w_depth * k, w_alpha * k,
                                ptr(partials), ptr(out[0:1]), st), "hgs_l1_loss_fwd")
        dmaps = torch.empty((3, P * 3), dtype=torch.float32, device=dev)
        check(L.hgs_ssim_fwd(ptr(render_colors), ptr(gt), C, H, W, D, ptr(dmaps), ptr(partials), ptr(out[1:2]), st),
              "hgs_ssim_fwd")
        ctx.save_for_backward(render_colors, gt, dmaps)
        ctx.cfg = (C, H, W, D, lam, w_depth * k, w_alpha * k, None if render_alphas is None else render_alphas.shape)
        return (1.0 - lam) * out[0] + lam * (1.0 - out[1])

    @staticmethod
    def backward(ctx, v_loss):
        render_colors, gt, dmaps = ctx.saved_tensors
        C, H, W, D, lam, wd, wa, a_shape = ctx.cfg
        L = _lib.lib()
        st = torch.cuda.current_stream().cuda_stream
        v = torch.stack([v_loss * (1.0 - lam), v_loss * (-lam)]).to(torch.float32).contiguous()
        v_rc = torch.empty_like(render_colors)
        v_ra = torch.empty(a_shape, dtype=torch.float32, device=render_colors.device) if a_shape is not None else None
        check(L.hgs_l1_loss_bwd(ptr(render_colors), ptr(gt), ptr(v[0:1]), C * H * W, D, wd, wa, ptr(v_rc), ptr(v_ra), st),
              "hgs_l1_loss_bwd")
        check(L.hgs_ssim_bwd(ptr(render_colors), ptr(gt), ptr(dmaps), ptr(v[1:2]), C, H, W, D, ptr(v_rc), st), "hgs_ssim_bwd")
        return v_rc, v_ra, None, None, None, None


def photometric_loss(render_colors: Tensor, gt: Tensor, lambda_dssim: float = 0.2, render_alphas: Optional[Tensor] = None,
                     w_depth: float = 0.0, w_alpha: float = 0.0) -> Tensor:
    """The reference's training loss (train.py:158-160):
        (1 - lambda_dssim) * l1_loss(image, gt) + lambda_dssim * (1 - ssim(image, gt))
    (utils/loss_utils.py:17-60; 11x11 Gaussian window, zero padding) on the channels-last images the rasterizer
    returns -- render_colors [C,H,W,3|4], gt [C,H,W,3] -- plus optional means of the depth channel / alpha map.
    Four kernels forward, two backward (csrc/loss.cu) instead of ~60 launches of convolutions and elementwise ops."""
    if not render_colors.is_cuda:
        raise ValueError("photometric_loss runs on CUDA tensors only (no CPU fallback)")
    assert render_colors.dim() == 4 and render_colors.shape[-1] in (3, 4), render_colors.shape
    assert gt.shape == render_colors.shape[:-1] + (3,), (render_colors.shape, gt.shape)
    assert render_colors.dtype == torch.float32 and gt.dtype == torch.float32
    assert 0.0 <= lambda_dssim <= 1.0
    if render_alphas is not None:
        assert render_alphas.numel() == render_colors.numel() // render_colors.shape[-1]
        render_alphas = render_alphas.contiguous()
    return _PhotometricL1SSIM.apply(render_colors.contiguous(), render_alphas, gt.contiguous(), float(lambda_dssim),
                                    float(w_depth), float(w_alpha))
