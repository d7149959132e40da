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
