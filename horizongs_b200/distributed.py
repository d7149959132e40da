"""View-sharded data parallelism for the rasterization path.

The reference trains one view per iteration in one process (no collective anywhere, section 2.1); the new
multi-GPU mode shards the per-step camera batch: rank r renders view r with replicated Gaussian parameters,
then ONE exchange step sums the parameter gradients and the densification statistics over ranks.
Densification statistics follow scene/basic_model.py:96-144: per-view gradient norms are computed locally
BEFORE any reduction (the norm is not linear), then SUM-reduced (mean mode, :136,:144) or MAX-reduced
(max mode, :138-139).

Plumbing only (torch.distributed: NCCL on GPUs, gloo in the CPU tests); no kernels here.
"""
from __future__ import annotations

from typing import Iterable, List, Optional, Sequence

import torch
import torch.distributed as dist


def shard_views(n_views: int, rank: int, world_size: int, step: int = 0) -> List[int]:
    """views rendered by `rank` at `step`: a rotating, disjoint, exhaustive split of range(n_views)."""
    return [v for v in range(n_views) if (v - step) % world_size == rank % world_size]


def allreduce_gradients(params: Iterable[torch.Tensor], group=None, average: bool = False,
                        async_op: bool = False):
    """SUM (or mean) the .grad of every parameter over the group, one collective per tensor (the tensors are
    large -- 12..108 bytes per Gaussian -- so bucketing buys nothing over NVLink)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return []
    handles = []
    world = dist.get_world_size(group)
    for p in params:
        if p.grad is None:
            continue
        if not p.grad.is_contiguous():
            p.grad = p.grad.contiguous()
        if average:
            p.grad.div_(world)
        handles.append(dist.all_reduce(p.grad, op=dist.ReduceOp.SUM, group=group, async_op=True))
    if not async_op:
        for h in handles:
            h.wait()
        return []
    return handles


class GradientExchange:
    """Starts the SUM all-reduce of each parameter's gradient the moment autograd has finished accumulating it
    (post-accumulate-grad hook), so the exchange of the early gradients (opacities after the blend backward, SH
    coefficients -- 71 % of the bytes -- after the SH backward) overlaps the rest of the backward pass.
    Call ``wait()`` after ``loss.backward()``; gradients must start as ``None`` each step."""

    def __init__(self, params: Sequence[torch.Tensor], group=None):
        self.group = group
        self.handles = []
        self.enabled = dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1
        self._hooks = []
        if self.enabled:
            for p in params:
                self._hooks.append(p.register_post_accumulate_grad_hook(self._on_grad))

    def _on_grad(self, p: torch.Tensor):
        if not p.grad.is_contiguous():
            p.grad = p.grad.contiguous()
        self.handles.append(dist.all_reduce(p.grad, op=dist.ReduceOp.SUM, group=self.group, async_op=True))

    def wait(self):
        for h in self.handles:
            h.wait()
        self.handles.clear()

    def close(self):
        for h in self._hooks:
            h.remove()
        self._hooks.clear()


def densification_statistics(means2d_grad: torch.Tensor, radii: torch.Tensor, width: int, height: int):
    """per-view statistics of one rank: (grad_norm[N], visible[N]) with the reference's scaling
    (scene/basic_model.py:131-134: pixel-unit gradient times (W/2, H/2))."""
    g = means2d_grad.reshape(-1, means2d_grad.shape[-2], 2).sum(0) if means2d_grad.dim() == 3 else means2d_grad
    norm = torch.sqrt((g[:, 0] * (0.5 * width)) ** 2 + (g[:, 1] * (0.5 * height)) ** 2)
    vis = (radii.reshape(-1, radii.shape[-1]) > 0).any(0)
    return norm, vis.to(norm.dtype)


def allreduce_densification(grad_norm: torch.Tensor, visible: torch.Tensor, max_radii: Optional[torch.Tensor] = None,
                            mode: str = "mean", group=None):
    """Reduce the per-view statistics over ranks: SUM of norms and visibility counts in 'mean' mode,
    MAX of norms (and of radii) in 'max' mode.  Returns the reduced tensors (in place)."""
    assert mode in ("mean", "max")
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return grad_norm, visible, max_radii
    if mode == "mean":
        buf = torch.stack([grad_norm, visible])
        dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=group)
        grad_norm.copy_(buf[0]); visible.copy_(buf[1])
    else:
        dist.all_reduce(grad_norm, op=dist.ReduceOp.MAX, group=group)
        dist.all_reduce(visible, op=dist.ReduceOp.SUM, group=group)
    if max_radii is not None:
        dist.all_reduce(max_radii, op=dist.ReduceOp.MAX, group=group)
    return grad_norm, visible, max_radii
