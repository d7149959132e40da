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


def bucketed_view(step: int, rank: int, n_views: int, n_kinds: int = 2) -> int:
    """Cost-bucketed view schedule for view-sharded data parallelism: the view `rank` renders at `step`.

    Views are stored interleaved by kind (index % n_kinds: Horizon-GS trains on aerial AND street cameras of one scene,
    train.py:107-125, and an aerial view costs ~1.6x a street view).  A step lasts as long as its slowest rank, so a
    step should hold views of ONE kind: step s renders kind s % n_kinds and rank r takes the (r + s // n_kinds)-th view
    of that kind.  Every rank visits every view, ranks never share a view within a step while
    world_size <= n_views / n_kinds, and every rank renders one view of each kind per n_kinds steps whatever the world
    size.  With one rank the schedule is 0, 1, 2, ..."""
    per_kind = n_views // n_kinds
    return n_kinds * ((rank + step // n_kinds) % per_kind) + (step % n_kinds)


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


class _PeerMailbox:
    """A zero-filled device allocation of `nbytes` on every rank, mapped into every peer through CUDA IPC handles
    traded once over torch.distributed (hgs_peer_* of the C ABI).  ptrs[r] is rank r's mailbox as seen from here."""

    def __init__(self, nbytes: int, group=None, device: Optional[torch.device] = None):
        import ctypes as C
        from . import _lib
        assert dist.is_available() and dist.is_initialized(), "peer-memory exchange needs an initialised process group"
        self._C, self._lib, self._check = C, _lib.lib(), _lib.check
        L = self._lib
        self.group = group
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else device
        self.nbytes = int(nbytes)
        self._local = C.c_void_p()
        self._check(L.hgs_peer_alloc(self.nbytes, C.byref(self._local)), "hgs_peer_alloc")
        handle = (C.c_ubyte * 64)()
        self._check(L.hgs_peer_export(self._local, handle), "hgs_peer_export")
        mine = torch.tensor(list(handle), dtype=torch.uint8, device=self.device)
        gathered = [torch.empty_like(mine) for _ in range(self.world)]
        dist.all_gather(gathered, mine, group=group)
        self.ptrs = []
        for r, h in enumerate(gathered):
            if r == self.rank:
                self.ptrs.append(self._local.value)
                continue
            buf = (C.c_ubyte * 64)(*h.cpu().tolist())
            p = C.c_void_p()
            self._check(L.hgs_peer_import(buf, C.byref(p)), f"hgs_peer_import(rank {r})")
            self.ptrs.append(p.value)
        self.ptrs_c = (C.c_void_p * self.world)(*self.ptrs)
        self.status = torch.zeros(1, dtype=torch.int32, device=self.device)
        self._status_host = torch.zeros(1, dtype=torch.int32).pin_memory()
        self._status_event = None
        dist.barrier(group=group)          # every mailbox is mapped everywhere before the first push

    @property
    def local(self):
        return self._local

    _STATUS_TEXT = {1: "timed out waiting for a peer's records",
                    2: "a rank's contribution exceeded the mailbox capacity (cap_rows); every rank skipped that step's "
                       "gradients -- construct the exchange with a larger cap_rows"}

    def _raise(self, code: int) -> None:
        from . import _lib
        raise _lib.HgsError("peer gradient exchange failed: " + self._STATUS_TEXT.get(code, f"status {code}"))

    def record_status(self) -> None:
        """enqueue an asynchronous copy of the sticky device status (after a reduce); poll_status() reads it later"""
        self._status_host.copy_(self.status, non_blocking=True)
        self._status_event = torch.cuda.Event()
        self._status_event.record()

    def poll_status(self) -> None:
        """raise if an EARLIER exchange failed on this or any other rank (timeout / overflow are outcomes every rank
        observes for the same step); never blocks: only a status copy that has already landed is looked at"""
        ev = self._status_event
        if ev is not None and ev.query():
            self._status_event = None
            code = int(self._status_host[0])
            if code != 0:
                self._raise(code)

    def check_status(self) -> None:
        """raise if any exchange so far failed (one blocking host read)"""
        code = int(self.status.item())
        if code != 0:
            self._raise(code)

    def close(self) -> None:
        if self._local is None:
            return
        torch.cuda.synchronize(self.device)
        dist.barrier(group=self.group)     # nobody unmaps while a peer may still push
        for r, p in enumerate(self.ptrs):
            if r != self.rank:
                self._lib.hgs_peer_close(self._C.c_void_p(p))
        self._lib.hgs_peer_free(self._local)
        self._local = None


class _SymmMailbox(_PeerMailbox):
    """The same mailbox, allocated through torch's symmetric memory (CUDA VMM + handle exchange done by torch): gives
    the peer pointers AND, on an NVSwitch fabric, a MULTICAST mapping of all ranks' mailboxes -- a store to
    multicast_ptr + offset lands at `offset` of every rank's mailbox after crossing this GPU's NVLink once."""

    def __init__(self, nbytes: int, group=None, device: Optional[torch.device] = None):
        import ctypes as C
        import torch.distributed._symmetric_memory as symm_mem
        from . import _lib
        assert dist.is_available() and dist.is_initialized(), "peer-memory exchange needs an initialised process group"
        self._C, self._lib, self._check = C, _lib.lib(), _lib.check
        self.group = dist.group.WORLD if group is None else group
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else device
        self.nbytes = int(nbytes)
        self._buf = symm_mem.empty((self.nbytes + 3) // 4, dtype=torch.float32, device=self.device)
        self._buf.zero_()
        self._hdl = symm_mem.rendezvous(self._buf, self.group)
        self.ptrs = [int(p) for p in self._hdl.buffer_ptrs]
        self.ptrs_c = (C.c_void_p * self.world)(*self.ptrs)
        self.multicast_ptr = int(self._hdl.multicast_ptr or 0)
        self._local = C.c_void_p(self.ptrs[self.rank])
        self.status = torch.zeros(1, dtype=torch.int32, device=self.device)
        self._status_host = torch.zeros(1, dtype=torch.int32).pin_memory()
        self._status_event = None
        torch.cuda.synchronize(self.device)
        dist.barrier(group=group)          # every mailbox is zeroed and mapped everywhere before the first push

    def close(self) -> None:
        if self._buf is None:
            return
        torch.cuda.synchronize(self.device)
        dist.barrier(group=self.group)     # nobody lets go while a peer may still push
        self._hdl = None
        self._buf = None
        self._local = None


def _make_mailbox(nbytes: int, group, device, want_multicast: bool):
    """symmetric-memory mailbox with multicast when the fabric offers it (and HGS_EXCHANGE_NO_MULTICAST is unset), else
    the CUDA-IPC mailbox.  The choice is collective: every rank takes the same path."""
    import os
    box = None
    if want_multicast and not os.environ.get("HGS_EXCHANGE_NO_MULTICAST"):
        try:
            box = _SymmMailbox(nbytes, group, device)
        except Exception:  # noqa: BLE001 -- symmetric memory unavailable on this build / fabric
            box = None
        ok = torch.tensor([1 if (box is not None and box.multicast_ptr) else 0], device=device)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=group)
        if int(ok.item()) == 0:
            if box is not None:
                box.close()
            box = None
    if box is None:
        box = _PeerMailbox(nbytes, group, device)
        box.multicast_ptr = 0
    return box


class PeerGradientExchange:
    """Sparse SUM all-reduce of view-sharded gradients over NVLink peer memory (csrc/exchange.cu).

    With one view per GPU only the rows of the Gaussians that view sees are non-zero, so instead of the dense
    NCCL all-reduce every rank stores its visible rows (one record = the rows of all tensors) straight into a
    mailbox in every peer's memory and then adds the W sources' records into its dense tensors in rank order --
    all replicas end with bit-identical sums.  torch.distributed is used once, at construction, to trade the
    CUDA IPC handles of the mailboxes; the per-step exchange is two C-ABI calls and no collective.

    widths: row widths of the tensors exchanged together, e.g. (3, 4, 3, 1, 27, 1, 1) for the gradients of
    means / quats / scales / opacities / SH coefficients and the two densification statistics.
    n_rows_total: N, the number of rows of every tensor; cap_rows: the largest number of rows one rank may
    contribute per step (its visible Gaussians).
    """

    def __init__(self, widths: Sequence[int], n_rows_total: int, cap_rows: int, group=None,
                 device: Optional[torch.device] = None):
        import ctypes as C
        from . import _lib
        L = _lib.lib()
        self._C, self._lib, self._check = C, L, _lib.check
        self.widths = [int(w) for w in widths]
        self._widths_c = (C.c_int * len(self.widths))(*self.widths)
        self.row = L.hgs_exchange_row_floats(self._widths_c, len(self.widths))
        if self.row <= 0:
            raise _lib.HgsError(f"unsupported tensor widths {self.widths}")
        self.cap_rows = (int(cap_rows) + 31) // 32 * 32
        self.n_ids = int(n_rows_total)
        world = dist.get_world_size(group)
        nbytes = L.hgs_exchange_mailbox_bytes(world, self.n_ids, self.cap_rows, self.row)
        if nbytes == 0:
            raise _lib.HgsError(f"unsupported exchange geometry: world {world}, N {self.n_ids}, cap {self.cap_rows}")
        self.box = _PeerMailbox(nbytes, group, device)
        self.world, self.rank = self.box.world, self.box.rank
        self.step = 0

    def _tensor_ptrs(self, tensors):
        assert len(tensors) == len(self.widths)
        for t, w in zip(tensors, self.widths):
            assert t.is_cuda and t.dtype == torch.float32 and t.is_contiguous(), "dense float32 CUDA tensors only"
            assert t.shape[0] == self.n_ids and t.numel() == self.n_ids * w, (t.shape, w)
        return (self._C.c_void_p * len(tensors))(*[t.data_ptr() for t in tensors])

    def push(self, tensors: Sequence[torch.Tensor], ids: torch.Tensor) -> None:
        """first half of exchange(): store this rank's records into every peer's mailbox and raise its flag"""
        C, L = self._C, self._lib
        assert ids.dtype == torch.int32 and ids.is_contiguous()
        n = int(ids.numel())
        # n > cap_rows is NOT raised here: the rank pushes an overflow marker instead, so that every rank fails the
        # same step (status 2, surfaced by poll_status / check_status) and the step counters stay in lock-step
        self.box.poll_status()
        self._check(L.hgs_exchange_push(self._tensor_ptrs(tensors), self._widths_c, len(tensors), self.n_ids,
                                        C.c_void_p(ids.data_ptr()) if n > 0 else None, n, self.cap_rows,
                                        self.box.ptrs_c, self.world, self.rank, self.step,
                                        torch.cuda.current_stream().cuda_stream), "hgs_exchange_push")

    def reduce(self, tensors: Sequence[torch.Tensor]) -> None:
        """second half: merge all ranks' records of this step into the dense tensors; ends the step"""
        C, L = self._C, self._lib
        self._check(L.hgs_exchange_reduce(self._tensor_ptrs(tensors), self._widths_c, len(tensors),
                                          self.n_ids, self.cap_rows, self.box.local, self.world, self.rank,
                                          self.step, C.c_void_p(self.box.status.data_ptr()),
                                          torch.cuda.current_stream().cuda_stream), "hgs_exchange_reduce")
        self.step += 1
        self.box.record_status()

    def exchange(self, tensors: Sequence[torch.Tensor], ids: torch.Tensor) -> None:
        """In place: tensors[k] ([N, widths[k]] float32, dense) become the SUM over ranks.  `ids` (int32, unique,
        ASCENDING) lists the rows of THIS rank that are non-zero (meta["visible_ids"] of a one-camera
        rasterization); rows a rank does not list must be zero on that rank."""
        self.push(tensors, ids)
        self.reduce(tensors)

    def _lib_error(self, msg):
        from . import _lib
        return _lib.HgsError(msg)

    def check_status(self) -> None:
        self.box.check_status()

    def close(self) -> None:
        self.box.close()


class FusedBackwardExchange:
    """The SH / projection backward of all ranks' views fused with the gradient exchange (csrc/exchange_vjp.cu).

    27 of the 38 gradient floats of a Gaussian are SH-coefficient gradients, and that part is rank one:
    basis(mean - camera position) x v_colour.  Every rank holds the means, so a rank runs the camera-specific
    backward of its own view (projection VJP, SH direction gradient, densification norm) inside the push kernel and
    trades one 64-byte record per visible Gaussian (2.5x less NVLink traffic than the 38 gradients); the reduce
    kernel expands the SH part and sums the pairs of every Gaussian in rank order: bit-identical gradients
    everywhere, and sh_bwd + project3d_bwd + densify_stats + the all-reduce of a step collapse into one push and
    one reduce kernel.  One camera per rank and step.

        ex = FusedBackwardExchange(N, cap_rows=N)   # N can never overflow; 2 x world x cap_rows x 64 B of HBM
        with ex.deferred():                       # backward stops after the blend backward
            rc, ra, meta = rasterization(means, quats, scales, opacities, colors, viewmat[None], K[None], W, H, ...)
            loss(rc, ra).backward()
        ex.finish(means, quats, scales, opacities, colors, grad_accum, denom)   # sets .grad of the five tensors
    """

    def __init__(self, n_gaussians: int, cap_rows: int, group=None, device: Optional[torch.device] = None):
        import ctypes as C
        from . import _lib
        L = _lib.lib()
        self._C, self._lib, self._check = C, L, _lib.check
        self.n_ids = int(n_gaussians)
        self.cap_rows = (int(cap_rows) + 31) // 32 * 32
        world = dist.get_world_size(group)
        nbytes = L.hgs_exchange_vjp_mailbox_bytes(world, self.n_ids, self.cap_rows)
        if nbytes == 0:
            raise _lib.HgsError(f"unsupported exchange geometry: world {world}, N {self.n_ids}, cap {self.cap_rows}")
        dev = torch.device("cuda", torch.cuda.current_device()) if device is None else device
        # NVSwitch multicast (one multimem.st per record instead of world - 1 unicast copies) is available as an option
        # (HGS_EXCHANGE_MULTICAST_MIN_RANKS=<n>) but OFF by default: every rank has to RECEIVE the other ranks' records
        # either way, so the push is bound by NVLink ingress (7 x 50 MB at 8 ranks), and measured on 8 x B200 the
        # multicast stores are slower (0.65 ms) than the per-peer TMA bulk stores (0.53 ms = 74 % of the ingress peak)
        import os
        mc_min = int(os.environ.get("HGS_EXCHANGE_MULTICAST_MIN_RANKS", "0"))
        self.box = _make_mailbox(nbytes, group, dev, want_multicast=mc_min > 0 and world >= mc_min)
        self.multicast = bool(self.box.multicast_ptr)
        self.world, self.rank = self.box.world, self.box.rank
        self.step = 0
        self.sink: dict = {}

    def deferred(self):
        from .cuda import _wrapper as W
        return W.deferred_backward(self.sink)

    def finish(self, means, quats, scales, opacities, colors, grad_accum: Optional[torch.Tensor] = None,
               denom: Optional[torch.Tensor] = None) -> None:
        """push this rank's blend-gradient rows, then compute the summed gradients of all ranks' views and store
        them in .grad of the five parameter tensors; grad_accum / denom ([N] float32, optional) receive += the
        densification statistics (scene/basic_model.py:131-144) of all views."""
        from . import _lib
        C, L, sk = self._C, self._lib, self.sink
        if "vpack" not in sk:
            raise _lib.HgsError("finish(): no deferred backward recorded (run rasterization + backward inside deferred())")
        N = self.n_ids
        vpack, ids = sk["vpack"], sk["vis_ids"]
        surfel = bool(sk.get("surfel", False))          # rasterization_2dgs: 24-float rows, surfel projection VJP
        assert vpack.shape == (1, N, 24 if surfel else 12) and vpack.is_contiguous() and sk["n"] == N
        n = int(ids.numel())
        # n > cap_rows is NOT raised here: the rank pushes an overflow marker, every rank's reduce then yields zero
        # gradients for this step and status 2 (surfaced one step later by poll_status, or by check_status)
        self.box.poll_status()
        sh_degree = sk["sh_degree"]
        cf = sk.get("colors_fwd")
        p = lambda t: None if t is None else C.c_void_p(t.data_ptr())  # noqa: E731
        for t in (means, quats, scales, opacities, colors):
            assert t.is_cuda and t.dtype == torch.float32 and t.is_contiguous() and t.shape[0] == N
        K = 1 if sh_degree is None else int(colors.shape[1])
        st = torch.cuda.current_stream().cuda_stream
        from .cuda import _wrapper as W
        W._mark("exchange_vjp_push", 0)
        deg = -1 if sh_degree is None else int(sh_degree)
        mc = C.c_void_p(self.box.multicast_ptr) if self.box.multicast_ptr else None
        dm, dq, ds, dc = means.detach(), quats.detach(), scales.detach(), colors.detach()
        if surfel:
            self._check(L.hgs_exchange_vjp_push_2dgs(deg, K, p(vpack), int(sk["has_depth"]), p(cf), p(sk["viewmats"]),
                                                     p(sk["Ks"]), p(sk["campos"]), p(dm), p(dq), p(ds), p(dc),
                                                     int(sk["width"]), int(sk["height"]), float(sk["near_plane"]),
                                                     float(sk["far_plane"]), N, p(ids) if n > 0 else None, n,
                                                     self.cap_rows, self.box.ptrs_c, mc, self.world, self.rank, self.step, st),
                        "hgs_exchange_vjp_push_2dgs")
        else:
            self._check(L.hgs_exchange_vjp_push(deg, K, p(vpack), p(cf), p(sk["viewmats"]), p(sk["Ks"]), p(sk["campos"]),
                                                p(dm), p(dq), p(ds), p(dc), int(sk["width"]), int(sk["height"]),
                                                float(sk["eps2d"]), float(sk["near_plane"]), float(sk["far_plane"]), N,
                                                p(ids) if n > 0 else None, n, self.cap_rows, self.box.ptrs_c, mc, self.world,
                                                self.rank, self.step, st), "hgs_exchange_vjp_push")
        W._mark("exchange_vjp_push", 1)
        outs = [torch.empty_like(t) for t in (means, quats, scales, opacities, colors)]
        W._mark("exchange_vjp_reduce", 0)
        self._check(L.hgs_exchange_vjp_reduce(deg, K, p(dm), N, self.cap_rows, self.box.local, self.world, self.rank,
                                              self.step, p(outs[0]), p(outs[1]), p(outs[2]), p(outs[3]), p(outs[4]),
                                              p(grad_accum), p(denom), p(self.box.status), st),
                    "hgs_exchange_vjp_reduce")
        W._mark("exchange_vjp_reduce", 1)
        self.step += 1
        self.box.record_status()
        for t, g in zip((means, quats, scales, opacities, colors), outs):
            t.grad = g
        sk.clear()

    def check_status(self) -> None:
        self.box.check_status()

    def close(self) -> None:
        self.box.close()
