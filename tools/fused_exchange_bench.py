#!/usr/bin/env python
"""Time FusedBackwardExchange.finish() (push + fused SH/projection backward of all ranks' rows) on bench.py's
6M scene.  Single process = world of one (the kernels still run: useful under ncu); or under torchrun."""
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import horizongs_b200 as hgs  # noqa: E402
from horizongs_b200 import distributed as D, scenes  # noqa: E402
from horizongs_b200.cuda import _wrapper as Wr  # noqa: E402


def main():
    if "RANK" not in os.environ:
        os.environ.update(RANK="0", WORLD_SIZE="1", LOCAL_RANK="0", MASTER_ADDR="127.0.0.1", MASTER_PORT="29534")
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    n = int(os.environ.get("EX_N", 6_000_000))
    sc, views, Ks, W, H = scenes.config4(n=n)
    sc = sc.to(dev)
    views, Ks = views.to(dev), Ks.to(dev)
    params = [t.requires_grad_() for t in (sc.means, sc.quats, sc.scales, sc.opacities, sc.colors)]
    ex = D.FusedBackwardExchange(n, cap_rows=n // 4, device=dev)
    stats = torch.zeros(2, n, device=dev)
    marks = []
    Wr.set_stage_hook(lambda name, ph: marks.append((name, ph, _rec())))

    def _rec():
        e = torch.cuda.Event(enable_timing=True)
        e.record()
        return e
    reps = int(os.environ.get("EX_REPS", 6))
    for s in range(reps):
        v = (rank + s) % 8
        with ex.deferred():
            rc, ra, meta = hgs.rasterization(*params, views[v:v + 1], Ks[v:v + 1], W, H, sh_degree=2, render_mode="RGB+ED")
            (rc.mean() + ra.mean()).backward()
        dist.barrier()
        torch.cuda.synchronize()
        ex.finish(*params, grad_accum=stats[0], denom=stats[1])
        torch.cuda.synchronize()
    ex.check_status()
    open_, out = {}, {}
    for name, ph, ev in marks:
        if ph == 0:
            open_[name] = ev
        else:
            out.setdefault(name, []).append(round(open_.pop(name).elapsed_time(ev), 4))
    if rank == 0:
        print(json.dumps({"world": world, "N": n, "push_ms": out.get("exchange_vjp_push"),
                          "reduce_ms": out.get("exchange_vjp_reduce")}), flush=True)
    ex.close()
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
