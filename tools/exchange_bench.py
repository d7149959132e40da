#!/usr/bin/env python
"""Time the peer-memory sparse gradient exchange against the dense NCCL all-reduce (torchrun, one rank per GPU).
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/exchange_bench.py
Prints one JSON line per method (rank 0): ms per exchange (max over ranks), bytes moved, GB/s per GPU."""
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from horizongs_b200 import distributed as D  # noqa: E402

N = int(os.environ.get("EX_N", 6_000_000))
FRAC = float(os.environ.get("EX_FRAC", 0.13))
WIDTHS = (3, 4, 3, 1, 27, 1, 1)


def main():
    if "RANK" not in os.environ:      # single process (e.g. under ncu): world of one, the kernels still run
        os.environ.update(RANK="0", WORLD_SIZE="1", LOCAL_RANK="0", MASTER_ADDR="127.0.0.1", MASTER_PORT="29533")
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    g = torch.Generator().manual_seed(rank)
    mask = torch.rand(N, generator=g) < FRAC
    ids = torch.nonzero(mask).flatten().to(torch.int32).to(dev)
    tensors = [torch.zeros((N, w) if w > 1 else (N,), device=dev) for w in WIDTHS]
    for t in tensors:
        t[ids.long()] = 1.0

    def timed(fn, reps=10, warm=3):
        for _ in range(warm):
            fn()
        dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1) / reps], device=dev)
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms)

    def nccl():
        hs = [dist.all_reduce(t, async_op=True) for t in tensors]
        for h in hs:
            h.wait()

    flat = torch.zeros(N * sum(WIDTHS), device=dev)

    def nccl_flat():
        dist.all_reduce(flat)

    ex = D.PeerGradientExchange(WIDTHS, N, cap_rows=int(N * min(1.0, FRAC * 1.5 + 0.01)), device=dev)

    def peer():
        ex.exchange(tensors, ids)

    res = {"nccl_per_tensor": timed(nccl), "nccl_flat": timed(nccl_flat), "peer_sparse": timed(peer)}
    # the two halves on their own (every rank pushes, then every rank reduces)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    dist.barrier()
    torch.cuda.synchronize()
    ev[0].record()
    ex.push(tensors, ids)
    ev[1].record()
    ex.reduce(tensors)
    ev[2].record()
    torch.cuda.synchronize()
    res["push_only"], res["reduce_only"] = ev[0].elapsed_time(ev[1]), ev[1].elapsed_time(ev[2])
    ex.check_status()
    if rank == 0:
        rows = int(ids.numel())
        dense = N * sum(WIDTHS) * 4
        print(json.dumps({"world": world, "N": N, "rows_per_rank": rows, "dense_bytes": dense,
                          "sparse_bytes_in_per_gpu": (world - 1) * rows * (sum(WIDTHS) + 1) * 4, "ms": res}), flush=True)
    ex.close()
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
