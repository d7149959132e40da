"""Multi-GPU parity of the peer-memory gradient exchange (csrc/exchange.cu): the sparse
all-reduce over NVLink mailboxes must equal the dense NCCL all-reduce of the same tensors, be bit-identical on
every rank, and survive many steps (double-buffered slots, monotonic flags), empty contributions and overlapping
row sets.  Needs >= 2 GPUs (skipped otherwise); run with `gpurun --gpus 2`."""
import os
import socket

import pytest
import torch

pytestmark = pytest.mark.gpu

WIDTHS = (3, 4, 3, 1, 27, 1, 1)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_dir, n, steps):
    import torch.distributed as dist
    from horizongs_b200 import distributed as D
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    ex = D.PeerGradientExchange(WIDTHS, n, cap_rows=n, device=dev)
    ok = True
    worst = 0.0
    for s in range(steps):
        g = torch.Generator().manual_seed(1000 * s + rank)
        # a different visible fraction per rank and step; step 3: rank 1 sees nothing; step 4: everybody sees everything
        frac = [0.13, 0.5, 0.02, 0.3][(s + rank) % 4]
        if s == 3 and rank == 1:
            frac = 0.0
        if s == 4:
            frac = 1.0
        mask = torch.rand(n, generator=g) < frac
        ids = torch.nonzero(mask).flatten().to(torch.int32).to(dev)
        tensors = []
        for w in WIDTHS:
            t = torch.zeros(n, w)
            t[mask] = torch.randn(int(mask.sum()), w, generator=g)
            tensors.append((t if w > 1 else t.flatten()).to(dev).contiguous())
        ref = [t.clone() for t in tensors]
        for t in ref:
            dist.all_reduce(t)
        ex.exchange(tensors, ids)
        torch.cuda.synchronize()
        ex.check_status()
        for t, r in zip(tensors, ref):
            err = float((t - r).abs().max())
            worst = max(worst, err / (float(r.abs().max()) + 1e-12))
        # bit-identical replicas: compare with rank 0's result
        for t in tensors:
            t0 = t.clone()
            dist.broadcast(t0, 0)
            ok = ok and bool(torch.equal(t0, t))
    torch.save({"ok": ok, "worst": worst}, os.path.join(out_dir, f"r{rank}.pt"))
    ex.close()
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(600)
def test_peer_exchange_equals_dense_allreduce(tmp_path):
    import torch.multiprocessing as mp
    world = min(torch.cuda.device_count(), 8)
    if world < 2:
        pytest.skip("needs >= 2 GPUs")
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path), 20011, 7), nprocs=world, join=True)
    for r in range(world):
        got = torch.load(os.path.join(str(tmp_path), f"r{r}.pt"))
        assert got["ok"], f"rank {r}: replicas are not bit-identical"
        assert got["worst"] < 1e-5, f"rank {r}: differs from the dense all-reduce by {got['worst']}"
