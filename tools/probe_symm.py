"""Development aid: does torch's symmetric memory give us peer pointers and an NVLink multicast pointer here?"""
import os
import torch
import torch.distributed as dist

rank = int(os.environ["RANK"])
torch.cuda.set_device(rank)
dev = torch.device("cuda", rank)
dist.init_process_group("nccl", device_id=dev)
import torch.distributed._symmetric_memory as symm_mem
try:
    t = symm_mem.empty(1 << 20, dtype=torch.float32, device=dev)
    h = symm_mem.rendezvous(t, dist.group.WORLD)
    print(rank, "buffer_ptrs", [hex(p) for p in h.buffer_ptrs], "multicast_ptr", hex(h.multicast_ptr) if h.multicast_ptr else 0,
          "signal_pad_ptrs", len(h.signal_pad_ptrs), flush=True)
except Exception as e:
    print(rank, "symm_mem failed:", repr(e)[:500], flush=True)
dist.barrier()
dist.destroy_process_group()
