"""Feasibility probe: torch symmetric memory (CUDA VMM peer mappings) between the ranks of one box."""
import os, sys, time
import torch, torch.distributed as dist
import torch.distributed._symmetric_memory as sm
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
try:
    t = sm.empty(1 << 20, dtype=torch.float32, device=torch.device("cuda", local))
    t.fill_(float(rank + 1))
    h = sm.rendezvous(t, dist.group.WORLD)
    print(rank, "ptrs", [hex(p) for p in h.buffer_ptrs], "pad", h.signal_pad_size, "multicast", h.has_multicast_support, flush=True)
    h.barrier()
    peer = h.get_buffer((rank + 1) % world, (16,), torch.float32)
    torch.cuda.synchronize()
    print(rank, "peer value", peer[:2].tolist(), flush=True)
    h.barrier()
    print(rank, "OK", flush=True)
except Exception as e:
    print(rank, "FAILED", repr(e)[:500], flush=True)
dist.barrier()
dist.destroy_process_group()
