"""Data-parallel gradient exchange over NVLink peer memory (one box).

The flat gradient buffer of every rank is allocated in symmetric memory (torch.distributed._symmetric_memory: CUDA VMM
allocations that all ranks of the group map), so a rank can read its peers' gradients with ordinary loads.  One kernel
(`stdadk_peer_allreduce`) then replaces the NCCL all-reduce of a step: two flag barriers through the same mappings and
a rank-ordered sum.  Being a plain kernel it is captured in the step's CUDA graph, which removes the graph boundary the
NCCL call needs.  If symmetric memory cannot be set up the trainer keeps the NCCL path.
"""
from __future__ import annotations

import ctypes as C

import torch
import torch.distributed as dist

from . import _lib as L
from . import ops

FLAG_WORDS = 64          # 2 * MAX_PEERS barrier words (+ padding), behind the gradient in the same allocation


class PeerExchange:
    def __init__(self, device):
        self.device = torch.device(device)
        self.rank, self.world = dist.get_rank(), dist.get_world_size()
        if self.world > L.MAX_PEERS:
            raise RuntimeError(f"peer exchange supports at most {L.MAX_PEERS} ranks, got {self.world}")
        self.buf = None

    def alloc(self, size: int) -> torch.Tensor:
        """Symmetric allocation holding `size` floats of gradient (+ the barrier words); returns the gradient view."""
        import torch.distributed._symmetric_memory as sm
        self.size = int(size)
        self.n_pad = (self.size + 3) // 4 * 4
        self.buf = sm.empty(self.n_pad + FLAG_WORDS, dtype=torch.float32, device=self.device)
        self.buf.zero_()
        self.handle = sm.rendezvous(self.buf, dist.group.WORLD)
        torch.cuda.synchronize(self.device)
        self.handle.barrier()                      # every rank's buffer (and its barrier words) is zero before first use
        self.out = torch.zeros(self.n_pad, dtype=torch.float32, device=self.device)
        self.ticket = torch.zeros(1, dtype=torch.int32, device=self.device)
        a = L.PeerAllreduceArgs()
        a.world, a.rank = self.world, self.rank
        for r, p in enumerate(self.handle.buffer_ptrs):
            a.src[r] = p
            a.flags[r] = p + 4 * self.n_pad
        a.out = self.out.data_ptr()
        a.n = self.n_pad
        a.ticket = self.ticket.data_ptr()
        self._args = a
        return self.buf[:self.size]

    def allreduce(self, step_count: torch.Tensor):
        """Sum over ranks of every rank's buf[:n_pad], written back into this rank's buffer.  `step_count` is the
        device-side step counter (identical on all ranks, incremented once per step after this call)."""
        self._args.step_count = step_count.data_ptr()
        ops.peer_allreduce(self._args)
        # the kernel returned only after every peer finished reading this rank's buffer: it may be overwritten now
        self.buf[:self.n_pad].copy_(self.out)
