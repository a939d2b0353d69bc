"""Data-parallel gradient exchange over NVLink peer memory (one box).

Every rank owns a receive area in symmetric memory (torch.distributed._symmetric_memory: CUDA VMM allocations that all
ranks of the group map).  One kernel (`stdadk_peer_allreduce`, low-latency protocol: each 8-byte packet carries its own
epoch flag) pushes the rank's flat gradient into every peer's area, reduces what the peers pushed into its own in rank
order -- in place, bit-identical on all ranks -- and can finish with the gradient norm of the reduced vector.  Being a
plain kernel it is captured in the step's single CUDA graph: no NCCL call, no graph boundary, no separate barrier.
If symmetric memory cannot be set up, or the gradient is too large for the double-buffered areas, the trainer keeps the
NCCL all-reduce (bandwidth-bound exchanges are NCCL's home ground; this path is for the latency-bound small model).
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional, Sequence

import torch
import torch.distributed as dist

from . import _lib as L
from . import ops

MAX_FLOATS = 1 << 22          # 4M floats: 16 * world * n bytes of receive area (512 MB at 8 ranks); beyond that: NCCL


class PeerExchange:
    def __init__(self, device, n_floats: int):
        import torch.distributed._symmetric_memory as sm
        self.device = torch.device(device)
        self.rank, self.world = dist.get_rank(), dist.get_world_size()
        if self.world > L.MAX_PEERS:
            raise RuntimeError(f"peer exchange supports at most {L.MAX_PEERS} ranks, got {self.world}")
        self.n = (int(n_floats) + 3) // 4 * 4
        if self.n > MAX_FLOATS:
            raise RuntimeError(f"gradient of {self.n} floats exceeds the peer-exchange limit ({MAX_FLOATS})")
        # [2 parities][world sources][n / 2] packets of 16 bytes = 4 floats
        self.recv = sm.empty(2 * self.world * (self.n // 2) * 4, dtype=torch.float32, device=self.device)
        self.recv.zero_()
        self.handle = sm.rendezvous(self.recv, dist.group.WORLD)
        torch.cuda.synchronize(self.device)
        self.handle.barrier()                      # every rank's area is zero (epoch 0 = "nothing yet") before first use
        self.ws = torch.zeros(148 * 8 + 8, dtype=torch.float32, device=self.device)
        a = L.PeerAllreduceArgs()
        a.world, a.rank, a.n = self.world, self.rank, self.n
        a.mode = int(os.environ.get("STDADK_PEER_MODE", "0"))     # 0 automatic, 1 one-shot, 2 two-phase (include/stdadk.h)
        for r, p in enumerate(self.handle.buffer_ptrs):
            a.recv[r] = p
        self._args = a
        self._keep = None

    def allreduce(self, g: torch.Tensor, step_count: torch.Tensor, group_end: Optional[Sequence[int]] = None,
                  sqnorms: Optional[torch.Tensor] = None):
        """g[:n] <- sum over ranks, in place (g must hold at least n floats, 16-byte aligned).  `step_count`: the
        device-side step counter (identical on all ranks, incremented once per step after this call).  With
        `group_end` / `sqnorms` the squared norm per group of the reduced gradient is produced by the same kernel."""
        a = self._args
        if g.numel() < self.n:
            raise RuntimeError("peer exchange: gradient buffer shorter than the padded exchange length")
        a.g = g.data_ptr()
        a.step_count = step_count.data_ptr()
        if group_end is not None:
            arr = (C.c_int64 * len(group_end))(*group_end)
            a.n_groups, a.group_end, a.sqnorms, a.workspace = len(group_end), arr, sqnorms.data_ptr(), self.ws.data_ptr()
            self._keep = arr
        else:
            a.n_groups = 0
        ops.peer_allreduce(a)
