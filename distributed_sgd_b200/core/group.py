"""Process-group plumbing: one process per GPU, `torch.distributed` for the control plane only.

Replaces the reference's cluster membership and channels (core/Master.scala:222-253; core/package.scala:
16-21): ranks exchange small byte strings (the NCCL unique id, IPC handles) and a few integers per
evaluation.  Gradients never travel through this class -- they are reduced on the device inside
libdsgd.so.  Works with the gloo backend on CPU (tests) and nccl/gloo on the GPU box.
"""
from __future__ import annotations

import pickle
from typing import List, Sequence


class Group:
    def __init__(self):
        import torch.distributed as dist
        self._dist = dist
        self.active = dist.is_available() and dist.is_initialized()
        self.rank = dist.get_rank() if self.active else 0
        self.world = dist.get_world_size() if self.active else 1
        self._device = None

    def _dev(self):
        import torch
        if self._device is None:
            backend = self._dist.get_backend() if self.active else "gloo"
            self._device = torch.device("cuda", torch.cuda.current_device()) if backend == "nccl" else torch.device("cpu")
        return self._device

    def barrier(self):
        if self.active:
            self._dist.barrier()

    def broadcast_bytes(self, payload: bytes, src: int = 0) -> bytes:
        if not self.active:
            return payload
        box = [payload if self.rank == src else None]
        self._dist.broadcast_object_list(box, src=src)
        return box[0]

    def all_gather_bytes(self, payload: bytes) -> List[bytes]:
        if not self.active:
            return [payload]
        out = [None] * self.world
        self._dist.all_gather_object(out, payload)
        return out

    def all_reduce_sum(self, values: Sequence[float]) -> List[float]:
        """Sum a few scalars over ranks (evaluation counters; exact for integers below 2^53)."""
        if not self.active:
            return list(values)
        import torch
        t = torch.tensor(list(values), dtype=torch.float64, device=self._dev())
        self._dist.all_reduce(t, op=self._dist.ReduceOp.SUM)
        return t.cpu().tolist()

    def all_reduce_max(self, value: float) -> float:
        if not self.active:
            return value
        import torch
        t = torch.tensor([value], dtype=torch.float64, device=self._dev())
        self._dist.all_reduce(t, op=self._dist.ReduceOp.MAX)
        return float(t.item())
