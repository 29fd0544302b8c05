"""Data-parallel plumbing of the train step: flat parameter / gradient buffers, batch sharding and the
single per-step gradient all-reduce (torch.distributed: NCCL over NVLink on the GPUs, gloo in CPU tests).

The reference has no parallelism at all; samples are independent in the conv-GAT path, so
the batch is split evenly over ranks, parameters are replicated, and the only exchange per step is ONE sum
all-reduce over the flat fp32 gradient buffer (43,936 .. 5.5 M elements: latency-bound, hence never
per-tensor).  The 1/world scale is folded into the Adam kernel.
"""
from __future__ import annotations

from typing import Iterable, List, Optional, Tuple

import torch
import torch.distributed as dist


def shard_range(n: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous [lo, hi) slice of a batch of ``n`` samples owned by ``rank`` (sizes differ by at most one)."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


class FlatParams:
    """Re-homes ``params`` into one flat buffer with ``p.data`` / ``p.grad`` as views (device agnostic)."""

    def __init__(self, named_params: Iterable[Tuple[str, torch.nn.Parameter]], dtype=torch.float32):
        self.named: List[Tuple[str, torch.nn.Parameter]] = list(named_params)
        if not self.named:
            raise ValueError("no parameters to flatten")
        dev = self.named[0][1].device
        total = sum(p.numel() for _, p in self.named)
        self.param = torch.empty(total, device=dev, dtype=dtype)
        self.grad = torch.zeros(total, device=dev, dtype=dtype)
        self.offsets = {}
        off = 0
        with torch.no_grad():
            for name, p in self.named:
                n = p.numel()
                self.param[off:off + n].copy_(p.detach().to(dtype).reshape(-1))
                p.data = self.param[off:off + n].view(p.shape)
                p.grad = self.grad[off:off + n].view(p.shape)
                self.offsets[name] = (off, n)
                off += n

    def zero_grad(self):
        self.grad.zero_()

    def all_reduce_grads(self, group: Optional[dist.ProcessGroup] = None) -> int:
        """ONE sum all-reduce of the whole gradient buffer; returns the world size (the caller scales)."""
        if not dist.is_initialized():
            return 1
        world = dist.get_world_size(group)
        if world > 1:
            dist.all_reduce(self.grad, op=dist.ReduceOp.SUM, group=group)
        return world

    def broadcast_params(self, src: int = 0, group: Optional[dist.ProcessGroup] = None):
        if dist.is_initialized() and dist.get_world_size(group) > 1:
            dist.broadcast(self.param, src=src, group=group)
