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
        # gradients plus a small scratch tail (loss scalars, per-step accumulators): ONE memset per step clears both
        self.total = total
        pad = (-total) % 4
        self.grad_all = torch.zeros(total + pad + self.SCRATCH, device=dev, dtype=dtype)
        self.grad = self.grad_all[:total]
        self.scratch = self.grad_all[total + pad:]
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

    SCRATCH = 512  # floats

    def zero_grad(self):
        self.grad_all.zero_()

    def all_reduce_grads(self, group: Optional[dist.ProcessGroup] = None) -> int:
        """ONE sum all-reduce of the whole gradient buffer; returns the world size (the caller scales)."""
        if not dist.is_initialized():
            return 1
        world = dist.get_world_size(group)
        if world > 1:
            dist.all_reduce(self.grad, op=dist.ReduceOp.SUM, group=group)
        return world

    def enable_p2p(self, group: Optional[dist.ProcessGroup] = None, max_elems: int = 1 << 20) -> bool:
        """Set up the peer-memory mailboxes of ``cgat_p2p_allreduce_adam`` (one NVLink-mapped buffer per rank, exchanged
        through torch's symmetric-memory rendezvous).  Returns False -- and leaves the NCCL path in place -- when the
        job is single-rank, the vector is large, or symmetric memory is unavailable on this system."""
        import ctypes

        from . import _lib

        self.p2p = None
        if not (dist.is_initialized() and self.param.is_cuda):
            return False
        world, rank = dist.get_world_size(group), dist.get_rank(group)
        n = self.param.numel()
        if world < 2 or world > 8 or n > max_elems:
            return False
        try:
            import torch.distributed._symmetric_memory as symm

            nbytes = _lib.lib().cgat_p2p_mailbox_bytes(n, world)
            box = symm.empty(nbytes, dtype=torch.uint8, device=self.param.device)
            box.zero_()
            hdl = symm.rendezvous(box, group if group is not None else dist.group.WORLD)
            ptrs = (ctypes.c_uint64 * world)(*[int(p) for p in hdl.buffer_ptrs])
            torch.cuda.synchronize()
            dist.barrier(group)  # every mailbox is zeroed before anyone pushes
        except Exception as e:  # noqa: BLE001 -- any failure here means "no peer memory": keep NCCL
            self.p2p_error = repr(e)
            return False
        n_pad = (n + 31) & ~31
        off = 2 * world * n_pad * 8  # the time-out marker follows the {value, epoch} words
        self.p2p = dict(box=box, hdl=hdl, ptrs=ptrs, rank=rank, world=world, n=n,
                        marker=box[off:off + 4].view(torch.int32),
                        marker_host=torch.zeros(1, dtype=torch.int32).pin_memory())
        return True

    def p2p_poll_timeout(self, refresh: bool) -> bool:
        """Non-blocking view of the exchange kernel's time-out marker: ``refresh`` enqueues an asynchronous copy of it
        into pinned host memory on the current stream; the return value is whatever the LAST completed copy saw (so a
        time-out surfaces a few steps late, without ever synchronising the training stream)."""
        p2p = getattr(self, "p2p", None)
        if p2p is None:
            return False
        seen = bool(p2p["marker_host"].item())
        if refresh:
            p2p["marker_host"].copy_(p2p["marker"], non_blocking=True)
        return seen

    def p2p_timed_out(self) -> bool:
        """True if the exchange kernel ever gave up waiting for a peer (host sync; for tests / diagnostics)."""
        p2p = getattr(self, "p2p", None)
        if p2p is None:
            return False
        return bool(p2p["marker"].item())

    def broadcast_params(self, src: int = 0, group: Optional[dist.ProcessGroup] = None):
        if dist.is_initialized() and dist.get_world_size(group) > 1:
            dist.broadcast(self.param, src=src, group=group)
