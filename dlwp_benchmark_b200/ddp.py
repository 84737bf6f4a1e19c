"""Data-parallel harness for the FNO train step: one process per GPU (torchrun), identical
weights, batch split evenly; the only collective is the gradient all-reduce.

``GradSync`` owns one flat fp32 gradient buffer; every ``param.grad`` is a view into it, so the
backward kernels (which accumulate into ``.grad`` in place) write straight into the all-reduce
buckets and no gather/scatter copies are needed.  Buckets are all-reduced (sum / world) with
NCCL over NVLink on a side stream so they overlap whatever is still running on the compute
stream; ``wait()`` joins the streams before the optimizer step.  Works with ``gloo`` on CPU
(used by the world_size-2 tests).
"""
from __future__ import annotations

from typing import Iterable, List

import torch
import torch.distributed as dist


class GradSync:
    def __init__(self, params: Iterable[torch.nn.Parameter], world_size: int, bucket_bytes: int = 64 << 20,
                 group=None):
        self.params: List[torch.nn.Parameter] = [p for p in params if p.requires_grad]
        self.world = int(world_size)
        self.group = group
        if not self.params:
            raise ValueError("GradSync: no trainable parameters")
        dev, dt = self.params[0].device, self.params[0].dtype
        total = sum(p.numel() for p in self.params)
        self.flat = torch.zeros(total, device=dev, dtype=dt)
        off = 0
        self.buckets: List[torch.Tensor] = []
        b_start = 0
        per = max(1, bucket_bytes // self.flat.element_size())
        for p in self.params:
            n = p.numel()
            p.grad = self.flat[off:off + n].view_as(p)
            off += n
            if off - b_start >= per:
                self.buckets.append(self.flat[b_start:off])
                b_start = off
        if b_start < off:
            self.buckets.append(self.flat[b_start:off])
        self.stream = torch.cuda.Stream(device=dev) if dev.type == "cuda" else None

    def zero(self):
        self.flat.zero_()

    def allreduce(self, async_op: bool = False):
        """Sum over ranks and divide by world (mean gradient, like DDP)."""
        if self.world <= 1:
            return
        if self.stream is not None:
            self.stream.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(self.stream):
                for b in self.buckets:
                    dist.all_reduce(b, op=dist.ReduceOp.SUM, group=self.group)
                    b.mul_(1.0 / self.world)
            if not async_op:
                self.wait()
        else:
            for b in self.buckets:
                dist.all_reduce(b, op=dist.ReduceOp.SUM, group=self.group)
                b.mul_(1.0 / self.world)

    def wait(self):
        if self.stream is not None:
            torch.cuda.current_stream().wait_stream(self.stream)


def shard_batch(n_items: int, rank: int, world: int):
    """Even split of ``n_items`` (samples / initial conditions) over ranks: [lo, hi)."""
    base, rem = divmod(n_items, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)
