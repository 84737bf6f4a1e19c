"""Data-parallel harness for the FNO train step: one process per GPU (torchrun), identical
weights, batch split evenly; the only collective is the gradient all-reduce.

``GradSync`` owns one flat fp32 gradient buffer; every ``param.grad`` is a (256-byte aligned) view into it.

* ``direct=True`` (CUDA, default there): every parameter also carries that view as its *gradient sink*.  The
  backward kernels of this package (``FusedChainFn``, ``TuckerReconstructManyFn``) write their weight / bias
  gradients straight into the sink through the C ABI and return ``None`` to autograd, so a step has no
  ``AccumulateGrad`` add per parameter and no ``flat.zero_()``: the sink is overwritten.  Parameters that some
  other autograd node differentiates (anything outside those Functions) still arrive through ``AccumulateGrad``;
  ``zero()`` clears exactly those ranges (learned from the first backward; before that, everything).
  Overwriting means one backward per step and no weight sharing between layers -- gradient accumulation over
  micro-batches needs ``direct=False``.
* The all-reduce is ONE ``ncclAllReduce`` with ``ReduceOp.AVG`` per bucket on a side stream (no separate scaling
  kernel); ``wait()`` joins the streams before the optimizer step.  The default bucket (64 MB) holds every
  configuration of the reference in a single collective: at 15-70 MB of gradients the NVLink time (~40-100 us)
  is far below the per-collective launch cost of splitting it, and the largest gradients (the spectral / Tucker
  weights) are produced last in the backward, so there is nothing left to overlap them with.
* Works with ``gloo`` on CPU (world_size-2 tests): SUM followed by a scale, since gloo has no AVG.
"""
from __future__ import annotations

from typing import Iterable, List, Optional

import torch
import torch.distributed as dist

_ALIGN = 64   # floats: 256-byte aligned views (vector stores / TMA-friendly destinations)


class GradSync:
    def __init__(self, params: Iterable[torch.nn.Parameter], world_size: int, bucket_bytes: int = 64 << 20,
                 group=None, direct: Optional[bool] = None, avg: Optional[bool] = None):
        self.params: List[torch.nn.Parameter] = [p for p in params if p.requires_grad]
        self.world = int(world_size)
        self.group = group
        if not self.params:
            raise ValueError("GradSync: no trainable parameters")
        dev, dt = self.params[0].device, self.params[0].dtype
        self.direct = (dev.type == "cuda") if direct is None else bool(direct)
        offs, off = [], 0
        for p in self.params:
            offs.append(off)
            off += (p.numel() + _ALIGN - 1) // _ALIGN * _ALIGN
        self.flat = torch.zeros(off, device=dev, dtype=dt)
        self.spans = []
        self.buckets: List[torch.Tensor] = []
        per = max(1, bucket_bytes // self.flat.element_size())
        b_start = 0
        for p, o in zip(self.params, offs):
            n = p.numel()
            view = self.flat[o:o + n]
            p.grad = view.view_as(p)
            self.spans.append((o, o + n))
            if self.direct:
                p._sb200_grad_sink = view
                p._sb200_sink_used = False
            end = o + (n + _ALIGN - 1) // _ALIGN * _ALIGN
            if end - b_start >= per:
                self.buckets.append(self.flat[b_start:end])
                b_start = end
        if b_start < off:
            self.buckets.append(self.flat[b_start:off])
        self.stream = torch.cuda.Stream(device=dev) if dev.type == "cuda" else None
        # NCCL has ReduceOp.AVG (gloo does not); avg=False = SUM followed by one scaling kernel
        self._avg = (dev.type == "cuda") if avg is None else (bool(avg) and dev.type == "cuda")

    def detach(self):
        """Remove the sinks (parameters keep their last gradients as ordinary tensors)."""
        for p in self.params:
            for a in ("_sb200_grad_sink", "_sb200_sink_used"):
                if hasattr(p, a):
                    delattr(p, a)
            if p.grad is not None:
                p.grad = p.grad.clone()

    def zero(self):
        """Clear what autograd will accumulate into: everything, minus the ranges the kernels overwrite."""
        if not self.direct:
            self.flat.zero_()
            return
        run = None
        for p, (a, b) in zip(self.params, self.spans):
            if getattr(p, "_sb200_sink_used", False):
                if run is not None:
                    self.flat[run[0]:run[1]].zero_()
                    run = None
            else:
                run = (run[0], b) if run is not None else (a, b)
        if run is not None:
            self.flat[run[0]:run[1]].zero_()

    def _reduce(self, b):
        if self._avg:
            dist.all_reduce(b, op=dist.ReduceOp.AVG, group=self.group)
        else:
            dist.all_reduce(b, op=dist.ReduceOp.SUM, group=self.group)
            b.mul_(1.0 / self.world)

    def allreduce(self, async_op: bool = False):
        """Mean over ranks (like DDP)."""
        if self.world <= 1:
            return
        if self.stream is not None:
            self.stream.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(self.stream):
                for b in self.buckets:
                    self._reduce(b)
            if not async_op:
                self.wait()
        else:
            for b in self.buckets:
                self._reduce(b)

    def wait(self):
        if self.stream is not None:
            torch.cuda.current_stream().wait_stream(self.stream)


def shard_batch(n_items: int, rank: int, world: int):
    """Even split of ``n_items`` (samples / initial conditions) over ranks: [lo, hi)."""
    base, rem = divmod(n_items, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)
