"""Autoregressive rollout around an ``FNO`` -- the loop of the reference wrappers
(src/nsbench/models/fno/fno.py:29-43 ``FNOModule.forward``; closed-loop evaluation in
src/nsbench/scripts/evaluate.py), SURVEY.md row f2.

``sequence_forward`` has the reference's exact semantics (teacher forcing for the first
``teacher_forcing_steps`` frames, closed loop afterwards, outputs stacked on dim 1) and stays
differentiable.  ``Rollout`` is the inference engine for long closed-loop rollouts: one model step
is captured once as a CUDA graph on static buffers and replayed, the frames are written straight
into the preallocated ``[B, T, C, H, W]`` result, and nothing is read back to the host inside the
loop (the reference synchronises every step through ``.cpu()``).  Initial conditions are independent,
so multi-GPU rollouts shard by initial condition (``ddp.shard_batch``) with no collective.
"""
from __future__ import annotations

from typing import Optional

import torch
import torch.nn as nn

from . import _lib


def sequence_forward(fno: nn.Module, x: torch.Tensor, teacher_forcing_steps: int = 50) -> torch.Tensor:
    """x [B, T, C, H, W] -> [B, T, C_out, H, W]; reference: src/nsbench/models/fno/fno.py:29-43."""
    outs = []
    x_t = None
    for t in range(x.shape[1]):
        x_t = x[:, t] if t < teacher_forcing_steps else x_t
        x_t = fno(x_t)
        outs.append(x_t)
    return torch.stack(outs, dim=1)


class Rollout:
    """Closed-loop inference ``x_{t+1} = fno(x_t)`` for ``steps`` steps from ``x0 [B, C, H, W]``.

    The model must map C channels to C channels.  ``graph=True`` (default) captures one step; the
    captured step is reused for every later call with the same input shape.
    """

    def __init__(self, fno: nn.Module, graph: bool = True):
        self.fno = fno
        self.use_graph = graph
        self._graph: Optional[torch.cuda.CUDAGraph] = None
        self._x: Optional[torch.Tensor] = None
        self._y: Optional[torch.Tensor] = None

    def _capture(self, x0: torch.Tensor):
        self._x = torch.empty_like(x0)
        self._x.copy_(x0)
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s), torch.no_grad():
            for _ in range(2):                         # warm-up: plans, workspaces, allocator pools
                self._y = self.fno(self._x)
        torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize()
        self._graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self._graph), torch.no_grad():
            self._y = self.fno(self._x)
        if self._y.shape != self._x.shape:
            raise _lib.SpectralB200Error(f"Rollout: the model maps {tuple(self._x.shape)} to {tuple(self._y.shape)}; "
                                         "a closed loop needs identical input and output shapes")

    @torch.no_grad()
    def __call__(self, x0: torch.Tensor, steps: int, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        if not x0.is_cuda:
            raise _lib.SpectralB200Error("Rollout got a CPU tensor: there is no CPU path")
        x0 = x0.contiguous().float()
        B = x0.shape[0]
        if out is None:
            out = torch.empty(B, steps, *x0.shape[1:], device=x0.device, dtype=torch.float32)
        if not self.use_graph:
            x = x0
            for t in range(steps):
                x = self.fno(x)
                out[:, t].copy_(x)
            return out
        if self._graph is None or self._x.shape != x0.shape or self._x.device != x0.device:
            self._capture(x0)
        self._x.copy_(x0)
        for t in range(steps):
            self._graph.replay()
            out[:, t].copy_(self._y)
            if t + 1 < steps:
                self._x.copy_(self._y)
        return out
