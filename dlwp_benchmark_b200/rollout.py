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


def _flat_tc(t: torch.Tensor) -> torch.Tensor:
    """einops "b t c h w -> b (t c) h w" on a contiguous tensor."""
    return t.reshape(t.shape[0], t.shape[1] * t.shape[2], *t.shape[3:])


def dlwp_sequence_forward(fno: nn.Module, constants: Optional[torch.Tensor], prescribed: Optional[torch.Tensor],
                          prognostic: torch.Tensor, context_size: int) -> torch.Tensor:
    """``FNO2DModule.forward`` of dlwpbench (src/dlwpbench/models/fno/fno.py:64-106), differentiable:

        for t in [ctx, T):  prognostic_t = last ctx frames of [true frames before ctx ; outputs so far]
                            x_t = cat(constants[:, 0], prescribed[:, t-ctx:t] (t c), prognostic_t (t c))
                            out_t = prognostic_t[:, -1] + fno(x_t)

    with the evident intent of the reference's two device slips (``list.to`` at :91-95, ``out.cpu()`` at :104): every
    frame stays on the device, nothing is read back inside the loop.  Returns [B, T-ctx, C_prog, H, W]."""
    ctx = int(context_size)
    outs = []
    T = prognostic.shape[1]
    for t in range(ctx, T):
        t_start = max(0, t - ctx)
        if t == ctx:
            prog_t = prognostic[:, t_start:t]
        else:
            prog_t = torch.cat([prognostic[:, t_start:ctx], torch.stack(outs, dim=1)[:, -ctx:]], dim=1)
        parts = []
        if constants is not None:
            parts.append(constants[:, 0])
        if prescribed is not None:
            parts.append(_flat_tc(prescribed[:, t - ctx:t]))
        parts.append(_flat_tc(prog_t))
        outs.append(prog_t[:, -1] + fno(torch.cat(parts, dim=1)))
    return torch.stack(outs, dim=1)


class DLWPRollout:
    """Inference engine for the dlwpbench loop (same semantics as ``dlwp_sequence_forward``): one model step is
    captured as a CUDA graph over a static input buffer ``[B, Cc + ctx*(Cp+Cg), H, W]``; per frame the prescribed
    window and the newest prognostic frame are copied into their channel slots on the device (the older
    prognostic frames are shifted inside the buffer), the graph is replayed, and the residual add writes the
    frame straight into the preallocated result.  Nothing is read back to the host inside the loop."""

    def __init__(self, fno: nn.Module, context_size: int, graph: bool = True):
        self.fno, self.ctx, self.use_graph = fno, int(context_size), graph
        self._graph: Optional[torch.cuda.CUDAGraph] = None
        self._x: Optional[torch.Tensor] = None
        self._y: Optional[torch.Tensor] = None

    def _step(self):
        if self._graph is not None:
            self._graph.replay()
        else:
            self._y = self.fno(self._x)

    def _prepare(self, shape, device):
        if self._x is not None and tuple(self._x.shape) == tuple(shape) and self._x.device == device:
            return
        self._x = torch.zeros(*shape, device=device, dtype=torch.float32)
        self._graph = None
        if not self.use_graph:
            return
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            for _ in range(2):
                self._y = self.fno(self._x)
        torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            self._y = self.fno(self._x)
        self._graph = g

    @torch.no_grad()
    def __call__(self, constants, prescribed, prognostic, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        if not prognostic.is_cuda:
            raise _lib.SpectralB200Error("DLWPRollout got a CPU tensor: there is no CPU path")
        ctx = self.ctx
        B, T, Cg, H, W = prognostic.shape
        Cc = constants.shape[2] if constants is not None else 0
        Cp = prescribed.shape[2] if prescribed is not None else 0
        cin = Cc + ctx * (Cp + Cg)
        self._prepare((B, cin, H, W), prognostic.device)
        x = self._x
        if out is None:
            out = torch.empty(B, T - ctx, Cg, H, W, device=prognostic.device, dtype=torch.float32)
        if Cc:
            x[:, :Cc].copy_(constants[:, 0])
        p0 = Cc + ctx * Cp                                  # first prognostic channel
        xg = x[:, p0:].view(B, ctx, Cg, H, W)
        xg.copy_(prognostic[:, :ctx])
        for i, t in enumerate(range(ctx, T)):
            if Cp:
                x[:, Cc:p0].copy_(_flat_tc(prescribed[:, t - ctx:t]))
            if i > 0:
                if ctx > 1:
                    xg[:, :-1].copy_(xg[:, 1:].clone())
                xg[:, -1].copy_(out[:, i - 1])
            self._step()
            torch.add(xg[:, -1], self._y, out=out[:, i])
        return out
