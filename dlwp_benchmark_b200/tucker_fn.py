"""Tucker reconstruction ``W = core x_0 U0 x_1 U1 x_2 U2 x_3 U3`` and its backward on the C-ABI strided complex
GEMM (``sb200_cgemm_grouped``): four mode products forward; per mode one factor gradient (split reduction) and
one core-side gradient backward.  No tensor is ever permuted: unfoldings are expressed as composite indices.
The L layers of an FNO have identical geometry, so each of those products is ONE launch for all layers.

Replaces tltorch ``TuckerTensor.to_tensor`` / neuralop ``_contract_tucker`` (einsum
``"abcd,fghi,bf,eg,ch,di->aecd"``; reached from ``TFNO(...)`` at src/dlwpbench/models/fno/fno.py:136-146) and the
autograd backward of that einsum.  Contracting the factors into a dense ``W`` first is mathematically the same
product, costs O(|W| r) independent of the batch, and lets the per-mode channel mix run unchanged.
"""
from __future__ import annotations

import torch

from . import _lib, ops

MAX_GROUP = 8


def _empty(dev, n, *shape):
    return [torch.empty(*shape, 2, device=dev, dtype=torch.float32) for _ in range(n)]


def _mode_products(cores, Us, s, r):
    """cores: list (per layer) of T0; Us[k]: list (per layer) of factor k.  Returns [T0, T1, T2, T3, T4 = W]
    (each a per-layer list); T_{k+1} = T_k x_k U_k with shape [s0..s_k, r_{k+1}..r3]."""
    dev, n = cores[0].device, len(cores)
    R, Q = r[1] * r[2] * r[3], r[2] * r[3]
    # mode 0:  T1[i, R] = sum_f U0[i,f] T0[f, R]                      rows m = R (contiguous), n = i
    T1 = _empty(dev, n, s[0], r[1], r[2], r[3])
    ops.cgemm(cores, Us[0], T1, M=R, N=s[0], K=r[0], sAm=1, sAk=R, sBk=1, sBn=r[0], sCm=1, sCn=R)
    # mode 1:  T2[i, o, Q] = sum_g U1[o,g] T1[i, g, Q]                rows m = (i, Q) composite, n = o
    T2 = _empty(dev, n, s[0], s[1], r[2], r[3])
    ops.cgemm(T1, Us[1], T2, M=s[0] * Q, M2=Q, N=s[1], K=r[1], sAm=(r[1] * Q, 1), sAk=Q, sBk=1, sBn=r[1],
              sCm=(s[1] * Q, 1), sCn=Q)
    # mode 2:  T3[io, p, j] = sum_h U2[p,h] T2[io, h, j]               rows m = (io, j) composite, n = p
    T3 = _empty(dev, n, s[0], s[1], s[2], r[3])
    ops.cgemm(T2, Us[2], T3, M=s[0] * s[1] * r[3], M2=r[3], N=s[2], K=r[2], sAm=(r[2] * r[3], 1), sAk=r[3],
              sBk=1, sBn=r[2], sCm=(s[2] * r[3], 1), sCn=r[3])
    # mode 3:  W[iop, q] = sum_j T3[iop, j] U3[q,j]
    W = _empty(dev, n, s[0], s[1], s[2], s[3])
    ops.cgemm(T3, Us[3], W, M=s[0] * s[1] * s[2], N=s[3], K=r[3], sAm=r[3], sAk=1, sBk=1, sBn=r[3], sCm=s[3], sCn=1)
    return [cores, T1, T2, T3, W]


class TuckerReconstructManyFn(torch.autograd.Function):
    """(n, core_0, U0_0, U1_0, U2_0, U3_0, core_1, ...) -> (W_0, ..., W_{n-1}); all layers share one geometry.
    Real views of complex64: core [r0,r1,r2,r3,2], Uk [sk,rk,2], W [s0,s1,s2,s3,2]."""

    @staticmethod
    @_lib.on_tensor_device
    def forward(ctx, n, *params):
        assert len(params) == 5 * n and 1 <= n <= MAX_GROUP
        if not params[0].is_cuda:
            raise _lib.SpectralB200Error("TuckerReconstructFn: CUDA tensors only (there is no CPU path)")
        cores = [params[5 * l].contiguous().float() for l in range(n)]
        Us = [[params[5 * l + 1 + k].contiguous().float() for l in range(n)] for k in range(4)]
        s = [Us[k][0].shape[0] for k in range(4)]
        r = [Us[k][0].shape[1] for k in range(4)]
        for l in range(n):
            assert tuple(cores[l].shape[:4]) == tuple(r) and all(tuple(Us[k][l].shape[:2]) == (s[k], r[k]) for k in range(4))
        Ts = _mode_products(cores, Us, s, r)
        ctx.n, ctx.s, ctx.r = n, s, r
        ctx.sinks = [ops.grad_sink(t) for t in params]      # ddp.GradSync(direct=True): gradients written in place
        ctx.save_for_backward(*Ts[0], *Ts[1], *Ts[2], *Ts[3], *Us[0], *Us[1], *Us[2], *Us[3])
        return tuple(Ts[4])

    @staticmethod
    @_lib.on_tensor_device
    def backward(ctx, *gWs):
        n, s, r = ctx.n, ctx.s, ctx.r
        sv = ctx.saved_tensors
        T0, T1, T2, T3, U0, U1, U2, U3 = (list(sv[i * n:(i + 1) * n]) for i in range(8))
        dev = T0[0].device
        sinks = ctx.sinks

        def dest(k, shape):
            """per-layer outputs of parameter slot k (0 = core, 1..4 = factors): the sink when there is one"""
            return [ops._out(sinks[5 * l + k], (*shape, 2), dev) for l in range(n)]

        g4 = [(g if g is not None else torch.zeros(*s, 2, device=dev)).contiguous().float() for g in gWs]
        R, Q, IO = r[1] * r[2] * r[3], r[2] * r[3], s[0] * s[1]
        # ---- mode 3 ----
        gU3 = dest(4, (s[3], r[3]))
        ops.cgemm(g4, T3, gU3, M=s[3], N=r[3], K=IO * s[2], sAm=1, sAk=s[3], sBk=r[3], sBn=1, sCm=r[3], sCn=1,
                  conjB=True)
        g3 = _empty(dev, n, s[0], s[1], s[2], r[3])
        ops.cgemm(g4, U3, g3, M=IO * s[2], N=r[3], K=s[3], sAm=s[3], sAk=1, sBk=r[3], sBn=1, sCm=r[3], sCn=1,
                  conjB=True)
        # ---- mode 2 ----  reduction index (io, j) composite
        gU2 = dest(3, (s[2], r[2]))
        ops.cgemm(g3, T2, gU2, M=s[2], N=r[2], K=IO * r[3], K2=r[3], sAm=r[3], sAk=(s[2] * r[3], 1),
                  sBk=(r[2] * r[3], 1), sBn=r[3], sCm=r[2], sCn=1, conjB=True)
        g2 = _empty(dev, n, s[0], s[1], r[2], r[3])
        ops.cgemm(g3, U2, g2, M=IO * r[3], M2=r[3], N=r[2], K=s[2], sAm=(s[2] * r[3], 1), sAk=r[3], sBk=r[2], sBn=1,
                  sCm=(r[2] * r[3], 1), sCn=r[3], conjB=True)
        # ---- mode 1 ----  reduction index (i, Q) composite
        gU1 = dest(2, (s[1], r[1]))
        ops.cgemm(g2, T1, gU1, M=s[1], N=r[1], K=s[0] * Q, K2=Q, sAm=Q, sAk=(s[1] * Q, 1), sBk=(r[1] * Q, 1), sBn=Q,
                  sCm=r[1], sCn=1, conjB=True)
        g1 = _empty(dev, n, s[0], r[1], r[2], r[3])
        ops.cgemm(g2, U1, g1, M=s[0] * Q, M2=Q, N=r[1], K=s[1], sAm=(s[1] * Q, 1), sAk=Q, sBk=r[1], sBn=1,
                  sCm=(r[1] * Q, 1), sCn=Q, conjB=True)
        # ---- mode 0 ----
        gU0 = dest(1, (s[0], r[0]))
        ops.cgemm(g1, T0, gU0, M=s[0], N=r[0], K=R, sAm=R, sAk=1, sBk=1, sBn=R, sCm=r[0], sCn=1, conjB=True)
        g0 = dest(0, (r[0], r[1], r[2], r[3]))
        ops.cgemm(g1, U0, g0, M=R, N=r[0], K=s[0], sAm=1, sAk=R, sBk=r[0], sBn=1, sCm=1, sCn=R, conjB=True)
        out = [None]
        for l in range(n):
            out += [t if sinks[5 * l + k] is None else None
                    for k, t in enumerate((g0[l], gU0[l], gU1[l], gU2[l], gU3[l]))]
        return tuple(out)


class TuckerReconstructFn:
    """Single-weight convenience wrapper: ``TuckerReconstructFn.apply(core, U0, U1, U2, U3) -> W``."""

    @staticmethod
    def apply(core, U0, U1, U2, U3):
        return TuckerReconstructManyFn.apply(1, core, U0, U1, U2, U3)[0]


def reconstruct_many(weights):
    """Dense real views of several ``TuckerWeight`` modules; layers of equal geometry share launches."""
    out = [None] * len(weights)
    groups = {}
    for idx, w in enumerate(weights):
        groups.setdefault((tuple(w.shape), tuple(w.rank)), []).append(idx)
    for idxs in groups.values():
        for c0 in range(0, len(idxs), MAX_GROUP):
            chunk = idxs[c0:c0 + MAX_GROUP]
            params = []
            for i in chunk:
                params += [weights[i].core, *list(weights[i].factors)]
            Ws = TuckerReconstructManyFn.apply(len(chunk), *params)
            for i, W in zip(chunk, Ws):
                out[i] = W
    return out
