"""``FusedChainFn``: the whole FNO (lifting MLP -> L FNO blocks -> projection MLP) as ONE autograd
node whose forward and backward are sequences of C-ABI kernel launches.

Every layer of an FNO has the same form

    z_l = [ SpectralConv(h_l ; W_l) ] + P_l h_l + b_l ,      h_{l+1} = gelu(z_l)  or  z_l

(the bracket only for the L FNO blocks; P_l is a 1x1 convolution), so one fused kernel
(``sb200_rowidft_pointwise``: row synthesis + channel mix + bias + GELU) serves all of them:

  forward   per layer: [analysis -> mode mix -> column synthesis] -> fused kernel writing h_{l+1}
            (and z_l when the layer has a GELU and gradients are needed)
  backward  per layer, in reverse: the producer of the incoming gradient already multiplied it by
            GELU'(z_l); weight grads (``pointwise_wgrad`` / ``wgrad_small`` + spectral ``modes_gemm``);
            data grad through the same fused kernel with transposed weights, adjoint tables and the
            GELU'(z_{l-1}) epilogue.

Reference call chain replaced: neuralop ``FNO.forward`` (lifting MLP, FNOBlocks x L, projection MLP)
as constructed at src/nsbench/models/fno/fno.py:19-27 / src/dlwpbench/models/fno/fno.py:38-47.
"""
from __future__ import annotations

from typing import List, Optional, Sequence

import torch

from . import _lib, ops
from .plan import fno_plan

SMALL = 8      # channel counts handled by the small-N forward kernel
SMALL_W = 16   # channel counts handled by the small weight-gradient kernel


class FusedChainFn(torch.autograd.Function):
    """args: x, n_modes_halved, spec (tuple of bool per layer: has a spectral branch),
    acts (tuple of bool per layer), then per layer (Wspec_l or None, Wp_l, b_l or None)."""

    @staticmethod
    def forward(ctx, x, n_modes_halved, spec, acts, *params):
        if not x.is_cuda:
            raise _lib.SpectralB200Error("FNO(B200) got a CPU tensor: there is no CPU / torch.fft path")
        x = x.contiguous().float()
        B, _, H, Wd = x.shape
        nl = len(spec)
        plan = fno_plan(x.device, H, Wd, n_modes_halved)
        need_grad = any(ctx.needs_input_grad)
        Ws, Ps, bs = [], [], []
        for l in range(nl):
            w, p, b = params[3 * l:3 * l + 3]
            Ws.append(w.contiguous().float() if w is not None else None)
            Ps.append(p.reshape(p.shape[0], p.shape[1]).contiguous().float())
            bs.append(b.reshape(-1).contiguous().float() if b is not None else None)
        hs, Xhs, zs = [], [], []
        h = x
        # fused head: the last two layers are a pointwise MLP  C -> 256 (GELU) -> 1: the hidden tensor stays on chip
        head = (nl >= 2 and not spec[nl - 2] and not spec[nl - 1] and acts[nl - 2] and not acts[nl - 1]
                and bs[nl - 2] is not None
                and ops.mlp_head_supported(Ps[nl - 2].shape[1], Ps[nl - 2].shape[0], Ps[nl - 1].shape[0], H * Wd))
        ctx.head = head
        # lifting tail: a 1-input-channel lifting MLP  x -> gelu(w1 x + b1) [256] -> C: its pre-activation is
        # recomputed from x in the backward (sb200_lift_tail_bwd), so z_0 is never written
        lift_tail = (nl >= 2 and need_grad and not ctx.needs_input_grad[0] and not spec[0] and not spec[1] and acts[0]
                     and tuple(Ps[0].shape) == (256, 1) and bs[0] is not None and Ps[1].shape[1] == 256
                     and ops.mlp_head_supported(Ps[1].shape[0], 256, 1, H * Wd))
        ctx.lift_tail = lift_tail
        # ... and when the shapes allow, the whole lifting MLP runs with its hidden tensor generated on chip
        lift_gen = (nl >= 2 and not spec[0] and not spec[1] and acts[0] and not acts[1] and tuple(Ps[0].shape) == (256, 1)
                    and bs[0] is not None and Ps[1].shape[1] == 256 and ops.lift_supported(Ps[1].shape[0], 256, H * Wd)
                    and (lift_tail or not need_grad))
        ctx.lift_gen = lift_gen
        for l in range(nl):
            if lift_gen and l == 0:
                hs.append(h); Xhs.append(None); zs.append(None)          # x itself is all the backward needs
                continue
            if lift_gen and l == 1:
                y = ops.lift_fwd(x, Ps[0].reshape(-1), bs[0], Ps[1], bs[1])
                hs.append(None); Xhs.append(None); zs.append(None)
                h = y
                continue
            if head and l == nl - 2:
                y = ops.mlp_head_fwd(h, Ps[l], bs[l], Ps[l + 1].reshape(-1), bs[l + 1])
                hs += [h, None]; Xhs += [None, None]; zs += [None, None]
                h = y
                break
            N, M = Ps[l].shape
            want_z = acts[l] and need_grad and not (lift_tail and l == 0)
            Xh = None
            if spec[l]:
                Xh = ops.analysis(plan, 0, h)
                Phi = ops.coldft_inv(plan, 0, ops.mix_fwd(Xh, Ws[l]))
                y, z = ops.rowidft_pointwise(plan, 0, Phi, h, Ps[l], M, 1, bs[l], None, B, M, N, 0, acts[l],
                                             want_z=want_z)
            elif N <= SMALL:
                y, z = ops.pointwise_small_n(h, Ps[l], bs[l], acts[l], want_z=want_z)
            else:
                y, z = ops.rowidft_pointwise(plan, 0, None, h, Ps[l], M, 1, bs[l], None, B, M, N, 0, acts[l],
                                             want_z=want_z)
            hs.append(h); Xhs.append(Xh); zs.append(z)
            h = y
        ctx.plan, ctx.nl, ctx.spec, ctx.acts = plan, nl, tuple(spec), tuple(acts)
        ctx.shapes = [(params[3 * l + 1].shape, params[3 * l + 2].shape if params[3 * l + 2] is not None else None)
                      for l in range(nl)]
        ctx.has_h = [t is not None for t in hs]
        saved = [t for t in hs if t is not None] + [t for t in Xhs if t is not None] + [t for t in zs if t is not None] \
            + [t for t in Ws if t is not None] + Ps + ([bs[nl - 2]] if head else []) + ([bs[0]] if lift_tail else [])
        ctx.has_z = [t is not None for t in zs]
        ctx.save_for_backward(*saved)
        return h

    @staticmethod
    def backward(ctx, gy):
        nl, spec, acts, plan = ctx.nl, ctx.spec, ctx.acts, ctx.plan
        sv = list(ctx.saved_tensors)
        it = iter(sv)
        hs = [next(it) if ctx.has_h[l] else None for l in range(nl)]
        Xhs = [next(it) if spec[l] else None for l in range(nl)]
        zs = [next(it) if ctx.has_z[l] else None for l in range(nl)]
        Ws = [next(it) if spec[l] else None for l in range(nl)]
        Ps = [next(it) for _ in range(nl)]
        head_b1 = next(it) if ctx.head else None
        lift_b1 = next(it) if ctx.lift_tail else None
        B = hs[0].shape[0]
        gz = gy.contiguous().float()
        if acts[nl - 1]:
            gz = ops.gelu_bwd(gz, zs[nl - 1])
        grads: List[Optional[torch.Tensor]] = [None] * (3 * nl)
        gx = None
        top = nl - 1
        bias_known = None
        if ctx.head:
            # stage 1 of the fused head backward: z1 recomputed on the tensor cores; gz1 = w2 gy gelu'(z1) and the
            # reductions for b1, w2, b2 come out of one kernel
            l1 = nl - 2
            gz, gb1, gw2, gb2 = ops.mlp_head_bwd(hs[l1], Ps[l1], head_b1, Ps[nl - 1].reshape(-1), gz,
                                                 want_gb2=ctx.shapes[nl - 1][1] is not None)
            grads[3 * (nl - 1) + 1] = gw2.reshape(ctx.shapes[nl - 1][0])
            if gb2 is not None:
                grads[3 * (nl - 1) + 2] = gb2.reshape(ctx.shapes[nl - 1][1])
            bias_known = gb1
            top = l1
        for l in range(top, -1, -1):
            N, M = Ps[l].shape
            has_b = ctx.shapes[l][1] is not None
            if bias_known is not None and l == top:
                grads[3 * l + 2] = bias_known.reshape(ctx.shapes[l][1])
                has_b = False
            # ---- weight gradients ----
            if ctx.lift_gen and l == 1:
                gP, gb = ops.lift_wgrad(gz, hs[0], Ps[0].reshape(-1), lift_b1, want_bias=has_b)
            elif N <= SMALL_W and N <= M:
                gP, gb, _ = ops.wgrad_small(gz, hs[l], False, has_b, False)
            elif M <= SMALL_W:
                gP, _, gb = ops.wgrad_small(hs[l], gz, True, False, has_b)
            else:
                gP, gb = ops.pointwise_wgrad(gz, hs[l], want_bias=has_b)
            grads[3 * l + 1] = gP.reshape(ctx.shapes[l][0])
            if has_b:
                grads[3 * l + 2] = gb.reshape(ctx.shapes[l][1])
            gYh = None
            if spec[l]:
                gYh = ops.analysis(plan, 1, gz)
                grads[3 * l] = ops.mix_bwd_weight(Xhs[l], gYh)
            if ctx.lift_tail and l == 1:
                # gz_0 = (P_1^T gz) gelu'(w1 x + b1) stays on chip; only its two pixel reductions leave
                gw1, gb1 = ops.lift_tail_bwd(gz, Ps[1], Ps[0].reshape(-1), lift_b1, hs[0])
                grads[1] = gw1.reshape(ctx.shapes[0][0])
                grads[2] = gb1.reshape(ctx.shapes[0][1])
                break
            # ---- data gradient (fused with GELU' of the previous layer) ----
            if l > 0 or ctx.needs_input_grad[0]:
                gPhi = ops.coldft_inv(plan, 1, ops.mix_bwd_input(gYh, Ws[l])) if spec[l] else None
                zprev = zs[l - 1] if (l > 0 and acts[l - 1]) else None
                gprev, _ = ops.rowidft_pointwise(plan, 1, gPhi, gz, Ps[l], 1, M, None, zprev, B, N, M, 1, False)
                if l > 0:
                    gz = gprev
                else:
                    gx = gprev
        return (gx, None, None, None, *grads)
