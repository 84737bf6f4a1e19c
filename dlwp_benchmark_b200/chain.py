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
    acts (tuple of bool per layer), grad_mode (``torch.is_grad_enabled()`` of the caller: inside ``forward`` autograd
    has already switched it off, and ``needs_input_grad`` stays True under ``no_grad``), then per layer
    (Wspec_l or None, Wp_l, b_l or None)."""

    @staticmethod
    @_lib.on_tensor_device
    def forward(ctx, x, n_modes_halved, spec, acts, grad_mode, *params):
        if not x.is_cuda:
            raise _lib.SpectralB200Error("FNO(B200) got a CPU tensor: there is no CPU / torch.fft path")
        x = x.contiguous().float()
        B, _, H, Wd = x.shape
        nl = len(spec)
        plan = fno_plan(x.device, H, Wd, n_modes_halved)
        need_grad = bool(grad_mode) and any(ctx.needs_input_grad)
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
            if not need_grad:
                hs.clear(); Xhs.clear(); zs.clear()         # inference: nothing is kept alive past its consumer
        if not need_grad:
            return h
        ctx.plan, ctx.nl, ctx.spec, ctx.acts = plan, nl, tuple(spec), tuple(acts)
        # direct gradient destinations (ddp.GradSync(direct=True)): written by the backward kernels themselves
        ctx.sinks = [ops.grad_sink(t) for t in params]
        ctx.shapes = [(params[3 * l + 1].shape, params[3 * l + 2].shape if params[3 * l + 2] is not None else None)
                      for l in range(nl)]
        ctx.has_h = [t is not None for t in hs]
        saved = [t for t in hs if t is not None] + [t for t in Xhs if t is not None] + [t for t in zs if t is not None] \
            + [t for t in Ws if t is not None] + Ps + ([bs[nl - 2]] if head else []) + ([bs[0]] if lift_tail else [])
        ctx.has_z = [t is not None for t in zs]
        ctx.save_for_backward(*saved)
        return h

    @staticmethod
    @_lib.on_tensor_device
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
        sinks = ctx.sinks

        def put(idx, t, shape):
            """hand gradient ``t`` of parameter slot ``idx`` to autograd -- unless the kernel wrote it into its sink"""
            if t is not None and sinks[idx] is None:
                grads[idx] = t.reshape(shape)

        gx = None
        top = nl - 1
        bias_known = False
        if ctx.head:
            # stage 1 of the fused head backward: z1 recomputed on the tensor cores; gz1 = w2 gy gelu'(z1) and the
            # reductions for b1, w2, b2 come out of one kernel
            l1, l2 = nl - 2, nl - 1
            gz, gb1, gw2, gb2 = ops.mlp_head_bwd(hs[l1], Ps[l1], head_b1, Ps[l2].reshape(-1), gz,
                                                 want_gb2=ctx.shapes[l2][1] is not None, out_gb1=sinks[3 * l1 + 2],
                                                 out_gw2=sinks[3 * l2 + 1], out_gb2=sinks[3 * l2 + 2])
            put(3 * l2 + 1, gw2, ctx.shapes[l2][0])
            put(3 * l2 + 2, gb2, ctx.shapes[l2][1])
            put(3 * l1 + 2, gb1, ctx.shapes[l1][1])
            bias_known = True
            top = l1
        for l in range(top, -1, -1):
            N, M = Ps[l].shape
            has_b = ctx.shapes[l][1] is not None and not (bias_known and l == top)
            sw, sb = sinks[3 * l + 1], (sinks[3 * l + 2] if has_b else None)
            # ---- weight gradients ----
            if ctx.lift_gen and l == 1:
                gP, gb = ops.lift_wgrad(gz, hs[0], Ps[0].reshape(-1), lift_b1, want_bias=has_b, out_w=sw, out_b=sb)
            elif N <= SMALL_W and N <= M:
                gP, gb, _ = ops.wgrad_small(gz, hs[l], False, has_b, False, out_dot=sw, out_small=sb)
            elif M <= SMALL_W:
                gP, _, gb = ops.wgrad_small(hs[l], gz, True, False, has_b, out_dot=sw, out_big=sb)
            else:
                gP, gb = ops.pointwise_wgrad(gz, hs[l], want_bias=has_b, out_w=sw, out_b=sb)
            put(3 * l + 1, gP, ctx.shapes[l][0])
            if has_b:
                put(3 * l + 2, gb, ctx.shapes[l][1])
            gYh = None
            if spec[l]:
                gYh = ops.analysis(plan, 1, gz)
                gW = ops.mix_bwd_weight(Xhs[l], gYh, out=sinks[3 * l])
                put(3 * l, gW, gW.shape)
            if ctx.lift_tail and l == 1:
                # gz_0 = (P_1^T gz) gelu'(w1 x + b1) stays on chip; only its two pixel reductions leave
                gw1, gb1 = ops.lift_tail_bwd(gz, Ps[1], Ps[0].reshape(-1), lift_b1, hs[0], out_gw1=sinks[1],
                                             out_gb1=sinks[2])
                put(1, gw1, ctx.shapes[0][0])
                put(2, gb1, ctx.shapes[0][1])
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
        return (gx, None, None, None, None, *grads)
