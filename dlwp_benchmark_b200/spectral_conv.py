"""``SpectralConv`` -- drop-in for ``neuralop.layers.spectral_convolution.SpectralConv``
(neuraloperator @05c01c3, the commit pinned by the reference's README.md:34-35) -- and the
``torch.autograd.Function``s that route its forward/backward through the C-ABI kernels.

Reference call chain replaced: ``FNO.forward`` -> ``FNOBlocks.forward`` ->
``SpectralConv.forward`` (rfftn -> fftshift -> slice -> einsum -> scatter -> fftshift ->
irfftn -> + bias) plus, in the fused functions, the 1x1 ``fno_skips`` conv, the add and the GELU.

Constructor arguments, parameter names / shapes (``weight.{l}.tensor`` real view
``[Cin,Cout,My,Mx,2]`` or Tucker ``weight.{l}.core`` + ``weight.{l}.factors.factor_{k}``;
``bias [L,Cout,1,1]``) and ``forward(x, indices=0, output_shape=None)`` follow neuralop.
"""
from __future__ import annotations

import math
from typing import List, Optional, Sequence

import torch
import torch.nn as nn

from . import _lib, ops
from .plan import fno_plan
from .tucker import TuckerWeight, DenseWeight


def _check_input(x: torch.Tensor):
    if not x.is_cuda:
        raise _lib.SpectralB200Error(
            "SpectralConv (B200) got a CPU tensor: this implementation has no CPU / torch.fft path")
    if x.dim() != 4:
        raise _lib.SpectralB200Error(f"only 2-D spectral convolutions are implemented (got {x.dim() - 2}-D input)")
    if x.shape[-1] % 4 != 0:
        raise _lib.SpectralB200Error(
            f"grid width W={x.shape[-1]} is not a multiple of 4: the row-synthesis kernels store 16-byte vectors "
            "(every grid the reference uses -- 64, 128, 256, 32x64 -- qualifies)")


# --------------------------------------------------------------------------------------
# autograd Functions
# --------------------------------------------------------------------------------------
class FNOBlockFn(torch.autograd.Function):
    """y = act( SpectralConv(x; W, bias) + Conv1x1(x; w_skip) )   -- one FNO block.

    ``w_skip`` / ``bias`` may be None (pure SpectralConv).  ``W`` is the dense complex weight as
    a real view [Cin,Cout,My,Mx,2] already sliced to the plan's retained block.
    """

    @staticmethod
    @_lib.on_tensor_device
    def forward(ctx, x, W, w_skip, bias, n_modes_halved, apply_act: bool, grad_mode: bool = True):
        _check_input(x)
        x = x.contiguous().float()
        W = W.contiguous().float()
        B, Cin, H, Wd = x.shape
        Cout = W.shape[1]
        plan = fno_plan(x.device, H, Wd, n_modes_halved)
        assert tuple(W.shape[2:]) == (plan.My, plan.Mx, 2), (W.shape, plan.My, plan.Mx)
        Xh = ops.analysis(plan, 0, x)
        Yh = ops.mix_fwd(Xh, W)
        Phi = ops.coldft_inv(plan, 0, Yh)
        ws = w_skip.reshape(Cout, Cin).contiguous().float() if w_skip is not None else None
        bv = bias.reshape(Cout).contiguous().float() if bias is not None else None
        need_z = apply_act and bool(grad_mode) and any(ctx.needs_input_grad)
        y, z = ops.rowidft_pointwise(plan, 0, Phi, x if ws is not None else None, ws, Cin, 1, bv, None,
                                     B, Cin, Cout, 0, apply_act, want_z=need_z)
        ctx.plan = plan
        ctx.apply_act = apply_act
        ctx.has_skip = ws is not None
        ctx.has_bias = bv is not None
        ctx.wskip_shape = w_skip.shape if w_skip is not None else None
        ctx.bias_shape = bias.shape if bias is not None else None
        ctx.save_for_backward(x, W, ws, Xh, z)
        return y

    @staticmethod
    @_lib.on_tensor_device
    def backward(ctx, gy):
        x, W, ws, Xh, z = ctx.saved_tensors
        plan = ctx.plan
        B, Cin = x.shape[:2]
        Cout = W.shape[1]
        gy = gy.contiguous().float()
        gz = ops.gelu_bwd(gy, z) if ctx.apply_act else gy
        gYh = ops.analysis(plan, 1, gz)
        gW = ops.mix_bwd_weight(Xh, gYh) if ctx.needs_input_grad[1] else None
        gx = None
        if ctx.needs_input_grad[0]:
            gXh = ops.mix_bwd_input(gYh, W)
            gPhi = ops.coldft_inv(plan, 1, gXh)
            gx, _ = ops.rowidft_pointwise(plan, 1, gPhi, gz if ctx.has_skip else None, ws, 1, Cin, None, None,
                                          B, Cout, Cin, 1, False)
        gws = gb = None
        if ctx.has_skip and (ctx.needs_input_grad[2] or (ctx.has_bias and ctx.needs_input_grad[3])):
            gws, gb = ops.pointwise_wgrad(gz, x, want_bias=ctx.has_bias)
            gws = gws.reshape(ctx.wskip_shape)
        elif ctx.has_bias and ctx.needs_input_grad[3]:
            gb = ops.channel_sum(gz)
        if gb is not None:
            gb = gb.reshape(ctx.bias_shape)
        return gx, gW, gws, gb, None, None, None


class FNOStackFn(torch.autograd.Function):
    """All L FNO blocks in one autograd node (the product path of ``FNO.forward``).

    Forward per layer: analysis -> mode mix -> column synthesis -> fused [row synthesis + skip
    conv + bias + GELU] writing y_l and the pre-activation z_l.  Backward per layer (reverse):
    the layer above's fused epilogue already multiplied by GELU'(z_l), so each layer costs one
    weighted analysis of gz, two mode contractions, a column synthesis, one fused
    [adjoint row synthesis + skip dgrad + GELU'(z_{l-1})] and one skip wgrad.

    args: x, n_modes_halved, L, grad_mode (the caller's ``torch.is_grad_enabled()``), then per layer
    (W_l, wskip_l, bias_l)
    """

    @staticmethod
    @_lib.on_tensor_device
    def forward(ctx, x, n_modes_halved, L, grad_mode, *params):
        _check_input(x)
        x = x.contiguous().float()
        B, C, H, Wd = x.shape
        plan = fno_plan(x.device, H, Wd, n_modes_halved)
        Ws = [params[3 * l].contiguous().float() for l in range(L)]
        wss = [params[3 * l + 1].reshape(C, C).contiguous().float() for l in range(L)]
        bvs = [params[3 * l + 2].reshape(C).contiguous().float() for l in range(L)]
        need_grad = bool(grad_mode) and any(ctx.needs_input_grad)
        hs, Xhs, zs = [], [], []
        h = x
        for l in range(L):
            act = l < L - 1
            Xh = ops.analysis(plan, 0, h)
            Yh = ops.mix_fwd(Xh, Ws[l])
            Phi = ops.coldft_inv(plan, 0, Yh)
            y, z = ops.rowidft_pointwise(plan, 0, Phi, h, wss[l], C, 1, bvs[l], None, B, C, C, 0, act,
                                         want_z=act and need_grad)
            hs.append(h); Xhs.append(Xh); zs.append(z)
            h = y
        ctx.plan, ctx.L, ctx.C = plan, L, C
        ctx.shapes = [(params[3 * l + 1].shape, params[3 * l + 2].shape) for l in range(L)]
        ctx.save_for_backward(*hs, *Xhs, *[z for z in zs if z is not None], *Ws, *wss)
        return h

    @staticmethod
    @_lib.on_tensor_device
    def backward(ctx, gy):
        L, C, plan = ctx.L, ctx.C, ctx.plan
        sv = list(ctx.saved_tensors)
        hs, Xhs = sv[:L], sv[L:2 * L]
        zs = sv[2 * L:3 * L - 1] + [None]
        Ws = sv[3 * L - 1:4 * L - 1]
        wss = sv[4 * L - 1:5 * L - 1]
        B = hs[0].shape[0]
        gz = gy.contiguous().float()
        grads: List[Optional[torch.Tensor]] = [None] * (3 * L)
        gx = None
        for l in range(L - 1, -1, -1):
            gYh = ops.analysis(plan, 1, gz)
            grads[3 * l] = ops.mix_bwd_weight(Xhs[l], gYh)
            gws, gb = ops.pointwise_wgrad(gz, hs[l], want_bias=True)
            grads[3 * l + 1] = gws.reshape(ctx.shapes[l][0])
            grads[3 * l + 2] = gb.reshape(ctx.shapes[l][1])
            if l > 0 or ctx.needs_input_grad[0]:
                gXh = ops.mix_bwd_input(gYh, Ws[l])
                gPhi = ops.coldft_inv(plan, 1, gXh)
                zprev = zs[l - 1] if l > 0 else None
                gprev, _ = ops.rowidft_pointwise(plan, 1, gPhi, gz, wss[l], 1, C, None, zprev, B, C, C, 1, False)
                if l > 0:
                    gz = gprev
                else:
                    gx = gprev
        return (gx, None, None, None, *grads)


# --------------------------------------------------------------------------------------
# the module
# --------------------------------------------------------------------------------------
class SpectralConv(nn.Module):
    """neuralop-compatible 2-D spectral convolution holding ``n_layers`` weights.

    Only the options the reference ever exercises are implemented: 2-D,
    ``separable=False``, ``fno_block_precision='full'``, ``output_scaling_factor=None``,
    ``fft_norm='forward'`` (what ``FNO`` passes), ``factorization in {None,'dense','tucker'}``,
    per-layer (non-joint) factorisation.  Anything else raises at construction.
    """

    def __init__(self, in_channels, out_channels, n_modes, max_n_modes=None, bias=True, n_layers=1, separable=False,
                 output_scaling_factor=None, fno_block_precision="full", rank=0.5, factorization=None,
                 implementation="reconstructed", fixed_rank_modes=False, joint_factorization=False,
                 decomposition_kwargs=None, init_std="auto", fft_norm="backward"):
        super().__init__()
        if isinstance(n_modes, int):
            n_modes = [n_modes]
        if len(n_modes) != 2:
            raise NotImplementedError("B200 SpectralConv: only 2-D (n_modes of length 2) is implemented")
        if separable or joint_factorization or output_scaling_factor is not None:
            raise NotImplementedError("separable / joint_factorization / output_scaling_factor are not implemented")
        if fno_block_precision != "full":
            raise NotImplementedError("fno_block_precision must be 'full'")
        if fft_norm != "forward":
            raise NotImplementedError("only fft_norm='forward' (what neuralop's FNO passes) is implemented")
        self.in_channels, self.out_channels = in_channels, out_channels
        self.order = 2
        self.n_modes = n_modes                      # property: halves the last dim
        if max_n_modes is None:
            max_n_modes = self.n_modes
        elif isinstance(max_n_modes, int):
            max_n_modes = [max_n_modes]
        self.max_n_modes = list(max_n_modes)
        self.n_layers = n_layers
        self.rank = rank
        self.implementation = implementation
        self.fft_norm = fft_norm
        if init_std == "auto":
            init_std = (2 / (in_channels + out_channels)) ** 0.5
        fac = "dense" if factorization is None else str(factorization).lower()
        if fac.startswith("complex"):
            fac = fac[len("complex"):]
        if fac not in ("dense", "tucker"):
            raise NotImplementedError(f"factorization={factorization!r} is not implemented (dense / tucker only)")
        self.factorization = fac
        shape = (in_channels, out_channels, *self.max_n_modes)
        if fac == "dense":
            self.weight = nn.ModuleList([DenseWeight(shape, init_std) for _ in range(n_layers)])
        else:
            self.weight = nn.ModuleList([TuckerWeight(shape, rank, init_std) for _ in range(n_layers)])
        if bias:
            self.bias = nn.Parameter(init_std * torch.randn(n_layers, out_channels, 1, 1))
        else:
            self.bias = None

    # neuralop: the last (rfft) dim keeps n//2+1 modes
    @property
    def n_modes(self):
        return self._n_modes

    @n_modes.setter
    def n_modes(self, n_modes):
        n_modes = [n_modes] if isinstance(n_modes, int) else list(n_modes)
        n_modes[-1] = n_modes[-1] // 2 + 1
        self._n_modes = n_modes

    def dense_weights_all(self, H: int, W: int):
        """``dense_weight`` of every layer; Tucker layers are reconstructed together (one launch per mode
        product for all layers, forward and backward)."""
        if self.factorization == "tucker" and self.weight[0].core.is_cuda and len(self.weight[0].shape) == 4:
            from .tucker_fn import reconstruct_many
            dense = reconstruct_many(list(self.weight))
            return [self._slice_modes(w, H, W) for w in dense]
        return [self.dense_weight(l, H, W) for l in range(self.n_layers)]

    def dense_weight(self, index: int, H: int, W: int) -> torch.Tensor:
        """Real view [Cin,Cout,My,Mx,2] of layer ``index`` sliced like SpectralConv.forward does."""
        return self._slice_modes(self.weight[index].to_dense_real(), H, W)   # [Cin,Cout,*max_n_modes,2]

    def _slice_modes(self, w: torch.Tensor, H: int, W: int) -> torch.Tensor:
        fft_size = [H, W // 2 + 1]
        kept = [min(s, n) for s, n in zip(fft_size, self.n_modes)]
        starts = [m - k for m, k in zip(self.max_n_modes, kept)]
        sl0 = slice(starts[0] // 2, -starts[0] // 2) if starts[0] else slice(None)
        sl1 = slice(None, -starts[1]) if starts[1] else slice(None)
        if starts[0] or starts[1]:
            w = w[:, :, sl0, sl1]
        return w

    def transform(self, x, layer_index=0, output_shape=None):
        if output_shape is not None and tuple(output_shape) != tuple(x.shape[2:]):
            raise NotImplementedError("output_shape resampling is not implemented")
        return x

    def forward(self, x, indices=0, output_shape=None):
        if output_shape is not None and tuple(output_shape) != tuple(x.shape[2:]):
            raise NotImplementedError("output_shape resampling is not implemented")
        _check_input(x)
        H, W = x.shape[-2:]
        w = self.dense_weight(indices, H, W)
        b = self.bias[indices] if self.bias is not None else None
        return FNOBlockFn.apply(x, w, None, b, tuple(self.n_modes), False, torch.is_grad_enabled())

    def get_conv(self, indices):
        if self.n_layers == 1:
            raise ValueError("A single convolution is parametrized, directly use the main class.")
        return SubConv(self, indices)

    def __getitem__(self, indices):
        return self.get_conv(indices)


class SubConv(nn.Module):
    """View on one layer of a multi-layer SpectralConv (neuralop API parity)."""

    def __init__(self, main_conv, indices):
        super().__init__()
        self.main_conv = main_conv
        self.indices = indices

    def forward(self, x, **kwargs):
        return self.main_conv.forward(x, self.indices, **kwargs)
