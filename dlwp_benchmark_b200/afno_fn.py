"""autograd.Function for AFNO2D: forward / backward through the channels-last C-ABI kernels.

forward  (reference fourcastnet.py:77-126):
    T = rowdft(x); Xh = coldft(T); O1 = relu(Xh*W1+b1); Yh = softshrink(O1*W2+b2);
    Phi = coldft_inv(Yh); y = rowidft(Phi) + x
backward: the adjoint transforms (plan tables of pass 1) around the two block-linear backward steps;
ReLU / softshrink masks are recovered from the saved layer outputs (O1 > 0, Yh != 0).
"""
from __future__ import annotations

import ctypes

import torch

from . import _lib
from .ops import _p, _req, _stream
from .plan import afno_plan


def _bl_fwd(inp, w, b, ntok, nb, Ni, No, act, lam):
    out = torch.empty(*inp.shape[:-2], nb * No, 2, device=inp.device, dtype=torch.float32)
    _lib.check(_lib.load().sb200_afno_blocklinear_fwd(_p(inp), _p(w), _p(b), _p(out), ntok, nb, Ni, No, act,
                                                       ctypes.c_float(lam), _stream()), "afno_blocklinear_fwd")
    return out


def _bl_dgrad(gout, fwd_out, kind, w, ntok, nb, Ni, No):
    gin = torch.empty(*gout.shape[:-2], nb * Ni, 2, device=gout.device, dtype=torch.float32)
    _lib.check(_lib.load().sb200_afno_blocklinear_dgrad(_p(gout), _p(fwd_out), kind, _p(w), _p(gin), ntok, nb, Ni, No,
                                                         _stream()), "afno_blocklinear_dgrad")
    return gin


def _bl_wgrad(a, gout, fwd_out, kind, ntok, nb, Ni, No):
    lib = _lib.load()
    ws = torch.empty(lib.sb200_afno_blocklinear_wgrad_workspace(ntok, nb, Ni, No), device=a.device, dtype=torch.float32)
    gw = torch.empty(2, nb, Ni, No, device=a.device, dtype=torch.float32)
    gb = torch.empty(2, nb, No, device=a.device, dtype=torch.float32)
    _lib.check(lib.sb200_afno_blocklinear_wgrad(_p(a), _p(gout), _p(fwd_out), kind, _p(gw), _p(gb), ntok, nb, Ni, No,
                                                _p(ws), _stream()), "afno_blocklinear_wgrad")
    return gw, gb


def afno_forward(x, w1c, b1c, w2c, b2c, nb: int, lam: float, frac: float, resid, resid2=None):
    """The AFNO2D filter on x [B,h,w,C] (fp32, contiguous): returns (y, saved) with
    y = synthesis(block_mlp(analysis(x))) + resid + resid2 (either may be None); ``saved`` feeds ``afno_backward``."""
    B, h, w, C = x.shape
    bs = C // nb
    bsf = w1c.shape[-1]
    plan = afno_plan(x.device, h, w, frac)
    lib = _lib.load()
    My, Mx = plan.My, plan.Mx
    ntok = B * My * Mx
    T = torch.empty(B, h, Mx, C, 2, device=x.device, dtype=torch.float32)
    _lib.check(lib.sb200_cl_rowdft_fwd(plan.handle, 0, _p(x), _p(T), B * h, C, _stream()), "cl_rowdft_fwd")
    Xh = torch.empty(B, My, Mx, C, 2, device=x.device, dtype=torch.float32)
    _lib.check(lib.sb200_cl_coldft_fwd(plan.handle, 0, _p(T), _p(Xh), B, C, _stream()), "cl_coldft_fwd")
    O1 = _bl_fwd(Xh, w1c, b1c, ntok, nb, bs, bsf, 1, 0.0)
    Yh = _bl_fwd(O1, w2c, b2c, ntok, nb, bsf, bs, 2, lam)
    Phi = torch.empty(B, h, Mx, C, 2, device=x.device, dtype=torch.float32)
    _lib.check(lib.sb200_cl_coldft_inv(plan.handle, 0, _p(Yh), _p(Phi), B, C, _stream()), "cl_coldft_inv")
    y = torch.empty_like(x)
    _lib.check(lib.sb200_cl_rowidft_res2(plan.handle, 0, _p(Phi), _p(resid), _p(resid2), _p(y), B * h, C, _stream()),
               "cl_rowidft_res")
    return y, (plan, (B, h, w, C, nb, bs, bsf, ntok), Xh, O1, Yh)


def afno_backward(saved, w1c, w2c, gy, resid_grad, need_gx: bool = True):
    """Backward of ``afno_forward``: gy [B,h,w,C] -> (gx or None, gw1, gb1, gw2, gb2); ``resid_grad`` (gy when the filter
    input was also its residual, else None) is added to gx inside the adjoint row synthesis."""
    plan, (B, h, w, C, nb, bs, bsf, ntok), Xh, O1, Yh = saved
    lib = _lib.load()
    My, Mx = plan.My, plan.Mx
    dev = gy.device
    T = torch.empty(B, h, Mx, C, 2, device=dev, dtype=torch.float32)
    _lib.check(lib.sb200_cl_rowdft_fwd(plan.handle, 1, _p(gy), _p(T), B * h, C, _stream()), "cl_rowdft_fwd")
    gYh = torch.empty(B, My, Mx, C, 2, device=dev, dtype=torch.float32)
    _lib.check(lib.sb200_cl_coldft_fwd(plan.handle, 1, _p(T), _p(gYh), B, C, _stream()), "cl_coldft_fwd")
    gw2, gb2 = _bl_wgrad(O1, gYh, Yh, 2, ntok, nb, bsf, bs)
    gO1 = _bl_dgrad(gYh, Yh, 2, w2c, ntok, nb, bsf, bs)
    gw1, gb1 = _bl_wgrad(Xh, gO1, O1, 1, ntok, nb, bs, bsf)
    gx = None
    if need_gx:
        gXh = _bl_dgrad(gO1, O1, 1, w1c, ntok, nb, bs, bsf)
        Phi = torch.empty(B, h, Mx, C, 2, device=dev, dtype=torch.float32)
        _lib.check(lib.sb200_cl_coldft_inv(plan.handle, 1, _p(gXh), _p(Phi), B, C, _stream()), "cl_coldft_inv")
        gx = torch.empty_like(gy)
        _lib.check(lib.sb200_cl_rowidft_res(plan.handle, 1, _p(Phi), _p(resid_grad), _p(gx), B * h, C, _stream()),
                   "cl_rowidft_res")
    return gx, gw1, gb1, gw2, gb2


class AFNO2DFn(torch.autograd.Function):
    @staticmethod
    @_lib.on_tensor_device
    def forward(ctx, x, w1, b1, w2, b2, num_blocks, lam, frac, add_residual=True):
        _req(x.contiguous(), "x")
        x = x.contiguous()
        w1c, b1c, w2c, b2c = (t.contiguous().float() for t in (w1, b1, w2, b2))
        y, saved = afno_forward(x, w1c, b1c, w2c, b2c, int(num_blocks), lam, frac, x if add_residual else None)
        ctx.add_residual = bool(add_residual)
        ctx.saved_misc = saved[:2]
        ctx.save_for_backward(saved[2], saved[3], saved[4], w1c, w2c)
        return y

    @staticmethod
    @_lib.on_tensor_device
    def backward(ctx, gy):
        Xh, O1, Yh, w1c, w2c = ctx.saved_tensors
        gy = gy.contiguous().float()
        saved = (*ctx.saved_misc, Xh, O1, Yh)
        gx, gw1, gb1, gw2, gb2 = afno_backward(saved, w1c, w2c, gy, gy if ctx.add_residual else None,
                                               need_gx=ctx.needs_input_grad[0])
        return gx, gw1, gb1, gw2, gb2, None, None, None, None
