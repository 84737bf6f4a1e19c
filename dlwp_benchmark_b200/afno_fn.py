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


class GemmDesc(ctypes.Structure):
    """Mirror of ``sb200_gemm_desc`` (include/spectral_b200.h)."""
    _fields_ = [(n, ctypes.c_int32) for n in ("M", "N", "K", "nbatch", "a_mn", "b_mn", "act", "split_k", "a_off", "b_off0",
                                              "b_off1", "reserved")] + \
               [(n, ctypes.c_int64) for n in ("lda", "ldb", "ldd", "a_ext0", "a_ext1", "b_ext0", "b_ext1", "d_off", "bias_off")] + \
               [("lam", ctypes.c_float), ("reserved2", ctypes.c_float)]


def _gemm_batched(A, B, D, *, M, N, K, nbatch, lda, ldb, ldd, a_ext, b_ext, a_mn=False, b_mn=False, a_off=0, b_off0=0, b_off1=0,
                  d_off=0, bias=None, bias_off=0, aux=None, act=0, lam=0.0, split_k=False):
    """``nbatch`` GEMMs of identical geometry in one launch (sb200_gemm_batched); see the header for the offsets."""
    d = GemmDesc()
    d.M, d.N, d.K, d.nbatch = M, N, K, nbatch
    d.a_mn, d.b_mn, d.act, d.split_k = int(a_mn), int(b_mn), act, int(split_k)
    d.a_off, d.b_off0, d.b_off1 = a_off, b_off0, b_off1
    d.lda, d.ldb, d.ldd = lda, ldb, ldd
    d.a_ext0, d.a_ext1 = a_ext
    d.b_ext0, d.b_ext1 = b_ext
    d.d_off, d.bias_off, d.lam = d_off, bias_off, lam
    lib = _lib.load()
    mode = _lib.tc_mode()
    nws = lib.sb200_gemm_batched_workspace(ctypes.byref(d), mode)
    ws = torch.empty(nws, device=A.device, dtype=torch.float32) if nws > 0 else None
    _lib.check(lib.sb200_gemm_batched(ctypes.byref(d), _p(A), _p(B), _p(D), _p(bias), _p(aux), _p(ws), _stream(), mode),
               "gemm_batched")
    return D


def _embed(w, nb, Ni, No):
    """w [2,nb,Ni,No] -> real embedding E [nb, 2*No, 2*Ni] of the complex block matrices"""
    E = torch.empty(nb, 2 * No, 2 * Ni, device=w.device, dtype=torch.float32)
    _lib.check(_lib.load().sb200_afno_embed(_p(w), _p(E), nb, Ni, No, _stream()), "afno_embed")
    return E


def _unembed(gE, nb, Ni, No):
    gw = torch.empty(2, nb, Ni, No, device=gE.device, dtype=torch.float32)
    _lib.check(_lib.load().sb200_afno_unembed(_p(gE), _p(gw), nb, Ni, No, _stream()), "afno_unembed")
    return gw


def _mlp_on_tensor_cores(bs, bsf):
    """The block MLP runs as batched real-embedded GEMMs on tcgen05 when the blocks tile (2*bs, 2*bsf multiples of 32)."""
    return _lib.tc_mode() != 0 and (2 * bs) % 32 == 0 and (2 * bsf) % 32 == 0


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
    tc = _mlp_on_tensor_cores(bs, bsf)
    if tc:
        # block-diagonal complex MLP (reference :95-121) as nb real-embedded GEMMs per layer in ONE tcgen05 launch each:
        # the token rows [ntok, 2C] are the K-major operand as they lie in memory, block b = columns [b*2bs, (b+1)*2bs)
        Cf = nb * bsf
        E1, E2 = _embed(w1c, nb, bs, bsf), _embed(w2c, nb, bsf, bs)
        be1 = b1c.permute(1, 2, 0).contiguous().view(-1)          # [nb, bsf, (re, im)]
        be2 = b2c.permute(1, 2, 0).contiguous().view(-1)
        O1 = torch.empty(B, My, Mx, Cf, 2, device=x.device, dtype=torch.float32)
        _gemm_batched(Xh, E1, O1, M=ntok, N=2 * bsf, K=2 * bs, nbatch=nb, lda=2 * C, ldb=2 * bs, ldd=2 * Cf,
                      a_ext=(2 * C, ntok), b_ext=(2 * bs, nb * 2 * bsf), a_off=2 * bs, b_off1=2 * bsf, d_off=2 * bsf,
                      bias=be1, bias_off=2 * bsf, act=3)
        Yh = torch.empty(B, My, Mx, C, 2, device=x.device, dtype=torch.float32)
        _gemm_batched(O1, E2, Yh, M=ntok, N=2 * bs, K=2 * bsf, nbatch=nb, lda=2 * Cf, ldb=2 * bsf, ldd=2 * C,
                      a_ext=(2 * Cf, ntok), b_ext=(2 * bsf, nb * 2 * bs), a_off=2 * bsf, b_off1=2 * bs, d_off=2 * bs,
                      bias=be2, bias_off=2 * bs, act=4, lam=lam)
        extra = (E1, E2)
    else:
        O1 = _bl_fwd(Xh, w1c, b1c, ntok, nb, bs, bsf, 1, 0.0)
        Yh = _bl_fwd(O1, w2c, b2c, ntok, nb, bsf, bs, 2, lam)
        extra = None
    Phi = torch.empty(B, h, Mx, C, 2, device=x.device, dtype=torch.float32)
    _lib.check(lib.sb200_cl_coldft_inv(plan.handle, 0, _p(Yh), _p(Phi), B, C, _stream()), "cl_coldft_inv")
    y = torch.empty_like(x)
    _lib.check(lib.sb200_cl_rowidft_res2(plan.handle, 0, _p(Phi), _p(resid), _p(resid2), _p(y), B * h, C, _stream()),
               "cl_rowidft_res")
    return y, (plan, (B, h, w, C, nb, bs, bsf, ntok, extra), Xh, O1, Yh)


def afno_backward(saved, w1c, w2c, gy, resid_grad, need_gx: bool = True):
    """Backward of ``afno_forward``: gy [B,h,w,C] -> (gx or None, gw1, gb1, gw2, gb2); ``resid_grad`` (gy when the filter
    input was also its residual, else None) is added to gx inside the adjoint row synthesis."""
    plan, (B, h, w, C, nb, bs, bsf, ntok, extra), Xh, O1, Yh = saved
    lib = _lib.load()
    My, Mx = plan.My, plan.Mx
    dev = gy.device
    T = torch.empty(B, h, Mx, C, 2, device=dev, dtype=torch.float32)
    _lib.check(lib.sb200_cl_rowdft_fwd(plan.handle, 1, _p(gy), _p(T), B * h, C, _stream()), "cl_rowdft_fwd")
    gYh = torch.empty(B, My, Mx, C, 2, device=dev, dtype=torch.float32)
    _lib.check(lib.sb200_cl_coldft_fwd(plan.handle, 1, _p(T), _p(gYh), B, C, _stream()), "cl_coldft_fwd")
    gXh = None
    if extra is not None:
        from . import ops
        E1, E2 = extra
        Cf = nb * bsf
        dY2 = torch.empty_like(gYh)                                     # gYh * (Yh != 0): soft-shrink backward
        _lib.check(lib.sb200_mask_mul(_p(gYh), _p(Yh), _p(dY2), gYh.numel(), 2, _stream()), "mask_mul")
        gE2 = torch.empty(nb, 2 * bs, 2 * bsf, device=dev, dtype=torch.float32)
        _gemm_batched(dY2, O1, gE2, M=2 * bs, N=2 * bsf, K=ntok, nbatch=nb, lda=2 * C, ldb=2 * Cf, ldd=2 * bsf, a_mn=True,
                      b_mn=True, a_ext=(2 * C, ntok), b_ext=(2 * Cf, ntok), a_off=2 * bs, b_off0=2 * bsf,
                      d_off=2 * bs * 2 * bsf, split_k=True)
        gb2 = ops.colsum(dY2.view(ntok, 2 * C)).view(nb, bs, 2).permute(2, 0, 1).contiguous()
        dZ1 = torch.empty(B, My, Mx, Cf, 2, device=dev, dtype=torch.float32)     # (dY2 W2^H) * (O1 > 0)
        _gemm_batched(dY2, E2, dZ1, M=ntok, N=2 * bsf, K=2 * bs, nbatch=nb, lda=2 * C, ldb=2 * bsf, ldd=2 * Cf, b_mn=True,
                      a_ext=(2 * C, ntok), b_ext=(2 * bsf, nb * 2 * bs), a_off=2 * bs, b_off1=2 * bs, d_off=2 * bsf,
                      aux=O1, act=5)
        gE1 = torch.empty(nb, 2 * bsf, 2 * bs, device=dev, dtype=torch.float32)
        _gemm_batched(dZ1, Xh, gE1, M=2 * bsf, N=2 * bs, K=ntok, nbatch=nb, lda=2 * Cf, ldb=2 * C, ldd=2 * bs, a_mn=True,
                      b_mn=True, a_ext=(2 * Cf, ntok), b_ext=(2 * C, ntok), a_off=2 * bsf, b_off0=2 * bs,
                      d_off=2 * bsf * 2 * bs, split_k=True)
        gb1 = ops.colsum(dZ1.view(ntok, 2 * Cf)).view(nb, bsf, 2).permute(2, 0, 1).contiguous()
        gw1, gw2 = _unembed(gE1, nb, bs, bsf), _unembed(gE2, nb, bsf, bs)
        if need_gx:
            gXh = torch.empty_like(gYh)
            _gemm_batched(dZ1, E1, gXh, M=ntok, N=2 * bs, K=2 * bsf, nbatch=nb, lda=2 * Cf, ldb=2 * bs, ldd=2 * C, b_mn=True,
                          a_ext=(2 * Cf, ntok), b_ext=(2 * bs, nb * 2 * bsf), a_off=2 * bsf, b_off1=2 * bsf, d_off=2 * bs)
    else:
        gw2, gb2 = _bl_wgrad(O1, gYh, Yh, 2, ntok, nb, bsf, bs)
        gO1 = _bl_dgrad(gYh, Yh, 2, w2c, ntok, nb, bsf, bs)
        gw1, gb1 = _bl_wgrad(Xh, gO1, O1, 1, ntok, nb, bs, bsf)
        if need_gx:
            gXh = _bl_dgrad(gO1, O1, 1, w1c, ntok, nb, bs, bsf)
    gx = None
    if need_gx:
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
