"""Tensor-level wrappers over the C ABI (one python function per exported stage).

Every function requires CUDA fp32 contiguous tensors and launches on torch's current stream of the CURRENT
device; the autograd Functions make the tensors' device current first (``_lib.on_tensor_device``).
"""
from __future__ import annotations

import ctypes
import math
from typing import Optional

import torch

from . import _lib
from .plan import Plan

_vp = ctypes.c_void_p


def _p(t: Optional[torch.Tensor]):
    return _vp(t.data_ptr()) if t is not None else None


def _stream():
    return _vp(torch.cuda.current_stream().cuda_stream)


def _req(t: torch.Tensor, name: str):
    if not t.is_cuda:
        raise _lib.SpectralB200Error(f"{name}: expected a CUDA tensor (there is no CPU path)")
    if t.dtype != torch.float32:
        raise _lib.SpectralB200Error(f"{name}: expected float32, got {t.dtype}")
    if not t.is_contiguous():
        raise _lib.SpectralB200Error(f"{name}: expected a contiguous tensor")


def grad_sink(p):
    """Direct gradient destination of a parameter (``ddp.GradSync(direct=True)``): a contiguous fp32 view with the
    parameter's numel into the flat all-reduce buffer, or None.  A backward that writes its result there returns
    None for that input, so autograd neither allocates nor accumulates (the sink is OVERWRITTEN, not added to)."""
    s = getattr(p, "_sb200_grad_sink", None) if p is not None else None
    if s is not None:
        owner = getattr(p, "_sb200_sink_owner", p)
        owner._sb200_sink_used = True
    return s


def _out(out, shape, dev):
    """``out`` viewed as ``shape`` (a grad sink) or a fresh tensor."""
    if out is None:
        return torch.empty(*shape, device=dev, dtype=torch.float32)
    assert out.is_contiguous() and out.dtype == torch.float32 and out.numel() == math.prod(shape), (out.shape, shape)
    return out.view(*shape)


def rowdft_fwd(plan: Plan, pas: int, x: torch.Tensor) -> torch.Tensor:
    """x [..., W] -> T [..., Mx, 2]"""
    _req(x, "x")
    assert x.shape[-1] == plan.W
    rows = x.numel() // plan.W
    T = torch.empty(*x.shape[:-1], plan.Mx, 2, device=x.device, dtype=torch.float32)
    _lib.check(_lib.load().sb200_rowdft_fwd(plan.handle, pas, _p(x), _p(T), rows, _stream(), _lib.tc_mode()), "rowdft_fwd")
    return T


def coldft_fwd(plan: Plan, pas: int, T: torch.Tensor) -> torch.Tensor:
    """T [..., H, Mx, 2] -> Xh [..., My, Mx, 2]"""
    _req(T, "T")
    assert T.shape[-3:] == (plan.H, plan.Mx, 2)
    nimg = T.numel() // (plan.H * plan.Mx * 2)
    Xh = torch.empty(*T.shape[:-3], plan.My, plan.Mx, 2, device=T.device, dtype=torch.float32)
    _lib.check(_lib.load().sb200_coldft_fwd(plan.handle, pas, _p(T), _p(Xh), nimg, _stream()), "coldft_fwd")
    return Xh


def coldft_inv(plan: Plan, pas: int, Yh: torch.Tensor) -> torch.Tensor:
    """Yh [..., My, Mx, 2] -> Phi [..., H, Mx, 2]"""
    _req(Yh, "Yh")
    assert Yh.shape[-3:] == (plan.My, plan.Mx, 2)
    nimg = Yh.numel() // (plan.My * plan.Mx * 2)
    Phi = torch.empty(*Yh.shape[:-3], plan.H, plan.Mx, 2, device=Yh.device, dtype=torch.float32)
    _lib.check(_lib.load().sb200_coldft_inv(plan.handle, pas, _p(Yh), _p(Phi), nimg, _stream()), "coldft_inv")
    return Phi


def analysis(plan: Plan, pas: int, x: torch.Tensor) -> torch.Tensor:
    """x [..., H, W] -> Xh [..., My, Mx, 2] (row + column stages; one fused kernel on small grids)."""
    _req(x, "x")
    assert x.shape[-2:] == (plan.H, plan.W)
    nimg = x.numel() // (plan.H * plan.W)
    lib = _lib.load()
    n = lib.sb200_analysis_scratch(plan.handle, nimg)
    if n == 0 and (x.data_ptr() % 16 != 0 or nimg * plan.H >= 2 ** 31):
        n = nimg * plan.H * plan.Mx * 2       # the fused kernel declines such inputs at launch: two-stage scratch
    scratch = torch.empty(n, device=x.device, dtype=torch.float32) if n > 0 else None
    Xh = torch.empty(*x.shape[:-2], plan.My, plan.Mx, 2, device=x.device, dtype=torch.float32)
    _lib.check(lib.sb200_analysis(plan.handle, pas, _p(x), _p(Xh), nimg, _p(scratch), _stream(), _lib.tc_mode()), "analysis")
    return Xh


def modes_gemm(A, sAr, sAp, B, sBr, sBq, out, sOp, sOq, P, Q, R, K, conj_flags=0):
    _req(A, "A"); _req(B, "B"); _req(out, "out")
    _lib.check(_lib.load().sb200_modes_gemm(_p(A), sAr, sAp, _p(B), sBr, sBq, _p(out), sOp, sOq,
                                            P, Q, R, K, conj_flags, _stream()), "modes_gemm")
    return out


def mix_fwd(Xh: torch.Tensor, Wc: torch.Tensor) -> torch.Tensor:
    """Yh[b,o,k] = sum_i Xh[b,i,k] W[i,o,k];  Xh [B,Cin,My,Mx,2], Wc [Cin,Cout,My,Mx,2]."""
    B, Cin = Xh.shape[:2]
    Cout = Wc.shape[1]
    M = Xh.shape[2] * Xh.shape[3]
    out = torch.empty(B, Cout, *Xh.shape[2:], device=Xh.device, dtype=torch.float32)
    return modes_gemm(Xh, M, Cin * M, Wc, Cout * M, M, out, Cout * M, M, B, Cout, Cin, M, 0)


def mix_bwd_input(gYh: torch.Tensor, Wc: torch.Tensor) -> torch.Tensor:
    """gXh[b,i,k] = sum_o conj(W[i,o,k]) gYh[b,o,k]"""
    B, Cout = gYh.shape[:2]
    Cin = Wc.shape[0]
    M = gYh.shape[2] * gYh.shape[3]
    out = torch.empty(B, Cin, *gYh.shape[2:], device=gYh.device, dtype=torch.float32)
    return modes_gemm(gYh, M, Cout * M, Wc, M, Cout * M, out, Cin * M, M, B, Cin, Cout, M, 2)


def mix_bwd_weight(Xh: torch.Tensor, gYh: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """gW[i,o,k] = sum_b conj(Xh[b,i,k]) gYh[b,o,k]"""
    B, Cin = Xh.shape[:2]
    Cout = gYh.shape[1]
    M = Xh.shape[2] * Xh.shape[3]
    out = _out(out, (Cin, Cout, *Xh.shape[2:]), Xh.device)
    return modes_gemm(Xh, Cin * M, M, gYh, Cout * M, M, out, Cout * M, M, Cin, Cout, B, M, 1)


def rowidft_pointwise(plan: Plan, pas: int, Phi, A, Wp, w_sn, w_sm, bias, zprev, B, M, N, mode, apply_act,
                      want_z: bool = False):
    """Fused row synthesis + pointwise mix + epilogue; returns (y, z or None)."""
    dev = (Phi if Phi is not None else A).device
    for t, n in ((Phi, "Phi"), (A, "A"), (Wp, "Wp"), (bias, "bias"), (zprev, "zprev")):
        if t is not None:
            _req(t, n)
    y = torch.empty(B, N, plan.H, plan.W, device=dev, dtype=torch.float32)
    z = torch.empty_like(y) if want_z else None
    _lib.check(_lib.load().sb200_rowidft_pointwise(plan.handle, pas, _p(Phi), _p(A), _p(Wp), w_sn, w_sm, _p(bias),
                                                   _p(zprev), _p(z), _p(y), B, M, N, mode, int(apply_act), _stream(), _lib.tc_mode()),
               "rowidft_pointwise")
    return y, z


def pointwise_wgrad(g: torch.Tensor, x: torch.Tensor, want_bias: bool = True, out_w=None, out_b=None):
    """gW[o,i] = sum g[b,o,p] x[b,i,p]; gbias[o] = sum g[b,o,p].  g [B,Cout,H,W], x [B,Cin,H,W]."""
    _req(g, "g"); _req(x, "x")
    B, Cout = g.shape[:2]
    Cin = x.shape[1]
    HW = g.shape[2] * g.shape[3]
    lib = _lib.load()
    n = lib.sb200_pointwise_wgrad_workspace(B, Cout, Cin, HW, _lib.tc_mode())
    ws = torch.empty(n, device=g.device, dtype=torch.float32)
    gW = _out(out_w, (Cout, Cin), g.device)
    gb = _out(out_b, (Cout,), g.device) if want_bias else None
    _lib.check(lib.sb200_pointwise_wgrad(_p(g), _p(x), _p(gW), _p(gb), B, Cout, Cin, HW, _p(ws), _stream(), _lib.tc_mode()),
               "pointwise_wgrad")
    return gW, gb


def channel_sum(g: torch.Tensor) -> torch.Tensor:
    """out[c] = sum_{b,h,w} g[b,c,h,w]"""
    _req(g, "g")
    B, C = g.shape[:2]
    HW = g.shape[2] * g.shape[3]
    lib = _lib.load()
    ws = torch.empty(lib.sb200_channel_sum_workspace(B, C), device=g.device, dtype=torch.float32)
    out = torch.empty(C, device=g.device, dtype=torch.float32)
    _lib.check(lib.sb200_channel_sum(_p(g), _p(out), B, C, HW, _p(ws), _stream()), "channel_sum")
    return out


def gelu_fwd(z: torch.Tensor) -> torch.Tensor:
    _req(z, "z")
    y = torch.empty_like(z)
    _lib.check(_lib.load().sb200_gelu_fwd(_p(z), _p(y), z.numel(), _stream()), "gelu_fwd")
    return y


def gelu_bwd(gy: torch.Tensor, z: torch.Tensor) -> torch.Tensor:
    _req(gy, "gy"); _req(z, "z")
    gz = torch.empty_like(z)
    _lib.check(_lib.load().sb200_gelu_bwd(_p(gy), _p(z), _p(gz), z.numel(), _stream()), "gelu_bwd")
    return gz


def pointwise_small_n(A: torch.Tensor, Wp: torch.Tensor, bias, apply_act: bool, want_z: bool = False):
    """out[b,n,p] = sum_m Wp[n,m] A[b,m,p] + bias[n] for N <= 8; returns (y, z or None)."""
    _req(A, "A"); _req(Wp, "Wp")
    B, M, H, W = A.shape
    N = Wp.shape[0]
    y = torch.empty(B, N, H, W, device=A.device, dtype=torch.float32)
    z = torch.empty_like(y) if want_z else None
    _lib.check(_lib.load().sb200_pointwise_small_n(_p(A), _p(Wp), _p(bias), _p(z), _p(y), B, M, N, H * W,
                                                   int(apply_act), _stream()), "pointwise_small_n")
    return y, z


def wgrad_small(small: torch.Tensor, big: torch.Tensor, transpose: bool, want_small_sum: bool, want_big_sum: bool,
                out_dot=None, out_small=None, out_big=None):
    """dot[s,l] = sum small[b,s,p] big[b,l,p] (returned as [l,s] if transpose); channel sums on request."""
    _req(small, "small"); _req(big, "big")
    B, S = small.shape[:2]
    L = big.shape[1]
    HW = small.shape[2] * small.shape[3]
    lib = _lib.load()
    ws = torch.empty(lib.sb200_wgrad_small_workspace(B, S, L, HW), device=small.device, dtype=torch.float32)
    dot = _out(out_dot, (L, S) if transpose else (S, L), small.device)
    ss = _out(out_small, (S,), small.device) if want_small_sum else None
    bs = _out(out_big, (L,), small.device) if want_big_sum else None
    _lib.check(lib.sb200_wgrad_small(_p(small), _p(big), _p(dot), _p(ss), _p(bs), B, S, L, HW, int(transpose), _p(ws),
                                     _stream()), "wgrad_small")
    return dot, ss, bs


class CgemmDesc(ctypes.Structure):
    """Mirror of ``sb200_cgemm_desc`` (include/spectral_b200.h)."""
    _fields_ = [(n, ctypes.c_int32) for n in ("M", "N", "K", "M2", "K2", "conjA", "conjB", "reserved")] + \
               [(n, ctypes.c_int64) for n in ("sAm1", "sAm2", "sAk1", "sAk2", "sBk1", "sBk2", "sBn",
                                              "sCm1", "sCm2", "sCn")]


def cgemm(A, B, C, *, M: int, N: int, K: int, sAm, sAk, sBk, sBn, sCm, sCn,
          M2: int = 1, K2: int = 1, conjA: bool = False, conjB: bool = False):
    """C[m,n] = sum_k opA(A[m,k]) opB(B[k,n]) on interleaved-complex fp32 storage; strides in complex elements.
    ``sAm``/``sCm`` are (outer, inner) pairs when ``M2 > 1``, ``sAk``/``sBk`` likewise when ``K2 > 1``.
    A, B, C may be lists (<= 8) of tensors of identical geometry: one launch computes all of them."""
    single = isinstance(A, torch.Tensor)
    As, Bs, Cs = ([A], [B], [C]) if single else (list(A), list(B), list(C))
    assert len(As) == len(Bs) == len(Cs) and 1 <= len(As) <= 8
    for t in As + Bs + Cs:
        _req(t, "cgemm operand")
    pair = lambda s: (int(s[0]), int(s[1])) if isinstance(s, (tuple, list)) else (0, int(s))
    d = CgemmDesc()
    d.M, d.N, d.K, d.M2, d.K2, d.conjA, d.conjB = M, N, K, M2, K2, int(conjA), int(conjB)
    d.sAm1, d.sAm2 = pair(sAm); d.sAk1, d.sAk2 = pair(sAk)
    d.sBk1, d.sBk2 = pair(sBk); d.sBn = int(sBn)
    d.sCm1, d.sCm2 = pair(sCm); d.sCn = int(sCn)
    lib = _lib.load()
    ng = len(As)
    nws = lib.sb200_cgemm_workspace(ctypes.byref(d), ng)
    ws = torch.empty(nws, device=As[0].device, dtype=torch.float32) if nws > 0 else None
    arr = lambda ts: (ctypes.c_void_p * ng)(*[t.data_ptr() for t in ts])
    _lib.check(lib.sb200_cgemm_grouped(ctypes.byref(d), ng, arr(As), arr(Bs), arr(Cs), _p(ws), _stream()), "cgemm")
    return C


def mlp_head_supported(M: int, N: int, n_out: int, HW: int) -> bool:
    """Shapes the fused head kernels cover (sb200_mlp_head_fwd / _bwd)."""
    return (n_out == 1 and N == 256 and M % 8 == 0 and (M <= 32 or M % 32 == 0) and HW % 4 == 0
            and _lib.tc_mode() != 0)


def mlp_head_fwd(h, W1, b1, w2, b2):
    """y[b,0,:,:] = w2 . gelu(W1 h + b1) + b2;  h [B,M,H,W], W1 [256,M], b1 [256], w2 [256], b2 [1] or None."""
    for t, n in ((h, "h"), (W1, "W1"), (b1, "b1"), (w2, "w2")):
        _req(t, n)
    B, M, H, W = h.shape
    y = torch.empty(B, 1, H, W, device=h.device, dtype=torch.float32)
    _lib.check(_lib.load().sb200_mlp_head_fwd(_p(h), _p(W1), _p(b1), _p(w2), _p(b2), _p(y), B, M, W1.shape[0], H * W,
                                              _stream(), _lib.tc_mode()), "mlp_head_fwd")
    return y


def mlp_head_bwd(h, W1, b1, w2, gy, want_gb2: bool = True, out_gb1=None, out_gw2=None, out_gb2=None):
    """Returns (gz1 [B,256,H,W], gb1 [256], gw2 [256], gb2 [1] or None)."""
    for t, n in ((h, "h"), (W1, "W1"), (b1, "b1"), (w2, "w2"), (gy, "gy")):
        _req(t, n)
    B, M, H, W = h.shape
    N = W1.shape[0]
    lib = _lib.load()
    dev = h.device
    gz1 = torch.empty(B, N, H, W, device=dev, dtype=torch.float32)
    gb1 = _out(out_gb1, (N,), dev)
    gw2 = _out(out_gw2, (N,), dev)
    gb2 = _out(out_gb2, (1,), dev) if want_gb2 else None
    ws = torch.empty(lib.sb200_mlp_head_bwd_workspace(), device=dev, dtype=torch.float32)
    _lib.check(lib.sb200_mlp_head_bwd(_p(h), _p(W1), _p(b1), _p(w2), _p(gy), _p(gz1), _p(gb1), _p(gw2), _p(gb2), _p(ws),
                                      B, M, N, H * W, _stream(), _lib.tc_mode()), "mlp_head_bwd")
    return gz1, gb1, gw2, gb2


def lift_tail_bwd(g, W2, w1, b1, x, out_gw1=None, out_gb1=None):
    """Backward of the first lifting layer for a 1-channel input: returns (gw1 [256], gb1 [256]).
    g [B,C,H,W] gradient wrt the lifting output, W2 [C,256], w1/b1 [256], x [B,1,H,W]."""
    for t, n in ((g, "g"), (W2, "W2"), (w1, "w1"), (b1, "b1"), (x, "x")):
        _req(t, n)
    B, C, H, W = g.shape
    N = W2.shape[1]
    lib = _lib.load()
    gw1 = _out(out_gw1, (N,), g.device)
    gb1 = _out(out_gb1, (N,), g.device)
    ws = torch.empty(lib.sb200_mlp_head_bwd_workspace(), device=g.device, dtype=torch.float32)
    _lib.check(lib.sb200_lift_tail_bwd(_p(g), _p(W2), _p(w1), _p(b1), _p(x), _p(gw1), _p(gb1), _p(ws), B, C, N, H * W,
                                       _stream(), _lib.tc_mode()), "lift_tail_bwd")
    return gw1, gb1


def lift_supported(C: int, N: int, HW: int) -> bool:
    """Shapes the generated-operand lifting kernels cover (1 input channel, hidden 256)."""
    return (N == 256 and C % 16 == 0 and 16 <= C < N and HW % 4 == 0 and _lib.tc_mode() != 0)


def lift_fwd(x, w1, b1, W2, b2):
    """y = W2 gelu(w1 x + b1) + b2 for x [B,1,H,W]; w1/b1 [256], W2 [C,256], b2 [C] or None -> y [B,C,H,W]."""
    for t, n in ((x, "x"), (w1, "w1"), (b1, "b1"), (W2, "W2")):
        _req(t, n)
    B, _, H, W = x.shape
    C, N = W2.shape
    y = torch.empty(B, C, H, W, device=x.device, dtype=torch.float32)
    _lib.check(_lib.load().sb200_lift_fwd(_p(x), _p(w1), _p(b1), _p(W2), _p(b2), _p(y), B, N, C, H * W, _stream(), _lib.tc_mode()),
               "lift_fwd")
    return y


def lift_wgrad(g, x, w1, b1, want_bias: bool = True, out_w=None, out_b=None):
    """(gW2 [C,256], gb2 [C] or None) with the hidden activations regenerated from x on chip."""
    for t, n in ((g, "g"), (x, "x"), (w1, "w1"), (b1, "b1")):
        _req(t, n)
    B, C, H, W = g.shape
    N = w1.numel()
    lib = _lib.load()
    ws = torch.empty(lib.sb200_pointwise_wgrad_workspace(B, C, N, H * W, _lib.tc_mode()), device=g.device, dtype=torch.float32)
    gW2 = _out(out_w, (C, N), g.device)
    gb2 = _out(out_b, (C,), g.device) if want_bias else None
    _lib.check(lib.sb200_lift_wgrad(_p(g), _p(x), _p(w1), _p(b1), _p(gW2), _p(gb2), _p(ws), B, C, N, H * W, _stream(), _lib.tc_mode()),
               "lift_wgrad")
    return gW2, gb2


# ---------------------------------------------------------------------------------------------
# channels-last token path (FourCastNet block remainder)
# ---------------------------------------------------------------------------------------------
def gemm(A: torch.Tensor, B: torch.Tensor, *, a_mn: bool = False, b_mn: bool = False, bias=None, act: int = 0, aux=None,
         resid=None, res_rows: int = 0, want_z: bool = False, out=None, split_k: bool = False, a_gelu: bool = False,
         b_gelu: bool = False, z_only: bool = False):
    """D[m,n] = sum_k A(m,k) B(n,k) on 2-D contiguous fp32 tensors.  ``a_mn=False``: A is [M,K] (K-major);
    ``a_mn=True``: A is [K,M] (used transposed, MN-major); likewise B ([N,K] or [K,N]).  Epilogue: + bias[n];
    z = pre-activation (returned when ``want_z``); act 1 = GELU, 2 = multiply by GELU'(aux[m,n]); + resid (row
    ``m % res_rows`` when ``res_rows > 0``).  ``split_k``: K split over CTAs (weight gradients), no epilogue.
    ``a_gelu`` / ``b_gelu``: the operand is GELU(A) / GELU(B), applied on chip.  ``z_only``: only the pre-activation
    z = A B^T + bias is produced (returned alone).  Returns D, or (D, z) when ``want_z``."""
    _req(A, "A"); _req(B, "B")
    assert A.dim() == 2 and B.dim() == 2
    (K, M) = A.shape if a_mn else A.shape[::-1]
    (Kb, N) = B.shape if b_mn else B.shape[::-1]
    assert K == Kb, (A.shape, B.shape, a_mn, b_mn)
    for t, n in ((bias, "bias"), (aux, "aux"), (resid, "resid")):
        if t is not None:
            _req(t, n)
    dev = A.device
    D = None if z_only else _out(out, (M, N), dev)
    z = torch.empty(M, N, device=dev, dtype=torch.float32) if (want_z or z_only) else None
    lib = _lib.load()
    nws = lib.sb200_gemm_workspace(M, N, K, int(b_mn), int(split_k), _lib.tc_mode())
    ws = torch.empty(nws, device=dev, dtype=torch.float32) if nws > 0 else None
    _lib.check(lib.sb200_gemm(_p(A), A.shape[1], int(a_mn), _p(B), B.shape[1], int(b_mn), _p(D), N, M, N, K, _p(bias), act,
                              _p(aux), N, _p(resid), N, res_rows, _p(z), N, int(a_gelu), int(b_gelu), int(split_k), _p(ws),
                              _stream(), _lib.tc_mode()), "gemm")
    if z_only:
        return z
    return (D, z) if want_z else D


def layernorm_fwd(x: torch.Tensor, gamma, beta, eps: float):
    """x [T, C] -> (y [T, C], mean [T], rstd [T])"""
    _req(x, "x"); _req(gamma, "gamma"); _req(beta, "beta")
    T, C = x.shape
    y = torch.empty_like(x)
    mean = torch.empty(T, device=x.device, dtype=torch.float32)
    rstd = torch.empty(T, device=x.device, dtype=torch.float32)
    _lib.check(_lib.load().sb200_layernorm_fwd(_p(x), _p(gamma), _p(beta), _p(y), _p(mean), _p(rstd), T, C,
                                               ctypes.c_float(eps), _stream()), "layernorm_fwd")
    return y, mean, rstd


def layernorm_bwd(dy, x, gamma, mean, rstd, dres=None, out_g=None, out_b=None):
    """-> (dx [T, C] (+ dres), dgamma [C], dbeta [C])"""
    for t, n in ((dy, "dy"), (x, "x"), (gamma, "gamma"), (mean, "mean"), (rstd, "rstd")):
        _req(t, n)
    if dres is not None:
        _req(dres, "dres")
    T, C = x.shape
    lib = _lib.load()
    ws = torch.empty(lib.sb200_layernorm_bwd_workspace(T, C), device=x.device, dtype=torch.float32)
    dx = torch.empty_like(x)
    dg = _out(out_g, (C,), x.device)
    db = _out(out_b, (C,), x.device)
    _lib.check(lib.sb200_layernorm_bwd(_p(dy), _p(x), _p(gamma), _p(mean), _p(rstd), _p(dres), _p(dx), _p(dg), _p(db), _p(ws),
                                       T, C, _stream()), "layernorm_bwd")
    return dx, dg, db


def colsum(a: torch.Tensor, out=None) -> torch.Tensor:
    """a [T, N] -> [N]"""
    _req(a, "a")
    T, N = a.shape
    lib = _lib.load()
    ws = torch.empty(lib.sb200_colsum_workspace(T, N), device=a.device, dtype=torch.float32)
    o = _out(out, (N,), a.device)
    _lib.check(lib.sb200_colsum(_p(a), N, _p(o), T, N, _p(ws), _stream()), "colsum")
    return o


def batch_sum(a: torch.Tensor, out=None) -> torch.Tensor:
    """a [B, ...] -> [...]"""
    _req(a, "a")
    B = a.shape[0]
    n = a.numel() // B
    o = _out(out, tuple(a.shape[1:]), a.device)
    _lib.check(_lib.load().sb200_batch_sum(_p(a), _p(o), B, n, _stream()), "batch_sum")
    return o
