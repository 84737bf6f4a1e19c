"""CPU oracle for FourCastNet's AFNO2D block.  TEST INFRASTRUCTURE ONLY.

Restates ``AFNO2D.forward`` of the reference
(``src/nsbench/models/fourcastnet/fourcastnet.py:77-126`` == ``src/dlwpbench/models/fourcastnet/fourcastnet.py:78-127``)
in two independent forms:

* ``afno2d_fft``      -- torch.fft statement (rfft2 'ortho' -> block-diagonal complex 2-layer
                         MLP on the kept rows/cols -> softshrink -> irfft2 'ortho' -> + x)
* ``afno2d_explicit`` -- explicit truncated-DFT statement (the kernels' math spec)

PARITY PINNED: ``tests/golden/afno2d_*.npz`` were produced by importing the *reference class
itself* in the build container (``oracle/make_golden.py``); ``tests/test_oracle_golden.py``
checks both restatements against those vectors (outputs and all gradients).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may
import this module; the product package never does.
"""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F


def kept_region(h: int, w: int, frac: float):
    """Rows r0:r1 (unshifted FFT order) and cols :kc actually written by the reference.

    reference :92-95 -- ``total = h//2+1; kept = int(total*frac)``; rows
    ``total-kept : total+kept`` (python slicing clips at h), cols ``:kept`` (clips at w//2+1).
    """
    total = h // 2 + 1
    kept = int(total * frac)
    r0 = max(total - kept, 0)
    r1 = min(total + kept, h)
    kc = min(kept, w // 2 + 1)
    return r0, r1, kc


def _block_mlp(xr, xi, w1, b1, w2, b2, lam):
    """Complex block-diagonal 2-layer MLP written in real arithmetic (reference :95-121)."""
    mm = lambda a, w: torch.einsum("...bi,bio->...bo", a, w)
    o1r = F.relu(mm(xr, w1[0]) - mm(xi, w1[1]) + b1[0])
    o1i = F.relu(mm(xi, w1[0]) + mm(xr, w1[1]) + b1[1])
    o2r = mm(o1r, w2[0]) - mm(o1i, w2[1]) + b2[0]
    o2i = mm(o1i, w2[0]) + mm(o1r, w2[1]) + b2[1]
    return F.softshrink(o2r, lam), F.softshrink(o2i, lam)


def afno2d_fft(x, w1, b1, w2, b2, num_blocks: int, sparsity_threshold: float = 0.01,
               hard_thresholding_fraction: float = 1.0):
    """x [B,h,w,C]; w1 [2,nb,bs,bs*f]; b1 [2,nb,bs*f]; w2 [2,nb,bs*f,bs]; b2 [2,nb,bs]."""
    B, h, w, C = x.shape
    bs = C // num_blocks
    resid = x
    xf = torch.fft.rfft2(x, dim=(1, 2), norm="ortho").reshape(B, h, w // 2 + 1, num_blocks, bs)
    r0, r1, kc = kept_region(h, w, hard_thresholding_fraction)
    outr = torch.zeros_like(xf.real)
    outi = torch.zeros_like(xf.real)
    sub = xf[:, r0:r1, :kc]
    sr, si = _block_mlp(sub.real, sub.imag, w1, b1, w2, b2, sparsity_threshold)
    outr[:, r0:r1, :kc] = sr
    outi[:, r0:r1, :kc] = si
    of = torch.complex(outr, outi).reshape(B, h, w // 2 + 1, C)
    y = torch.fft.irfft2(of, s=(h, w), dim=(1, 2), norm="ortho")
    return y + resid


def afno2d_explicit(x, w1, b1, w2, b2, num_blocks: int, sparsity_threshold: float = 0.01,
                    hard_thresholding_fraction: float = 1.0):
    """Same operator via explicit truncated DFT matrices (scale 1/sqrt(hw) each way,
    Hermitian column weights 1/2/1 on synthesis)."""
    B, h, w, C = x.shape
    bs = C // num_blocks
    rd = x.dtype
    cd = torch.complex128 if rd == torch.float64 else torch.complex64
    r0, r1, kc = kept_region(h, w, hard_thresholding_fraction)
    s = 1.0 / math.sqrt(h * w)
    ky = torch.arange(r0, r1, dtype=rd)
    kx = torch.arange(kc, dtype=rd)
    EH = torch.exp(-2j * math.pi * ky[:, None] * torch.arange(h, dtype=rd)[None] / h).to(cd)
    EW = torch.exp(-2j * math.pi * kx[:, None] * torch.arange(w, dtype=rd)[None] / w).to(cd)
    Xh = s * torch.einsum("ky,byxc,qx->bkqc", EH, x.to(cd), EW)
    Xh = Xh.reshape(B, r1 - r0, kc, num_blocks, bs)
    sr, si = _block_mlp(Xh.real, Xh.imag, w1, b1, w2, b2, sparsity_threshold)
    Yh = torch.complex(sr, si).reshape(B, r1 - r0, kc, C)
    wk = torch.full((kc,), 2.0, dtype=rd)
    wk[0] = 1.0
    if w % 2 == 0 and kc == w // 2 + 1:
        wk[-1] = 1.0
    Phi = torch.einsum("ky,bkqc->byqc", EH.conj(), Yh)
    y = s * torch.einsum("byqc,qx,q->byxc", Phi, EW.conj(), wk.to(cd)).real
    return y + x
