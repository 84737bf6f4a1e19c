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


# --------------------------------------------------------------------------------------
# FourCastNet block remainder: restatement of the reference's ``Block`` / ``Mlp`` / ``PatchEmbed`` /
# ``AFNONet.forward_features`` + head in plain torch on a reference-layout state_dict.  PINNED like the filter:
# tests/golden/fcn_*.npz come from the reference's own classes (oracle/make_golden.py::make_fourcastnet).
# --------------------------------------------------------------------------------------
def block_forward(sd, prefix, x, num_blocks, sparsity_threshold=0.01, hard_thresholding_fraction=1.0, double_skip=True,
                  eps=1e-6):
    """reference ``Block.forward`` (src/dlwpbench/models/fourcastnet/fourcastnet.py:181-193); x [B,h,w,C]."""
    g = lambda k: sd[prefix + k].to(x.dtype)
    C = x.shape[-1]
    residual = x
    h = F.layer_norm(x, (C,), g("norm1.weight"), g("norm1.bias"), eps)
    h = afno2d_fft(h, g("filter.w1"), g("filter.b1"), g("filter.w2"), g("filter.b2"), num_blocks, sparsity_threshold,
                   hard_thresholding_fraction)
    if double_skip:
        h = h + residual
        residual = h
    h = F.layer_norm(h, (C,), g("norm2.weight"), g("norm2.bias"), eps)
    h = F.linear(F.gelu(F.linear(h, g("mlp.fc1.weight"), g("mlp.fc1.bias"))), g("mlp.fc2.weight"), g("mlp.fc2.bias"))
    return h + residual


def afnonet_step(sd, x_t, patch_size, depth, num_blocks, use_pos_embed=True, sparsity_threshold=0.01,
                 hard_thresholding_fraction=1.0, prefix=""):
    """PatchEmbed (+ pos_embed) -> blocks -> head -> pixel shuffle (reference :283-293, :343-353); x_t [B,Cin,H,W]
    -> [B,Cout,H,W] (the network proper, without the ``prognostic_t[:, -1] +`` residual of the loop)."""
    g = lambda k: sd[prefix + k].to(x_t.dtype)
    p1, p2 = patch_size
    B, _, H, W = x_t.shape
    h, w = H // p1, W // p2
    t = F.conv2d(x_t, g("patch_embed.proj.weight"), g("patch_embed.proj.bias"), stride=(p1, p2)).flatten(2).transpose(1, 2)
    if use_pos_embed:
        t = t + g("pos_embed")
    E = t.shape[-1]
    t = t.reshape(B, h, w, E)
    for i in range(depth):
        t = block_forward(sd, f"{prefix}blocks.{i}.", t, num_blocks, sparsity_threshold, hard_thresholding_fraction)
    y = F.linear(t, g("head.weight"))
    c_out = y.shape[-1] // (p1 * p2)
    y = y.reshape(B, h, w, p1, p2, c_out).permute(0, 5, 1, 3, 2, 4)
    return y.reshape(B, c_out, h * p1, w * p2)
