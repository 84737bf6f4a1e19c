"""CPU oracle for the FNO / TFNO spectral-convolution hot path.  TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` leg may import this module.  The product package
(``dlwp_benchmark_b200``) never imports it and has no CPU path.

What this restates
------------------
The reference (amazon-science/dlwp-benchmark) wraps ``neuralop.models.FNO/TFNO``
(reference call sites: ``src/nsbench/models/fno/fno.py:19-27,56-65,205-215``,
``src/dlwpbench/models/fno/fno.py:38-47,136-146``,
``src/dlwpbench/models/fourcastnet/fourcastnet.py:134-143``).  The arithmetic lives in
the third-party package ``neuraloperator`` pinned by the reference README
(``README.md:34-35``, git commit ``05c01c3``) which pulls ``tensorly`` +
``tensorly-torch`` (unpinned).  None of the three is vendored under ``/root/reference``
nor installed in this image, so this file restates the *published* algorithm of that
commit (``neuralop/layers/spectral_convolution.py::SpectralConv.forward``,
``neuralop/layers/fno_block.py::FNOBlocks.forward_with_postactivation``,
``neuralop/layers/mlp.py::MLP``, ``neuralop/models/fno.py::FNO.forward``,
``tltorch`` Tucker reconstruction, ``tensorly.validate_tucker_rank``) with plain
``torch.fft`` / ``torch.einsum``.

PARITY UNPINNED for this file: the reference holds no test, golden vector or fixture
for the FNO path and neuralop cannot be executed here.  The only
in-tree corroboration is the parameter-count labels of the reference's own sweeps
(checked in ``tests/test_oracle_golden.py::test_reference_size_labels``).  The AFNO2D oracle (``afno_oracle.py``)
*is* pinned against the reference class executed in this container.

Every function works in the dtype of its inputs (use float64 for a tight oracle,
float32 to mimic the reference's own rounding).
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F


# --------------------------------------------------------------------------------------
# mode bookkeeping (neuralop SpectralConv.n_modes setter + forward slicing)
# --------------------------------------------------------------------------------------
def halve_last_mode(n_modes: Sequence[int]) -> List[int]:
    """neuralop ``SpectralConv.n_modes`` setter: the last (rfft) dim keeps n//2+1 modes."""
    m = list(n_modes)
    m[-1] = m[-1] // 2 + 1
    return m


def _centre_slices(sizes: Sequence[int], kept: Sequence[int]) -> List[slice]:
    """The ``starts -> slice(start//2, -start//2)`` rule of SpectralConv.forward.

    ``sizes`` / ``kept`` are per spectral dim (last one is the rfft dim, kept from 0).
    Python's ``-start//2`` is ``floor(-start/2)`` which matters for odd ``start``.
    """
    starts = [s - min(s, k) for s, k in zip(sizes, kept)]
    out = []
    for st in starts[:-1]:
        out.append(slice(st // 2, -st // 2) if st else slice(st, None))
    out.append(slice(None, -starts[-1]) if starts[-1] else slice(None))
    return out


def retained_rows(H: int, n0: int) -> Tuple[int, int]:
    """(first shifted row, count) of the retained ky block; shifted row r <-> ky = r - H//2."""
    sl = _centre_slices([H, 1], [n0, 1])[0]
    lo, hi, _ = sl.indices(H)
    return lo, hi - lo


# --------------------------------------------------------------------------------------
# SpectralConv forward (dense weight)  -- neuralop SpectralConv.forward, fftshift era
# --------------------------------------------------------------------------------------
def spectral_conv_dense(x: torch.Tensor, weight: torch.Tensor, bias: Optional[torch.Tensor],
                        n_modes: Sequence[int], max_n_modes: Optional[Sequence[int]] = None,
                        fft_norm: str = "forward") -> torch.Tensor:
    """x [B,Cin,H,W] real; weight [Cin,Cout,*max_n_modes] complex; bias [Cout,1,1] or None.

    ``n_modes`` is the *already halved* per-dim mode count (what ``SpectralConv.n_modes``
    returns); ``max_n_modes`` the weight's spectral shape (defaults to ``n_modes``).
    """
    B, Cin, H, W = x.shape
    n_modes = list(n_modes)
    max_n_modes = list(max_n_modes) if max_n_modes is not None else n_modes
    fft_size = [H, W // 2 + 1]
    xf = torch.fft.rfftn(x, norm=fft_norm, dim=(-2, -1))
    xf = torch.fft.fftshift(xf, dim=(-2,))
    out_fft = torch.zeros(B, weight.shape[1], *fft_size, dtype=xf.dtype, device=xf.device)
    # slice the weight down to min(size, n_modes) out of max_n_modes
    w_sl = _centre_slices(max_n_modes, [min(s, n) for s, n in zip(fft_size, n_modes)])
    w = weight[(slice(None), slice(None), *w_sl)]
    x_sl = _centre_slices(fft_size, list(w.shape[2:]))
    idx = (slice(None), slice(None), *x_sl)
    out_fft[idx] = torch.einsum("bixy,ioxy->boxy", xf[idx], w.to(xf.dtype))
    out_fft = torch.fft.fftshift(out_fft, dim=(-2,))
    y = torch.fft.irfftn(out_fft, s=(H, W), dim=(-2, -1), norm=fft_norm)
    if bias is not None:
        y = y + bias
    return y


# --------------------------------------------------------------------------------------
# Tucker (TFNO): rank rule, reconstruction, factorised contraction
# --------------------------------------------------------------------------------------
def tucker_rank(shape: Sequence[int], rank: float) -> Tuple[int, ...]:
    """tensorly ``validate_tucker_rank`` for a float rank (fraction of parameters kept).

    Solve  f^n * prod(s) + f * sum(s_k^2) = rank * prod(s)  for f in [0, max(rank,1)]
    and round s_k * f (``round``), floor at 1.  (SURVEY.md 8c known answers:
    (64,64,16,9), 0.8 -> (60,60,15,8); (32,32,12,7), 0.8 -> (30,30,11,7); rank 1.0 -> full.)
    """
    shape = [int(s) for s in shape]
    if isinstance(rank, int) and not isinstance(rank, bool):
        return tuple(min(rank, s) for s in shape)
    n = len(shape)
    prod = float(math.prod(shape))
    sq = float(sum(s * s for s in shape))
    target = float(rank) * prod

    def fun(f):
        return prod * f ** n + sq * f - target

    lo, hi = 0.0, max(float(rank), 1.0)
    for _ in range(200):
        mid = 0.5 * (lo + hi)
        if fun(mid) > 0:
            hi = mid
        else:
            lo = mid
    f = 0.5 * (lo + hi)
    return tuple(max(int(round(s * f)), 1) for s in shape)


def tucker_to_dense(core: torch.Tensor, factors: Sequence[torch.Tensor]) -> torch.Tensor:
    """W[i,o,p,q] = sum core[f,g,h,j] U0[i,f] U1[o,g] U2[p,h] U3[q,j]  (all complex)."""
    return torch.einsum("fghj,if,og,ph,qj->iopq", core, *factors)


def spectral_conv_tucker(x, core, factors, bias, n_modes, fft_norm="forward"):
    """neuralop ``_contract_tucker`` ('abcd,fghi,bf,eg,ch,di->aecd'), never materialising W."""
    B, Cin, H, W = x.shape
    fft_size = [H, W // 2 + 1]
    xf = torch.fft.fftshift(torch.fft.rfftn(x, norm=fft_norm, dim=(-2, -1)), dim=(-2,))
    Cout = factors[1].shape[0]
    out_fft = torch.zeros(B, Cout, *fft_size, dtype=xf.dtype, device=xf.device)
    kept = [min(s, n) for s, n in zip(fft_size, n_modes)]
    w_sl = _centre_slices([factors[2].shape[0], factors[3].shape[0]], kept)
    f2, f3 = factors[2][w_sl[0]], factors[3][w_sl[1]]
    x_sl = _centre_slices(fft_size, [f2.shape[0], f3.shape[0]])
    idx = (slice(None), slice(None), *x_sl)
    # 'abcd,fghi,bf,eg,ch,di->aecd' evaluated pairwise in the order a contraction-path optimiser
    # (opt_einsum, which tensorly's einsum uses when present) picks; a naive left-to-right
    # torch.einsum would build an 8-index outer product.
    cd = xf.dtype
    xs = xf[idx]
    t1 = torch.einsum("abcd,bf->afcd", xs, factors[0].to(cd))
    t2 = torch.einsum("fghi,ch,di->fgcd", core.to(cd), f2.to(cd), f3.to(cd))
    t3 = torch.einsum("afcd,fgcd->agcd", t1, t2)
    out_fft[idx] = torch.einsum("agcd,eg->aecd", t3, factors[1].to(cd))
    out_fft = torch.fft.fftshift(out_fft, dim=(-2,))
    y = torch.fft.irfftn(out_fft, s=(H, W), dim=(-2, -1), norm=fft_norm)
    if bias is not None:
        y = y + bias
    return y


# --------------------------------------------------------------------------------------
# Explicit-DFT statement of the same operator (the kernels' math spec)
# --------------------------------------------------------------------------------------
def explicit_tables(H: int, W: int, n_modes: Sequence[int], dtype=torch.float64):
    """Returns (ky [My], kx [Mx], w_kx [Mx]) for the retained block (even H)."""
    lo, My = retained_rows(H, n_modes[0])
    Mx = min(n_modes[1], W // 2 + 1)
    ky = torch.arange(lo, lo + My, dtype=dtype) - (H // 2)
    kx = torch.arange(Mx, dtype=dtype)
    w = torch.full((Mx,), 2.0, dtype=dtype)
    w[0] = 1.0
    if W % 2 == 0 and Mx == W // 2 + 1:
        w[-1] = 1.0
    return ky, kx, w


def spectral_conv_explicit(x, weight, bias, n_modes):
    """Direct O(HW*M) evaluation of SpectralConv (norm='forward', even H). Small cases only.
    ``n_modes`` is the already-halved mode tuple, as for ``spectral_conv_dense``."""
    B, Cin, H, W = x.shape
    rd = x.dtype
    cd = torch.complex128 if rd == torch.float64 else torch.complex64
    ky, kx, w = explicit_tables(H, W, n_modes, dtype=rd)
    yy = torch.arange(H, dtype=rd)
    xx = torch.arange(W, dtype=rd)
    EH = torch.exp(-2j * math.pi * (ky[:, None] * yy[None, :]) / H).to(cd)   # [My,H]
    EW = torch.exp(-2j * math.pi * (kx[:, None] * xx[None, :]) / W).to(cd)   # [Mx,W]
    Xh = torch.einsum("ky,bcyx,qx->bckq", EH, x.to(cd), EW) / (H * W)
    Yh = torch.einsum("bikq,iokq->bokq", Xh, weight.to(cd))
    Phi = torch.einsum("ky,bokq->boyq", EH.conj(), Yh)                       # [B,O,H,Mx]
    y = torch.einsum("boyq,qx,q->boyx", Phi, EW.conj(), w.to(cd)).real
    if bias is not None:
        y = y + bias
    return y


# --------------------------------------------------------------------------------------
# FNOBlocks / MLP / FNO forward on a neuralop-layout state_dict
# --------------------------------------------------------------------------------------
def _conv1x1(x, w, b=None):
    return F.conv2d(x, w, b)


def fno_forward(sd: Dict[str, torch.Tensor], x: torch.Tensor, n_modes: Sequence[int], n_layers: int,
                prefix: str = "", return_intermediates: bool = False):
    """neuralop ``FNO.forward`` (defaults: postactivation, linear skip, no norm/mlp, GELU,
    fft_norm='forward', lifting/projection = 2-layer 1x1-conv MLPs).

    ``sd`` uses the neuralop key layout: ``lifting.fcs.{0,1}.{weight,bias}``,
    ``fno_blocks.convs.weight.{l}.tensor`` (real view [...,2]) or ``.core`` / ``.factors.factor_{k}``,
    ``fno_blocks.convs.bias`` [L,C,1,1], ``fno_blocks.fno_skips.{l}.weight``,
    ``projection.fcs.{0,1}.{weight,bias}``.  ``n_modes`` is the user-facing (un-halved) tuple.
    """
    g = lambda k: sd[prefix + k].to(x.dtype)
    nm = halve_last_mode(n_modes)
    inter = []
    h = _conv1x1(x, g("lifting.fcs.0.weight"), g("lifting.fcs.0.bias"))
    h = F.gelu(h)
    h = _conv1x1(h, g("lifting.fcs.1.weight"), g("lifting.fcs.1.bias"))
    inter.append(h)
    bias = g("fno_blocks.convs.bias")
    for l in range(n_layers):
        skip = _conv1x1(h, g(f"fno_blocks.fno_skips.{l}.weight"))
        if prefix + f"fno_blocks.convs.weight.{l}.tensor" in sd:
            w = torch.view_as_complex(g(f"fno_blocks.convs.weight.{l}.tensor").contiguous())
            spec = spectral_conv_dense(h, w, bias[l], nm, list(w.shape[2:]))
        else:
            core = torch.view_as_complex(g(f"fno_blocks.convs.weight.{l}.core").contiguous())
            facs = [torch.view_as_complex(g(f"fno_blocks.convs.weight.{l}.factors.factor_{k}").contiguous())
                    for k in range(4)]
            spec = spectral_conv_tucker(h, core, facs, bias[l], nm)
        h = spec + skip
        if l < n_layers - 1:
            h = F.gelu(h)
        inter.append(h)
    o = _conv1x1(h, g("projection.fcs.0.weight"), g("projection.fcs.0.bias"))
    o = F.gelu(o)
    o = _conv1x1(o, g("projection.fcs.1.weight"), g("projection.fcs.1.bias"))
    return (o, inter) if return_intermediates else o


def fno_param_count(n_modes, in_channels, hidden, lifting, projection, out_channels, n_layers) -> int:
    """Real-parameter count of a dense neuralop FNO (used for the reference's size labels)."""
    nm = halve_last_mode(n_modes)
    spec = 2 * hidden * hidden * math.prod(nm)
    per_layer = spec + hidden + hidden * hidden              # weight(real view) + bias + 1x1 skip
    lift = in_channels * lifting + lifting + lifting * hidden + hidden
    proj = hidden * projection + projection + projection * out_channels + out_channels
    return n_layers * per_layer + lift + proj


def rollout(sd, x0, n_modes, n_layers, steps: int, prefix: str = ""):
    """Closed-loop rollout x_{t+1} = FNO(x_t) (src/nsbench/models/fno/fno.py:36-41 with
    teacher_forcing_steps=1; requires in_channels == out_channels)."""
    outs = []
    x = x0
    for _ in range(steps):
        x = fno_forward(sd, x, n_modes, n_layers, prefix)
        outs.append(x)
    return torch.stack(outs, dim=1)


def dlwp_rollout(sd, constants, prescribed, prognostic, context_size: int, n_modes, n_layers, prefix: str = ""):
    """``FNO2DModule.forward`` of dlwpbench (src/dlwpbench/models/fno/fno.py:64-106) with the evident intent of its
    two device bugs (``[...].to(device=...)`` on a python list at :91-95 and ``outs.append(out.cpu())`` at :104, which
    make the reference raise for ``T > context_size + 1`` / on CUDA): every frame stays on the input's device.

    constants [B,1,Cc,H,W] or None, prescribed [B,T,Cp,H,W] or None, prognostic [B,T,Cg,H,W]
    -> [B, T - context_size, Cg, H, W];   frame t:  out_t = prognostic_t[:, -1] + FNO(x_t),
    x_t = cat(constants[:, 0], prescribed[:, t-ctx:t] as (t c), prognostic_t as (t c)) on the channel axis,
    prognostic_t = the last ``ctx`` frames of [true frames before ctx ; model outputs so far].
    """
    ctx = int(context_size)
    outs = []
    B, T = prognostic.shape[:2]
    flat = lambda t: t.reshape(t.shape[0], t.shape[1] * t.shape[2], *t.shape[3:])
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
            parts.append(flat(prescribed[:, t - ctx:t]))
        parts.append(flat(prog_t))
        x_t = torch.cat(parts, dim=1)
        outs.append(prog_t[:, -1] + fno_forward(sd, x_t, n_modes, n_layers, prefix))
    return torch.stack(outs, dim=1)
