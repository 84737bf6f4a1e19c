"""``FNO`` / ``TFNO`` with the constructor, ``state_dict`` layout and ``forward`` of
``neuralop.models.FNO`` / ``TFNO`` (neuraloperator @05c01c3), i.e. what the reference's wrappers
construct at src/nsbench/models/fno/fno.py:19-27,56-65,205-215,
src/dlwpbench/models/fno/fno.py:38-47,136-146 and
src/dlwpbench/models/fourcastnet/fourcastnet.py:134-143.

state_dict keys:
  lifting.fcs.{0,1}.{weight,bias}            1x1 Conv2d, hidden = lifting_channels
  fno_blocks.convs.weight.{l}.tensor         [C,C,My,Mx,2]   (dense)
  fno_blocks.convs.weight.{l}.core / .factors.factor_{k}     (Tucker)
  fno_blocks.convs.bias                      [L,C,1,1]
  fno_blocks.fno_skips.{l}.weight            [C,C,1,1]
  projection.fcs.{0,1}.{weight,bias}

``FNO.forward`` runs lifting MLP + FNO blocks + projection MLP as ONE autograd node (``FusedChainFn``) whose
forward / backward are sequences of C-ABI kernel launches; a stand-alone ``MLP`` / ``FNOBlocks`` call goes
through the same kernels.  No torch.matmul / cuBLAS / cuFFT sits on any of these paths.
"""
from __future__ import annotations

from functools import partial
from typing import Optional

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import _lib
from .spectral_conv import SpectralConv as _SpectralConv, FNOBlockFn, FNOStackFn
from .chain import FusedChainFn


def _check_grid(x):
    if not x.is_cuda:
        raise _lib.SpectralB200Error("FNO(B200) got a CPU tensor: there is no CPU / torch.fft path")
    if x.dim() != 4 or x.shape[-1] % 4 != 0 or x.shape[-2] % 2 != 0:
        raise _lib.SpectralB200Error(
            f"FNO(B200): input {tuple(x.shape)} -- need [B,C,H,W] with W a multiple of 4 (16-byte vector stores) and "
            "H even (the reference's double-fftshift quirk for odd H is not reproduced); every grid the reference "
            "uses (64, 128, 256, 32x64) qualifies")


class MLP(nn.Module):
    """neuralop.layers.mlp.MLP for n_dim=2: a stack of 1x1 convolutions with a non-linearity between."""

    def __init__(self, in_channels, out_channels=None, hidden_channels=None, n_layers=2, n_dim=2,
                 non_linearity=F.gelu, dropout=0.0, **kwargs):
        super().__init__()
        if n_dim != 2:
            raise NotImplementedError("only n_dim=2")
        self.n_layers = n_layers
        self.in_channels = in_channels
        self.out_channels = in_channels if out_channels is None else out_channels
        self.hidden_channels = in_channels if hidden_channels is None else hidden_channels
        if non_linearity is not F.gelu:
            raise NotImplementedError("MLP(B200): only non_linearity=F.gelu is implemented (fused epilogue)")
        if dropout:
            raise NotImplementedError("MLP(B200): dropout is not implemented (the reference never sets it)")
        self.non_linearity = non_linearity
        self.dropout = None
        self.fcs = nn.ModuleList()
        for i in range(n_layers):
            if i == 0 and i == n_layers - 1:
                self.fcs.append(nn.Conv2d(self.in_channels, self.out_channels, 1))
            elif i == 0:
                self.fcs.append(nn.Conv2d(self.in_channels, self.hidden_channels, 1))
            elif i == n_layers - 1:
                self.fcs.append(nn.Conv2d(self.hidden_channels, self.out_channels, 1))
            else:
                self.fcs.append(nn.Conv2d(self.hidden_channels, self.hidden_channels, 1))

    def forward(self, x):
        """The 1x1-conv stack as a chain of pointwise C-ABI launches (FusedChainFn without spectral branches)."""
        _check_grid(x)
        n = len(self.fcs)
        params = []
        for fc in self.fcs:
            params += [None, fc.weight, fc.bias]
        return FusedChainFn.apply(x, (2, 2), (False,) * n, tuple(i < n - 1 for i in range(n)),
                                  torch.is_grad_enabled(), *params)


class FNOBlocks(nn.Module):
    """neuralop.layers.fno_block.FNOBlocks restricted to the defaults the reference uses:
    post-activation, linear (1x1 conv, no bias) fno_skip, no norm, no MLP, GELU."""

    def __init__(self, in_channels, out_channels, n_modes, output_scaling_factor=None, n_layers=1, max_n_modes=None,
                 fno_block_precision="full", use_mlp=False, mlp_dropout=0, mlp_expansion=0.5, non_linearity=F.gelu,
                 stabilizer=None, norm=None, ada_in_features=None, preactivation=False, fno_skip="linear",
                 mlp_skip="soft-gating", separable=False, factorization=None, rank=1.0, SpectralConv=_SpectralConv,
                 joint_factorization=False, fixed_rank_modes=False, implementation="factorized",
                 decomposition_kwargs=None, fft_norm="forward", **kwargs):
        super().__init__()
        if isinstance(n_modes, int):
            n_modes = [n_modes]
        unsupported = dict(use_mlp=use_mlp, stabilizer=stabilizer, norm=norm, preactivation=preactivation,
                           output_scaling_factor=output_scaling_factor)
        for k, v in unsupported.items():
            if v:
                raise NotImplementedError(f"FNOBlocks(B200): {k}={v!r} is not implemented")
        if fno_skip != "linear":
            raise NotImplementedError("FNOBlocks(B200): only fno_skip='linear' is implemented")
        if non_linearity is not F.gelu:
            raise NotImplementedError("FNOBlocks(B200): only non_linearity=F.gelu is implemented")
        if in_channels != out_channels:
            raise NotImplementedError("FNOBlocks(B200): in_channels must equal out_channels")
        self._n_modes = list(n_modes)
        self.n_dim = len(n_modes)
        self.n_layers = n_layers
        self.in_channels, self.out_channels = in_channels, out_channels
        self.non_linearity = non_linearity
        self.convs = SpectralConv(in_channels, out_channels, self._n_modes, max_n_modes=max_n_modes, rank=rank,
                                  fixed_rank_modes=fixed_rank_modes, implementation=implementation,
                                  separable=separable, factorization=factorization,
                                  decomposition_kwargs=decomposition_kwargs, joint_factorization=joint_factorization,
                                  n_layers=n_layers, fft_norm=fft_norm, fno_block_precision=fno_block_precision)
        self.fno_skips = nn.ModuleList([nn.Conv2d(in_channels, out_channels, kernel_size=1, bias=False)
                                        for _ in range(n_layers)])

    @property
    def n_modes(self):
        return self._n_modes

    def forward(self, x, index=0, output_shape=None):
        """One block (neuralop API): act(convs(x, index) + fno_skips[index](x)), act skipped on the last."""
        if output_shape is not None and tuple(output_shape) != tuple(x.shape[2:]):
            raise NotImplementedError("output_shape resampling is not implemented")
        H, W = x.shape[-2:]
        w = self.convs.dense_weight(index, H, W)
        b = self.convs.bias[index] if self.convs.bias is not None else None
        act = index < self.n_layers - 1
        return FNOBlockFn.apply(x, w, self.fno_skips[index].weight, b, tuple(self.convs.n_modes), act,
                                torch.is_grad_enabled())

    def forward_all(self, x):
        """All blocks through one fused autograd node (what ``FNO.forward`` uses)."""
        if self.convs.bias is None:
            y = x
            for l in range(self.n_layers):
                y = self.forward(y, l)
            return y
        H, W = x.shape[-2:]
        params = []
        dense = self.convs.dense_weights_all(H, W)
        for l in range(self.n_layers):
            params += [dense[l], self.fno_skips[l].weight, self.convs.bias[l]]
        return FNOStackFn.apply(x, tuple(self.convs.n_modes), self.n_layers, torch.is_grad_enabled(), *params)


class FNO(nn.Module):
    """N-D (here: 2-D) Fourier Neural Operator with neuralop's signature; unknown keyword
    arguments are accepted and ignored exactly like neuralop's ``**kwargs``."""

    def __init__(self, n_modes, hidden_channels, in_channels=3, out_channels=1, lifting_channels=256,
                 projection_channels=256, n_layers=4, output_scaling_factor=None, max_n_modes=None,
                 fno_block_precision="full", use_mlp=False, mlp_dropout=0, mlp_expansion=0.5, non_linearity=F.gelu,
                 stabilizer=None, norm=None, preactivation=False, fno_skip="linear", mlp_skip="soft-gating",
                 separable=False, factorization=None, rank=1.0, joint_factorization=False, fixed_rank_modes=False,
                 implementation="factorized", decomposition_kwargs=dict(), domain_padding=None,
                 domain_padding_mode="one-sided", fft_norm="forward", SpectralConv=_SpectralConv, **kwargs):
        super().__init__()
        self.n_dim = len(n_modes)
        if self.n_dim != 2:
            raise NotImplementedError("FNO(B200): only 2-D FNOs are implemented (SURVEY.md row f4 is out of scope)")
        if domain_padding:
            raise NotImplementedError("FNO(B200): domain_padding is not implemented")
        self._n_modes = list(n_modes)
        self.hidden_channels = hidden_channels
        self.lifting_channels = lifting_channels
        self.projection_channels = projection_channels
        self.in_channels, self.out_channels = in_channels, out_channels
        self.n_layers = n_layers
        self.non_linearity = non_linearity
        self.rank = rank
        self.factorization = factorization
        self.fno_blocks = FNOBlocks(
            in_channels=hidden_channels, out_channels=hidden_channels, n_modes=self._n_modes,
            output_scaling_factor=output_scaling_factor, use_mlp=use_mlp, mlp_dropout=mlp_dropout,
            mlp_expansion=mlp_expansion, non_linearity=non_linearity, stabilizer=stabilizer, norm=norm,
            preactivation=preactivation, fno_skip=fno_skip, mlp_skip=mlp_skip, max_n_modes=max_n_modes,
            fno_block_precision=fno_block_precision, rank=rank, fft_norm=fft_norm,
            fixed_rank_modes=fixed_rank_modes, implementation=implementation, separable=separable,
            factorization=factorization, decomposition_kwargs=decomposition_kwargs,
            joint_factorization=joint_factorization, SpectralConv=SpectralConv, n_layers=n_layers)
        if lifting_channels:
            self.lifting = MLP(in_channels, hidden_channels, hidden_channels=lifting_channels, n_layers=2, n_dim=2)
        else:
            self.lifting = MLP(in_channels, hidden_channels, hidden_channels=hidden_channels, n_layers=1, n_dim=2)
        self.projection = MLP(hidden_channels, out_channels, hidden_channels=projection_channels, n_layers=2,
                              n_dim=2, non_linearity=non_linearity)
        self.fused = True   # route the whole model through FusedChainFn (False: per-stage autograd nodes)

    @property
    def n_modes(self):
        return self._n_modes

    def forward(self, x, output_shape=None, **kwargs):
        if output_shape is not None:
            raise NotImplementedError("FNO(B200): output_shape is only used by the 3-D wrappers (out of scope)")
        _check_grid(x)
        if self.fused and self.fno_blocks.convs.bias is not None:
            return self._forward_fused(x)
        x = self.lifting(x)
        x = self.fno_blocks.forward_all(x)
        x = self.projection(x)
        return x

    def _forward_fused(self, x):
        """lifting + blocks + projection as one chain of C-ABI kernel launches (FusedChainFn)."""
        H, W = x.shape[-2:]
        spec, acts, params = [], [], []
        nl_lift = len(self.lifting.fcs)
        for i, fc in enumerate(self.lifting.fcs):
            spec.append(False); acts.append(i < nl_lift - 1)
            params += [None, fc.weight, fc.bias]
        blocks = self.fno_blocks
        dense = blocks.convs.dense_weights_all(H, W)
        bias = blocks.convs.bias
        bsink = getattr(bias, "_sb200_grad_sink", None)        # ddp.GradSync(direct=True)
        for l in range(self.n_layers):
            spec.append(True); acts.append(l < self.n_layers - 1)
            b_l = bias[l]
            if bsink is not None:                              # layer l's slice of the bias gradient destination
                b_l._sb200_grad_sink = bsink.view(self.n_layers, -1)[l]
                b_l._sb200_sink_owner = bias
            params += [dense[l], blocks.fno_skips[l].weight, b_l]
        nl_proj = len(self.projection.fcs)
        for i, fc in enumerate(self.projection.fcs):
            spec.append(False); acts.append(i < nl_proj - 1)
            params += [None, fc.weight, fc.bias]
        return FusedChainFn.apply(x, tuple(blocks.convs.n_modes), tuple(spec), tuple(acts), torch.is_grad_enabled(),
                                  *params)


class TFNO(FNO):
    """``partialclass('TFNO', FNO, factorization='Tucker')`` in neuralop."""

    def __init__(self, *args, **kwargs):
        kwargs.setdefault("factorization", "Tucker")
        super().__init__(*args, **kwargs)
