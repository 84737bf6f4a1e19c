"""``AFNO2D`` -- drop-in for the reference's FourCastNet spectral filter
(src/nsbench/models/fourcastnet/fourcastnet.py:59-126 ==
src/dlwpbench/models/fourcastnet/fourcastnet.py:60-127): same constructor, same parameters
(``w1 [2,nb,bs,bs*f]``, ``b1 [2,nb,bs*f]``, ``w2 [2,nb,bs*f,bs]``, ``b2 [2,nb,bs]``, init
0.02*randn), same ``forward(x[B,h,w,C]) -> [B,h,w,C]`` (any input dtype, fp32 inside, cast back).
The arithmetic runs in the channels-last kernels of libspectral_b200.so (afno_kernels.cu).
"""
from __future__ import annotations

import torch
import torch.nn as nn

from . import _lib


class AFNO2D(nn.Module):
    def __init__(self, hidden_size, num_blocks=8, sparsity_threshold=0.01, hard_thresholding_fraction=1,
                 hidden_size_factor=1):
        super().__init__()
        assert hidden_size % num_blocks == 0, \
            f"hidden_size {hidden_size} should be divisble by num_blocks {num_blocks}"
        self.hidden_size = hidden_size
        self.sparsity_threshold = sparsity_threshold
        self.num_blocks = num_blocks
        self.block_size = self.hidden_size // self.num_blocks
        self.hard_thresholding_fraction = hard_thresholding_fraction
        self.hidden_size_factor = hidden_size_factor
        self.scale = 0.02
        bs, f = self.block_size, self.hidden_size_factor
        self.w1 = nn.Parameter(self.scale * torch.randn(2, self.num_blocks, bs, bs * f))
        self.b1 = nn.Parameter(self.scale * torch.randn(2, self.num_blocks, bs * f))
        self.w2 = nn.Parameter(self.scale * torch.randn(2, self.num_blocks, bs * f, bs))
        self.b2 = nn.Parameter(self.scale * torch.randn(2, self.num_blocks, bs))

    def forward(self, x):
        if not x.is_cuda:
            raise _lib.SpectralB200Error("AFNO2D (B200) got a CPU tensor: there is no CPU / torch.fft path")
        from .afno_fn import AFNO2DFn
        dtype = x.dtype
        fp32_in = dtype == torch.float32
        # fp32 input (every shipped config): the residual add is fused into the row-synthesis kernel.  Other dtypes:
        # the reference casts the filter output first and adds the residual in the input dtype
        # (fourcastnet.py:124-126) -- same order here, so the rounding is the reference's.
        y = AFNO2DFn.apply(x.float(), self.w1, self.b1, self.w2, self.b2, self.num_blocks,
                           float(self.sparsity_threshold), float(self.hard_thresholding_fraction), fp32_in)
        return y if fp32_in else y.type(dtype) + x
