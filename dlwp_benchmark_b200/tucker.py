"""Spectral weight containers with the parameter names of tltorch's complex tensors, so that
``state_dict`` keys match a neuralop FNO / TFNO:

    dense  : ``...convs.weight.{l}.tensor``                      real view [Cin,Cout,My,Mx,2]
    tucker : ``...convs.weight.{l}.core`` [r0,r1,r2,r3,2] and
             ``...convs.weight.{l}.factors.factor_{k}`` [dim_k, r_k, 2]

The rank rule restates ``tensorly.validate_tucker_rank`` for a float rank (fraction of the
dense parameter count), the init restates ``tltorch.factorized_tensors.init.tucker_init``.
tensorly / tltorch are not installed here and are not vendored by the reference ([RECALLED]).
"""
from __future__ import annotations

import math
from typing import Sequence, Tuple

import torch
import torch.nn as nn


def tucker_rank(shape: Sequence[int], rank) -> Tuple[int, ...]:
    """Solve f^n prod(s) + f sum(s_k^2) = rank prod(s) on [0, max(rank,1)]; r_k = max(round(s_k f), 1)."""
    shape = [int(s) for s in shape]
    if isinstance(rank, (tuple, list)):
        return tuple(int(r) for r in rank)
    if isinstance(rank, int) and not isinstance(rank, bool):
        return tuple(min(rank, s) for s in shape)
    n = len(shape)
    prod = float(math.prod(shape))
    sq = float(sum(s * s for s in shape))
    target = float(rank) * prod
    lo, hi = 0.0, max(float(rank), 1.0)
    for _ in range(200):                       # bisection (monotone polynomial)
        mid = 0.5 * (lo + hi)
        if prod * mid ** n + sq * mid - target > 0:
            hi = mid
        else:
            lo = mid
    f = 0.5 * (lo + hi)
    return tuple(max(int(round(s * f)), 1) for s in shape)


class DenseWeight(nn.Module):
    """Complex dense weight stored as its real view (tltorch ``ComplexDenseTensor``)."""

    def __init__(self, shape, init_std: float):
        super().__init__()
        self.shape = tuple(shape)
        t = torch.empty(*self.shape, dtype=torch.cfloat).normal_(0, init_std)
        self.tensor = nn.Parameter(torch.view_as_real(t).clone())

    def to_dense_real(self) -> torch.Tensor:
        return self.tensor

    def to_dense_complex(self) -> torch.Tensor:
        return torch.view_as_complex(self.tensor)


class FactorList(nn.Module):
    """Parameter container whose entries are named ``factor_{i}`` (tltorch ``FactorList``)."""

    def __init__(self, factors):
        super().__init__()
        self._n = len(factors)
        for i, f in enumerate(factors):
            self.register_parameter(f"factor_{i}", nn.Parameter(f))

    def __len__(self):
        return self._n

    def __getitem__(self, i):
        return getattr(self, f"factor_{i}")

    def __iter__(self):
        return (self[i] for i in range(self._n))


class TuckerWeight(nn.Module):
    """Complex Tucker-factorised weight (tltorch ``ComplexTuckerTensor``), real-view storage."""

    def __init__(self, shape, rank, init_std: float):
        super().__init__()
        self.shape = tuple(shape)
        self.rank = tucker_rank(self.shape, rank)
        r = math.prod(math.sqrt(x) for x in self.rank)
        std_f = (init_std / r) ** (1.0 / (len(self.shape) + 1))
        cplx = lambda *s: torch.view_as_real(torch.empty(*s, dtype=torch.cfloat).normal_(0, std_f)).clone()
        self.core = nn.Parameter(cplx(*self.rank))
        self.factors = FactorList([cplx(s, rk) for s, rk in zip(self.shape, self.rank)])

    def to_dense_complex(self) -> torch.Tensor:
        return torch.view_as_complex(self.to_dense_real())

    def to_dense_real(self) -> torch.Tensor:
        if self.core.is_cuda and len(self.shape) == 4:
            from .tucker_fn import TuckerReconstructFn       # C-ABI kernels, forward and backward
            return TuckerReconstructFn.apply(self.core, *list(self.factors))
        # host-side inspection of a CPU module only (the layers themselves refuse CPU inputs)
        c = torch.view_as_complex(self.core)
        f = [torch.view_as_complex(x) for x in self.factors]
        letters = "ijklmn"[:len(self.shape)]
        ranks = "uvwxyz"[:len(self.shape)]
        eq = ranks + "," + ",".join(a + b for a, b in zip(letters, ranks)) + "->" + letters
        return torch.view_as_real(torch.einsum(eq, c, *f).contiguous())
