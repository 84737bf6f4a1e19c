#!/usr/bin/env python
"""Small-shape pass through every kernel family of the library, meant to run under compute-sanitizer
(`--tool memcheck` / `--tool racecheck`): whole FNO / TFNO train steps (tcgen05 row synthesis, wgrad, fused head,
generated lifting, fused analysis, two-stage tensor-core analysis, mode GEMMs, Tucker cgemm) and AFNO2D fwd+bwd."""
import os
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import dlwp_benchmark_b200 as pkg  # noqa: E402

dev = "cuda"
torch.manual_seed(0)
cases = [
    ("FNO 32x32 fused analysis", pkg.FNO, dict(n_modes=(8, 8), hidden_channels=32, in_channels=1, out_channels=1,
                                               lifting_channels=256, projection_channels=256, n_layers=2), (2, 1, 32, 32)),
    ("TFNO 64x64 tucker", pkg.TFNO, dict(n_modes=(8, 8), hidden_channels=16, in_channels=1, out_channels=1, rank=0.8,
                                         lifting_channels=256, projection_channels=256, n_layers=2), (1, 1, 64, 64)),
    ("FNO 128x128 two-stage tcgen05 analysis", pkg.FNO, dict(n_modes=(16, 16), hidden_channels=16, in_channels=3,
                                                             out_channels=2, lifting_channels=32, projection_channels=32,
                                                             n_layers=1), (1, 3, 128, 128)),
    ("FNO 12x20 CUDA-core kernels", pkg.FNO, dict(n_modes=(6, 6), hidden_channels=5, in_channels=2, out_channels=2,
                                                  lifting_channels=7, projection_channels=7, n_layers=1), (1, 2, 12, 20)),
]
for name, cls, kw, shape in cases:
    m = cls(**kw).to(dev)
    x = torch.randn(*shape, device=dev)
    y = m(x)
    F.mse_loss(y, torch.randn_like(y)).backward()
    torch.cuda.synchronize()
    print("ok:", name, tuple(y.shape), flush=True)
a = pkg.AFNO2D(32, num_blocks=4).to(dev)
x = torch.randn(1, 8, 16, 32, device=dev, requires_grad=True)
a(x).square().mean().backward()
torch.cuda.synchronize()
print("ok: AFNO2D", flush=True)
