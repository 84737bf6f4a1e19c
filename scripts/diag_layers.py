#!/usr/bin/env python
"""Layer-by-layer error of the CUDA path vs the fp64 oracle (and of the oracle in fp32 vs fp64, i.e. what plain
fp32 PyTorch would do) on a whole model:  python scripts/diag_layers.py cfg5 2 [tc_mode]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from dlwp_benchmark_b200 import _lib  # noqa: E402
from oracle import spectral_oracle as so  # noqa: E402


def rel(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return (torch.linalg.norm(a - b) / torch.linalg.norm(b)).item()


name, B = sys.argv[1], int(sys.argv[2])
mode = int(sys.argv[3]) if len(sys.argv) > 3 else 3
wl = bench.WORKLOADS[name]
m = bench.build_model(wl)
sd = {k: v.detach().clone() for k, v in m.state_dict().items()}
x = torch.randn(B, 1, wl["H"], wl["W"], generator=torch.Generator().manual_seed(41))
with torch.no_grad():
    y64, i64 = so.fno_forward({k: v.double() for k, v in sd.items()}, x.double(), wl["n_modes"], wl["L"], return_intermediates=True)
    y32, i32 = so.fno_forward(sd, x, wl["n_modes"], wl["L"], return_intermediates=True)
    _lib.set_tc_mode(mode)
    m = m.cuda()
    xd = x.cuda()
    h = m.lifting(xd)
    got = [h]
    for l in range(wl["L"]):
        h = m.fno_blocks(h, l)
        got.append(h)
    out = m.projection(h)
    fused = m(xd)
    # each CUDA stage fed with the ORACLE's input of that stage (isolates the stage's own error)
    iso = [rel(m.lifting(xd), i64[0])]
    for l in range(wl["L"]):
        iso.append(rel(m.fno_blocks(i64[l].float().cuda(), l), i64[l + 1]))
    iso.append(rel(m.projection(i64[-1].float().cuda()), y64))
print(f"{name} B={B} tc_mode={mode}")
print("stage            cuda(chained)  cuda(isolated)  torch-fp32(chained)   rms(mean)")
names = ["lifting"] + [f"block{l}" for l in range(wl["L"])] + ["projection"]
for i, n in enumerate(names):
    ref = i64[i] if i < len(i64) else y64
    g = got[i] if i < len(got) else out
    t = i32[i] if i < len(i32) else y32
    print(f"{n:14s} {rel(g, ref):12.3e} {iso[i]:14.3e} {rel(t, ref):18.3e}   {ref.pow(2).mean().sqrt().item():.4f} ({ref.mean().item():.4f})")
print("fused forward vs oracle:", rel(fused, y64), " vs chained:", rel(fused, out))
