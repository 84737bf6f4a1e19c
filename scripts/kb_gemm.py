#!/usr/bin/env python
"""Per-shape timing of the tcgen05 GEMM at the cfg4 (FourCastNet) shapes: T = 32768 tokens, C = 256, hidden 1024."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dlwp_benchmark_b200 import ops, _lib  # noqa: E402

dev = "cuda"
T, C, Hd = 32768, 256, 1024
x, z, gy, gz = (torch.randn(T, n, device=dev) for n in (C, Hd, C, Hd))
W1, W2 = torch.randn(Hd, C, device=dev) * 0.05, torch.randn(C, Hd, device=dev) * 0.05
b1, b2 = torch.randn(Hd, device=dev), torch.randn(C, device=dev)
flush = torch.empty(192 * 1024 * 1024 // 4, device=dev)
cases = {
    "fc1 fwd  z=xW1^T+b (z only)   [T,256]x[1024,256]": (lambda: ops.gemm(x, W1, bias=b1, z_only=True), 2 * T * C * Hd),
    "fc2 fwd  y=gelu(z)W2^T+b+res  [T,1024]x[256,1024]": (lambda: ops.gemm(z, W2, bias=b2, resid=x, a_gelu=True), 2 * T * C * Hd),
    "fc2 dgrad gz=(gy W2)*gelu'(z) [T,256]x[256,1024]mn": (lambda: ops.gemm(gy, W2, b_mn=True, act=2, aux=z), 2 * T * C * Hd),
    "fc1 dgrad gx=gz W1            [T,1024]x[1024,256]mn": (lambda: ops.gemm(gz, W1, b_mn=True), 2 * T * C * Hd),
    "fc2 wgrad gy^T gelu(z) splitK [256,T]x[1024,T]": (lambda: ops.gemm(gy, z, a_mn=True, b_mn=True, b_gelu=True, split_k=True), 2 * T * C * Hd),
    "fc1 wgrad gz^T x splitK       [1024,T]x[256,T]": (lambda: ops.gemm(gz, x, a_mn=True, b_mn=True, split_k=True), 2 * T * C * Hd),
}
lib = _lib.load()
for mode in (3, 1):
    _lib.set_tc_mode(mode)
    for name, (fn, fl) in cases.items():
        for _ in range(2):
            fn()
        ts = []
        for _ in range(5):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); fn(); e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1) * 1e3)
        t = sorted(ts)[len(ts) // 2]
        print(f"mode {mode}  {name:55s} {t:8.1f} us  {fl / t / 1e6:7.1f} TFLOP/s (fp32-equivalent)", flush=True)
_lib.set_tc_mode(3)
a = torch.randn(T, C, device=dev)
g, b = torch.ones(C, device=dev), torch.zeros(C, device=dev)
y, m, r = ops.layernorm_fwd(a, g, b, 1e-6)
for name, fn in (("ln_fwd", lambda: ops.layernorm_fwd(a, g, b, 1e-6)), ("ln_bwd(+dres)", lambda: ops.layernorm_bwd(y, a, g, m, r, dres=y)),
                 ("colsum [T,1024]", lambda: ops.colsum(gz)), ("colsum [T,256]", lambda: ops.colsum(gy))):
    fn()
    flush.zero_()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); fn(); e1.record(); torch.cuda.synchronize()
    print(f"{name:20s} {e0.elapsed_time(e1) * 1e3:8.1f} us")
