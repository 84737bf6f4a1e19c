#!/usr/bin/env python
"""Block-diagonal complex MLP of AFNO2D at cfg4 shapes (16 x 32x17 retained modes = 8704 tokens, C 256, 8 blocks of 32):
batched tcgen05 GEMMs (sb200_gemm_batched) vs the FFMA blocklinear kernels."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dlwp_benchmark_b200 import afno_fn as af, _lib
dev = "cuda"
B, My, Mx, C, nb = 16, 32, 17, 256, 8
bs = bsf = C // nb
ntok = B * My * Mx
Xh = torch.randn(B, My, Mx, C, 2, device=dev)
w1 = torch.randn(2, nb, bs, bsf, device=dev) * 0.2
b1 = torch.randn(2, nb, bsf, device=dev) * 0.1
E1 = af._embed(w1, nb, bs, bsf)
be1 = b1.permute(1, 2, 0).contiguous().view(-1)
O1 = torch.empty(B, My, Mx, C, 2, device=dev)
gE = torch.empty(nb, 2 * bsf, 2 * bs, device=dev)
flush = torch.empty(192 * 1024 * 1024 // 4, device=dev)
cases = {
    "tc fwd (relu)": lambda: af._gemm_batched(Xh, E1, O1, M=ntok, N=2 * bsf, K=2 * bs, nbatch=nb, lda=2 * C, ldb=2 * bs, ldd=2 * C,
                                              a_ext=(2 * C, ntok), b_ext=(2 * bs, nb * 2 * bsf), a_off=2 * bs, b_off1=2 * bsf, d_off=2 * bsf,
                                              bias=be1, bias_off=2 * bsf, act=3),
    "tc dgrad (mask)": lambda: af._gemm_batched(Xh, E1, O1, M=ntok, N=2 * bs, K=2 * bsf, nbatch=nb, lda=2 * C, ldb=2 * bs, ldd=2 * C, b_mn=True,
                                                a_ext=(2 * C, ntok), b_ext=(2 * bs, nb * 2 * bsf), a_off=2 * bsf, b_off1=2 * bsf, d_off=2 * bs,
                                                aux=Xh, act=5),
    "tc wgrad (split-K)": lambda: af._gemm_batched(O1, Xh, gE, M=2 * bsf, N=2 * bs, K=ntok, nbatch=nb, lda=2 * C, ldb=2 * C, ldd=2 * bs, a_mn=True,
                                                   b_mn=True, a_ext=(2 * C, ntok), b_ext=(2 * C, ntok), a_off=2 * bsf, b_off0=2 * bs,
                                                   d_off=2 * bsf * 2 * bs, split_k=True),
    "ffma fwd": lambda: af._bl_fwd(Xh, w1, b1, ntok, nb, bs, bsf, 1, 0.0),
    "ffma dgrad": lambda: af._bl_dgrad(Xh, O1, 1, w1, ntok, nb, bs, bsf),
    "ffma wgrad": lambda: af._bl_wgrad(Xh, O1, O1, 1, ntok, nb, bs, bsf),
    "embed": lambda: af._embed(w1, nb, bs, bsf),
    "unembed": lambda: af._unembed(gE, nb, bs, bsf),
}
for name, fn in cases.items():
    for _ in range(2):
        fn()
    ts = []
    for _ in range(5):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    print(f"{name:22s} {sorted(ts)[2]:8.1f} us", flush=True)
