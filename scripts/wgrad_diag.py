import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dlwp_benchmark_b200 import ops, _lib
lib = _lib.load()
dev = "cuda"
def rel(a, b): return ((a.double().cpu() - b).norm() / b.norm()).item()
for (B, Co, Ci, H, W) in ((64, 64, 64, 64, 64), (8, 64, 64, 64, 64), (64, 256, 64, 64, 64)):
    g = torch.randn(B, Co, H, W, generator=torch.Generator().manual_seed(5))
    x = torch.randn(B, Ci, H, W, generator=torch.Generator().manual_seed(6))
    ref = torch.einsum("bop,bip->oi", g.double().flatten(2), x.double().flatten(2))
    ref32 = torch.einsum("bop,bip->oi", g.flatten(2), x.flatten(2)).double()
    out = {}
    for mode in (0, 3, 1):
        _lib.set_tc_mode(mode)
        gW, gb = ops.pointwise_wgrad(g.to(dev), x.to(dev))
        out[mode] = rel(gW, ref)
    _lib.set_tc_mode(3)
    print((B, Co, Ci, H, W), "ffma", out[0], "3xtf32", out[3], "tf32", out[1], "torch-cpu-fp32", rel(ref32, ref))
