"""One FNO block (cfg2 shapes) forward + backward, for ncu captures:  python scripts/run_block.py [B C H W n]"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dlwp_benchmark_b200.spectral_conv import FNOBlockFn
B, C, H, W, n = [int(v) for v in (sys.argv[1:6] if len(sys.argv) >= 6 else (64, 64, 64, 64, 16))]
dev = "cuda"
torch.manual_seed(0)
x = torch.randn(B, C, H, W, device=dev, requires_grad=True)
w = (0.1 * torch.randn(C, C, n, n // 2 + 1, 2, device=dev)).requires_grad_(True)
ws = (0.1 * torch.randn(C, C, 1, 1, device=dev)).requires_grad_(True)
b = torch.randn(C, 1, 1, device=dev, requires_grad=True)
for _ in range(3):
    y = FNOBlockFn.apply(x, w, ws, b, (n, n // 2 + 1), True)
    y.backward(torch.ones_like(y))
torch.cuda.synchronize()
print("ok")
