import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dlwp_benchmark_b200 import ops, _lib
mode = int(sys.argv[1]) if len(sys.argv) > 1 else 3
_lib.set_tc_mode(mode)
T, C, Hd = 32768, 256, 1024
x = torch.randn(T, C, device="cuda"); W1 = torch.randn(Hd, C, device="cuda") * 0.05; b1 = torch.randn(Hd, device="cuda")
for _ in range(3):
    ops.gemm(x, W1, bias=b1, z_only=True)
torch.cuda.synchronize()
