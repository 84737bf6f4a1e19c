"""Time sb200_pointwise_wgrad alone (events around N back-to-back calls on >L2 data)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dlwp_benchmark_b200 import ops, _lib
if os.environ.get('SB200_TC_MODE'):
    _lib.set_tc_mode(int(os.environ['SB200_TC_MODE']))
B, C, H, W = 64, 64, 64, 64
gs = [torch.randn(B, C, H, W, device='cuda') for _ in range(4)]
xs = [torch.randn(B, C, H, W, device='cuda') for _ in range(4)]
for i in range(4): ops.pointwise_wgrad(gs[i], xs[i])
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for r in range(5):
    for i in range(4): ops.pointwise_wgrad(gs[i], xs[i])
e1.record(); torch.cuda.synchronize()
t = e0.elapsed_time(e1) / 20 * 1e3
print({k: v for k, v in os.environ.items() if k.startswith('SB200')}, f"{t:.1f} us per call, {2*B*C*H*W*4/t*1e-3:.0f} GB/s")
