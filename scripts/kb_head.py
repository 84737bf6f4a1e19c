"""Time sb200_mlp_head_fwd alone (cfg2 shapes) over > L2 inputs; with a bring-up build SB200_TP_TRACE=n dumps a clock64 trace of call n."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dlwp_benchmark_b200 import ops
B, C, H, W, LP = 64, 64, 64, 64, 256
dev = torch.device('cuda')
xs = [torch.randn(B, C, H, W, device=dev) for _ in range(4)]
Wh = torch.randn(LP, C, device=dev) * 0.1; b1 = torch.randn(LP, device=dev) * 0.1
wo = torch.randn(LP, device=dev) * 0.05; bo = torch.zeros(1, device=dev)
for x in xs: ops.mlp_head_fwd(x, Wh, b1, wo, bo)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for r in range(5):
    for x in xs: ops.mlp_head_fwd(x, Wh, b1, wo, bo)
e1.record(); torch.cuda.synchronize()
print(f"mlp_head_fwd: {e0.elapsed_time(e1) / 20 * 1e3:.1f} us per call")
