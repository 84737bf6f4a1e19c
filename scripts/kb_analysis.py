"""Time sb200_analysis (x -> Xh) alone on > L2 data (CUDA-graph replay: no host launch overhead)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dlwp_benchmark_b200 import ops
from dlwp_benchmark_b200.plan import fno_plan
B, C, H, W, n = [int(v) for v in (sys.argv[1:6] if len(sys.argv) >= 6 else (64, 64, 64, 64, 16))]
plan = fno_plan(torch.device('cuda'), H, W, [n, n // 2 + 1])
xs = [torch.randn(B, C, H, W, device='cuda') for _ in range(4)]
def run():
    for x in xs: ops.analysis(plan, 0, x)
run(); torch.cuda.synchronize()
g = torch.cuda.CUDAGraph(); s = torch.cuda.Stream()
with torch.cuda.stream(s):
    run(); torch.cuda.synchronize()
    with torch.cuda.graph(g): run()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for r in range(5): g.replay()
e1.record(); torch.cuda.synchronize()
t = e0.elapsed_time(e1) / 20 * 1e3
print(f"analysis {B}x{C}x{H}x{W} n{n}: {t:.1f} us per call, {B*C*H*W*4/t*1e-3:.0f} GB/s")
