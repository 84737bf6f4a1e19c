"""Time the Tucker reconstruction (forward, forward+backward) at the cfg2 weight shape."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dlwp_benchmark_b200.tucker import TuckerWeight
shape = tuple(int(v) for v in sys.argv[1:5]) if len(sys.argv) >= 5 else (64, 64, 16, 9)
tw = TuckerWeight(shape, 0.8, 0.1).cuda()
gW = torch.randn(*shape, 2, device='cuda')
def fwd():
    with torch.no_grad(): return tw.to_dense_real()
def fb():
    W = tw.to_dense_real(); W.backward(gW)
if os.environ.get("KB_ONCE"):
    fb(); fb(); torch.cuda.synchronize(); sys.exit(0)
for name, fn in (("fwd", fwd), ("fwd+bwd", fb)):
    for _ in range(3): fn()
    g = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        fn()
        torch.cuda.synchronize()
        with torch.cuda.graph(g): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): g.replay()
    e1.record(); torch.cuda.synchronize()
    print(f"tucker {shape} rank {tw.rank} {name}: {e0.elapsed_time(e1) / 20 * 1e3:.1f} us (graph replay)")
