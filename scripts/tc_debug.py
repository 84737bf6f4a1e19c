import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dlwp_benchmark_b200 import ops, _lib
from dlwp_benchmark_b200.plan import fno_plan
dev = "cuda"
B, M, N, H, W = 1, 64, 64, 16, 16
plan = fno_plan(dev, H, W, [4, 3])
g = torch.Generator().manual_seed(0)
A = torch.randn(B, M, H, W, generator=g).to(dev)
Wp = torch.randn(N, M, generator=g).to(dev)
ref = torch.einsum("nm,bmhw->bnhw", Wp.double(), A.double())
_lib.load().sb200_set_tc_mode(int(os.environ.get("TCMODE", "1")))
y, _ = ops.rowidft_pointwise(plan, 0, None, A, Wp, M, 1, None, None, B, M, N, 0, False)
torch.cuda.synchronize()
dbg = os.environ.get("SB200_TC_DEBUG", "0")
print("debug", dbg, "mode", os.environ.get("TCMODE", "1"))
if dbg == "1":
    print("A dump max err", (y[:, :M] - A).abs().max().item(), "nonzero frac", (y != 0).float().mean().item())
elif dbg == "2":
    # y[b,n,px] = Wp[n, px % 64]
    px = torch.arange(H * W, device=dev) % 64
    want = Wp[:, px].reshape(1, N, H, W)
    print("B dump max err", (y - want).abs().max().item())
else:
    err = (y.double() - ref.to(dev)).norm() / ref.norm()
    print("rel err", err.item(), "y absmax", y.abs().max().item(), "ref absmax", ref.abs().max().item())
    # correlation diagnostics: is y a permutation / partial sum?
    print("y[0,0,0,:8]", y[0, 0, 0, :8].tolist())
    print("ref[0,0,0,:8]", ref[0, 0, 0, :8].tolist())
