import os, sys, ctypes, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dlwp_benchmark_b200 import _lib
lib = _lib.load()
vp = ctypes.c_void_p
dev = "cuda"
for (N, K) in ((64, 8), (64, 40), (16, 32), (256, 64)):
    g = torch.Generator().manual_seed(N + K)
    A = torch.randn(128, K, generator=g).to(dev)
    B = torch.randn(N, K, generator=g).to(dev)
    ref = A.double() @ B.double().t()
    for a_layout in (0, 1, 3, 16 + 0, 16 + 3, 16 + 1):
        for use_mask in (0,):
            D = torch.full((128, N), 777.0, device=dev)
            info = torch.zeros(8, dtype=torch.int32, device=dev)
            rc = lib.sb200_tc_selftest(vp(A.data_ptr()), vp(B.data_ptr()), vp(D.data_ptr()), N, K, a_layout, use_mask,
                                       vp(info.data_ptr()), vp(torch.cuda.current_stream().cuda_stream))
            torch.cuda.synchronize()
            err = ((D.double() - ref).norm() / ref.norm()).item()
            print(f"N={N} K={K} a_layout={a_layout} mask={use_mask} rc={rc} rel_err={err:.3e} absmax={D.abs().max().item():.3f} "
                  f"info={[hex(v & 0xffffffff) for v in info.tolist()[:6]]}")
