"""Per-stage timing table for everything a train step launches (not only the spectral kernels kbench.py covers):
   python scripts/kb_all.py [cfg2|cfg3|cfg1|cfg2d]
Each entry is timed alone with CUDA events on the launching stream, L2 flushed between repetitions, and reported
with its algorithmic bytes (HBM-bound stages) or flops (FFMA-bound stages).  Library tunables are read from the
environment (SB200_*), SB200_TC_MODE selects the tensor-core mode.  Needs a GPU; nothing here is on the product path."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
import dlwp_benchmark_b200 as pkg  # noqa: E402
from dlwp_benchmark_b200 import _lib, ops  # noqa: E402

PEAK_GBS = 6650.0
dev = torch.device("cuda")
flush = torch.empty(192 * 1024 * 1024 // 4, device=dev)


def timed(fn, reps=8, warm=2):
    for _ in range(warm):
        fn()
    ts = []
    for _ in range(reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    ts.sort()
    return ts[len(ts) // 2]


def row(table, name, us, nbytes=None, flops=None):
    r = {"us": round(us, 1)}
    if nbytes:
        r["GB/s"] = round(nbytes / us * 1e-3, 0)
        r["hbm_frac"] = round(nbytes / us * 1e-3 / PEAK_GBS, 3)
    if flops:
        r["TFLOP/s"] = round(flops / us * 1e-6, 2)
    table[name] = r


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
    wl = bench.WORKLOADS[name]
    if os.environ.get("SB200_TC_MODE"):
        _lib.set_tc_mode(int(os.environ["SB200_TC_MODE"]))
    table = {}
    B, C, H, W = wl["batch"], wl["hidden"], wl["H"], wl["W"]
    P = B * C * H * W
    HW = H * W
    # ---- spectral-path kernels (same cases as bench.py's roofline block) ----
    for k, v in bench.kernel_rooflines(wl, PEAK_GBS, reps=8).items():
        row(table, k, v["ms"] * 1e3, nbytes=v["alg_bytes"])
    # ---- lifting MLP 1 -> 256 -> C and projection head C -> 256 -> 1 ----
    x1 = torch.randn(B, 1, H, W, device=dev)
    w1, b1 = torch.randn(256, device=dev) * 0.5, torch.randn(256, device=dev) * 0.1
    W2, b2 = torch.randn(C, 256, device=dev) * 0.05, torch.randn(C, device=dev) * 0.1
    g = torch.randn(B, C, H, W, device=dev)
    h = torch.randn(B, C, H, W, device=dev)
    Wh, bh = torch.randn(256, C, device=dev) * 0.1, torch.randn(256, device=dev) * 0.1
    wo, bo = torch.randn(256, device=dev) * 0.1, torch.randn(1, device=dev)
    gy = torch.randn(B, 1, H, W, device=dev)
    hid = 256 * B * HW
    try:
        if ops.lift_supported(C, 256, HW):
            row(table, "lift_fwd (hidden generated on chip)", timed(lambda: ops.lift_fwd(x1, w1, b1, W2, b2)),
                nbytes=4 * P + 4 * B * HW, flops=2.0 * hid * C)
            row(table, "lift_wgrad (hidden regenerated)", timed(lambda: ops.lift_wgrad(g, x1, w1, b1)),
                nbytes=4 * P + 4 * B * HW, flops=2.0 * hid * C)
            row(table, "lift_tail_bwd", timed(lambda: ops.lift_tail_bwd(g, W2, w1, b1, x1)),
                nbytes=4 * P + 4 * B * HW, flops=2.0 * hid * C)
        if ops.mlp_head_supported(C, 256, 1, HW):
            row(table, "mlp_head_fwd", timed(lambda: ops.mlp_head_fwd(h, Wh, bh, wo, bo)),
                nbytes=4 * P + 4 * B * HW, flops=2.0 * hid * C)
            row(table, "mlp_head_bwd (writes gz1)", timed(lambda: ops.mlp_head_bwd(h, Wh, bh, wo, gy)),
                nbytes=4 * P + 4 * hid, flops=2.0 * hid * C)
            gz1 = torch.randn(B, 256, H, W, device=dev)
            row(table, "head wgrad (tc_wgrad 256xC)", timed(lambda: ops.pointwise_wgrad(gz1, h)),
                nbytes=4 * hid + 4 * P, flops=2.0 * hid * C)
    except Exception as ex:  # noqa: BLE001
        table["lifting/projection"] = {"error": repr(ex)}
    # ---- Tucker reconstruction (all layers per launch) forward and backward ----
    try:
        if wl["rank"] > 0:
            m = bench.build_model(wl).to(dev)
            convs = m.fno_blocks.convs
            row(table, "tucker reconstruct fwd (4 layers)", timed(lambda: convs.dense_weights_all(H, W)))

            def fb():
                ws = convs.dense_weights_all(H, W)
                torch.autograd.backward(ws, [torch.ones_like(w) for w in ws])
            row(table, "tucker reconstruct fwd+bwd (4 layers)", timed(fb))
    except Exception as ex:  # noqa: BLE001
        table["tucker"] = {"error": repr(ex)}
    # ---- one AFNO2D layer at BASELINE configs[3] shapes ----
    try:
        a = bench.WORKLOADS["cfg4"]
        layer = pkg.AFNO2D(a["embed"], num_blocks=a["nb"]).to(dev)
        xa = torch.randn(a["batch"], a["H"], a["W"], a["embed"], device=dev, requires_grad=True)
        Pa = xa.numel()
        with torch.no_grad():
            row(table, "AFNO2D layer fwd", timed(lambda: layer(xa)), nbytes=8 * Pa)

        def afb():
            layer.zero_grad(set_to_none=True)
            xa.grad = None
            layer(xa).sum().backward()
        row(table, "AFNO2D layer fwd+bwd", timed(afb), nbytes=20 * Pa)
    except Exception as ex:  # noqa: BLE001
        table["afno"] = {"error": repr(ex)}
    env = {k: v for k, v in os.environ.items() if k.startswith("SB200_")}
    print(json.dumps({"workload": name, "env": env, "stages": table}, indent=1))


if __name__ == "__main__":
    main()
