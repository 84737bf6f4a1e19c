"""Generate tests/golden/*.npz by executing the REFERENCE's own AFNO2D class.

Runs only in the build container (needs /root/reference, which does not exist on the GPU
box).  The reference module is imported unmodified from where it lies; two import-time
dependencies that are absent from this image are stubbed (``timm.models.layers`` ->
DropPath = Identity, trunc_normal_ = torch's; ``numpy.lib.arraypad`` was removed in NumPy 2
and is only imported, never used -- reference fourcastnet.py:13).

    python oracle/make_golden.py

FNO / TFNO fixtures (``fno_*.npz``) are produced by the *restatement* in
``spectral_oracle.py`` in float64 -- they pin the CUDA path against the oracle on a box
where /root/reference is absent, but they do NOT pin the oracle to the reference
("parity unpinned", see spectral_oracle.py header).
"""
from __future__ import annotations

import importlib.util
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = "/root/reference"
OUT = os.path.join(ROOT, "tests", "golden")
sys.path.insert(0, ROOT)


def load_reference_fourcastnet(flavour: str = "nsbench"):
    """Import <REF>/src/<flavour>/models/fourcastnet/fourcastnet.py with minimal stubs."""
    if "timm" not in sys.modules:
        timm = types.ModuleType("timm")
        tm = types.ModuleType("timm.models")
        tl = types.ModuleType("timm.models.layers")
        tl.DropPath = lambda *a, **k: torch.nn.Identity()
        tl.trunc_normal_ = torch.nn.init.trunc_normal_
        timm.models = tm
        tm.layers = tl
        sys.modules.update({"timm": timm, "timm.models": tm, "timm.models.layers": tl})
    if "numpy.lib.arraypad" not in sys.modules:
        ap = types.ModuleType("numpy.lib.arraypad")
        ap.pad = np.pad
        sys.modules["numpy.lib.arraypad"] = ap
    if flavour == "dlwpbench":
        for name in ("neuralop", "neuralop.models", "torch_harmonics", "torch_harmonics.examples",
                     "torch_harmonics.examples.sfno"):
            if name not in sys.modules:
                sys.modules[name] = types.ModuleType(name)
        sys.modules["neuralop.models"].FNO = object
        sys.modules["neuralop.models"].TFNO = object
        sys.modules["torch_harmonics.examples.sfno"].SphericalFourierNeuralOperatorNet = object
    path = os.path.join(REF, "src", flavour, "models", "fourcastnet", "fourcastnet.py")
    spec = importlib.util.spec_from_file_location(f"_ref_fourcastnet_{flavour}", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


AFNO_CASES = [
    # name, B, h, w, C, num_blocks, frac, factor
    ("afno2d_sq8", 2, 8, 8, 16, 4, 1.0, 1),          # square grid: Nyquist column retained
    ("afno2d_8x16", 2, 8, 16, 16, 4, 1.0, 1),        # w = 2h: column truncation quirk
    ("afno2d_32x64", 1, 32, 64, 32, 8, 1.0, 1),      # cfg4 token grid, narrow channels
    ("afno2d_frac", 2, 16, 16, 8, 2, 0.5, 1),        # hard_thresholding_fraction < 1
    ("afno2d_fac2", 1, 8, 8, 8, 2, 1.0, 2),          # hidden_size_factor 2
]


def make_afno():
    ref = load_reference_fourcastnet("nsbench")
    for name, B, h, w, C, nb, frac, fac in AFNO_CASES:
        torch.manual_seed(1234)
        m = ref.AFNO2D(C, num_blocks=nb, sparsity_threshold=0.01, hard_thresholding_fraction=frac,
                       hidden_size_factor=fac)
        # the shipped init (0.02*randn) leaves almost everything inside the softshrink dead zone;
        # scale up so the fixture exercises ReLU / softshrink on both sides of their kinks
        with torch.no_grad():
            for p in m.parameters():
                p.mul_(12.0)
        x = torch.randn(B, h, w, C, requires_grad=True)
        gy = torch.randn(B, h, w, C)
        y = m(x)
        y.backward(gy)
        np.savez_compressed(
            os.path.join(OUT, name + ".npz"),
            x=x.detach().numpy(), gy=gy.numpy(), y=y.detach().numpy(), gx=x.grad.numpy(),
            w1=m.w1.detach().numpy(), b1=m.b1.detach().numpy(), w2=m.w2.detach().numpy(),
            b2=m.b2.detach().numpy(), gw1=m.w1.grad.numpy(), gb1=m.b1.grad.numpy(),
            gw2=m.w2.grad.numpy(), gb2=m.b2.grad.numpy(),
            meta=np.array([B, h, w, C, nb, fac], dtype=np.int64), frac=np.array(frac),
        )
        print("wrote", name)


FNO_CASES = [
    # name, B, Cin, hidden, out, H, W, n_modes, layers, lifting/projection, tucker rank (0 = dense)
    ("fno_cfg1_small", 2, 1, 8, 1, 16, 16, (6, 6), 2, 16, 0.0),
    ("fno_rect", 2, 3, 8, 2, 16, 32, (8, 12), 2, 16, 0.0),
    ("fno_fullmodes", 1, 2, 4, 2, 8, 8, (8, 8), 1, 8, 0.0),
    ("tfno_small", 2, 1, 8, 1, 16, 16, (6, 6), 2, 16, 0.8),
]


def make_fno():
    from oracle import spectral_oracle as so
    for name, B, cin, hid, cout, H, W, nm, L, lp, rank in FNO_CASES:
        g = torch.Generator().manual_seed(1234)
        rn = lambda *s: torch.randn(*s, generator=g, dtype=torch.float64)
        half = so.halve_last_mode(nm)
        sd = {
            "lifting.fcs.0.weight": rn(lp, cin, 1, 1) * 0.5, "lifting.fcs.0.bias": rn(lp) * 0.1,
            "lifting.fcs.1.weight": rn(hid, lp, 1, 1) * 0.3, "lifting.fcs.1.bias": rn(hid) * 0.1,
            "projection.fcs.0.weight": rn(lp, hid, 1, 1) * 0.3, "projection.fcs.0.bias": rn(lp) * 0.1,
            "projection.fcs.1.weight": rn(cout, lp, 1, 1) * 0.3, "projection.fcs.1.bias": rn(cout) * 0.1,
            "fno_blocks.convs.bias": rn(L, hid, 1, 1) * 0.2,
        }
        for l in range(L):
            sd[f"fno_blocks.fno_skips.{l}.weight"] = rn(hid, hid, 1, 1) * 0.3
            if rank == 0.0:
                sd[f"fno_blocks.convs.weight.{l}.tensor"] = rn(hid, hid, half[0], half[1], 2) * 0.5
            else:
                r = so.tucker_rank((hid, hid, half[0], half[1]), rank)
                sd[f"fno_blocks.convs.weight.{l}.core"] = rn(*r, 2) * 0.6
                for k, (s, rk) in enumerate(zip((hid, hid, half[0], half[1]), r)):
                    sd[f"fno_blocks.convs.weight.{l}.factors.factor_{k}"] = rn(s, rk, 2) * 0.6
        leaves = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
        x = rn(B, cin, H, W).requires_grad_(True)
        gy = rn(B, cout, H, W)
        y = so.fno_forward(leaves, x, nm, L)
        y.backward(gy)
        out = {"x": x.detach().numpy(), "gy": gy.numpy(), "y": y.detach().numpy(), "gx": x.grad.numpy(),
               "meta": np.array([B, cin, hid, cout, H, W, nm[0], nm[1], L, lp], dtype=np.int64),
               "rank": np.array(rank)}
        for k, v in leaves.items():
            out["p:" + k] = v.detach().numpy()
            out["g:" + k] = v.grad.numpy()
        np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)
        print("wrote", name)


def _dump_module(path, m, inputs, out, gout, extra):
    """inputs: dict name -> tensor (requires_grad where a gradient is wanted); runs backward and writes everything"""
    out.backward(gout)
    d = {"y": out.detach().numpy(), "gy": gout.numpy(), **extra}
    for k, v in inputs.items():
        d["in:" + k] = v.detach().numpy()
        if v.grad is not None:
            d["gin:" + k] = v.grad.numpy()
    for k, v in m.state_dict().items():
        d["p:" + k] = v.detach().numpy()
    for k, p in m.named_parameters():
        if p.grad is not None:
            d["g:" + k] = p.grad.numpy()
    np.savez_compressed(path, **d)


def make_fourcastnet():
    """FourCastNet block remainder from the reference's own classes: one ``Block`` (both flavours share
    it), a depth-2 dlwpbench ``AFNONet`` (single step: the reference loop raises for T > context_size + 1, see
    spectral_oracle.dlwp_rollout) and a depth-2 nsbench ``AFNONet`` with patch (2,2), context 2, teacher forcing."""
    from functools import partial
    ref = load_reference_fourcastnet("dlwpbench")
    torch.manual_seed(1234)
    blk = ref.Block(dim=32, mlp_ratio=4., norm_layer=partial(torch.nn.LayerNorm, eps=1e-6), num_blocks=4)
    with torch.no_grad():
        for n, p in blk.named_parameters():
            if n.startswith("filter."):
                p.mul_(12.0)                       # exercise ReLU / softshrink on both sides of their kinks
            elif "norm" in n:
                p.add_(0.3 * torch.randn_like(p))  # non-trivial LayerNorm affine parameters
            else:
                p.add_(0.1 * torch.randn_like(p))
    x = torch.randn(2, 8, 16, 32, requires_grad=True)
    y = blk(x)
    _dump_module(os.path.join(OUT, "fcn_block.npz"), blk, {"x": x}, y, torch.randn_like(y),
                 {"meta": np.array([2, 8, 16, 32, 4], dtype=np.int64)})
    print("wrote fcn_block")

    torch.manual_seed(4321)
    net = ref.AFNONet(img_height=16, img_width=32, patch_size=(1, 1), constant_channels=4, prescribed_channels=1,
                      prognostic_channels=8, embed_dim=32, depth=2, mlp_ratio=4., num_blocks=4, context_size=1,
                      use_pos_embed=True)
    with torch.no_grad():
        for n, p in net.named_parameters():
            if ".filter." in n:
                p.mul_(12.0)
            elif n.endswith("bias") or "norm" in n:
                p.add_(0.2 * torch.randn_like(p))
            elif "mlp" in n or "head" in n:
                p.mul_(6.0)
    B, T = 2, 2
    c = torch.randn(B, 1, 4, 16, 32)
    pr = torch.randn(B, T, 1, 16, 32)
    pg = torch.randn(B, T, 8, 16, 32)
    y = net(constants=c, prescribed=pr, prognostic=pg)
    _dump_module(os.path.join(OUT, "fcn_dlwp_net.npz"), net, {"constants": c, "prescribed": pr, "prognostic": pg}, y,
                 torch.randn_like(y), {"meta": np.array([16, 32, 1, 1, 4, 1, 8, 32, 2, 4, 1], dtype=np.int64)})
    print("wrote fcn_dlwp_net")

    refn = load_reference_fourcastnet("nsbench")
    torch.manual_seed(999)
    net = refn.AFNONet(img_height=16, img_width=16, patch_size=(2, 2), in_chans=2, out_chans=2, embed_dim=32, depth=2,
                       mlp_ratio=4., num_blocks=4, context_size=2)
    with torch.no_grad():
        for n, p in net.named_parameters():
            if ".filter." in n:
                p.mul_(12.0)
            elif n.endswith("bias") or "norm" in n:
                p.add_(0.2 * torch.randn_like(p))
            elif "mlp" in n or "head" in n:
                p.mul_(6.0)
    x = torch.randn(2, 5, 2, 16, 16)
    y = net(x, teacher_forcing_steps=3)
    _dump_module(os.path.join(OUT, "fcn_ns_net.npz"), net, {"x": x}, y, torch.randn_like(y),
                 {"meta": np.array([16, 16, 2, 2, 2, 2, 32, 2, 4, 2, 3], dtype=np.int64)})
    print("wrote fcn_ns_net")


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    which = sys.argv[1:] or ["afno", "fno", "fcn"]
    if "afno" in which:
        make_afno()
    if "fno" in which:
        make_fno()
    if "fcn" in which:
        make_fourcastnet()
