"""GPU parity of the FourCastNet block remainder: Block / AFNONet (both flavours) through the C-ABI
kernels vs (a) vectors produced by the reference's own classes (tests/golden/fcn_*.npz, oracle/make_golden.py) and
(b) the oracle restatement at the cfg4 width.  rel-L2 <= 1e-5 on outputs, 2e-5 on gradients (fp32 vectors)."""
import os

import numpy as np
import pytest
import torch

import dlwp_benchmark_b200 as pkg
from dlwp_benchmark_b200 import fourcastnet as fcn
from oracle import afno_oracle as ao
from conftest import rel_l2

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _load(golden_dir, name):
    d = np.load(os.path.join(golden_dir, name + ".npz"))
    sd = {k[2:]: torch.tensor(d[k]) for k in d.files if k.startswith("p:")}
    return d, sd


def _check_param_grads(m, d, tol=2e-5):
    for k, p in m.named_parameters():
        if "g:" + k in d.files:
            assert p.grad is not None, k
            e = rel_l2(p.grad, torch.tensor(d["g:" + k]))
            assert e < tol, (k, e)
        else:
            assert p.grad is None or float(p.grad.abs().max()) == 0.0, k


def test_block_against_reference_vectors(golden_dir):
    d, sd = _load(golden_dir, "fcn_block")
    from functools import partial
    blk = fcn.Block(dim=32, mlp_ratio=4., norm_layer=partial(torch.nn.LayerNorm, eps=1e-6), num_blocks=4)
    blk.load_state_dict(sd, strict=True)
    blk = blk.to(DEV)
    x = torch.tensor(d["in:x"]).to(DEV).requires_grad_(True)
    y = blk(x)
    y.backward(torch.tensor(d["gy"]).to(DEV))
    assert rel_l2(y, torch.tensor(d["y"])) < 1e-5
    assert rel_l2(x.grad, torch.tensor(d["gin:x"])) < 1e-5
    _check_param_grads(blk, d)


def test_dlwpbench_afnonet_against_reference_vectors(golden_dir):
    d, sd = _load(golden_dir, "fcn_dlwp_net")
    H, W, p1, p2, cc, cp, cg, E, depth, nb, ctx = [int(v) for v in d["meta"]]
    net = fcn.AFNONet(img_height=H, img_width=W, patch_size=(p1, p2), constant_channels=cc, prescribed_channels=cp,
                      prognostic_channels=cg, embed_dim=E, depth=depth, mlp_ratio=4., num_blocks=nb, context_size=ctx)
    net.load_state_dict(sd, strict=True)
    net = net.to(DEV)
    c, pr, pg = (torch.tensor(d["in:" + k]).to(DEV) for k in ("constants", "prescribed", "prognostic"))
    y = net(constants=c, prescribed=pr, prognostic=pg)
    y.backward(torch.tensor(d["gy"]).to(DEV))
    assert rel_l2(y, torch.tensor(d["y"])) < 1e-5
    _check_param_grads(net, d)


def test_nsbench_afnonet_against_reference_vectors(golden_dir):
    d, sd = _load(golden_dir, "fcn_ns_net")
    H, W, p1, p2, ci, co, E, depth, nb, ctx, tf = [int(v) for v in d["meta"]]
    net = fcn.AFNONetNS(img_height=H, img_width=W, patch_size=(p1, p2), in_chans=ci, out_chans=co, embed_dim=E, depth=depth,
                        mlp_ratio=4., num_blocks=nb, context_size=ctx)
    net.load_state_dict(sd, strict=True)
    net = net.to(DEV)
    x = torch.tensor(d["in:x"]).to(DEV)
    y = net(x, teacher_forcing_steps=tf)
    y.backward(torch.tensor(d["gy"]).to(DEV))
    assert rel_l2(y, torch.tensor(d["y"])) < 1e-5
    _check_param_grads(net, d, tol=3e-5)


def test_cfg4_width_vs_oracle():
    """dlwpbench AFNONet at the cfg4 width (32x64 tokens, patch 1x1, 13 in / 8 out, embed 256, num_blocks 8),
    depth 2, batch 2: network step fwd + all gradients against the fp64 oracle restatement."""
    torch.manual_seed(11)
    net = fcn.AFNONet(img_height=32, img_width=64, patch_size=(1, 1), constant_channels=4, prescribed_channels=1,
                      prognostic_channels=8, embed_dim=256, depth=2, mlp_ratio=4., num_blocks=8, context_size=1)
    with torch.no_grad():
        for n, p in net.named_parameters():
            if ".filter." in n:
                p.mul_(10.0)
    sd = {k: v.detach().clone() for k, v in net.state_dict().items()}
    g = torch.Generator().manual_seed(5)
    x_t = torch.randn(2, 13, 32, 64, generator=g)
    gy = torch.randn(2, 8, 32, 64, generator=g)
    leaves = {k: v.double().requires_grad_(True) for k, v in sd.items()}
    yo = ao.afnonet_step(leaves, x_t.double(), (1, 1), 2, 8)
    yo.backward(gy.double())
    net = net.to(DEV)
    y = net.step(x_t.to(DEV))
    y.backward(gy.to(DEV))
    assert rel_l2(y, yo) < 1e-5
    for k, p in net.named_parameters():
        if leaves[k].grad is not None and k != "norm.weight" and k != "norm.bias":
            e = rel_l2(p.grad, leaves[k].grad)
            assert e < 2e-5, (k, e)


def test_state_dict_keys_equal_the_reference_layout(golden_dir):
    d, sd = _load(golden_dir, "fcn_dlwp_net")
    net = fcn.AFNONet(img_height=16, img_width=32, patch_size=(1, 1), constant_channels=4, prescribed_channels=1,
                      prognostic_channels=8, embed_dim=32, depth=2, mlp_ratio=4., num_blocks=4, context_size=1)
    assert set(net.state_dict().keys()) == set(sd.keys())
    for k, v in net.state_dict().items():
        assert tuple(v.shape) == tuple(sd[k].shape), k
