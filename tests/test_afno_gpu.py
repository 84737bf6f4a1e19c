"""GPU parity tests of AFNO2D: CUDA kernels (through the C ABI) vs the reference's own vectors
(tests/golden/afno2d_*.npz, produced by the reference class) and vs the oracle at cfg4 shapes."""
import os

import numpy as np
import pytest
import torch

import dlwp_benchmark_b200 as pkg
from oracle import afno_oracle as ao
from conftest import rel_l2

pytestmark = pytest.mark.gpu
DEV = "cuda"
TOL = 1e-5
AFNO = ["afno2d_sq8", "afno2d_8x16", "afno2d_32x64", "afno2d_frac", "afno2d_fac2"]


@pytest.mark.parametrize("name", AFNO)
def test_afno_against_reference_vectors(golden_dir, name):
    d = np.load(os.path.join(golden_dir, name + ".npz"))
    B, h, w, C, nb, fac = [int(v) for v in d["meta"]]
    m = pkg.AFNO2D(C, num_blocks=nb, sparsity_threshold=0.01, hard_thresholding_fraction=float(d["frac"]),
                   hidden_size_factor=fac)
    m.load_state_dict({k: torch.tensor(d[k]) for k in ("w1", "b1", "w2", "b2")}, strict=True)
    m = m.to(DEV)
    x = torch.tensor(d["x"]).to(DEV).requires_grad_(True)
    y = m(x)
    y.backward(torch.tensor(d["gy"]).to(DEV))
    assert rel_l2(y, torch.tensor(d["y"])) < TOL
    assert rel_l2(x.grad, torch.tensor(d["gx"])) < TOL
    for k in ("w1", "b1", "w2", "b2"):
        assert rel_l2(getattr(m, k).grad, torch.tensor(d["g" + k])) < 2e-5, k


@pytest.mark.parametrize("B,h,w,C,nb", [(16, 32, 64, 256, 8), (2, 32, 64, 64, 4), (3, 16, 16, 24, 3), (1, 8, 12, 10, 5)])
def test_afno_vs_oracle(B, h, w, C, nb):
    """cfg4 token grid (32x64, embed 256, 8 blocks, batch 16) and ragged shapes (odd C/nb, w%8!=0)."""
    torch.manual_seed(3)
    m = pkg.AFNO2D(C, num_blocks=nb)
    with torch.no_grad():
        for p in m.parameters():
            p.mul_(10.0)            # move activations across the ReLU / softshrink kinks
    x = torch.randn(B, h, w, C)
    gy = torch.randn(B, h, w, C)
    xo = x.double().requires_grad_(True)
    ps = [p.detach().double().requires_grad_(True) for p in (m.w1, m.b1, m.w2, m.b2)]
    yo = ao.afno2d_fft(xo, *ps, nb, 0.01, 1.0)
    yo.backward(gy.double())
    m = m.to(DEV)
    xc = x.to(DEV).requires_grad_(True)
    yc = m(xc)
    yc.backward(gy.to(DEV))
    assert rel_l2(yc, yo) < TOL
    # gradients: some of the ~10^6 pre-activations sit within fp32 rounding of a ReLU / softshrink
    # kink and flip their mask relative to the fp64 oracle (each flip is an O(1) local change), so
    # only a loose bound is meaningful here; the 1e-5 bar on gradients is enforced on the
    # reference's own vectors above and on the kink-free full-size case below.
    assert rel_l2(xc.grad, xo.grad) < 1e-2
    for p, po in zip((m.w1, m.b1, m.w2, m.b2), ps):
        assert rel_l2(p.grad, po.grad) < 1e-2


def test_afno_cfg4_kink_free_gradients():
    """cfg4 shapes with every ReLU active (large b1) and softshrink lambda = 0: the block is smooth,
    so outputs AND all gradients must match the oracle to 1e-5."""
    B, h, w, C, nb = 16, 32, 64, 256, 8
    torch.manual_seed(4)
    m = pkg.AFNO2D(C, num_blocks=nb, sparsity_threshold=0.0)
    with torch.no_grad():
        for p in m.parameters():
            p.mul_(5.0)
        m.b1.add_(60.0)
    x = torch.randn(B, h, w, C)
    gy = torch.randn(B, h, w, C)
    xo = x.double().requires_grad_(True)
    ps = [p.detach().double().requires_grad_(True) for p in (m.w1, m.b1, m.w2, m.b2)]
    yo = ao.afno2d_fft(xo, *ps, nb, 0.0, 1.0)
    yo.backward(gy.double())
    m = m.to(DEV)
    xc = x.to(DEV).requires_grad_(True)
    yc = m(xc)
    yc.backward(gy.to(DEV))
    assert rel_l2(yc, yo) < TOL
    assert rel_l2(xc.grad, xo.grad) < TOL
    for p, po in zip((m.w1, m.b1, m.w2, m.b2), ps):
        assert rel_l2(p.grad, po.grad) < 2e-5


def test_afno_dtype_roundtrip():
    m = pkg.AFNO2D(16, num_blocks=4).to(DEV)
    x = torch.randn(1, 8, 8, 16, device=DEV, dtype=torch.float16)
    assert m(x).dtype == torch.float16


def test_afno_linear_when_mlp_is_dead():
    """Size-independent property: with zero biases and the default 0.02 init every spectral value of a
    small input falls in the softshrink dead zone, so AFNO2D(x) == x exactly (as in the reference)."""
    torch.manual_seed(0)
    m = pkg.AFNO2D(256, num_blocks=8).to(DEV)
    with torch.no_grad():
        m.b2.zero_(); m.b1.zero_()
    x = 1e-3 * torch.randn(2, 32, 64, 256, device=DEV)
    assert torch.equal(m(x), x)


@pytest.mark.parametrize("B,h,w,C,nb,fac,frac", [(2, 16, 32, 64, 2, 1, 1.0), (1, 32, 64, 128, 8, 1, 1.0), (2, 8, 8, 64, 4, 2, 1.0),
                                                  (3, 16, 16, 256, 8, 1, 0.5)])
def test_block_mlp_on_tensor_cores_vs_oracle(B, h, w, C, nb, fac, frac):
    """Blocks of >= 16 channels: the block-diagonal complex MLP runs as batched real-embedded tcgen05 GEMMs
    (sb200_gemm_batched) -- forward, input gradient and all four parameter gradients against the fp64 oracle."""
    from oracle import afno_oracle as ao
    torch.manual_seed(3)
    m = pkg.AFNO2D(C, num_blocks=nb, hard_thresholding_fraction=frac, hidden_size_factor=fac)
    with torch.no_grad():
        for p in m.parameters():
            p.mul_(10.0)
    g = torch.Generator().manual_seed(4)
    x = torch.randn(B, h, w, C, generator=g)
    gy = torch.randn(B, h, w, C, generator=g)
    leaves = [p.detach().double().requires_grad_(True) for p in (m.w1, m.b1, m.w2, m.b2)]
    xo = x.double().requires_grad_(True)
    yo = ao.afno2d_fft(xo, *leaves, nb, 0.01, frac)
    yo.backward(gy.double())
    m = m.to(DEV)
    xd = x.to(DEV).requires_grad_(True)
    y = m(xd)
    y.backward(gy.to(DEV))
    assert rel_l2(y, yo) < 1e-5
    assert rel_l2(xd.grad, xo.grad) < 1e-5
    for p, q in zip((m.w1, m.b1, m.w2, m.b2), leaves):
        assert rel_l2(p.grad, q.grad) < 2e-5
