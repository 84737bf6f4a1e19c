"""GPU parity tests of the FNO / TFNO path: CUDA kernels (through the C ABI) vs the oracle.

Tolerance (north_star): relative L2 <= 1e-5 for outputs and every gradient in the fp32 path.
"""
import math
import os

import numpy as np
import pytest
import torch

import dlwp_benchmark_b200 as pkg
from dlwp_benchmark_b200 import ops
from dlwp_benchmark_b200.plan import fno_plan
from dlwp_benchmark_b200.spectral_conv import FNOBlockFn
from oracle import spectral_oracle as so
from conftest import rel_l2

pytestmark = pytest.mark.gpu
TOL = 1e-5
DEV = "cuda"


def _rand(*s, seed=0, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(*s, generator=g) * scale


STAGE_SHAPES = [
    # B, C, H, W, n_modes
    (2, 3, 16, 16, (6, 6)),
    (1, 5, 16, 32, (8, 12)),
    (2, 4, 8, 8, (8, 8)),        # all modes incl. Nyquist column
    (3, 7, 64, 64, (12, 12)),    # cfg1 grid
    (2, 8, 32, 64, (12, 12)),    # dlwpbench grid
    (1, 2, 64, 128, (32, 32)),
    (1, 2, 12, 20, (6, 7)),      # W % 8 != 0, odd mode counts
    (1, 2, 256, 256, (32, 32)),  # cfg3 grid
    (1, 1, 128, 256, (16, 16)),
    (1, 1, 256, 64, (32, 32)),
    (2, 1, 128, 128, (32, 32)),  # cfg5 grid
]


@pytest.mark.parametrize("B,C,H,W,nm", STAGE_SHAPES)
def test_analysis_stages(B, C, H, W, nm):
    half = so.halve_last_mode(nm)
    plan = fno_plan(DEV, H, W, half)
    x = _rand(B, C, H, W)
    xd = x.double()
    full = torch.fft.fftshift(torch.fft.rfftn(xd, norm="forward", dim=(-2, -1)), dim=(-2,))
    lo, My = so.retained_rows(H, half[0])
    ref = full[:, :, lo:lo + My, :plan.Mx]
    T = ops.rowdft_fwd(plan, 0, x.to(DEV))
    refT = torch.fft.rfft(xd, dim=-1)[..., :plan.Mx]
    assert rel_l2(torch.view_as_complex(T), refT) < TOL
    Xh = ops.coldft_fwd(plan, 0, T)
    assert rel_l2(torch.view_as_complex(Xh), ref) < TOL


TC_ROWDFT_SHAPES = [
    # B, C, H, W, n_modes        (rows = B*C*H; tiles of 128 rows)
    (1, 3, 50, 64, (12, 12)),    # 150 rows: partial last tile, N = 16
    (2, 5, 128, 128, (32, 32)),  # cfg5 grid, N = 48 (Mx = 17), 10 tiles
    (1, 9, 256, 256, (32, 32)),  # cfg3 grid: 8 K chunks through a 3-stage ring
    (3, 2, 64, 256, (16, 16)),   # N = 32 (Mx = 9)
    (1, 1, 8, 256, (8, 60)),     # fewer rows than one tile, N = 64 (Mx = 31): 128 KB of twiddles, 2-stage ring
]


@pytest.mark.parametrize("mode,tol", [(3, 1e-5), (1, 3e-3)])
@pytest.mark.parametrize("pas", [0, 1])
@pytest.mark.parametrize("B,C,H,W,nm", TC_ROWDFT_SHAPES)
def test_tc_rowdft(B, C, H, W, nm, pas, mode, tol):
    """tcgen05 row DFT (tc_rowdft.cu) against torch.fft (pass 0) and against the exact-fp32 FFMA kernel (both passes)."""
    half = so.halve_last_mode(nm)
    plan = fno_plan(DEV, H, W, half)
    x = _rand(B, C, H, W, seed=21)
    xg = x.to(DEV)
    try:
        _set_tc(0)
        exact = ops.rowdft_fwd(plan, pas, xg)
        _set_tc(mode)
        T = ops.rowdft_fwd(plan, pas, xg)
    finally:
        _set_tc(3)
    assert rel_l2(T, exact) < tol
    if pas == 0:
        refT = torch.fft.rfft(x.double(), dim=-1)[..., :plan.Mx]
        assert rel_l2(torch.view_as_complex(T), refT) < tol


ANALYSIS_SHAPES = STAGE_SHAPES + [
    (5, 3, 64, 64, (16, 16)),    # fused kernel, 15 images: partial last group (G = 4)
    (64, 8, 64, 64, (16, 16)),   # fused kernel, more groups than one wave per SM pair
    (3, 5, 32, 64, (24, 24)),    # fused, G = 8, Mx = 13
    (2, 3, 32, 32, (32, 32)),    # fused, W = 32, Mx = 17 (Nyquist column retained)
    (1, 5, 128, 64, (32, 8)),    # fused, G = 2, Mx = 5, My = 32
    (7, 1, 16, 32, (12, 12)),    # fused, G = 16, My = 12, Mx = 7
    (1, 3, 64, 64, (20, 12)),    # fused, My = 20
    (1, 3, 64, 64, (18, 16)),    # fused, My not a multiple of 4
    (2, 4, 64, 64, (10, 12)),    # fused, first retained frequency odd (ky0 = -5): the radix-2 column fold flips its sign
    (2, 4, 32, 32, (6, 8)),      # fused, W = 32 (two folded row chunks), ky0 = -3
    (1, 2, 64, 32, (14, 16)),    # fused, W = 32, H = 64, Mx = 9, ky0 = -7
]


@pytest.mark.parametrize("pas", [0, 1])
@pytest.mark.parametrize("B,C,H,W,nm", ANALYSIS_SHAPES)
def test_analysis_single_call(B, C, H, W, nm, pas):
    """sb200_analysis (one fused kernel on small grids) == row stage + column stage == torch.fft."""
    half = so.halve_last_mode(nm)
    plan = fno_plan(DEV, H, W, half)
    x = _rand(B, C, H, W, seed=11)
    Xh = ops.analysis(plan, pas, x.to(DEV))
    two = ops.coldft_fwd(plan, pas, ops.rowdft_fwd(plan, pas, x.to(DEV)))
    assert rel_l2(Xh, two) < TOL
    if pas == 0:
        full = torch.fft.fftshift(torch.fft.rfftn(x.double(), norm="forward", dim=(-2, -1)), dim=(-2,))
        lo, My = so.retained_rows(H, half[0])
        assert rel_l2(torch.view_as_complex(Xh), full[:, :, lo:lo + My, :plan.Mx]) < TOL


@pytest.mark.parametrize("B,C,H,W,nm", STAGE_SHAPES)
def test_synthesis_stages(B, C, H, W, nm):
    half = so.halve_last_mode(nm)
    plan = fno_plan(DEV, H, W, half)
    lo, My = so.retained_rows(H, half[0])
    Yh = torch.view_as_complex(_rand(B, C, My, plan.Mx, 2, seed=3).double())
    full = torch.zeros(B, C, H, W // 2 + 1, dtype=torch.complex128)
    full[:, :, lo:lo + My, :plan.Mx] = Yh
    ref = torch.fft.irfftn(torch.fft.fftshift(full, dim=(-2,)), s=(H, W), dim=(-2, -1), norm="forward")
    Phi = ops.coldft_inv(plan, 0, torch.view_as_real(Yh).float().contiguous().to(DEV))
    y, _ = ops.rowidft_pointwise(plan, 0, Phi, None, None, 0, 0, None, None, B, C, C, 0, False)
    assert rel_l2(y, ref) < TOL


@pytest.mark.parametrize("P,Q,R,K", [(5, 7, 3, 10), (64, 64, 64, 144), (32, 32, 32, 84), (8, 33, 65, 7)])
def test_modes_gemm_all_three_uses(P, Q, R, K):
    A = torch.view_as_complex(_rand(P, R, K, 2, seed=1).double())     # Xh [b,i,k]
    Wc = torch.view_as_complex(_rand(R, Q, K, 2, seed=2).double())    # W  [i,o,k]
    G = torch.view_as_complex(_rand(P, Q, K, 2, seed=3).double())     # gYh [b,o,k]
    f = lambda t: torch.view_as_real(t).float().contiguous().to(DEV).reshape(*t.shape[:2], K, 1, 2)
    Yh = ops.mix_fwd(f(A), f(Wc))
    assert rel_l2(torch.view_as_complex(Yh).squeeze(-1), torch.einsum("bik,iok->bok", A, Wc)) < TOL
    gX = ops.mix_bwd_input(f(G), f(Wc))
    assert rel_l2(torch.view_as_complex(gX).squeeze(-1), torch.einsum("iok,bok->bik", Wc.conj(), G)) < TOL
    gW = ops.mix_bwd_weight(f(A), f(G))
    assert rel_l2(torch.view_as_complex(gW).squeeze(-1), torch.einsum("bik,bok->iok", A.conj(), G)) < TOL


@pytest.mark.parametrize("B,Co,Ci,H,W", [(2, 5, 3, 8, 8), (3, 64, 64, 16, 16), (2, 70, 33, 16, 32), (1, 8, 8, 64, 64)])
def test_pointwise_wgrad(B, Co, Ci, H, W):
    g = _rand(B, Co, H, W, seed=5)
    x = _rand(B, Ci, H, W, seed=6)
    gW, gb = ops.pointwise_wgrad(g.to(DEV), x.to(DEV))
    assert rel_l2(gW, torch.einsum("bop,bip->oi", g.double().flatten(2), x.double().flatten(2))) < TOL
    assert rel_l2(gb, g.double().sum(dim=(0, 2, 3))) < TOL


def test_gelu_helpers():
    z = _rand(1000, seed=9, scale=2.0)
    zd = z.double().requires_grad_(True)
    yd = torch.nn.functional.gelu(zd)
    yd.backward(torch.ones_like(yd))
    assert rel_l2(ops.gelu_fwd(z.to(DEV)), yd) < 1e-6
    assert rel_l2(ops.gelu_bwd(torch.ones(1000, device=DEV), z.to(DEV)), zd.grad) < 1e-6


BLOCK_SHAPES = [
    # B, Cin, Cout, H, W, n_modes, act, skip
    (2, 3, 3, 16, 16, (6, 6), True, True),
    (2, 4, 6, 16, 32, (8, 12), False, False),    # pure SpectralConv, Cin != Cout
    (1, 8, 8, 64, 64, (12, 12), True, True),
    (2, 5, 5, 8, 8, (8, 8), True, True),
    (1, 70, 70, 16, 16, (6, 6), True, True),     # more than one channel tile
]


@pytest.mark.parametrize("B,Ci,Co,H,W,nm,act,skip", BLOCK_SHAPES)
def test_block_forward_backward(B, Ci, Co, H, W, nm, act, skip):
    half = so.halve_last_mode(nm)
    lo, My = so.retained_rows(H, half[0])
    Mx = min(half[1], W // 2 + 1)
    x = _rand(B, Ci, H, W, seed=1)
    w = _rand(Ci, Co, My, Mx, 2, seed=2, scale=0.5)
    ws = _rand(Co, Ci, 1, 1, seed=3, scale=0.4) if skip else None
    b = _rand(Co, 1, 1, seed=4, scale=0.3)
    gy = _rand(B, Co, H, W, seed=5)
    # oracle, fp64
    xo, wo, bo = x.double().requires_grad_(True), w.double().requires_grad_(True), b.double().requires_grad_(True)
    wso = ws.double().requires_grad_(True) if skip else None
    yo = so.spectral_conv_dense(xo, torch.view_as_complex(wo), bo, [My, Mx])
    if skip:
        yo = yo + torch.nn.functional.conv2d(xo, wso)
    if act:
        yo = torch.nn.functional.gelu(yo)
    yo.backward(gy.double())
    # CUDA
    xc, wc, bc = (t.to(DEV).requires_grad_(True) for t in (x, w, b))
    wsc = ws.to(DEV).requires_grad_(True) if skip else None
    yc = FNOBlockFn.apply(xc, wc, wsc, bc, tuple(half), act)
    yc.backward(gy.to(DEV))
    assert rel_l2(yc, yo) < TOL
    assert rel_l2(xc.grad, xo.grad) < TOL
    assert rel_l2(wc.grad, wo.grad) < TOL
    assert rel_l2(bc.grad, bo.grad) < TOL
    if skip:
        assert rel_l2(wsc.grad, wso.grad) < TOL


def _load_fixture(golden_dir, name):
    d = np.load(os.path.join(golden_dir, name + ".npz"))
    B, cin, hid, cout, H, W, n0, n1, L, lp = [int(v) for v in d["meta"]]
    rank = float(d["rank"])
    cls = pkg.TFNO if rank > 0 else pkg.FNO
    m = cls(n_modes=(n0, n1), hidden_channels=hid, in_channels=cin, out_channels=cout, lifting_channels=lp,
            projection_channels=lp, n_layers=L, rank=rank if rank > 0 else 1.0)
    sd = {k[2:]: torch.tensor(d[k]).float() for k in d.files if k.startswith("p:")}
    missing = m.load_state_dict(sd, strict=True)
    return d, m.to(DEV), (n0, n1), L


@pytest.mark.parametrize("name", ["fno_cfg1_small", "fno_rect", "fno_fullmodes", "tfno_small"])
def test_model_against_golden_fixture(golden_dir, name):
    """Whole FNO / TFNO fwd+bwd on the committed vectors (state_dict loads with strict=True)."""
    d, m, nm, L = _load_fixture(golden_dir, name)
    x = torch.tensor(d["x"]).float().to(DEV).requires_grad_(True)
    y = m(x)
    y.backward(torch.tensor(d["gy"]).float().to(DEV))
    assert rel_l2(y, torch.tensor(d["y"])) < TOL
    assert rel_l2(x.grad, torch.tensor(d["gx"])) < TOL
    for k, p in m.named_parameters():
        assert rel_l2(p.grad, torch.tensor(d["g:" + k])) < 2e-5, k


def test_per_block_api_equals_fused_stack(golden_dir):
    d, m, nm, L = _load_fixture(golden_dir, "fno_cfg1_small")
    x = torch.tensor(d["x"]).float().to(DEV)
    with torch.no_grad():
        h = m.lifting(x)
        a = m.fno_blocks.forward_all(h)
        b = h
        for l in range(L):
            b = m.fno_blocks(b, l)
    assert rel_l2(a, b) < 1e-6


def test_cfg1_full_size_vs_oracle():
    """BASELINE configs[0]: FNO2D 64x64, in 1, width 32, 12 modes, 4 layers, batch 8, fwd+bwd."""
    torch.manual_seed(1234)
    m = pkg.FNO(n_modes=(12, 12), hidden_channels=32, in_channels=1, out_channels=1, lifting_channels=256,
                projection_channels=256, n_layers=4)
    sd = {k: v.detach().clone() for k, v in m.state_dict().items()}
    x = _rand(8, 1, 64, 64, seed=7)
    tgt = _rand(8, 1, 64, 64, seed=8)
    leaves = {k: v.double().requires_grad_(True) for k, v in sd.items()}
    yo = so.fno_forward(leaves, x.double(), (12, 12), 4)
    lo = torch.nn.functional.mse_loss(yo, tgt.double())
    lo.backward()
    m = m.to(DEV)
    yc = m(x.to(DEV))
    lc = torch.nn.functional.mse_loss(yc, tgt.to(DEV))
    lc.backward()
    assert rel_l2(yc, yo) < TOL
    assert abs(lc.item() - lo.item()) / abs(lo.item()) < TOL
    for k, p in m.named_parameters():
        assert rel_l2(p.grad, leaves[k].grad) < 2e-5, k


def test_rollout_20_steps_vs_oracle():
    """north_star: parity checked over a 20-step closed-loop rollout."""
    torch.manual_seed(5)
    m = pkg.FNO(n_modes=(12, 12), hidden_channels=16, in_channels=1, out_channels=1, lifting_channels=32,
                projection_channels=32, n_layers=4)
    sd = {k: v.detach().double() for k, v in m.state_dict().items()}
    x0 = _rand(2, 1, 32, 32, seed=11)
    ref = so.rollout(sd, x0.double(), (12, 12), 4, 20)
    m = m.to(DEV)
    outs, x = [], x0.to(DEV)
    with torch.no_grad():
        for _ in range(20):
            x = m(x)
            outs.append(x)
    got = torch.stack(outs, dim=1)
    assert rel_l2(got, ref) < TOL
    assert rel_l2(got[:, -1], ref[:, -1]) < 5e-5


def test_rollout_engine_graph_and_sequence_forward():
    """the captured-graph closed loop equals the eager loop and the oracle; the reference wrapper's
    teacher-forcing loop (src/nsbench/models/fno/fno.py:29-43) is reproduced by sequence_forward."""
    torch.manual_seed(7)
    m = pkg.FNO(n_modes=(12, 12), hidden_channels=16, in_channels=1, out_channels=1, lifting_channels=32,
                projection_channels=32, n_layers=4)
    sd = {k: v.detach().double() for k, v in m.state_dict().items()}
    x0 = _rand(3, 1, 32, 32, seed=12)
    ref = so.rollout(sd, x0.double(), (12, 12), 4, 12)
    m = m.to(DEV)
    eng = pkg.Rollout(m, graph=True)
    got = eng(x0.to(DEV), 12)
    assert got.shape == ref.shape
    assert rel_l2(got, ref) < TOL
    again = eng(x0.to(DEV), 12)                       # second call replays the captured step
    assert torch.equal(got, again)
    eager = pkg.Rollout(m, graph=False)(x0.to(DEV), 12)
    assert rel_l2(eager, got) < 1e-6
    # teacher forcing for 3 frames, closed loop for the remaining 3
    xs = _rand(2, 6, 1, 32, 32, seed=13)
    outs, x_t = [], None
    for t in range(6):
        x_t = xs[:, t].double() if t < 3 else x_t
        x_t = so.fno_forward(sd, x_t, (12, 12), 4)
        outs.append(x_t)
    want = torch.stack(outs, dim=1)
    with torch.no_grad():
        seq = pkg.sequence_forward(m, xs.to(DEV), teacher_forcing_steps=3)
    assert rel_l2(seq, want) < TOL


def test_cfg2_shapes_vs_oracle():
    """BASELINE configs[1] shapes (64x64, width 64, 16 modes, batch 64): one block fwd+bwd against the
    fp64 CPU oracle."""
    B, C, H, W, nm = 64, 64, 64, 64, (16, 16)
    half = so.halve_last_mode(nm)
    x = _rand(B, C, H, W, seed=1)
    w = _rand(C, C, 16, 9, 2, seed=2, scale=0.2)
    ws = _rand(C, C, 1, 1, seed=3, scale=0.1)
    b = _rand(C, 1, 1, seed=4, scale=0.3)
    gy = _rand(B, C, H, W, seed=5)
    xo, wo, wso, bo = (t.double().requires_grad_(True) for t in (x, w, ws, b))
    yo = torch.nn.functional.gelu(so.spectral_conv_dense(xo, torch.view_as_complex(wo), bo, half)
                                  + torch.nn.functional.conv2d(xo, wso))
    yo.backward(gy.double())
    xc, wc, wsc, bc = (t.to(DEV).requires_grad_(True) for t in (x, w, ws, b))
    yc = FNOBlockFn.apply(xc, wc, wsc, bc, tuple(half), True)
    yc.backward(gy.to(DEV))
    assert rel_l2(yc, yo) < TOL
    assert rel_l2(xc.grad, xo.grad) < TOL
    assert rel_l2(wc.grad, wo.grad) < 2e-5
    assert rel_l2(wsc.grad, wso.grad) < 2e-5
    assert rel_l2(bc.grad, bo.grad) < 2e-5


def test_linearity_at_cfg3_tile():
    """Size-independent property at 256x256 / 32 modes: SpectralConv is linear in x; plus parity with
    the CPU oracle.  (The checker runs on the CPU on purpose: for the non-Hermitian spectra this layer
    produces, torch.fft.irfftn on CUDA (cuFFT multi-dim C2R) is implementation-defined and at 256^2
    differs from the CPU result by ~10%; the oracle, like neuralop evaluated on CPU, means
    "inverse complex FFT over H, then C2R over W".  See DESIGN.md.)"""
    half = so.halve_last_mode((32, 32))
    m = pkg.SpectralConv(8, 8, (32, 32), fft_norm="forward", bias=False).to(DEV)
    a, b = _rand(2, 8, 256, 256, seed=1).to(DEV), _rand(2, 8, 256, 256, seed=2).to(DEV)
    with torch.no_grad():
        lhs = m(2.0 * a - 3.0 * b)
        rhs = 2.0 * m(a) - 3.0 * m(b)
        ref = so.spectral_conv_dense(a.double().cpu(), m.weight[0].to_dense_complex().cpu().to(torch.complex128),
                                     None, half)
        got = m(a)
    assert rel_l2(lhs, rhs) < 1e-5
    assert rel_l2(got, ref) < TOL


def test_odd_height_is_rejected_loudly():
    m = pkg.SpectralConv(2, 2, (4, 4), fft_norm="forward").to(DEV)
    with pytest.raises(pkg._lib.SpectralB200Error):
        m(torch.randn(1, 2, 15, 16, device=DEV))


# ----------------------------------------------------------------------------------------------
# tensor-core (tcgen05) pointwise kernel vs the CUDA-core kernel and the fp64 reference
# ----------------------------------------------------------------------------------------------
def _set_tc(mode):
    pkg._lib.set_tc_mode(mode)


TC_SHAPES = [
    # B, M(in), N(out), H, W
    (2, 64, 64, 16, 16),
    (3, 64, 64, 64, 64),
    (2, 256, 64, 32, 32),     # lifting fc2: K loop of 4 chunks
    (2, 64, 256, 32, 32),     # projection fc1: N = 256
    (1, 32, 48, 12, 20),      # KC < 64, N = 48, ragged pixel tail (HW = 240)
    (1, 8, 16, 8, 8),
]


@pytest.mark.parametrize("B,M,N,H,W", TC_SHAPES)
@pytest.mark.parametrize("mode,tol", [(3, 1e-5), (1, 2e-3)])
def test_tc_pointwise(B, M, N, H, W, mode, tol):
    plan = fno_plan(DEV, H, W, [min(4, H), min(3, W // 2 + 1)])
    A = _rand(B, M, H, W, seed=1)
    Wp = _rand(N, M, seed=2, scale=0.3)
    bias = _rand(N, seed=3)
    zprev = _rand(B, N, H, W, seed=4)
    ref_z = torch.einsum("nm,bmhw->bnhw", Wp.double(), A.double()) + bias.double().view(1, -1, 1, 1)
    ref_y = torch.nn.functional.gelu(ref_z)
    zp = zprev.double().requires_grad_(True)
    torch.nn.functional.gelu(zp).backward(torch.ones_like(zp))
    ref_b = (ref_z - bias.double().view(1, -1, 1, 1)) * zp.grad
    try:
        _set_tc(mode)
        y, z = ops.rowidft_pointwise(plan, 0, None, A.to(DEV), Wp.to(DEV), M, 1, bias.to(DEV), None, B, M, N, 0, True,
                                     want_z=True)
        gb, _ = ops.rowidft_pointwise(plan, 1, None, A.to(DEV), Wp.to(DEV), M, 1, None, zprev.to(DEV), B, M, N, 1, False)
        # transposed-weight access (what the data-gradient uses): Wp given as [M, N] with strides (1, N)
        yt, _ = ops.rowidft_pointwise(plan, 0, None, A.to(DEV), Wp.t().contiguous().to(DEV), 1, N, None, None, B, M, N, 0,
                                      False)
    finally:
        _set_tc(3)
    assert rel_l2(z, ref_z) < tol
    assert rel_l2(y, ref_y) < tol
    assert rel_l2(gb, ref_b) < tol
    assert rel_l2(yt, ref_z - bias.double().view(1, -1, 1, 1)) < tol


def test_tc_mode_0_is_exact_fp32_path():
    plan = fno_plan(DEV, 16, 16, [4, 3])
    A, Wp = _rand(2, 64, 16, 16, seed=1).to(DEV), _rand(64, 64, seed=2).to(DEV)
    try:
        _set_tc(0)
        y0, _ = ops.rowidft_pointwise(plan, 0, None, A, Wp, 64, 1, None, None, 2, 64, 64, 0, False)
        _set_tc(3)
        y3, _ = ops.rowidft_pointwise(plan, 0, None, A, Wp, 64, 1, None, None, 2, 64, 64, 0, False)
    finally:
        _set_tc(3)
    assert rel_l2(y3, y0) < 2e-6


@pytest.mark.parametrize("B,Co,Ci,H,W", [(3, 64, 64, 16, 16), (2, 256, 64, 32, 32), (2, 64, 256, 32, 32), (64, 64, 64, 64, 64),
                                          (1, 128, 32, 12, 20), (2, 64, 64, 8, 8)])
@pytest.mark.parametrize("mode,tol", [(3, 1e-5), (1, 2e-3)])
def test_tc_wgrad(B, Co, Ci, H, W, mode, tol):
    g = _rand(B, Co, H, W, seed=5)
    x = _rand(B, Ci, H, W, seed=6)
    try:
        _set_tc(mode)
        gW, gb = ops.pointwise_wgrad(g.to(DEV), x.to(DEV))
    finally:
        _set_tc(3)
    assert rel_l2(gW, torch.einsum("bop,bip->oi", g.double().flatten(2), x.double().flatten(2))) < tol
    assert rel_l2(gb, g.double().sum(dim=(0, 2, 3))) < 1e-5


TC_BLOCK_SHAPES = [
    # B, Cin, Cout, H, W, n_modes, act, skip
    (2, 16, 16, 64, 64, (12, 12), True, True),      # R = 2 rows per tile
    (1, 32, 32, 256, 256, (32, 32), True, True),    # V = 2 x-offset variants (cfg3 grid)
    (2, 16, 32, 32, 32, (8, 8), False, False),      # R = 4, pure SpectralConv (no pointwise operand)
    (1, 16, 16, 128, 128, (16, 16), True, True),    # cfg5 grid
    (2, 64, 64, 32, 64, (12, 12), True, True),      # dlwpbench grid
]


@pytest.mark.parametrize("B,Ci,Co,H,W,nm,act,skip", TC_BLOCK_SHAPES)
@pytest.mark.parametrize("mode,tol", [(3, 1e-5), (1, 3e-3)])
def test_tc_fno_block(B, Ci, Co, H, W, nm, act, skip, mode, tol):
    """FNO block (spectral + skip + GELU) fwd+bwd with the row-synthesis stage on tcgen05."""
    half = so.halve_last_mode(nm)
    lo, My = so.retained_rows(H, half[0])
    Mx = min(half[1], W // 2 + 1)
    x = _rand(B, Ci, H, W, seed=1)
    w = _rand(Ci, Co, My, Mx, 2, seed=2, scale=0.5)
    ws = _rand(Co, Ci, 1, 1, seed=3, scale=0.2) if skip else None
    b = _rand(Co, 1, 1, seed=4, scale=0.3)
    gy = _rand(B, Co, H, W, seed=5)
    xo, wo, bo = x.double().requires_grad_(True), w.double().requires_grad_(True), b.double().requires_grad_(True)
    wso = ws.double().requires_grad_(True) if skip else None
    yo = so.spectral_conv_dense(xo, torch.view_as_complex(wo), bo, [My, Mx])
    if skip:
        yo = yo + torch.nn.functional.conv2d(xo, wso)
    if act:
        yo = torch.nn.functional.gelu(yo)
    yo.backward(gy.double())
    xc, wc, bc = (t.to(DEV).requires_grad_(True) for t in (x, w, b))
    wsc = ws.to(DEV).requires_grad_(True) if skip else None
    try:
        _set_tc(mode)
        yc = FNOBlockFn.apply(xc, wc, wsc, bc, tuple(half), act)
        yc.backward(gy.to(DEV))
    finally:
        _set_tc(3)
    assert rel_l2(yc, yo) < tol
    assert rel_l2(xc.grad, xo.grad) < tol
    assert rel_l2(wc.grad, wo.grad) < 2 * tol
    assert rel_l2(bc.grad, bo.grad) < 2 * tol
    if skip:
        assert rel_l2(wsc.grad, wso.grad) < 2 * tol


@pytest.mark.parametrize("B,M,N,H,W", [(2, 1, 256, 16, 16), (3, 13, 64, 32, 64), (2, 8, 70, 8, 12), (1, 16, 33, 16, 16)])
def test_small_m_pointwise(B, M, N, H, W):
    """lifting fc1 (few input channels): streaming kernel, forward with GELU + z and backward with GELU'."""
    plan = fno_plan(DEV, H, W, [min(4, H), min(3, W // 2 + 1)])
    A = _rand(B, M, H, W, seed=1)
    Wp = _rand(N, M, seed=2, scale=0.5)
    bias = _rand(N, seed=3)
    zprev = _rand(B, N, H, W, seed=4)
    ref_z = torch.einsum("nm,bmhw->bnhw", Wp.double(), A.double()) + bias.double().view(1, -1, 1, 1)
    zp = zprev.double().requires_grad_(True)
    torch.nn.functional.gelu(zp).backward(torch.ones_like(zp))
    y, z = ops.rowidft_pointwise(plan, 0, None, A.to(DEV), Wp.to(DEV), M, 1, bias.to(DEV), None, B, M, N, 0, True, want_z=True)
    g, _ = ops.rowidft_pointwise(plan, 1, None, A.to(DEV), Wp.t().contiguous().to(DEV), 1, N, None, zprev.to(DEV), B, M, N, 1, False)
    assert rel_l2(z, ref_z) < TOL
    assert rel_l2(y, torch.nn.functional.gelu(ref_z)) < TOL
    assert rel_l2(g, (ref_z - bias.double().view(1, -1, 1, 1)) * zp.grad) < TOL


@pytest.mark.parametrize("B,M,H,W", [(2, 64, 64, 64), (3, 32, 16, 32), (1, 16, 8, 16), (5, 64, 32, 64)])
def test_fused_mlp_head(B, M, H, W):
    """sb200_mlp_head_fwd / _bwd (projection C -> 256 -> 1 with the hidden tensor on chip) vs fp64 torch."""
    N = 256
    h = _rand(B, M, H, W, seed=1)
    W1 = _rand(N, M, seed=2, scale=M ** -0.5)
    b1 = _rand(N, seed=3, scale=0.3)
    w2 = _rand(N, seed=4, scale=N ** -0.5)
    b2 = _rand(1, seed=5)
    gy = _rand(B, 1, H, W, seed=6)
    ho, W1o, b1o, w2o, b2o = (t.double().requires_grad_(True) for t in (h, W1, b1, w2, b2))
    z1 = torch.einsum("nm,bmhw->bnhw", W1o, ho) + b1o.view(1, -1, 1, 1)
    z1.retain_grad()
    yo = torch.einsum("n,bnhw->bhw", w2o, torch.nn.functional.gelu(z1)).unsqueeze(1) + b2o
    yo.backward(gy.double())
    d = lambda t: t.to(DEV).contiguous()
    yc = ops.mlp_head_fwd(d(h), d(W1), d(b1), d(w2), d(b2))
    assert rel_l2(yc, yo) < TOL
    gz1, gb1, gw2, gb2 = ops.mlp_head_bwd(d(h), d(W1), d(b1), d(w2), d(gy))
    assert rel_l2(gz1, z1.grad) < TOL
    assert rel_l2(gb1, b1o.grad) < 2e-5
    assert rel_l2(gw2, w2o.grad) < 2e-5
    assert rel_l2(gb2, b2o.grad) < 2e-5


def test_model_with_fused_head_vs_oracle():
    """Whole FNO (projection_channels = 256 -> fused head path) forward + all gradients against the oracle."""
    torch.manual_seed(7)
    m = pkg.FNO(n_modes=(8, 8), hidden_channels=32, in_channels=1, out_channels=1, lifting_channels=256,
                projection_channels=256, n_layers=2)
    sd = {k: v.detach().double() for k, v in m.state_dict().items()}
    x = _rand(3, 1, 32, 32, seed=1)
    tgt = _rand(3, 1, 32, 32, seed=2)
    leaves = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    yo = so.fno_forward(leaves, x.double(), (8, 8), 2)
    torch.nn.functional.mse_loss(yo, tgt.double()).backward()
    m = m.to(DEV)
    yc = m(x.to(DEV))
    torch.nn.functional.mse_loss(yc, tgt.to(DEV)).backward()
    assert rel_l2(yc, yo) < TOL
    for k, p in m.named_parameters():
        assert rel_l2(p.grad, leaves[k].grad) < 2e-5, k


@pytest.mark.parametrize("B,C,H,W", [(2, 64, 64, 64), (3, 32, 16, 32), (1, 16, 8, 16)])
def test_lift_tail_bwd(B, C, H, W):
    """sb200_lift_tail_bwd: gradients of (w1, b1) of a 1-input-channel lifting MLP with gz1 kept on chip."""
    N = 256
    x = _rand(B, 1, H, W, seed=1)
    w1 = _rand(N, seed=2)
    b1 = _rand(N, seed=3, scale=0.3)
    W2 = _rand(C, N, seed=4, scale=N ** -0.5)
    g = _rand(B, C, H, W, seed=5)
    w1o, b1o = w1.double().requires_grad_(True), b1.double().requires_grad_(True)
    h1 = torch.nn.functional.gelu(w1o.view(1, -1, 1, 1) * x.double() + b1o.view(1, -1, 1, 1))
    out = torch.einsum("cn,bnhw->bchw", W2.double(), h1)
    out.backward(g.double())
    d = lambda t: t.to(DEV).contiguous()
    gw1, gb1 = ops.lift_tail_bwd(d(g), d(W2), d(w1), d(b1), d(x))
    assert rel_l2(gw1, w1o.grad) < 2e-5
    assert rel_l2(gb1, b1o.grad) < 2e-5


@pytest.mark.parametrize("B,C,H,W", [(2, 64, 64, 64), (3, 32, 16, 32), (1, 16, 8, 16), (2, 64, 12, 20)])
def test_lift_fwd_and_wgrad_generated_operand(B, C, H, W):
    """sb200_lift_fwd / sb200_lift_wgrad: the 256-channel hidden tensor is generated on chip, never stored."""
    N = 256
    x = _rand(B, 1, H, W, seed=1)
    w1 = _rand(N, seed=2)
    b1 = _rand(N, seed=3, scale=0.3)
    W2 = _rand(C, N, seed=4, scale=N ** -0.5)
    b2 = _rand(C, seed=6)
    g = _rand(B, C, H, W, seed=5)
    W2o, b2o = W2.double().requires_grad_(True), b2.double().requires_grad_(True)
    h1 = torch.nn.functional.gelu(w1.double().view(1, -1, 1, 1) * x.double() + b1.double().view(1, -1, 1, 1))
    out = torch.einsum("cn,bnhw->bchw", W2o, h1) + b2o.view(1, -1, 1, 1)
    out.backward(g.double())
    d = lambda t: t.to(DEV).contiguous()
    y = ops.lift_fwd(d(x), d(w1), d(b1), d(W2), d(b2))
    assert rel_l2(y, out) < TOL
    gW2, gb2 = ops.lift_wgrad(d(g), d(x), d(w1), d(b1))
    assert rel_l2(gW2, W2o.grad) < 2e-5
    assert rel_l2(gb2, b2o.grad) < 2e-5
