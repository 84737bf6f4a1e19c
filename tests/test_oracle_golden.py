"""CPU: the oracle against the committed golden vectors and known answers."""
import math
import os

import numpy as np
import pytest
import torch

from oracle import afno_oracle as ao
from oracle import spectral_oracle as so
from conftest import rel_l2

AFNO = ["afno2d_sq8", "afno2d_8x16", "afno2d_32x64", "afno2d_frac", "afno2d_fac2"]
FNO = ["fno_cfg1_small", "fno_rect", "fno_fullmodes", "tfno_small"]


@pytest.mark.parametrize("name", AFNO)
@pytest.mark.parametrize("fn", [ao.afno2d_fft, ao.afno2d_explicit])
def test_afno_oracle_matches_reference_vectors(golden_dir, name, fn):
    """Golden vectors were produced by the reference's own AFNO2D class (oracle/make_golden.py)."""
    d = np.load(os.path.join(golden_dir, name + ".npz"))
    B, h, w, C, nb, fac = [int(v) for v in d["meta"]]
    t = lambda k: torch.tensor(d[k])
    x = t("x").requires_grad_(True)
    ps = [t(k).requires_grad_(True) for k in ("w1", "b1", "w2", "b2")]
    y = fn(x, *ps, nb, 0.01, float(d["frac"]))
    y.backward(t("gy"))
    tol = 1e-6 if fn is ao.afno2d_fft else 2e-5
    assert rel_l2(y, t("y")) < tol
    assert rel_l2(x.grad, t("gx")) < tol
    for p, k in zip(ps, ("w1", "b1", "w2", "b2")):
        assert rel_l2(p.grad, t("g" + k)) < tol, k


def test_afno_kept_region_quirk():
    # at 32x64 tokens only 17 of 33 rfft columns are processed, all 32 rows
    assert ao.kept_region(32, 64, 1.0) == (0, 32, 17)
    assert ao.kept_region(8, 8, 1.0) == (0, 8, 5)
    assert ao.kept_region(16, 16, 0.5) == (5, 13, 4)


@pytest.mark.parametrize("name", FNO)
def test_fno_oracle_reproduces_fixture(golden_dir, name):
    d = np.load(os.path.join(golden_dir, name + ".npz"))
    B, cin, hid, cout, H, W, n0, n1, L, lp = [int(v) for v in d["meta"]]
    sd = {k[2:]: torch.tensor(d[k]).requires_grad_(True) for k in d.files if k.startswith("p:")}
    x = torch.tensor(d["x"]).requires_grad_(True)
    y = so.fno_forward(sd, x, (n0, n1), L)
    y.backward(torch.tensor(d["gy"]))
    assert rel_l2(y, torch.tensor(d["y"])) < 1e-12
    assert rel_l2(x.grad, torch.tensor(d["gx"])) < 1e-12
    for k, v in sd.items():
        assert rel_l2(v.grad, torch.tensor(d["g:" + k])) < 1e-12, k


@pytest.mark.parametrize("shape", [(2, 3, 16, 32, (8, 12)), (1, 2, 8, 8, (8, 8)), (2, 2, 12, 20, (6, 7)),
                                   (1, 1, 64, 64, (12, 12))])
def test_fft_and_explicit_statements_agree(shape):
    B, C, H, W, nm = shape
    torch.manual_seed(0)
    half = so.halve_last_mode(nm)
    x = torch.randn(B, C, H, W, dtype=torch.float64)
    w = torch.randn(C, C + 1, *[min(a, b) for a, b in zip(half, (H, W // 2 + 1))], dtype=torch.complex128)
    b = torch.randn(C + 1, 1, 1, dtype=torch.float64)
    a = so.spectral_conv_dense(x, w, b, half)
    e = so.spectral_conv_explicit(x, w, b, half)
    assert (a - e).abs().max().item() < 1e-12


def test_retained_rows_match_survey():
    # 64^2/n12 -> rows 26..37 (ky -6..5); n16 -> ky -8..7; 256^2/n32 -> ky -16..15
    assert so.retained_rows(64, 12) == (26, 12)
    assert so.retained_rows(64, 16) == (24, 16)
    assert so.retained_rows(256, 32) == (112, 32)
    assert so.retained_rows(128, 32) == (48, 32)


def test_backward_spec_of_survey_8c():
    """The adjoint formulas the kernels implement against torch autograd in fp64."""
    torch.manual_seed(1)
    B, C, O, H, W, nm = 2, 3, 4, 16, 16, (6, 6)
    half = so.halve_last_mode(nm)
    x = torch.randn(B, C, H, W, dtype=torch.float64, requires_grad=True)
    wr = torch.randn(C, O, *half, 2, dtype=torch.float64, requires_grad=True)
    gy = torch.randn(B, O, H, W, dtype=torch.float64)
    y = so.spectral_conv_dense(x, torch.view_as_complex(wr), None, half)
    y.backward(gy)
    ky, kx, wk = so.explicit_tables(H, W, half)
    yy, xx = torch.arange(H, dtype=torch.float64), torch.arange(W, dtype=torch.float64)
    EH = torch.exp(-2j * math.pi * ky[:, None] * yy[None] / H)
    EW = torch.exp(-2j * math.pi * kx[:, None] * xx[None] / W)
    Xh = torch.einsum("ky,bcyx,qx->bckq", EH, x.detach().to(torch.complex128), EW) / (H * W)
    gYh = torch.einsum("ky,boyx,qx,q->bokq", EH, gy.to(torch.complex128), EW, wk.to(torch.complex128))
    Wc = torch.view_as_complex(wr.detach())
    gW = torch.einsum("bikq,bokq->iokq", Xh.conj(), gYh)
    gXh = torch.einsum("iokq,bokq->bikq", Wc.conj(), gYh)
    gx = torch.einsum("ky,bikq,qx->biyx", EH.conj(), gXh, EW.conj()).real / (H * W)
    assert rel_l2(torch.view_as_real(gW), wr.grad) < 1e-12
    assert rel_l2(gx, x.grad) < 1e-12


def test_tucker_rank_known_answers():
    assert so.tucker_rank((64, 64, 16, 9), 0.8) == (60, 60, 15, 8)
    assert so.tucker_rank((32, 32, 12, 7), 0.8) == (30, 30, 11, 7)
    assert so.tucker_rank((64, 64, 16, 9), 1.0) == (64, 64, 16, 9)


def test_tucker_contraction_equals_dense():
    torch.manual_seed(2)
    r = (3, 4, 3, 2)
    core = torch.randn(*r, dtype=torch.complex128)
    facs = [torch.randn(s, k, dtype=torch.complex128) for s, k in zip((4, 5, 6, 4), r)]
    x = torch.randn(2, 4, 16, 16, dtype=torch.float64)
    a = so.spectral_conv_tucker(x, core, facs, None, [6, 4])
    b = so.spectral_conv_dense(x, so.tucker_to_dense(core, facs), None, [6, 4])
    assert (a - b).abs().max().item() < 1e-12


def test_reference_size_labels():
    """In-tree corroboration of the recalled weight layout: the nsbench sweep
    hidden_channels -> the plot labels 5k..32M (src/nsbench/scripts/train_commands.txt:83-91,
    src/nsbench/scripts/plot_results.py:58)."""
    want = {2: 0.007, 8: 0.051, 27: 0.510, 38: 0.999, 54: 2.002, 77: 4.051, 108: 7.944, 154: 16.115, 217: 31.948}
    for hc, m in want.items():
        got = so.fno_param_count((12, 12), 10, hc, 256, 256, 1, 4) / 1e6
        assert abs(got - m) < 6e-4, (hc, got)


# --------------------------------------------------------------------------------------
# FourCastNet block remainder (row f3): oracle restatement vs vectors from the reference's own classes
# --------------------------------------------------------------------------------------
def _fcn(golden_dir, name):
    d = np.load(os.path.join(golden_dir, name + ".npz"))
    sd = {k[2:]: torch.tensor(d[k]) for k in d.files if k.startswith("p:")}
    return d, sd


def test_fcn_block_oracle_matches_reference_class(golden_dir):
    from oracle import afno_oracle as ao
    d, sd = _fcn(golden_dir, "fcn_block")
    leaves = {k: v.double().requires_grad_(True) for k, v in sd.items()}
    x = torch.tensor(d["in:x"]).double().requires_grad_(True)
    y = ao.block_forward(leaves, "", x, int(d["meta"][4]))
    y.backward(torch.tensor(d["gy"]).double())
    assert rel_l2(y, torch.tensor(d["y"])) < 2e-6          # the vectors are fp32 outputs of the reference
    assert rel_l2(x.grad, torch.tensor(d["gin:x"])) < 5e-6
    for k, v in leaves.items():
        assert rel_l2(v.grad, torch.tensor(d["g:" + k])) < 2e-5, k


def test_fcn_dlwp_net_oracle_matches_reference_class(golden_dir):
    from oracle import afno_oracle as ao
    d, sd = _fcn(golden_dir, "fcn_dlwp_net")
    H, W, p1, p2, cc, cp, cg, E, depth, nb, ctx = [int(v) for v in d["meta"]]
    leaves = {k: v.double().requires_grad_(True) for k, v in sd.items()}
    c, pr, pg = (torch.tensor(d["in:" + k]).double() for k in ("constants", "prescribed", "prognostic"))
    step = lambda x_t: ao.afnonet_step(leaves, x_t, (p1, p2), depth, nb)
    x_t = torch.cat([c[:, 0], pr[:, 0:1].flatten(1, 2), pg[:, 0:1].flatten(1, 2)], dim=1)
    y = (pg[:, 0] + step(x_t)).unsqueeze(1)
    y.backward(torch.tensor(d["gy"]).double())
    assert rel_l2(y, torch.tensor(d["y"])) < 2e-6
    for k, v in leaves.items():
        if "g:" + k in d.files:
            assert rel_l2(v.grad, torch.tensor(d["g:" + k])) < 3e-5, k
    assert "g:norm.weight" not in d.files                    # the final norm is unused by the reference forward
