"""The radix-2 folds used by analysis_fused / coldft_inv2 / the AFNO channels-last transforms, restated in numpy and checked
against the plain truncated DFT the oracle computes (CPU only).  The kernels build the same table entries
(dlwp_benchmark_b200/csrc/plan.cu: rowF[x][k] = a (cos, -sin)(2 pi k x / W), colF[j][y] = s exp(-2 pi i ky_j y / H),
colI[y][j] = s exp(+2 pi i ky_j y / H), ky_j = (ky0 + j) mod H), so an identity that holds here holds for them."""
import numpy as np
import pytest


def _rowF(W, Mx):
    x, k = np.arange(W)[:, None], np.arange(Mx)[None, :]
    th = 2 * np.pi * ((k * x) % W) / W
    return np.cos(th), -np.sin(th)


@pytest.mark.parametrize("W,Mx", [(64, 9), (64, 17), (32, 5), (32, 17), (64, 33), (20, 7)])
def test_row_stage_real_input_fold_and_second_fold(W, Mx):
    rng = np.random.default_rng(W * 100 + Mx)
    x = rng.standard_normal(W)
    C, S = _rowF(W, Mx)
    ref_re, ref_im = x @ C, x @ S
    # fold 1 (all kernels): n and W - n share the cosine and negate the sine
    nh = W // 2
    e, o = np.zeros(nh + 1), np.zeros(nh + 1)
    for n in range(nh + 1):
        paired = n > 0 and 2 * n != W
        e[n] = x[n] + (x[W - n] if paired else 0.0)
        o[n] = x[n] - x[W - n] if paired else 0.0
    assert np.allclose(e @ C[:nh + 1], ref_re) and np.allclose(o @ S[:nh + 1], ref_im)
    if W % 4:
        return
    # fold 2 (analysis_fused): m and W/2 - m share their twiddles up to (-1)^k
    q = W // 4
    re, im = np.zeros(Mx), np.zeros(Mx)
    for m in range(q + 1):
        if m == q:                                   # its own partner: both parities see e[q], o[q]
            ee = eo = e[q]; oe = od = o[q]
        else:
            ee, eo = e[m] + e[nh - m], e[m] - e[nh - m]
            oe, od = o[m] - o[nh - m], o[m] + o[nh - m]
        for k in range(Mx):
            re[k] += (eo if k & 1 else ee) * C[m, k]
            im[k] += (od if k & 1 else oe) * S[m, k]
    assert np.allclose(re, ref_re) and np.allclose(im, ref_im)


@pytest.mark.parametrize("H,ky0,My", [(64, -8, 16), (64, -5, 10), (32, 0, 32), (16, 5, 8), (64, -9, 18)])
def test_column_stage_folds(H, ky0, My):
    rng = np.random.default_rng(H + My)
    ky = (ky0 + np.arange(My)) % H
    y = np.arange(H)
    colF = np.exp(-2j * np.pi * np.outer(ky, y) / H)            # [My][H]
    colI = np.exp(+2j * np.pi * np.outer(y, ky) / H)            # [H][My]
    first_odd = ky0 % 2 == 1
    # analysis: inputs y and y + H/2 folded; even / odd frequencies see the sum / difference
    T = rng.standard_normal(H) + 1j * rng.standard_normal(H)
    s, d = T[:H // 2] + T[H // 2:], T[:H // 2] - T[H // 2:]
    out = np.zeros(My, complex)
    for j in range(My):
        odd = (j & 1) ^ int(first_odd)
        out[j] = colF[j, :H // 2] @ (d if odd else s)
    assert np.allclose(out, colF @ T)
    # synthesis: rows y and y + H/2 from the sums over the even- / odd-indexed modes
    Y = rng.standard_normal(My) + 1j * rng.standard_normal(My)
    E = colI[:H // 2, 0::2] @ Y[0::2]
    O = colI[:H // 2, 1::2] @ Y[1::2]
    fs = -1.0 if first_odd else 1.0
    assert np.allclose(np.concatenate([E + O, fs * (E - O)]), colI @ Y)


@pytest.mark.parametrize("W,Mx", [(64, 17), (32, 9), (21, 5)])
def test_row_synthesis_hermitian_output_fold(W, Mx):
    rng = np.random.default_rng(W + Mx)
    k, w = np.arange(Mx)[:, None], np.arange(W)[None, :]
    th = 2 * np.pi * ((k * w) % W) / W
    tx, ty = np.cos(th), -np.sin(th)                               # rowI[k][w]
    P = rng.standard_normal(Mx) + 1j * rng.standard_normal(Mx)
    ref = P.real @ tx + P.imag @ ty
    yv = np.zeros(W)
    for ww in range(W // 2 + 1):
        A, B = P.real @ tx[:, ww], P.imag @ ty[:, ww]
        yv[ww] = A + B
        if ww > 0 and 2 * ww != W:
            yv[W - ww] = A - B
    assert np.allclose(yv, ref)
