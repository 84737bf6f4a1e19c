"""GPU: the channels-last token kernels (tcgen05 GEMM in all four operand-major combinations, split-K, fused
epilogues; LayerNorm fwd / bwd; column / batch sums) against float64 torch on the CPU."""
import pytest
import torch
import torch.nn.functional as F

from dlwp_benchmark_b200 import _lib, ops
from conftest import rel_l2

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _rand(*s, seed=0, scale=1.0):
    return torch.randn(*s, generator=torch.Generator().manual_seed(seed)) * scale


GEMM_SHAPES = [
    # M, N, K
    (128, 128, 32), (256, 64, 64), (300, 200, 96), (1000, 256, 256), (128, 1024, 256), (384, 16, 256), (200, 8, 64),
    (512, 256, 1024), (128, 48, 8), (130, 36, 40),
]


@pytest.mark.parametrize("mode,tol", [(3, 2e-6), (1, 2e-3), (0, 2e-6)])
@pytest.mark.parametrize("a_mn,b_mn", [(False, False), (False, True), (True, False), (True, True)])
@pytest.mark.parametrize("M,N,K", GEMM_SHAPES)
def test_gemm_all_majors(M, N, K, a_mn, b_mn, mode, tol):
    lib = _lib.load()
    if (a_mn and M % 4) or (b_mn and N % 4) or (not a_mn and K % 4) or (not b_mn and K % 4):
        pass                                   # unaligned leading dims take the CUDA-core kernel: still must be right
    A = _rand(M, K, seed=1)
    B = _rand(N, K, seed=2)
    want = A.double() @ B.double().t()
    Ad = (A.t().contiguous() if a_mn else A).to(DEV)
    Bd = (B.t().contiguous() if b_mn else B).to(DEV)
    old = _lib.tc_mode()
    try:
        _lib.set_tc_mode(mode)
        got = ops.gemm(Ad, Bd, a_mn=a_mn, b_mn=b_mn)
    finally:
        _lib.set_tc_mode(old)
    assert rel_l2(got, want) < tol


@pytest.mark.parametrize("M,N,K", [(512, 256, 128), (300, 72, 64), (256, 1024, 256)])
def test_gemm_epilogues(M, N, K):
    A, B = _rand(M, K, seed=3), _rand(N, K, seed=4, scale=0.1)
    bias, aux, res = _rand(N, seed=5), _rand(M, N, seed=6), _rand(M, N, seed=7)
    z = A.double() @ B.double().t() + bias.double()
    Ad, Bd = A.to(DEV), B.to(DEV)
    y, zz = ops.gemm(Ad, Bd, bias=bias.to(DEV), act=1, resid=res.to(DEV), want_z=True)
    assert rel_l2(zz, z) < 2e-6
    assert rel_l2(y, F.gelu(z) + res.double()) < 2e-6
    a64 = aux.double().requires_grad_(True)
    F.gelu(a64).sum().backward()
    y2 = ops.gemm(Ad, Bd, act=2, aux=aux.to(DEV))
    assert rel_l2(y2, (A.double() @ B.double().t()) * a64.grad) < 2e-6
    pos = _rand(64, N, seed=8)
    y3 = ops.gemm(Ad, Bd, bias=bias.to(DEV), resid=pos.to(DEV), res_rows=64)
    assert rel_l2(y3, z + pos.double().repeat(M // 64 + 1, 1)[:M]) < 2e-6


@pytest.mark.parametrize("M,N,K", [(1024, 256, 32768), (256, 1024, 8192), (8, 256, 4096), (256, 16, 20000)])
def test_gemm_split_k_weight_gradient_shape(M, N, K):
    """dW[n_out, n_in] = dY^T X: both operands MN-major, the contraction runs over the tokens."""
    dY, X = _rand(K, M, seed=9), _rand(K, N, seed=10)
    want = dY.double().t() @ X.double()
    got = ops.gemm(dY.to(DEV), X.to(DEV), a_mn=True, b_mn=True, split_k=True)
    assert rel_l2(got, want) < 3e-6
    again = ops.gemm(dY.to(DEV), X.to(DEV), a_mn=True, b_mn=True, split_k=True)
    assert torch.equal(got, again)             # fixed-order reduction


@pytest.mark.parametrize("T,C", [(1000, 256), (77, 64), (4096, 1024), (33, 272), (512, 384)])
def test_layernorm_fwd_bwd(T, C):
    x, dy, dres = _rand(T, C, seed=11) * 2 + 0.5, _rand(T, C, seed=12), _rand(T, C, seed=13)
    g, b = _rand(C, seed=14) * 0.5 + 1, _rand(C, seed=15)
    xo, go, bo = (t.double().requires_grad_(True) for t in (x, g, b))
    yo = F.layer_norm(xo, (C,), go, bo, eps=1e-6)
    yo.backward(dy.double())
    y, mean, rstd = ops.layernorm_fwd(x.to(DEV), g.to(DEV), b.to(DEV), 1e-6)
    assert rel_l2(y, yo) < 2e-6
    dx, dg, db = ops.layernorm_bwd(dy.to(DEV), x.to(DEV), g.to(DEV), mean, rstd, dres=dres.to(DEV))
    assert rel_l2(dx, xo.grad + dres.double()) < 3e-6
    assert rel_l2(dg, go.grad) < 3e-6 and rel_l2(db, bo.grad) < 3e-6


def test_colsum_and_batch_sum():
    a = _rand(5000, 272, seed=16)
    assert rel_l2(ops.colsum(a.to(DEV)), a.double().sum(0)) < 2e-6
    b = _rand(16, 2048, 40, seed=17)
    assert rel_l2(ops.batch_sum(b.to(DEV)), b.double().sum(0)) < 1e-6


def test_gemm_gelu_operand_transform_and_z_only():
    """h = GELU(z) formed on chip: as the A operand (fc2 forward) and as the B operand (fc2 weight gradient)."""
    T, C, Hd = 640, 64, 256
    z, W2, gy = _rand(T, Hd, seed=21), _rand(C, Hd, seed=22, scale=0.1), _rand(T, C, seed=23)
    h = F.gelu(z.double())
    y = ops.gemm(z.to(DEV), W2.to(DEV), a_gelu=True)
    assert rel_l2(y, h @ W2.double().t()) < 2e-6
    gW2 = ops.gemm(gy.to(DEV), z.to(DEV), a_mn=True, b_mn=True, b_gelu=True, split_k=True)
    assert rel_l2(gW2, gy.double().t() @ h) < 3e-6
    x, W1, b1 = _rand(T, C, seed=24), _rand(Hd, C, seed=25, scale=0.2), _rand(Hd, seed=26)
    zz = ops.gemm(x.to(DEV), W1.to(DEV), bias=b1.to(DEV), z_only=True)
    assert rel_l2(zz, x.double() @ W1.double().t() + b1.double()) < 2e-6
    lib = _lib.load()
    old = _lib.tc_mode()
    try:
        for mode, tol in ((1, 2e-3), (0, 2e-6)):
            _lib.set_tc_mode(mode)
            assert rel_l2(ops.gemm(z.to(DEV), W2.to(DEV), a_gelu=True), h @ W2.double().t()) < tol
    finally:
        _lib.set_tc_mode(old)
