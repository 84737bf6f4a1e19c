"""GPU parity tests of the strided complex GEMM (sb200_cgemm) and the Tucker reconstruction built on it,
against fp64 einsum on the CPU (oracle.spectral_oracle.tucker_to_dense).  Tolerance: rel-L2 <= 1e-5."""
import pytest
import torch

from dlwp_benchmark_b200 import ops
from dlwp_benchmark_b200.tucker import TuckerWeight
from dlwp_benchmark_b200.tucker_fn import TuckerReconstructFn
from oracle import spectral_oracle as so
from conftest import rel_l2

pytestmark = pytest.mark.gpu
TOL = 1e-5
DEV = "cuda"


def _crand(*s, seed=0):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(*s, 2, generator=g)


CGEMM_SHAPES = [
    # M, N, K  (tile pickers: 16x16, 32x32, 256x16, 128x32, 64x64; split-K when few tiles and long K)
    (9, 8, 5000), (16, 15, 333), (30, 17, 4100), (1000, 9, 8), (777, 16, 15), (500, 30, 33), (130, 70, 60),
    (64, 60, 7680), (1, 1, 1), (65, 129, 17),
    (5000, 9, 8), (4100, 16, 15), (3001, 30, 33), (2048, 5, 64),     # skinny kernel (thread owns two rows)
]


@pytest.mark.parametrize("conjA,conjB", [(False, False), (True, False), (False, True)])
@pytest.mark.parametrize("M,N,K", CGEMM_SHAPES)
def test_cgemm_plain(M, N, K, conjA, conjB):
    """Both storage orders of A and B (m-fast / k-fast loaders) against complex128 matmul."""
    for a_t in (False, True):
        for b_t in (False, True):
            A = _crand(*((K, M) if a_t else (M, K)), seed=1)
            B = _crand(*((N, K) if b_t else (K, N)), seed=2)
            Ac = torch.view_as_complex(A.double()); Bc = torch.view_as_complex(B.double())
            Am = Ac.T if a_t else Ac
            Bm = Bc.T if b_t else Bc
            ref = (Am.conj() if conjA else Am) @ (Bm.conj() if conjB else Bm)
            C = torch.empty(M, N, 2, device=DEV)
            ops.cgemm(A.to(DEV), B.to(DEV), C, M=M, N=N, K=K,
                      sAm=1 if a_t else K, sAk=M if a_t else 1, sBk=1 if b_t else N, sBn=K if b_t else 1,
                      sCm=N, sCn=1, conjA=conjA, conjB=conjB)
            assert rel_l2(torch.view_as_complex(C), ref) < TOL, (a_t, b_t)


def test_cgemm_composite_indices():
    """Mode-1 product and mode-1 factor gradient of a 3-way tensor via composite m / k, no permutes."""
    I, G, Q, O = 7, 11, 13, 9
    T = _crand(I, G, Q, seed=3); U = _crand(O, G, seed=4)
    Tc, Uc = torch.view_as_complex(T.double()), torch.view_as_complex(U.double())
    ref = torch.einsum("og,igq->ioq", Uc, Tc)
    out = torch.empty(I, O, Q, 2, device=DEV)
    ops.cgemm(T.to(DEV), U.to(DEV), out, M=I * Q, M2=Q, N=O, K=G, sAm=(G * Q, 1), sAk=Q, sBk=1, sBn=G,
              sCm=(O * Q, 1), sCn=Q)
    assert rel_l2(torch.view_as_complex(out), ref) < TOL
    Gr = _crand(I, O, Q, seed=5)
    Gc = torch.view_as_complex(Gr.double())
    refU = torch.einsum("ioq,igq->og", Gc, Tc.conj())
    gU = torch.empty(O, G, 2, device=DEV)
    ops.cgemm(Gr.to(DEV), T.to(DEV), gU, M=O, N=G, K=I * Q, K2=Q, sAm=Q, sAk=(O * Q, 1), sBk=(G * Q, 1), sBn=Q,
              sCm=G, sCn=1, conjB=True)
    assert rel_l2(torch.view_as_complex(gU), refU) < TOL


TUCKER_SHAPES = [
    ((64, 64, 16, 9), 0.8),      # cfg2: ranks (60, 60, 15, 8)
    ((32, 32, 12, 7), 0.8),      # cfg1 width
    ((5, 3, 4, 3), 0.5),
    ((8, 8, 6, 4), 1.0),
    ((16, 24, 32, 17), 0.3),
]


@pytest.mark.parametrize("shape,rank", TUCKER_SHAPES)
def test_tucker_reconstruct_forward_backward(shape, rank):
    torch.manual_seed(0)
    tw = TuckerWeight(shape, rank, init_std=0.1)
    core = tw.core.detach().clone()
    facs = [f.detach().clone() for f in tw.factors]
    gW = _crand(*shape, seed=9)
    # fp64 reference on the CPU
    co = core.double().requires_grad_(True)
    fo = [f.double().requires_grad_(True) for f in facs]
    Wo = torch.view_as_real(so.tucker_to_dense(torch.view_as_complex(co), [torch.view_as_complex(f) for f in fo]))
    Wo.backward(gW.double())
    cc = core.to(DEV).requires_grad_(True)
    fc = [f.to(DEV).requires_grad_(True) for f in facs]
    Wc = TuckerReconstructFn.apply(cc, *fc)
    Wc.backward(gW.to(DEV))
    assert rel_l2(Wc, Wo) < TOL
    assert rel_l2(cc.grad, co.grad) < TOL
    for a, b in zip(fc, fo):
        assert rel_l2(a.grad, b.grad) < TOL


def test_tucker_grouped_layers_equal_single():
    """Four layers reconstructed in grouped launches == each layer on its own (forward and gradients)."""
    from dlwp_benchmark_b200.tucker_fn import reconstruct_many
    torch.manual_seed(1)
    shape = (12, 10, 6, 4)
    tws = [TuckerWeight(shape, 0.7, init_std=0.1).to(DEV) for _ in range(4)]
    gWs = [_crand(*shape, seed=20 + i).to(DEV) for i in range(4)]
    Ws = reconstruct_many(tws)
    sum((W * g).sum() for W, g in zip(Ws, gWs)).backward()
    for tw, W, g in zip(tws, Ws, gWs):
        core = tw.core.detach().clone().requires_grad_(True)
        facs = [f.detach().clone().requires_grad_(True) for f in tw.factors]
        W1 = TuckerReconstructFn.apply(core, *facs)
        (W1 * g).sum().backward()
        assert torch.equal(W1, W)
        assert torch.equal(core.grad, tw.core.grad)
        for a, b in zip(facs, tw.factors):      # the split of the reduction depends on the group count
            assert rel_l2(a.grad, b.grad) < 1e-6
