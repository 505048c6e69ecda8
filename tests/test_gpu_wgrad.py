"""Tensor-core weight gradient (mil_wgrad_tc_f32): dW = G^T X, db = colsum(G) against fp64, over the shapes the path produces."""
import pytest
import torch

import cases

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def K():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import mhimk
    return mhimk.ops


@pytest.mark.parametrize("M,N,Kd", [(256, 128, 256), (1000, 512, 1024), (7765, 512, 1024), (10000, 512, 1536), (9700, 128, 512),
                                    (1941, 1024, 512), (50000, 512, 1024), (50433, 1536, 512), (4099, 384, 512), (33, 128, 256)])
def test_wgrad_matches_fp64(K, M, N, Kd):
    g = torch.Generator().manual_seed(M + N)
    G = torch.randn(M, N, generator=g) * 1e-3                     # gradient-sized values
    X = torch.randn(M, Kd, generator=g)
    want_W, want_b = G.double().t() @ X.double(), G.double().sum(0)
    Gc, Xc = G.cuda(), X.cuda()
    gW, gb = K.weight_grad(Gc, Xc, True)
    if M >= K.TC_MIN_ROWS:
        assert gb is not None                                    # the tensor-core path (bias gradient from the same pass)
        assert cases.rel_err(gb, want_b) < 1e-5
    assert cases.rel_err(gW, want_W) < 3e-5                        # bf16 hi+lo: 7.6e-6 unit roundoff
    gW2, _ = K.weight_grad(Gc, Xc, True)
    assert torch.equal(gW, gW2)                                    # deterministic slice reduction


def test_wgrad_row_tail_and_leading_dimension(K):
    """M not a multiple of 32 (TMA zero-fills the tail rows) and operands that are row slices of larger buffers."""
    M, N, Kd = 1000 + 17, 128, 256
    g = torch.Generator().manual_seed(3)
    G, X = torch.randn(M + 5, N, generator=g).cuda(), torch.randn(M + 5, Kd, generator=g).cuda()
    gW, gb = K.weight_grad(G[:M], X[:M], True)                     # rows beyond M hold data that must not leak in
    assert cases.rel_err(gW, G[:M].double().t() @ X[:M].double()) < 3e-5
    assert cases.rel_err(gb, G[:M].double().sum(0)) < 1e-5


def test_linear_act_backward_uses_it_and_matches_the_exact_path(K):
    M, Kd, N = 3000, 1024, 512
    g = torch.Generator().manual_seed(5)
    x, W, b, go = torch.randn(M, Kd, generator=g).cuda(), (torch.randn(N, Kd, generator=g) * 0.03).cuda(), torch.zeros(N).cuda(), torch.randn(M, N, generator=g).cuda()
    res = {}
    for tc in (True, False):
        K.WGRAD_TC = tc
        Wc, bc = W.clone().requires_grad_(), b.clone().requires_grad_()
        K.linear_act(x, Wc, bc, "gelu").backward(go)
        res[tc] = (Wc.grad, bc.grad)
    K.WGRAD_TC = True
    assert cases.rel_err(res[True][0], res[False][0]) < 3e-5 and cases.rel_err(res[True][1], res[False][1]) < 1e-5
