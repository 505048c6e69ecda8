"""GEMV-shaped Linear layers (mil_skinny_*): forward and backward against torch autograd in fp64, and bit-stable across repeats."""
import pytest
import torch

import cases
from oracle import mil_oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def K():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import mhimk
    return mhimk.ops


@pytest.mark.parametrize("act", ["none", "gelu", "tanh", "relu"])
@pytest.mark.parametrize("M,Kd,N,bias", [(7765, 128, 1, False), (10000, 512, 2, True), (9700, 128, 2, False), (1, 512, 2, True), (5, 512, 512, False),
                                         (5, 512, 512, True), (2, 512, 128, True), (2, 128, 128, True), (300, 256, 8, True), (4000, 1536, 1, True), (8, 1536, 700, True),
                                         (33, 96, 3, True)])
def test_skinny_linear_fwd_bwd(K, act, M, Kd, N, bias):
    g = torch.Generator().manual_seed(M + N + Kd)
    x, W = torch.randn(M, Kd, generator=g), torch.randn(N, Kd, generator=g) * 0.05
    b = torch.randn(N, generator=g) * 0.1 if bias else None
    go = torch.randn(M, N, generator=g)
    xr, Wr = x.double().requires_grad_(), W.double().requires_grad_()
    br = b.double().requires_grad_() if bias else None
    pre = xr @ Wr.t() + (br if bias else 0)
    xc, Wc = x.cuda().requires_grad_(), W.cuda().requires_grad_()
    bc = b.cuda().requires_grad_() if bias else None
    assert K._skinny(M, N, Kd)
    y = K.linear_act(xc, Wc, bc, act)
    y.backward(go.cuda())
    if act == "relu":                      # the kernel's own gates (exact fp32 here, but fp64 may still differ on a ~1e-8 pre-activation)
        yr = pre * (y.detach().cpu() > 0).double()
    else:
        yr = O.apply_act(pre, act)
    yr.backward(go.double())
    tol = 2e-6 if Kd <= 512 else 1e-5                  # fp32 accumulation over K terms
    assert cases.rel_err(y, yr) < tol
    assert cases.rel_err(xc.grad, xr.grad) < tol and cases.rel_err(Wc.grad, Wr.grad) < 5e-6
    if bias:
        assert cases.rel_err(bc.grad, br.grad) < 5e-6
    # same results as the general fp32 GEMM path, and deterministic
    K.SKINNY = False
    try:
        xc2, Wc2 = x.cuda().requires_grad_(), W.cuda().requires_grad_()
        y2 = K.linear_act(xc2, Wc2, b.cuda() if bias else None, act)
        y2.backward(go.cuda())
    finally:
        K.SKINNY = True
    assert cases.rel_err(y, y2) < tol and cases.rel_err(Wc.grad, Wc2.grad) < 5e-6
    xc3, Wc3 = x.cuda().requires_grad_(), W.cuda().requires_grad_()
    K.linear_act(xc3, Wc3, b.cuda() if bias else None, act).backward(go.cuda())
    assert torch.equal(Wc3.grad, Wc.grad) and torch.equal(xc3.grad, xc.grad)
