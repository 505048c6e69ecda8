"""GPU: the tensor-core Linear(+bias+act) forward (mil_linear_act_tc_f32) against fp64, incl. multi-block widths and autograd."""
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


@pytest.mark.parametrize("M,N,Kd", [(256, 64, 32), (1000, 128, 512), (777, 192, 96), (3000, 512, 1024), (5000, 1536, 512), (2000, 1024, 512), (300, 320, 64)])
@pytest.mark.parametrize("act", ["none", "relu", "gelu", "tanh"])
def test_linear_forward_tc(K, M, N, Kd, act):
    g = torch.Generator().manual_seed(M + N)
    x, W, b = torch.randn(M, Kd, generator=g), torch.randn(N, Kd, generator=g) * 0.05, torch.randn(N, generator=g) * 0.1
    pre = torch.empty(M, N, device="cuda")
    y = K.linear_forward(x.cuda(), W.cuda(), b.cuda(), act, pre)
    ref_pre = x.double() @ W.double().t() + b.double()
    assert cases.rel_err(pre, ref_pre) < 2e-5
    assert cases.rel_err(y, O.apply_act(ref_pre, act)) < 5e-5      # bf16x3 contraction (<= 1e-5) + the MUFU-based activation (<= 1e-6 abs)
    y2 = K.linear_forward(x.cuda(), W.cuda(), None, act)
    assert cases.rel_err(y2, O.apply_act(x.double() @ W.double().t(), act)) < 5e-5


def test_linear_act_autograd_through_tc(K):
    g = torch.Generator().manual_seed(4)
    M, N, Kd = 1500, 512, 1024
    x, W, b = torch.randn(M, Kd, generator=g), torch.randn(N, Kd, generator=g) * 0.03, torch.randn(N, generator=g) * 0.1
    go = torch.randn(M, N, generator=g)
    for act in ("gelu", "relu"):
        Wd, bd = W.cuda().requires_grad_(True), b.cuda().requires_grad_(True)
        yd = K.linear_act(x.cuda(), Wd, bd, act)
        (yd * go.cuda()).sum().backward()
        Wr, br = W.double().requires_grad_(True), b.double().requires_grad_(True)
        pre_ref = x.double() @ Wr.t() + br
        if act == "relu":
            # ReLU's derivative is discontinuous: an element with |pre| ~ 1e-6 may sit on the other side of 0 in fp64 and flips a
            # whole 1/M share of a gradient row.  Check the backward kernels against the gate the forward actually used.
            gate = (yd.detach().cpu() > 0).double()
            assert float((gate != (pre_ref.detach() > 0).double()).double().mean()) < 1e-4
            ((pre_ref * gate) * go.double()).sum().backward()
        else:
            (O.apply_act(pre_ref, act) * go.double()).sum().backward()
        assert cases.rel_err(Wd.grad, Wr.grad) < 1e-4 and cases.rel_err(bd.grad, br.grad) < 1e-4
        with torch.no_grad():                                   # cached weight image must follow in-place updates
            Wd.mul_(1.5)
        y = K.linear_act(x.cuda(), Wd, bd, act)
        assert cases.rel_err(y, O.apply_act(x.double() @ (1.5 * W.double()).t() + b.double(), act)) < 2e-5


def test_column_split_and_whole_width_share_one_weight_image(K):
    """A 512-wide Linear over <= 74 row tiles runs as two 256-column blocks on separate CTAs, over more rows as whole-width tiles; both read
    the SAME cached weight image (256-row blocks), so alternating bag sizes on one weight must stay correct; dropout bits and the
    pre-activation output follow the column block."""
    g = torch.Generator().manual_seed(12)
    N, Kd = 512, 1024
    W, b = (torch.randn(N, Kd, generator=g) * 0.05).cuda(), (torch.randn(N, generator=g) * 0.1).cuda()
    for M in (3000, 20000, 129, 9472, 9473, 3000):
        x = torch.randn(M, Kd, generator=g)
        pre = torch.empty(M, N, device="cuda")
        y = K.linear_forward(x.cuda(), W, b, "gelu", pre)
        ref_pre = x.double() @ W.cpu().double().t() + b.cpu().double()
        assert cases.rel_err(pre, ref_pre) < 2e-5, M
        assert cases.rel_err(y, O.apply_act(ref_pre, "gelu")) < 5e-5, M
    M = 2500
    x = torch.randn(M, Kd, generator=g)
    keep = torch.rand(M, N, generator=g) > 0.25
    spec = K.DropSpec(p=0.25, keep_bits=K.pack_keep_bits(keep.cuda()))
    y = K.linear_forward(x.cuda(), W, b, "gelu", None, dropout=spec)
    ref = O.apply_act(x.double() @ W.cpu().double().t() + b.cpu().double(), "gelu") * keep.double() / 0.75
    assert cases.rel_err(y, ref) < 5e-5
