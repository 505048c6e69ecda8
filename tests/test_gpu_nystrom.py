"""The Nystrom / TransMIL forward on the library's own kernels (ops.nystrom_attention_forward and its pieces) against the CPU oracle
(== the live reference to 5e-6, tests/test_oracle_vs_reference.py) and against torch for the individual kernels."""
import pytest
import torch

import cases
from oracle import mil_oracle as O

pytestmark = pytest.mark.gpu
TOL = 1e-4


@pytest.fixture(scope="module")
def K():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import mhimk
    return mhimk.ops


def test_layernorm_rowsoftmax_bmm(K):
    g = torch.Generator().manual_seed(1)
    x, w, b = torch.randn(777, 512, generator=g) * 3 + 1, torch.randn(512, generator=g), torch.randn(512, generator=g)
    assert cases.rel_err(K.layernorm(x.cuda(), w.cuda(), b.cuda(), 1e-5), torch.nn.functional.layer_norm(x.double(), (512,), w.double(), b.double(), 1e-5)) < 2e-6
    s = torch.randn(1000, 256, generator=g) * 4
    sc = s.cuda()
    from mhimk import _lib
    _lib.check(_lib.lib().mil_row_softmax_f32(_lib.ptr(sc), 1000, 256, _lib.stream_ptr()), "row softmax")
    assert cases.rel_err(sc, torch.softmax(s.double(), -1)) < 2e-6
    a, bb = torch.randn(8, 256, 256, generator=g), torch.randn(8, 256, 64, generator=g)
    assert cases.rel_err(K.bmm(a.cuda(), bb.cuda()), a.double() @ bb.double()) < 2e-6
    c = torch.randn(8, 100, 64, generator=g)
    assert cases.rel_err(K.bmm_nt(bb.cuda(), c.cuda()), bb.double() @ c.double().transpose(1, 2)) < 2e-6
    x2 = torch.softmax(torch.randn(8, 256, 256, generator=g), -1)
    assert cases.rel_err(K.pinv_iter(x2.cuda(), 6), O.pinv_iter(x2.double(), 6)) < 1e-4


@pytest.mark.parametrize("n,m,dh", [(1024, 256, 64), (50432, 256, 64), (300, 256, 64), (5000, 100, 32)])
def test_colsoftmax_pool(K, n, m, dh):
    from mhimk import _lib
    g = torch.Generator().manual_seed(n)
    S, Vw = torch.randn(n, m, generator=g) * 3, torch.randn(n, 3 * dh, generator=g)
    V = Vw[:, dh:2 * dh]                                                     # a column block of a wider buffer (leading dimension 3 dh)
    P = torch.softmax(S.double(), 0)
    want = P.t() @ V.double()
    Sc, Vc = S.cuda(), Vw.cuda()
    out, M, L = torch.empty(m, dh, device="cuda"), torch.empty(m, device="cuda"), torch.empty(m, device="cuda")
    lib = _lib.lib()
    ws = torch.empty(lib.mil_colsoftmax_pool_workspace_bytes(n, m), dtype=torch.uint8, device="cuda")
    _lib.check(lib.mil_colsoftmax_pool_f32(_lib.ptr(Sc), _lib.c_void_p(Vc.data_ptr() + 4 * dh), 3 * dh, n, m, dh, _lib.ptr(out), _lib.ptr(M), _lib.ptr(L), _lib.ptr(ws),
                                           ws.numel(), _lib.stream_ptr()), "colsoftmax_pool")
    assert cases.rel_err(out, want) < 5e-6
    assert torch.equal(M.cpu(), S.max(0).values)
    assert cases.rel_err(L, torch.exp(S.double() - S.double().max(0).values).sum(0)) < 5e-6
    w = torch.randn(m, generator=g)
    o = torch.empty(n, device="cuda")
    _lib.check(lib.mil_expdot_rows_f32(_lib.ptr(Sc), n, m, _lib.ptr(M), _lib.ptr(w.cuda()), _lib.ptr(o), _lib.stream_ptr()), "expdot")
    assert cases.rel_err(o, torch.exp(S.double() - S.double().max(0).values) @ w.double()) < 5e-6


def test_dwconv_and_ppeg(K):
    from mhimk import _lib
    g = torch.Generator().manual_seed(3)
    rows, heads, dh = 1000, 8, 64
    v, w = torch.randn(rows, heads * dh, generator=g), torch.randn(heads, 1, 33, 1, generator=g) * 0.1
    want = torch.nn.functional.conv2d(v.reshape(rows, heads, dh).permute(1, 0, 2)[None].double(), w.double(), padding=(16, 0), groups=heads)[0].permute(1, 0, 2).reshape(rows, -1)
    out = torch.ones(rows, heads * dh, device="cuda")
    _lib.check(_lib.lib().mil_dwconv_tokens_f32(_lib.ptr(v.cuda()), heads * dh, rows, heads, dh, _lib.ptr(w.reshape(heads, 33).cuda()), 33, _lib.ptr(out), heads * dh, 1,
                                                _lib.stream_ptr()), "dwconv")
    assert cases.rel_err(out - 1, want) < 5e-6
    H = W = 23
    C = 512
    tok = torch.randn(H * W, C, generator=g)
    convs = [torch.nn.Conv2d(C, C, k, 1, k // 2, groups=C) for k in (7, 5, 3)]
    gg = tok.t().reshape(1, C, H, W)
    with torch.no_grad():
        want = (convs[0](gg) + gg + convs[1](gg) + convs[2](gg)).flatten(2).transpose(1, 2)[0]
        got = K.ppeg_forward(tok.cuda(), H, W, [c.cuda() for c in convs])
    assert cases.rel_err(got, want) < 5e-6


@pytest.mark.parametrize("n,ret,no_norm", [(600, False, False), (600, True, False), (600, True, True), (257, True, False), (1024, False, False),
                                           (5001, True, False)])
def test_nystrom_layer_matches_oracle(K, n, ret, no_norm):
    sd = cases.mhim_state(71, "selfattn")
    p = "online_encoder.layer1.attn."
    x = torch.randn(n, 512, generator=torch.Generator().manual_seed(n))
    with torch.no_grad():
        ref = O.nystrom_attention(sd, x, p, return_attn=ret, no_norm=no_norm)
    c = {k: v.cuda() for k, v in sd.items()}
    got = K.nystrom_attention_forward(x.cuda(), c[p + "to_qkv.weight"], c[p + "to_out.0.weight"], c[p + "to_out.0.bias"], c[p + "res_conv.weight"], 8, 256, 6,
                                      64 ** -0.5, return_attn=ret, no_norm=no_norm)
    if not ret:
        assert cases.rel_err(got, ref) < TOL
        return
    for a, b in zip(got, ref):
        assert cases.rel_err(a, b) < TOL
