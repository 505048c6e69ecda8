"""Row selection by a permutation and Merge's cross-attention (mil_take_rows / mil_scatter_rows / mil_mca_*), forward and backward,
against torch indexing / the oracle's MCA in fp64."""
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


@pytest.mark.parametrize("rows,cols,n_keep", [(10, 512, 3), (9700, 512, 7760), (2000, 512, 1999), (257, 64, 1)])
def test_split_and_take_rows(K, rows, cols, n_keep):
    g = torch.Generator().manual_seed(rows)
    x = torch.randn(rows, cols, generator=g)
    perm = torch.randperm(rows, generator=g)
    ga, gb = torch.randn(n_keep, cols, generator=g), torch.randn(rows - n_keep, cols, generator=g)
    xr = x.clone().requires_grad_()
    (xr[perm[:n_keep]] * ga).sum().backward(retain_graph=True)
    g_take = xr.grad.clone()
    xr.grad = None
    ((xr[perm[:n_keep]] * ga).sum() + (xr[perm[n_keep:]] * gb).sum()).backward()
    xc = x.cuda().requires_grad_()
    t = K.take_rows(xc, perm.cuda(), n_keep)
    assert torch.equal(t.cpu(), x[perm[:n_keep]])
    (t * ga.cuda()).sum().backward()
    assert torch.equal(xc.grad.cpu(), g_take)                       # masked rows: exact zeros
    xc.grad = None
    a, b = K.split_rows(xc, perm.cuda(), n_keep)
    assert torch.equal(a.cpu(), x[perm[:n_keep]]) and torch.equal(b.cpu(), x[perm[n_keep:]])
    ((a * ga.cuda()).sum() + (b * gb.cuda()).sum()).backward()
    assert torch.equal(xc.grad.cpu(), xr.grad)


@pytest.mark.parametrize("L,kq,drop", [(1940, 5, False), (1940, 5, True), (7, 1, False), (9700, 8, True), (300, 3, False)])
def test_mca_attend_fwd_bwd(K, L, kq, drop):
    heads, dh = 8, 64
    inner = heads * dh
    g = torch.Generator().manual_seed(L + kq)
    q, kv = torch.randn(kq, inner, generator=g), torch.randn(L, 2 * inner, generator=g)
    go = torch.randn(kq, inner, generator=g)
    pmask = torch.nn.functional.dropout(torch.ones(heads, kq, L), 0.1, True) if drop else None
    scale = dh ** -0.5
    qr, kvr = q.double().requires_grad_(), kv.double().requires_grad_()
    split = lambda t: t.reshape(t.shape[0], heads, dh).permute(1, 0, 2)
    attn = torch.softmax(split(qr) @ split(kvr[:, :inner]).transpose(-1, -2) * scale, -1)
    if drop:
        attn = attn * pmask.double()
    ref = (attn @ split(kvr[:, inner:])).permute(1, 0, 2).reshape(kq, inner)
    ref.backward(go.double())
    qc, kvc = q.cuda().requires_grad_(), kv.cuda().requires_grad_()
    out = K.mca_attend(qc, kvc, heads, scale, pmask.cuda() if drop else None)
    out.backward(go.cuda())
    assert cases.rel_err(out, ref) < 2e-6
    assert cases.rel_err(qc.grad, qr.grad) < 5e-6 and cases.rel_err(kvc.grad, kvr.grad) < 5e-6
    out2 = K.mca_attend(qc, kvc, heads, scale, pmask.cuda() if drop else None)
    assert torch.equal(out, out2)
