"""GPU parity of the fp32 CUDA-core kernels and the selection kernels against torch fp64 / the oracle."""
import math

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


def dev(t):
    return t.cuda().contiguous()


@pytest.mark.parametrize("M,N,Kd", [(1, 1, 1), (7, 5, 3), (128, 128, 8), (257, 130, 77), (1000, 512, 1024), (333, 2, 512), (50, 1, 128)])
@pytest.mark.parametrize("act", ["none", "relu", "gelu", "tanh", "sigmoid"])
def test_sgemm_nt(K, M, N, Kd, act):
    g = torch.Generator().manual_seed(M * 31 + N)
    x, W, b = torch.randn(M, Kd, generator=g), torch.randn(N, Kd, generator=g) * 0.1, torch.randn(N, generator=g)
    ref = O.apply_act(x.double() @ W.double().t() + b.double(), act)
    got = K.sgemm(dev(x), Kd, 1, dev(W), Kd, 1, M, N, Kd, bias=dev(b), act=act)      # the exact-fp32 CUDA-core path
    assert cases.rel_err(got, ref) < (2e-5 if act in ('tanh', 'sigmoid') else 3e-6)   # fp32 accumulation over K, then a saturating act


def test_sgemm_forms_splitk_and_rowids(K):
    g = torch.Generator().manual_seed(5)
    M, N, Kd = 3001, 96, 200
    gp, x, W = torch.randn(M, N, generator=g), torch.randn(M, Kd, generator=g), torch.randn(N, Kd, generator=g)
    for sk in (1, 4, 13):
        gW = K.sgemm(dev(gp), 1, N, dev(x), 1, Kd, N, Kd, M, splitk=sk)                 # TN
        assert cases.rel_err(gW, gp.double().t() @ x.double()) < 3e-6
    gx = K.sgemm(dev(gp), N, 1, dev(W), 1, Kd, M, Kd, N)                                 # NN
    assert cases.rel_err(gx, gp.double() @ W.double()) < 3e-6
    ids = torch.randperm(M, generator=g)[:777]
    y = K.linear_act_rows(dev(x), dev(W), None, "relu", dev(ids))
    assert cases.rel_err(y, torch.relu(x[ids].double() @ W.double().t())) < 3e-6


@pytest.mark.parametrize("act", ["relu", "gelu", "tanh", "sigmoid", "none"])
def test_linear_act_autograd(K, act):
    g = torch.Generator().manual_seed(9)
    M, N, Kd = 515, 130, 96
    x, W, b = torch.randn(M, Kd, generator=g), torch.randn(N, Kd, generator=g) * 0.2, torch.randn(N, generator=g) * 0.1
    go = torch.randn(M, N, generator=g)
    xr, Wr, br = (t.double().requires_grad_(True) for t in (x, W, b))
    (O.apply_act(xr @ Wr.t() + br, act) * go.double()).sum().backward()
    xd, Wd, bd = (dev(t).requires_grad_(True) for t in (x, W, b))
    (K.linear_act(xd, Wd, bd, act) * dev(go)).sum().backward()
    assert cases.rel_err(xd.grad, xr.grad) < 5e-6
    assert cases.rel_err(Wd.grad, Wr.grad) < 5e-6
    assert cases.rel_err(bd.grad, br.grad) < 5e-6


@pytest.mark.parametrize("L", [1, 2, 63, 64, 65, 1000, 20011])
def test_softmax_pool_fwd_bwd(K, L):
    g = torch.Generator().manual_seed(L)
    H = 512
    s, h, gp = torch.randn(L, generator=g) * 3, torch.randn(L, H, generator=g), torch.randn(H, generator=g)
    keep = (torch.rand(L, generator=g) > 0.3).to(torch.uint8)
    keep[0] = 1
    for kp in (None, keep):
        sr, hr = s.double().requires_grad_(True), h.double().requires_grad_(True)
        sm = sr if kp is None else sr.masked_fill(kp == 0, float("-inf"))
        a = torch.softmax(sm, 0)
        p = a @ hr
        (p * gp.double()).sum().backward()
        sd, hd = dev(s).requires_grad_(True), dev(h).requires_grad_(True)
        pd, ad = K.softmax_pool(sd, hd, None if kp is None else dev(kp))
        (pd * dev(gp)).sum().backward()
        assert cases.rel_err(pd, p) < 3e-6
        assert cases.rel_err(ad, a) < 3e-6
        assert float((sd.grad.cpu().double() - sr.grad).abs().max()) < 2e-5 * max(1.0, float(sr.grad.abs().max()))
        assert cases.rel_err(hd.grad, hr.grad) < 3e-6
    # strided logits column (DSMIL: one softmax per class column)
    s2 = torch.randn(L, 2, generator=g)
    pd, _ = K.softmax_pool(dev(s2)[:, 1], dev(h))
    assert cases.rel_err(pd, torch.softmax(s2[:, 1].double(), 0) @ h.double()) < 3e-6


def test_pool_merge_matches_oracle(K):
    g = torch.Generator().manual_seed(1)
    s, h = torch.randn(5000, generator=g) * 4, torch.randn(5000, 512, generator=g)
    parts, o = [], 0
    for c in (1, 999, 2000, 1500, 500):
        m, l, P = O.pool_partial(s[o:o + c].double(), h[o:o + c].double())
        parts.append(torch.cat([m[None], l[None], P]).float())
        o += c
    parts.append(torch.zeros(514))             # an idle shard (l == 0) is ignored
    stats, pooled = K.pool_merge(dev(torch.stack(parts)))
    assert cases.rel_err(pooled, O.softmax_pool(s.double(), h.double())[0]) < 3e-6


def tie_aware_equal(score, idx, ref_idx, k):
    """Same k-th value; everything strictly better is included in both; the rest are ties at the boundary."""
    v = score[idx]
    vr = score[ref_idx]
    assert torch.equal(v.sort(descending=True).values, vr.sort(descending=True).values)
    thr = vr.min()
    strict = set(torch.nonzero(score > thr).flatten().tolist())
    assert strict <= set(idx.tolist()) and len(set(idx.tolist())) == k
    assert all(score[i] >= thr for i in idx.tolist())


@pytest.mark.parametrize("N,k", [(1, 1), (5, 1), (3, 3), (1000, 30), (1000, 8), (4099, 205), (10000, 1024), (10000, 1025), (50000, 1500), (50001, 300),
                                 (200000, 6000), (200000, 16), (3000, 3000), (70000, 70000)])
def test_topk_exact_on_tie_free_scores(K, N, k):
    g = torch.Generator().manual_seed(N)
    score = torch.randperm(N, generator=g).float() / N - 0.3          # distinct values, both signs
    for largest in (True, False):
        ref = torch.topk(score, k, largest=largest).indices
        got = K.topk(dev(score), k, largest).cpu()
        assert torch.equal(got, ref)                                     # bit-exact incl. order


@pytest.mark.parametrize("N,k", [(50000, 1500), (50000, 700), (3000, 64)])
def test_topk_ties_lowest_index_first(K, N, k):
    """k > 1024 orders the winners by the stable radix sort, k <= 1024 by counting ranks: both must put equal scores in index order."""
    g = torch.Generator().manual_seed(3)
    score = (0.5 + torch.randint(0, 900, (N,), generator=g).float() * 5.9604645e-08)   # ~900 distinct values around 0.5
    got = K.topk(dev(score), k, True).cpu()
    key = score.double() * 1e6 * N - torch.arange(N).double() / 1.0
    order = sorted(range(N), key=lambda i: (-score[i].item(), i))[:k]
    assert got.tolist() == order
    tie_aware_equal(score, got, torch.topk(score, k).indices, k)


@pytest.mark.parametrize("ps,ratio,hr", [(1000, 0.03, 1.0), (4099, 0.05, 1.0), (255, 0.01, 1.0), (10000, 0.03, 1.0), (1000, 0.03, 0.5), (2, 0.03, 1.0)])
def test_mask_ids_bit_exact_vs_oracle(K, ps, ratio, hr):
    g = torch.Generator().manual_seed(ps)
    score = torch.randperm(ps, generator=g).float()[None] / ps
    torch.manual_seed(11)
    lk_ref, ids_ref = O.select_mask(ps, score, True, ratio, len_keep_other=ps, random_ratio=hr)
    k = O.topk_count(ps, ratio / hr if ratio / hr <= 1 else 1.0)
    idx = K.topk(dev(score), k, True)
    if hr < 1.0:
        torch.manual_seed(11)
        perm = torch.randperm(k)
        idx = idx[dev(perm[: int(math.ceil(k * hr))])]
    mask_ids, keep, len_keep = K.mask_from_indices(idx, ps)
    assert int(len_keep.item()) == lk_ref
    assert torch.equal(mask_ids.cpu(), ids_ref)
    assert int(keep.sum().item()) == lk_ref


def test_cam_score(K):
    g = torch.Generator().manual_seed(2)
    L = 3000
    s, h = torch.randn(L, generator=g), torch.randn(L, 512, generator=g)
    Wp, bp = torch.randn(2, 512, generator=g) * 0.05, torch.randn(2, generator=g)
    a = torch.softmax(s.double(), 0)
    ref = O.pseudo_score(Wp.double(), bp.double(), h.double(), a)
    m = s.max()
    stats = torch.stack([m, torch.exp(s - m).sum()])
    got = K.cam_score(dev(s), dev(h @ Wp.t()), dev(stats), float(bp[0]))
    assert cases.rel_err(got, ref) < 1e-6
    got_dev = K.cam_score(dev(s), dev(h @ Wp.t()), dev(stats), dev(bp))      # bias read on the device (no host sync): same bits
    assert torch.equal(got_dev, got)
