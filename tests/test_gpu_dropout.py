"""Dropout inside the kernels: the reference's trainer keeps the teacher in train() mode with dropout 0.25 on
all N rows (engines/base_engine.py:36-37, modules/mhim.py:193-194), so the fused pass must apply it itself.
  * mask-in parity: the reference-style Bernoulli mask is fed as keep bits to both the kernel and the oracle (1e-4);
  * the in-kernel Philox stream equals its host restatement (tests/philox_ref.py) bit for bit, and passes statistical checks;
  * MHIM teacher / student / DAttention in train mode with dropout run through the fused kernel / the fused-dropout GEMM."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

import cases
import philox_ref
from oracle import mil_oracle as O

pytestmark = pytest.mark.gpu
TOL = 1e-4


@pytest.fixture(scope="module")
def K():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import mhimk
    return mhimk.ops


def unpack(bits, ncols):
    """int32 keep words [rows, ncols/32] -> bool [rows, ncols] (CPU)."""
    w = bits.cpu().to(torch.int64) & 0xFFFFFFFF
    return ((w[:, :, None] >> torch.arange(32)) & 1).reshape(bits.shape[0], ncols).bool()


@pytest.mark.parametrize("rows,ncols,p,seed,off", [(1, 32, 0.25, 1, 0), (1000, 512, 0.25, 2021, 7), (4099, 128, 0.1, 2 ** 40 + 5, 2 ** 33 + 1),
                                                   (257, 512, 0.5, 0, 2 ** 63)])
def test_philox_bits_match_host_restatement(K, rows, ncols, p, seed, off):
    spec = K.DropSpec(p, seed, off)
    got = unpack(K.dropout_bits(rows, ncols, spec, "cuda"), ncols)
    want = torch.from_numpy(philox_ref.keep_mask(rows, ncols, p, seed, off))
    assert torch.equal(got, want)
    assert torch.equal(unpack(K.pack_keep_bits(want.cuda()), ncols), want)           # the packing helper round-trips


def test_philox_stream_statistics(K):
    """Keep rate = 1 - p within 4 sigma overall, per row and per column; successive offsets are independent streams."""
    rows, ncols, p = 20000, 512, 0.25
    a = unpack(K.dropout_bits(rows, ncols, K.DropSpec(p, 2021, 1), "cuda"), ncols).float()
    b = unpack(K.dropout_bits(rows, ncols, K.DropSpec(p, 2021, 2), "cuda"), ncols).float()
    n = rows * ncols
    sig = (p * (1 - p)) ** 0.5
    assert abs(a.mean().item() - (1 - p)) < 4 * sig / n ** 0.5
    assert (a.mean(1) - (1 - p)).abs().max().item() < 5.5 * sig / ncols ** 0.5       # 20 000 rows: 5.5 sigma
    assert (a.mean(0) - (1 - p)).abs().max().item() < 5 * sig / rows ** 0.5
    corr = ((a - a.mean()) * (b - b.mean())).mean().item() / (p * (1 - p))
    assert abs(corr) < 5 / n ** 0.5                                                  # different offsets: uncorrelated
    lag = ((a[:, 1:] - (1 - p)) * (a[:, :-1] - (1 - p))).mean().item() / (p * (1 - p))
    assert abs(lag) < 5 / n ** 0.5                                                   # neighbouring columns: uncorrelated


def run_oracle(sd, x, act, mask):
    sd64 = {k: v.double() for k, v in sd.items()}
    h = O.apply_act(O.affine(x[0].double(), sd64["feature.0.weight"], sd64["feature.0.bias"]), act) * mask.double()
    u = torch.tanh(O.affine(h, sd64["attention.0.weight"], sd64["attention.0.bias"]))
    s = O.affine(u, sd64["attention.2.weight"], sd64["attention.2.bias"])[:, 0]
    return h, s, torch.softmax(s, 0) @ h


@pytest.mark.parametrize("pipe", ["pair", "single"])
@pytest.mark.parametrize("mode", ["bits", "philox"])
@pytest.mark.parametrize("N,act", [(1, "relu"), (129, "gelu"), (4099, "relu"), (10000, "gelu")])
def test_fused_forward_with_dropout(K, pipe, mode, N, act):
    """mask-in parity at 1e-4: the kernel and the fp64 oracle see the same keep mask (mode 'bits': torch's own Bernoulli draw, as the
    reference makes it; mode 'philox': the in-kernel stream, read back through mil_dropout_bits)."""
    sd, x = cases.abmil_state(300 + N), cases.make_bag(400 + N, N, 1024)
    p = 0.25
    if mode == "bits":
        torch.manual_seed(N)
        mask = F.dropout(torch.ones(N, 512), p, True)
        spec = K.DropSpec(p, keep_bits=K.pack_keep_bits((mask > 0).cuda()))
    else:
        spec = K.DropSpec(p, 77, N)
        mask = unpack(K.dropout_bits(N, 512, spec, "cuda"), 512).float() / (1 - p)
    h_ref, s_ref, p_ref = run_oracle(sd, x, act, mask)
    c = {k: v.cuda() for k, v in sd.items()}
    Wp = torch.randn(2, 512, generator=torch.Generator().manual_seed(1)) * 0.05
    out = K.abmil_fused_forward(x[0].cuda(), c["feature.0.weight"], c["feature.0.bias"], act, c["attention.0.weight"], c["attention.0.bias"],
                                c["attention.2.weight"], c["attention.2.bias"], "tanh", Wp=Wp.cuda(), want_scores=True, want_h=True,
                                Wcls=c["classifier.weight"], bcls=c["classifier.bias"], pipeline=pipe, dropout=spec)
    assert cases.rel_err(out["h"], h_ref) < TOL
    assert bool(((out["h"].cpu() == 0) | (mask > 0)).all())                          # dropped entries are exact zeros
    assert cases.rel_err(out["pooled"], p_ref) < TOL
    assert cases.rel_err(out["s"], s_ref) < 3 * TOL
    assert cases.rel_err(out["t"], h_ref @ Wp.double().t()) < 3 * TOL
    ref_logits = p_ref @ sd["classifier.weight"].double().t() + sd["classifier.bias"].double()
    assert cases.rel_err(out["logits"][0], ref_logits) < TOL


@pytest.mark.parametrize("act", ["relu", "gelu", "tanh", "none"])
@pytest.mark.parametrize("M,K_,N", [(1000, 1024, 512), (300, 512, 128), (100, 96, 64)])
def test_linear_act_dropout_fwd_bwd(K, act, M, K_, N):
    """drop(act(x W^T + b)) with the dropout in the GEMM epilogue (tensor-core path for M >= 256, CUDA-core path below) and its
    backward (mask regenerated from the Philox stream) against torch autograd in fp64 with the same mask."""
    g = torch.Generator().manual_seed(M + N)
    x, W, b = torch.randn(M, K_, generator=g), torch.randn(N, K_, generator=g) * 0.05, torch.randn(N, generator=g) * 0.1
    spec = K.DropSpec(0.25, 5, M)
    mask = unpack(K.dropout_bits(M, N, spec, "cuda"), N).double() / 0.75
    xc, Wc, bc = x.cuda().requires_grad_(), W.cuda().requires_grad_(), b.cuda().requires_grad_()
    y = K.linear_act(xc, Wc, bc, act, dropout=spec)
    go = torch.randn(M, N, generator=g)
    y.backward(go.cuda())
    xr, Wr, br = x.double().requires_grad_(), W.double().requires_grad_(), b.double().requires_grad_()
    pre = xr @ Wr.t() + br
    if act == "relu":
        # the gate of a pre-activation within the contraction's rounding error of zero (a few of the 5e5 elements) is decided by that
        # error: take the kernel's own gates so that the comparison checks the arithmetic, not which side of 0 a 1e-6 value fell on
        gate = torch.where(mask > 0, y.detach().cpu().double() > 0, pre.detach() > 0).double()
        flipped = (pre.detach() > 0).double() != gate
        assert int(flipped.sum()) <= 8 and (int(flipped.sum()) == 0 or float(pre.detach()[flipped].abs().max()) < 1e-4)
        yr = pre * gate * mask
    else:
        yr = O.apply_act(pre, act) * mask
    yr.backward(go.double())
    assert cases.rel_err(y, yr) < TOL
    assert cases.rel_err(xc.grad, xr.grad) < TOL and cases.rel_err(Wc.grad, Wr.grad) < TOL and cases.rel_err(bc.grad, br.grad) < TOL


def _hook_from_masks(K, masks):
    """DROPOUT_HOOK that hands out the given keep masks (one per dropout call, in order) as keep bits."""
    it = iter(masks)

    def hook(rows, ncols, p, device):
        m = next(it)
        assert tuple(m.shape) == (rows, ncols)
        return K.DropSpec(p, keep_bits=K.pack_keep_bits((m > 0).to(device)))
    return hook


def test_mhim_train_mode_teacher_takes_the_fused_kernel(K, monkeypatch):
    """The reference's training configuration: MHIM(dropout=0.25), teacher in train() under no_grad.  It must launch the fused
    kernel (not the composed path) and match the oracle fed with the same mask; the student's feature GEMM carries its dropout too."""
    from mhimk import modules as M
    n, d, seed = 3000, 1024, 51
    kw = dict(cases.MHIM_KW, baseline="attn", input_dim=d, dropout=0.25)
    cfg = O.MHIMConfig(**dict(cases.MHIM_KW, baseline="attn", input_dim=d))
    tea, stu = M.MHIM(**kw).cuda().train(), M.MHIM(**kw).cuda().train()
    sd_t, sd_s = cases.mhim_state(seed + 1, "attn", D=d), cases.mhim_state(seed, "attn", D=d)
    tea.load_state_dict({k: v.cuda() for k, v in sd_t.items()}, strict=True)
    stu.load_state_dict({k: v.cuda() for k, v in sd_s.items()}, strict=True)
    for m in (tea, stu):
        for name, mod in m.named_modules():
            if isinstance(mod, torch.nn.Dropout) and name != "dp":
                mod.p = 0.0
    x = cases.make_bag(seed + 1000, n, d)
    torch.manual_seed(3)
    mask_t = F.dropout(torch.ones(n, 512), 0.25, True)
    mask_s = F.dropout(torch.ones(n, 512), 0.25, True)
    calls = []
    real = K.abmil_fused_forward
    monkeypatch.setattr(K, "abmil_fused_forward", lambda *a, **k: (calls.append(k.get("dropout")), real(*a, **k))[1])
    monkeypatch.setattr(K, "DROPOUT_HOOK", _hook_from_masks(K, [mask_t, mask_s]))
    cls_tea, score = tea.forward_teacher(x.cuda())
    assert len(calls) == 1 and calls[0] is not None and calls[0].p == 0.25        # the fused kernel ran, with dropout
    with torch.no_grad():
        rc, rs = O.mhim_forward_teacher(cfg, sd_t, x, drop_mask=mask_t)
    assert cases.rel_err(cls_tea, rc) < TOL and cases.rel_err(score, rs) < TOL
    # student pass on the reference's scores, same merge noise on both sides
    torch.manual_seed(11)
    stu.merge._noise = lambda L, dev: torch.rand(L).to(dev)                      # the CPU random stream the oracle consumes (merge.py:164)
    logits, loss, ps, lk = stu(x.cuda(), rs.cuda(), rc.cuda(), i=0)
    sd_ref = {k: v.clone().requires_grad_(v.is_floating_point()) for k, v in sd_s.items()}
    torch.manual_seed(11)
    olg, oloss, ops_, olk, newq, ids = O.mhim_forward(cfg, sd_ref, x, rs, rc, i=0, training=True, drop_mask=mask_s)
    assert (ps, lk) == (ops_, olk)
    assert cases.rel_err(logits, olg) < TOL and cases.rel_err(loss, oloss) < TOL
    (F.cross_entropy(logits, torch.tensor([1]).cuda()) + 0.5 * loss).backward()
    (F.cross_entropy(olg, torch.tensor([1])) + 0.5 * oloss).backward()
    for k, p_ in stu.named_parameters():
        if p_.grad is not None and sd_ref[k].grad is not None and float(sd_ref[k].grad.abs().max()) > 0:
            assert cases.rel_err(p_.grad, sd_ref[k].grad) < TOL, k


def test_mhim_philox_teacher_is_deterministic_under_manual_seed(K):
    from mhimk import modules as M
    kw = dict(cases.MHIM_KW, baseline="attn", input_dim=1024, dropout=0.25)
    tea = M.MHIM(**kw).cuda().train()
    x = cases.make_bag(1, 2000, 1024).cuda()
    torch.manual_seed(5)
    a = tea.forward_teacher(x)[1].clone()
    b = tea.forward_teacher(x)[1].clone()
    torch.manual_seed(5)
    c = tea.forward_teacher(x)[1].clone()
    assert torch.equal(a, c) and not torch.equal(a, b)
    tea.eval()
    assert torch.equal(tea.forward_teacher(x)[1], tea.forward_teacher(x)[1])      # eval: no dropout


def test_dattention_train_mode_dropout(K, monkeypatch):
    """abmil.DAttention(dropout=True) in train mode: hard-coded p = 0.25 after the feature activation (abmil.py:188-189)."""
    from mhimk import modules as M
    n = 1500
    sd, x = cases.abmil_state(9), cases.make_bag(10, n, 1024)
    # gelu: a ReLU gate that flips (pre ~ 0 within the 7.6e-6 unit roundoff of the hi/lo arithmetic) in one high-attention instance moves
    # the feature gradient by 5e-3 -- seed-dependent, and true of any two fp32 implementations (SURVEY 9.9 "gate flips")
    m = M.DAttention(1024, 2, dropout=True, act="gelu").cuda().train()
    m.load_state_dict({k: v.cuda() for k, v in sd.items()}, strict=True)
    torch.manual_seed(1)
    mask = F.dropout(torch.ones(n, 512), 0.25, True)
    monkeypatch.setattr(K, "DROPOUT_HOOK", _hook_from_masks(K, [mask, mask]))
    with torch.no_grad():
        lg = m(x.cuda())                                                         # fused kernel with dropout
    sd_ref = {k: v.clone().requires_grad_() for k, v in sd.items()}
    ref = O.abmil_dattention(sd_ref, x, "gelu", drop_mask=mask)
    assert cases.rel_err(lg, ref) < TOL
    lg2 = m(x.cuda())                                                            # composed path with CUDA backward
    assert cases.rel_err(lg2, ref) < TOL
    F.cross_entropy(lg2, torch.tensor([1]).cuda()).backward()
    F.cross_entropy(ref, torch.tensor([1])).backward()
    for k, p_ in m.named_parameters():
        if k == "attention.2.bias":                                              # mathematically zero (softmax is shift-invariant)
            assert float(p_.grad.abs().max()) < 1e-6
            continue
        assert cases.rel_err(p_.grad, sd_ref[k].grad) < TOL, k
