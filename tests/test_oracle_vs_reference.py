"""Oracle against the LIVE reference classes (skipped where /root/reference is absent, e.g. the GPU box)."""
import pytest
import torch
import torch.nn.functional as F

import cases
from _refload import have_reference, load_reference, zero_dropout
from oracle import mil_oracle as O

pytestmark = pytest.mark.skipif(not have_reference(), reason="/root/reference not present")


@pytest.fixture(scope="module")
def R():
    return load_reference()


@pytest.mark.parametrize("N", [2, 5, 33, 255, 256, 257, 1000])
@pytest.mark.parametrize("act", ["relu", "gelu"])
def test_abmil_and_gated(R, N, act):
    x = cases.make_bag(N, N, 1024)
    sd = cases.abmil_state(N + 1)
    m = R.abmil.DAttention(1024, 2, dropout=0.0, act=act).eval()
    m.load_state_dict(sd, strict=True)
    ref = m(x.clone(), return_attn=True, return_act=True)
    got = O.abmil_dattention(sd, x, act, return_attn=True, return_act=True)
    for a, b in zip(got, ref):
        assert cases.rel_err(a, b) <= 1e-6
    sg = cases.gated_state(N + 2)
    mg = R.abmil.AttentionGated(1024, 2, act=act, dropout=0.0).eval()
    mg.load_state_dict(sg, strict=True)
    assert cases.rel_err(O.abmil_gated(sg, x, act), mg(x.clone())) <= 1e-6


@pytest.mark.parametrize("base,N", [("attn", 5), ("attn", 33), ("attn", 4099), ("dsmil", 257), ("selfattn", 255), ("selfattn", 513)])
def test_mhim_entry_points(R, base, N):
    d = 1536 if base == "dsmil" else 1024
    kw = dict(cases.MHIM_KW, baseline=base, input_dim=d, dropout=0.0)
    cfg = O.MHIMConfig(**dict(cases.MHIM_KW, baseline=base, input_dim=d))
    sd_s, sd_t = cases.mhim_state(N, base, D=d), cases.mhim_state(N + 1, base, D=d)
    stu, tea = zero_dropout(R.mhim.MHIM(**kw)), zero_dropout(R.mhim.MHIM(**kw))
    stu.load_state_dict(sd_s, strict=True)
    tea.load_state_dict(sd_t, strict=True)
    stu.train(), tea.train()
    x = cases.make_bag(N + 5, N, d)
    tol = 5e-6 if base == "selfattn" else 1e-6
    ct, sc = tea.forward_teacher(x)
    oct_, osc = O.mhim_forward_teacher(cfg, sd_t, x)
    assert cases.rel_err(oct_, ct) <= tol and cases.rel_err(osc, sc) <= tol
    tcf = ct[0] if base == "dsmil" else ct
    torch.manual_seed(9)
    lg, loss, ps, lk = stu(x, sc, tcf, i=0)
    torch.manual_seed(9)
    olg, oloss, ops, olk, newq, ids = O.mhim_forward(cfg, sd_s, x, sc, tcf, i=0, training=True)
    assert (ps, lk) == (ops, olk)
    pairs = zip(olg, lg) if base == "dsmil" else [(olg, lg)]
    for a, b in pairs:
        assert cases.rel_err(a, b) <= tol
    assert cases.rel_err(oloss, loss) <= tol
    assert cases.rel_err(newq, stu.merge.global_q_mm.data) <= 1e-6


@pytest.mark.parametrize("ps,ratio,hr,largest", [(1000, 0.03, 1.0, True), (1000, 0.03, 0.3, True), (4099, 0.01, 1.0, True),
                                                 (257, 0.9, 0.5, True), (100, 0.05, 1.0, False), (2, 0.03, 1.0, True)])
def test_select_mask(R, ps, ratio, hr, largest):
    attn = torch.rand(1, ps, generator=torch.Generator().manual_seed(ps))
    torch.manual_seed(1)
    lk, ids = R.masking.select_mask_fn(ps, attn, largest, ratio, len_keep_other=ps, random_ratio=hr)
    torch.manual_seed(1)
    olk, oids = O.select_mask(ps, attn, largest, ratio, len_keep_other=ps, random_ratio=hr)
    assert lk == olk and torch.equal(ids[0, lk:], oids[0, olk:])          # masked ids: exact, same order
    if lk * 4 >= ps:
        assert torch.equal(ids, oids)
    else:
        # When most of the bag is masked (only reachable through the ratio/hr > 1 clamp, masking.py:33-35) the
        # reference's kept ids come out in CPython set-table order (hash mod table size), not ascending.  The
        # oracle and the product always return them ascending; same set.
        assert torch.equal(ids[0, :lk].sort().values, oids[0, :olk])


def test_vote_path_and_score_identity(R):
    attn = torch.rand(1, 8, 600, generator=torch.Generator().manual_seed(2))
    lk, ids = R.masking.select_mask_fn(600, attn, True, 0.03, len_keep_other=600, random_ratio=1.0)
    olk, oids = O.select_mask(600, attn, True, 0.03, len_keep_other=600, random_ratio=1.0)
    assert lk == olk and torch.equal(ids, oids)
    # for C=2 the CAM score is sigmoid(|a_n h_n.(W0-W1)|); the predictor bias never matters
    h, a = torch.randn(300, 512), torch.softmax(torch.randn(300), 0)
    w, b = torch.randn(2, 512) * 0.05, torch.randn(2)
    s = O.pseudo_score(w, b, h, a)
    s2 = torch.sigmoid(((h * a[:, None]) @ (w[0] - w[1])).abs())
    assert cases.rel_err(s2, s) < 1e-6
    assert torch.equal(O.pseudo_score(w, b * 0, h, a), s) or cases.rel_err(O.pseudo_score(w, b * 0, h, a), s) < 1e-6


def test_transmil_and_milnet(R):
    sd, x = cases.transmil_state(3), cases.make_bag(4, 1000, 1024)
    m = zero_dropout(R.transmil.TransMIL(1024, 2, dropout=0.0, act="relu")).eval()
    m.load_state_dict(sd, strict=True)
    assert cases.rel_err(O.transmil_forward(sd, x, "relu"), m(x)) <= 5e-6
    sd, x = cases.milnet_state(5), cases.make_bag(6, 400, 1536)
    d = R.dsmil.MILNet(2, 0.0, "gelu", input_dim=1536).eval()
    d.load_state_dict(sd, strict=True)
    rp, rc = d(x)
    op, oc, _, _ = O.milnet_forward(sd, x, "gelu")
    assert cases.rel_err(op, rp) <= 1e-6 and cases.rel_err(oc, rc) <= 1e-6


@pytest.mark.parametrize("base,N", [("attn", 257), ("dsmil", 129)])
def test_feature_dropout_mask_semantics(R, base, N):
    """The reference's own training configuration (dropout=0.25, teacher kept in train(), engines/base_engine.py:36-37): with the
    global torch seed fixed, `self.dp` draws exactly F.dropout(ones(1,N,512)) -- the oracle's `drop_mask` argument reproduces the
    train-mode teacher and student passes of the live classes."""
    d = 1536 if base == "dsmil" else 1024
    kw = dict(cases.MHIM_KW, baseline=base, input_dim=d, dropout=0.25)
    cfg = O.MHIMConfig(**dict(cases.MHIM_KW, baseline=base, input_dim=d))
    sd_s, sd_t = cases.mhim_state(N, base, D=d), cases.mhim_state(N + 1, base, D=d)
    stu, tea = R.mhim.MHIM(**kw), R.mhim.MHIM(**kw)
    for m in (stu, tea):                                               # keep the feature dropout, neutralise the others (merge / attention)
        for name, mod in m.named_modules():
            if isinstance(mod, torch.nn.Dropout) and name != "dp":
                mod.p = 0.0
    stu.load_state_dict(sd_s, strict=True)
    tea.load_state_dict(sd_t, strict=True)
    stu.train(), tea.train()
    x = cases.make_bag(N + 5, N, d)
    torch.manual_seed(123)
    ct, sc = tea.forward_teacher(x)
    torch.manual_seed(123)
    mask = torch.nn.functional.dropout(torch.ones(1, N, 512), 0.25, True)[0]
    oct_, osc = O.mhim_forward_teacher(cfg, sd_t, x, drop_mask=mask)
    assert cases.rel_err(oct_, ct) <= 1e-6 and cases.rel_err(osc, sc) <= 1e-6
    assert 0.70 < float((mask > 0).float().mean()) < 0.80 and torch.all((mask == 0) | ((mask - 1 / 0.75).abs() < 1e-6))
    tcf = ct[0] if base == "dsmil" else ct
    torch.manual_seed(9)
    lg, loss, ps, lk = stu(x, sc, tcf, i=0)
    torch.manual_seed(9)
    mask_s = torch.nn.functional.dropout(torch.ones(1, N, 512), 0.25, True)[0]      # the student's dp draws first (mhim.py:331-336)
    olg, oloss, ops, olk, newq, ids = O.mhim_forward(cfg, sd_s, x, sc, tcf, i=0, training=True, drop_mask=mask_s)
    pairs = zip(olg, lg) if base == "dsmil" else [(olg, lg)]
    for a, b in pairs:
        assert cases.rel_err(a, b) <= 1e-6
    assert cases.rel_err(oloss, loss) <= 1e-6


@pytest.mark.parametrize("base,N", [("attn", 2000), ("dsmil", 700)])
def test_full_gradient_tensors_against_the_live_reference(R, base, N):
    """Every gradient tensor of the student pass, element for element, against the live reference's autograd -- including
    merge.norm.weight, whose reference value is taken with the ALREADY EMA-updated global_q (the in-forward `.data` write of
    merge.py:127-129 precedes backward); the oracle reproduces that (oracle/mil_oracle.py:_LayerNormInputReadAtBackward)."""
    import torch.nn.functional as F
    d = 1536 if base == "dsmil" else 1024
    kw = dict(cases.MHIM_KW, baseline=base, input_dim=d, dropout=0.0)
    cfg = O.MHIMConfig(**dict(cases.MHIM_KW, baseline=base, input_dim=d))
    sd_s, sd_t = cases.mhim_state(N, base, D=d), cases.mhim_state(N + 1, base, D=d)
    stu, tea = zero_dropout(R.mhim.MHIM(**kw)), zero_dropout(R.mhim.MHIM(**kw))
    stu.load_state_dict(sd_s, strict=True)
    tea.load_state_dict(sd_t, strict=True)
    stu.train(), tea.train()
    x = cases.make_bag(N + 5, N, d)
    ct, sc = tea.forward_teacher(x)
    tcf = ct[0] if base == "dsmil" else ct
    torch.manual_seed(9)
    lg, loss, _, _ = stu(x, sc, tcf, i=0)
    lt = 0.5 * lg[0].view(1, -1) + 0.5 * lg[1].view(1, -1) if base == "dsmil" else lg
    (F.cross_entropy(lt, torch.tensor([1])) + 0.5 * loss).backward()
    sd_ref = {k: v.clone().requires_grad_(v.is_floating_point()) for k, v in sd_s.items()}
    torch.manual_seed(9)
    olg, oloss, *_ = O.mhim_forward(cfg, sd_ref, x, sc, tcf, i=0, training=True)
    olt = 0.5 * olg[0].view(1, -1) + 0.5 * olg[1].view(1, -1) if base == "dsmil" else olg
    (F.cross_entropy(olt, torch.tensor([1])) + 0.5 * oloss).backward()
    n = 0
    for k, p in stu.named_parameters():
        if p.grad is None or k not in sd_ref or sd_ref[k].grad is None or float(p.grad.abs().max()) == 0:
            continue
        assert cases.rel_err(sd_ref[k].grad, p.grad) <= 2e-5, (k, cases.rel_err(sd_ref[k].grad, p.grad))
        n += 1
    assert n >= 10


@pytest.mark.parametrize("mil_norm,pos_", [("ln", 0), ("ln", 1), ("bn", 0), ("bn", 1)])
def test_dattention_mil_norm_and_sincos(R, mil_norm, pos_):
    """abmil.DAttention's legal non-default configurations (abmil.py:162-178, 207-223): mil_norm bn / ln at both positions (eval mode: the
    reference's BatchNorm of a single pooled vector raises in train mode) and pos='sincos'."""
    N = 333
    x = cases.make_bag(N, N, 1024)
    sd = cases.abmil_norm_state(N + pos_, mil_norm, pos_)
    m = R.abmil.DAttention(1024, 2, dropout=0.0, act="gelu", mil_norm=mil_norm, embed_norm_pos=pos_).eval()
    m.load_state_dict(sd, strict=True)
    ref = m(x.clone(), return_attn=True, return_act=True)
    got = O.abmil_dattention(sd, x, "gelu", return_attn=True, return_act=True, mil_norm=mil_norm, embed_norm_pos=pos_)
    for a, b in zip(got, ref):
        assert cases.rel_err(a, b) <= 2e-6
    if mil_norm == "ln" and pos_ == 0:
        sd0, pos = cases.abmil_state(N + 9), cases.sincos_pos(5, N)
        ms = R.abmil.DAttention(1024, 2, dropout=0.0, act="relu", pos="sincos").eval()
        ms.load_state_dict(sd0, strict=True)
        assert cases.rel_err(O.abmil_dattention(sd0, x, "relu", pos=pos), ms(x.clone(), pos=pos)) <= 2e-6


@pytest.mark.parametrize("distill", ["AFS", "MaxS", "MaxMinS"])
@pytest.mark.parametrize("N", [7, 333, 2000])
def test_dtfd_train_and_test_forward(R, distill, N):
    """DTFD-MIL (modules/dtfd.py:148-272), both forwards, all three distillations, dropout neutralised; eval mode shuffles with python's RNG."""
    import importlib
    import random
    D = importlib.import_module("modules.dtfd")
    sd, x = cases.dtfd_state(N), cases.make_bag(N + 3, N, 1024)[0]
    m = zero_dropout(D.DTFD(torch.device("cpu"), 1e-4, 1e-5, 10, distill=distill))
    m.load_state_dict(sd, strict=True)
    m.train()
    assert cases.rel_err(O.dtfd_forward(sd, x, True, distill=distill), m(x[None])) <= 2e-6
    m.eval()
    random.seed(5)
    ref = m(x[None])
    random.seed(5)
    ids = list(range(N))
    random.shuffle(ids)
    assert cases.rel_err(O.dtfd_forward(sd, x, False, distill=distill, test_ids=ids), ref) <= 2e-6


@pytest.mark.parametrize("mb", [False, True])
@pytest.mark.parametrize("gate,subtyping,n_cls", [(True, False, 2), (True, True, 3), (False, False, 2)])
def test_clam_forward_instance_loss_and_gradients(R, mb, gate, subtyping, n_cls):
    """CLAM_SB / CLAM_MB (modules/clam.py:93-331) incl. the instance-level branch with the SmoothTop1SVM loss (modules/topk/svm.py:84-108): logits,
    instance loss and EVERY gradient tensor of `CE(logits) + instance_loss` against the live classes (dropout 0; act relu and gelu)."""
    from _refload import load_clam
    clam = load_clam()
    for act, N, seed in (("relu", 300, 3), ("gelu", 1200, 4)):
        sd = cases.clam_state(seed, mb, C=n_cls, gate=gate)
        x = cases.make_bag(seed + 10, N, 1024)[0]
        cls = clam.CLAM_MB if mb else clam.CLAM_SB
        m = cls(input_dim=1024, gate=gate, n_classes=n_cls, subtyping=subtyping, act=act, dropout=0.0)
        m.load_state_dict(sd, strict=True)
        m.eval()
        lg_ref = m(x[None])
        lg, _, araw = O.clam_forward(sd, x, mb, n_cls, gate, act, subtyping=subtyping)
        assert cases.rel_err(lg, lg_ref) <= 2e-6
        assert cases.rel_err(araw, m(x[None], attention_only=True).reshape(araw.shape)) <= 2e-6
        m.train()
        label = torch.tensor([1])
        lg_ref, il_ref, ps = m(x[None], label=label, instance_eval=True)
        (F.cross_entropy(lg_ref, label) + il_ref).backward()
        sdl = {k: v.clone().requires_grad_(v.is_floating_point()) for k, v in sd.items()}
        lg, il, _ = O.clam_forward(sdl, x, mb, n_cls, gate, act, subtyping=subtyping, label=1)
        (F.cross_entropy(lg, label) + il).backward()
        assert ps == N and cases.rel_err(lg, lg_ref) <= 2e-6 and cases.rel_err(il, il_ref) <= 2e-6
        for k, p_ in m.named_parameters():
            if p_.grad is None:
                assert sdl[k].grad is None or float(sdl[k].grad.abs().max()) == 0.0, k
            elif float(p_.grad.abs().max()) < 1e-6:            # analytically zero (a bias under a softmax over N): rounding noise on both sides
                assert float(sdl[k].grad.abs().max()) < 1e-6, k
            else:
                assert cases.rel_err(sdl[k].grad, p_.grad) <= 2e-5, k


def test_smooth_top1_svm_hard_and_smooth_rows(R):
    """Rows whose top-2 gap exceeds tau * log(1e3) take the hard max-margin form, the others the smooth one (utils.py:36-42)."""
    from _refload import load_clam
    load_clam()
    import importlib
    svm = importlib.import_module("modules.topk.svm")
    g = torch.Generator().manual_seed(0)
    x = torch.randn(64, 2, generator=g) * 6
    y = torch.randint(0, 2, (64,), generator=g)
    ref = svm.SmoothTop1SVM(2)(x, y)
    assert cases.rel_err(O.smooth_top1_svm(x, y), ref) <= 1e-6
    assert bool(((x.max(1).values - x.min(1).values) >= 6.9077).any()) and bool(((x.max(1).values - x.min(1).values) < 6.9077).any())


@pytest.mark.parametrize("mb", [False, True])
@pytest.mark.parametrize("kw", [dict(), dict(dropout=0.25), dict(gate=False, dropout=0.25), dict(size_arg="big", n_classes=4), dict(gate=False)])
def test_clam_dropin_has_the_reference_state_dict(R, mb, kw):
    """A reference CLAM checkpoint loads into the drop-in with strict=True (same keys and shapes, `instance_loss_fn.labels` included)."""
    import mhimk  # noqa: F401
    from mhimk.modules import clam as mine
    from _refload import load_clam
    ref = load_clam()
    a = (ref.CLAM_MB if mb else ref.CLAM_SB)(input_dim=1024, **kw)
    b = (mine.CLAM_MB if mb else mine.CLAM_SB)(input_dim=1024, **kw)
    sa, sb = a.state_dict(), b.state_dict()
    assert list(sa.keys()) == list(sb.keys())
    assert all(sa[k].shape == sb[k].shape and sa[k].dtype == sb[k].dtype for k in sa)
    b.load_state_dict(sa, strict=True)
