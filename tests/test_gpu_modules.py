"""GPU parity of the drop-in modules (same constructors / forward signatures / state_dict keys as the reference) against the
committed golden vectors made from the live reference and against the CPU oracle."""
import os
import types

import pytest
import torch
import torch.nn.functional as F

import cases
from oracle import mil_oracle as O

pytestmark = pytest.mark.gpu
G = torch.load(os.path.join(os.path.dirname(__file__), "golden", "golden_v1.pt"), weights_only=False)
TOL = 1e-4            # north-star gate: logits and gradients within 1e-4 relative (max|d| / max|ref| per tensor)


@pytest.fixture(scope="module")
def M():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import mhimk
    from mhimk import modules
    return modules


def cuda_sd(sd):
    return {k: v.cuda() for k, v in sd.items()}


def check_grads(model, golden, tol=TOL):
    for k, p in model.named_parameters():
        if k not in golden:
            continue
        g = golden[k]
        assert p.grad is not None, k
        if g["norm"] < 1e-6:
            assert p.grad.norm().item() < 1e-4, k
            continue
        assert abs(p.grad.double().norm().item() - g["norm"]) <= tol * g["norm"], (k, p.grad.double().norm().item(), g["norm"])
        scale = max(g["head"].abs().max().item(), g["norm"] / (p.numel() ** 0.5))
        assert float((p.grad.flatten()[:8].cpu() - g["head"]).abs().max()) <= 10 * tol * scale, k


@pytest.mark.parametrize("name", list(G["abmil"]))
def test_dattention_eval_fused_and_train(M, name):
    act, n, kind = name.split("_")
    i = list(G["abmil"]).index(name)
    g = G["abmil"][name]
    sd, x = cases.abmil_state(11 + i), cases.make_bag(11 + i + 1000, int(n), 1024, kind)
    m = M.DAttention(1024, 2, dropout=0.0, act=act).cuda()
    m.load_state_dict(cuda_sd(sd), strict=True)
    m.eval()
    with torch.no_grad():                                             # fused tcgen05 path
        logits, attn, actv = m(x.cuda(), return_attn=True, return_act=True)
        lg2, feat = m(x.cuda(), return_img_feat=True)
    assert cases.rel_err(logits, g["logits"]) < TOL and cases.rel_err(lg2, g["logits"]) < TOL
    assert cases.rel_err(feat, g["pooled"]) < TOL
    assert cases.rel_err(attn[0, :16], g["attn_head"]) < 3 * TOL
    assert tuple(actv.shape) == (1, int(n), 512)
    m.train()                                                          # composed path with CUDA backward
    logits = m(x.cuda())
    assert cases.rel_err(logits, g["logits"]) < TOL
    F.cross_entropy(logits, torch.tensor([1]).cuda()).backward()
    check_grads(m, g["grads"])


@pytest.mark.parametrize("name", list(G["gated"]))
def test_attention_gated_fwd_bwd(M, name):
    """BASELINE config 0: ABMIL gated-attention fwd/bwd on a synthetic bag (N=1024, D=1024)."""
    act, n, kind = name.split("_")
    i = list(G["gated"]).index(name)
    g = G["gated"][name]
    sd, x = cases.gated_state(31 + i), cases.make_bag(31 + i + 1000, int(n), 1024, kind)
    m = M.AttentionGated(1024, 2, act=act, dropout=0.0).cuda()
    m.load_state_dict(cuda_sd(sd), strict=True)
    m.train()
    x2 = x[0].cuda()                                                   # 2-D input is unsqueezed in place like the reference
    logits = m(x2)
    assert x2.dim() == 3
    assert cases.rel_err(logits, g["logits"]) < TOL
    F.cross_entropy(logits, torch.tensor([1]).cuda()).backward()
    check_grads(m, g["grads"])


MHIM_CASES = {"attn_2000": ("attn", 2000, 1024, 51), "attn_33": ("attn", 33, 1024, 52), "dsmil_1000": ("dsmil", 1000, 1536, 61),
              "selfattn_600": ("selfattn", 600, 1024, 71)}


def build_mhim(M, base, d, seed):
    m = M.MHIM(**dict(cases.MHIM_KW, baseline=base, input_dim=d, dropout=0.0)).cuda()
    m.load_state_dict(cuda_sd(cases.mhim_state(seed, base, D=d)), strict=True)
    for mod in m.modules():
        if isinstance(mod, torch.nn.Dropout):
            mod.p = 0.0                                                # golden was made with every dropout neutralised
    return m


@pytest.mark.parametrize("name", list(MHIM_CASES))
def test_mhim_teacher_student_test_pure(M, name):
    base, n, d, seed = MHIM_CASES[name]
    g = G["mhim"][name]
    tol = TOL
    stu, tea = build_mhim(M, base, d, seed), build_mhim(M, base, d, seed + 1)
    stu.train(), tea.train()
    x = cases.make_bag(seed + 1000, n, d).cuda()
    cls_tea, score = tea.forward_teacher(x)
    assert cases.rel_err(cls_tea, g["cls_tea"]) < tol
    assert cases.rel_err(score, g["score"]) < tol
    # index parity is defined on equal inputs: feed the reference's own fp32 scores
    score_ref = g["score"].cuda()
    lk, ids = stu.get_mask(n, 0, score_ref)
    assert lk == g["mask_len_keep"] and cases.tensor_digest(ids) == g["mask_ids_digest"]      # bit-exact mask indices
    torch.manual_seed(seed + 7)
    stu.merge._noise = lambda L, dev: torch.rand(L).to(dev)            # the CPU random stream the reference consumed
    tcf = g["cls_tea"].cuda()
    tcf = tcf[0] if base == "dsmil" else tcf
    logits, loss, ps, len_keep = stu(x, score_ref, tcf, i=0)
    assert (ps, len_keep) == (g["ps"], g["len_keep"])
    if base == "dsmil":
        for a, b in zip(logits, g["logits"]):
            assert cases.rel_err(a, b) < tol
        lt = 0.5 * logits[0].view(1, -1) + 0.5 * logits[1].view(1, -1)
    else:
        assert cases.rel_err(logits, g["logits"]) < tol
        lt = logits
    assert cases.rel_err(loss, g["loss"]) < tol
    assert cases.rel_err(stu.merge.global_q_mm.data[0, :, :8], g["new_global_q_head"]) < tol
    (F.cross_entropy(lt, torch.tensor([1]).cuda()) + 0.5 * loss).backward()
    check_grads(stu, g["grads"], tol=TOL)
    stu.eval()
    stu.merge.global_q_mm.data.copy_(cases.mhim_state(seed, base, D=d)["merge.global_q_mm"].cuda())
    ft, pu = stu.forward_test(x), stu.pure(x)
    if base == "dsmil":
        for a, b in zip(ft[0], g["forward_test"]):
            assert cases.rel_err(a, b) < tol
        for a, b in zip(pu, g["pure_eval"]):
            assert cases.rel_err(a, b) < tol
    else:
        assert cases.rel_err(ft, g["forward_test"]) < tol and cases.rel_err(pu, g["pure_eval"]) < tol


def test_mhim_teacher_scores_topk_tie_aware(M):
    """End to end the CAM scores sit on a few hundred fp32 values around 0.5: the kept set computed from OUR
    scores must agree with the one from the reference scores wherever values are strictly separated."""
    base, n, d, seed = MHIM_CASES["attn_2000"]
    g = G["mhim"]["attn_2000"]
    tea = build_mhim(M, base, d, seed + 1).eval()
    _, score = tea.forward_teacher(cases.make_bag(seed + 1000, n, d).cuda())
    ref = g["score"][0]
    k = O.topk_count(n, 0.03)
    mine = set(torch.topk(score[0].cpu(), k).indices.tolist())
    thr = torch.topk(ref, k).values.min()
    margin = 4 * 5.96e-8
    must_have = set(torch.nonzero(ref > thr + margin).flatten().tolist())
    may_have = set(torch.nonzero(ref >= thr - margin).flatten().tolist())
    assert must_have <= mine <= may_have


def test_transmil_and_milnet_eval(M):
    g = G["transmil"]["700"]
    t = M.TransMIL(1024, 2, dropout=0.0, act="relu").cuda().eval()
    t.load_state_dict(cuda_sd(cases.transmil_state(81)), strict=True)
    for mod in t.modules():
        if isinstance(mod, torch.nn.Dropout):
            mod.p = 0.0
    with torch.no_grad():
        logits, attn, v = t(cases.make_bag(1081, 700, 1024).cuda(), return_attn=True, return_act=True)
    assert cases.rel_err(logits, g["logits"]) < TOL
    assert cases.rel_err(attn[0][0, :, :8], g["attn0_head"]) < TOL and cases.rel_err(v[0, :, :2, :4], g["v_head"]) < TOL
    g = G["milnet"]["500"]
    d = M.MILNet(2, 0.0, "relu", input_dim=1536).cuda().eval()
    d.load_state_dict(cuda_sd(cases.milnet_state(85)), strict=True)
    with torch.no_grad():
        pred, classes = d(cases.make_bag(1085, 500, 1536).cuda())
    assert cases.rel_err(pred, g["pred"]) < TOL and cases.rel_err(classes, g["classes"]) < TOL


def test_common_mil_adapter_tuple(M):
    import mhimk
    from mhimk.engines import CommonMIL
    base, n, d, seed = "attn", 500, 1024, 5
    stu, tea = build_mhim(M, base, d, seed), build_mhim(M, base, d, seed + 1)
    stu.train(), tea.train()
    args = types.SimpleNamespace(model="mhim", baseline="attn", aux_alpha=0.5)
    bag, label = cases.make_bag(9, n, d).cuda(), torch.tensor([1]).cuda()
    out = CommonMIL(args).forward_func(args, stu, tea, bag, label, torch.nn.CrossEntropyLoss(), 1, 0, 0, 0, None)
    logits, lab, aux, patch_num, keep_num, pad_ratio, kn_std = out
    assert tuple(logits.shape) == (1, 2) and patch_num == n and keep_num == int((n - O.topk_count(n, 0.03)) * 0.8) + 5
    assert float(aux) > 0 and pad_ratio == 0.0 and kn_std == 0.0
    stu.eval()
    lg, _ = CommonMIL(args).validate_func(args, stu, bag, label, None, 1, 0, None)
    assert tuple(lg.shape) == (1, 2)
    args2 = types.SimpleNamespace(model="abmil", baseline="attn", aux_alpha=0.0)
    ab = M.DAttention(1024, 2, dropout=0.0, act="relu").cuda().eval()
    out2 = CommonMIL(args2).forward_func(args2, ab, None, bag, label, None, 1, 0, 0, 0, None)
    assert tuple(out2[0].shape) == (1, 2) and out2[3] == n


def test_fused_refuses_stale_weight_images(M):
    """The cached 16-bit weight images must follow in-place weight updates (optimizer steps)."""
    m = M.DAttention(1024, 2, dropout=0.0, act="relu").cuda().eval()
    x = cases.make_bag(3, 300, 1024).cuda()
    with torch.no_grad():
        a = m(x).clone()
        m.feature[0].weight.mul_(0.5)
        b = m(x)
        sd = {k: v.detach().cpu() for k, v in m.state_dict().items()}
    assert cases.rel_err(b, O.abmil_dattention(sd, x.cpu(), "relu")) < TOL
    assert cases.rel_err(a, b) > 1e-3


def _ema_through_data(student, teacher, mm):
    """The reference's teacher update, verbatim in effect (engines/base_engine.py:166-167): in-place writes through `.data`,
    which autograd's version counter does not see."""
    for param_q, param_k in zip(student.parameters(), teacher.parameters()):
        param_k.data.mul_(mm).add_(param_q.data, alpha=1. - mm)


@pytest.mark.parametrize("base", ["attn", "dsmil"])
def test_teacher_follows_ema_updates_written_through_data(M, base):
    """Cached 16-bit weight images must not survive the EMA update: train-mode forwards rebuild them, and the train->eval switch
    invalidates them (the last EMA update of an epoch comes after the last train-mode forward)."""
    _, n, d, seed = MHIM_CASES["attn_2000" if base == "attn" else "dsmil_1000"]
    cfg = O.MHIMConfig(**dict(cases.MHIM_KW, baseline=base, input_dim=d))
    stu, tea = build_mhim(M, base, d, seed).train(), build_mhim(M, base, d, seed + 1).train()
    x = cases.make_bag(seed + 1000, n, d).cuda()

    def oracle_teacher():
        sd = {k: v.detach().cpu() for k, v in tea.state_dict().items()}
        with torch.no_grad():
            return O.mhim_forward_teacher(cfg, sd, x.cpu()), O.mhim_forward_test(cfg, sd, x.cpu())

    before, _ = tea.forward_teacher(x)
    before = before.clone()
    _ema_through_data(stu, tea, 0.5)                                   # a large step so that stale images would be far off
    (cls_ref, score_ref), _ = oracle_teacher()
    cls_tea, score = tea.forward_teacher(x)                            # train mode: images rebuilt
    assert cases.rel_err(cls_tea, cls_ref) < TOL and cases.rel_err(score, score_ref) < TOL
    assert cases.rel_err(before, cls_ref) > 1e-2
    _ema_through_data(stu, tea, 0.5)                                   # the epoch's last update, then validation with the teacher
    _, test_ref = oracle_teacher()
    tea.eval()
    got = tea.forward_test(x)
    if base == "dsmil":
        for a, b in zip(got[0], test_ref[0]):
            assert cases.rel_err(a, b) < TOL
    else:
        assert cases.rel_err(got, test_ref) < TOL


def test_eval_mode_data_write_needs_weights_touched(M):
    """In eval mode the images are cached across calls (the headline path); a write through `.data` there is invisible until
    ops.weights_touched() (documented in INTEGRATION.md)."""
    import mhimk
    m = M.DAttention(1024, 2, dropout=0.0, act="relu").cuda().eval()
    x = cases.make_bag(3, 300, 1024).cuda()
    with torch.no_grad():
        m(x)
        m.feature[0].weight.data.mul_(0.5)
        mhimk.ops.weights_touched()
        b = m(x)
    sd = {k: v.detach().cpu() for k, v in m.state_dict().items()}
    assert cases.rel_err(b, O.abmil_dattention(sd, x.cpu(), "relu")) < TOL


def same_fma_result(got, want):
    """Bit-equal, except that the oracle's fused multiply-add goes through a double (two roundings): allow a couple of 1-ulp cases."""
    return int((got != want).sum()) <= 2 and torch.allclose(got, want, rtol=2.5e-7, atol=1e-37)


@pytest.mark.parametrize("mm", [0.9999, 0.5])
def test_ema_update_one_launch(M, mm):
    """mhimk.engines.ema_update == the reference's per-parameter loop (oracle, bit for bit), and the teacher's next forward sees
    the new weights even in eval mode (the helper notifies the weight-image caches)."""
    from mhimk.engines import ema_update
    _, n, d, seed = MHIM_CASES["attn_2000"]
    stu, tea = build_mhim(M, "attn", d, seed), build_mhim(M, "attn", d, seed + 1).eval()
    x = cases.make_bag(seed + 1000, n, d).cuda()
    tea.forward_test(x)                                                  # fills the image cache with the old weights
    want = O.ema_update(list(stu.parameters()), list(tea.parameters()), mm)
    ema_update(stu, tea, mm)
    for p, w in zip(tea.parameters(), want):
        assert same_fma_result(p.detach().cpu(), w)
    cfg = O.MHIMConfig(**dict(cases.MHIM_KW, baseline="attn", input_dim=d))
    sd = {k: v.detach().cpu() for k, v in tea.state_dict().items()}
    with torch.no_grad():
        ref = O.mhim_forward_test(cfg, sd, x.cpu())
    assert cases.rel_err(tea.forward_test(x), ref) < TOL
    with pytest.raises(AssertionError):
        ema_update(stu, tea, 1.5)


def test_ema_update_odd_sizes_and_alignment(M):
    """Segments that are not multiples of 4 elements or not 16-byte aligned take the scalar path; > 32768 elements span CTAs."""
    from mhimk.engines import ema_update

    class Bag(torch.nn.Module):
        def __init__(self, seed):
            super().__init__()
            g = torch.Generator().manual_seed(seed)
            self.a = torch.nn.Parameter(torch.randn(70001, generator=g))
            self.b = torch.nn.Parameter(torch.randn(3, generator=g))
            self.c = torch.nn.Parameter(torch.randn(1, generator=g))
            self.d = torch.nn.Parameter(torch.randn(257, 129, generator=g))

    q, k = Bag(1).cuda(), Bag(2).cuda()
    k.a.data = torch.randn(70002, generator=torch.Generator().manual_seed(3)).cuda()[1:]      # 4-byte aligned only
    want = O.ema_update(list(q.parameters()), list(k.parameters()), 0.99)
    ema_update(q, k, 0.99)
    for p, w in zip(k.parameters(), want):
        assert same_fma_result(p.detach().cpu(), w)


@pytest.mark.parametrize("mil_norm,pos_", [("ln", 0), ("ln", 1), ("bn", 0), ("bn", 1)])
def test_dattention_mil_norm_variants(M, mil_norm, pos_):
    """The reference's legal non-default DAttention configurations (abmil.py:162-178, 207-223) load with strict=True under the
    reference's keys (the input LayerNorm is feature.0, the Linear feature.1) and match the oracle, eval forward and ('ln') backward."""
    N = 1333
    x = cases.make_bag(N, N, 1024)
    sd = cases.abmil_norm_state(N + pos_, mil_norm, pos_)
    m = M.DAttention(1024, 2, dropout=0.0, act="gelu", mil_norm=mil_norm, embed_norm_pos=pos_).cuda().eval()
    m.load_state_dict(cuda_sd(sd), strict=True)
    with torch.no_grad():
        got = m(x.cuda(), return_attn=True, return_act=True)
        ref = O.abmil_dattention(sd, x, "gelu", return_attn=True, return_act=True, mil_norm=mil_norm, embed_norm_pos=pos_)
    for a, b in zip(got, ref):
        assert cases.rel_err(a, b) < TOL
    if mil_norm == "ln":
        m.train()
        lg = m(x.cuda())
        sd_ref = {k: v.clone().requires_grad_() for k, v in sd.items()}
        rl = O.abmil_dattention(sd_ref, x, "gelu", mil_norm=mil_norm, embed_norm_pos=pos_)
        F.cross_entropy(lg, torch.tensor([1]).cuda()).backward()
        F.cross_entropy(rl, torch.tensor([1])).backward()
        for k, p in m.named_parameters():
            if k != "attention.2.bias":                                # mathematically zero
                assert cases.rel_err(p.grad, sd_ref[k].grad) < TOL, k


def test_dattention_sincos_and_amp(M):
    N = 777
    x, pos = cases.make_bag(N, N, 1024), cases.sincos_pos(5, N)
    sd = cases.abmil_state(N)
    m = M.DAttention(1024, 2, dropout=0.0, act="relu", pos="sincos").cuda().eval()
    m.load_state_dict(cuda_sd(sd), strict=True)
    with torch.no_grad():
        assert cases.rel_err(m(x.cuda(), pos=pos.cuda()), O.abmil_dattention(sd, x, "relu", pos=pos)) < TOL
    # --amp (engines/base_engine.py:78): autocast hands fp16 tensors to the kernels; they compute in fp32 instead of raising
    plain = M.DAttention(1024, 2, dropout=0.0, act="relu").cuda().eval()
    plain.load_state_dict(cuda_sd(sd), strict=True)
    with torch.no_grad(), torch.autocast("cuda", dtype=torch.float16):
        a = plain(x.cuda().half())
        b = plain(x.cuda())
    ref = O.abmil_dattention(sd, x, "relu")
    assert cases.rel_err(b, ref) < TOL and cases.rel_err(a, O.abmil_dattention(sd, x.half().float(), "relu")) < TOL


def test_mhim_configs_outside_the_fused_kernel_take_the_composed_path(M):
    """`--act none` (options.py:84), a free-form `--da_act` and n_classes > 4 are legal upstream; the fused kernel is instantiated for
    act in {relu, gelu}, att_act in {tanh, relu, gelu}, C <= 4 only.  Such models must fall back to the composed path, not raise."""
    n, d = 700, 1024
    x = cases.make_bag(5, n, d)
    for kw, tag in ((dict(act="none"), "act none"), (dict(n_classes=6), "6 classes"), (dict(da_act="sigmoid_like_unknown"), "unknown da_act")):
        full = dict(cases.MHIM_KW, baseline="attn", input_dim=d, dropout=0.0)
        full.update(kw)
        m = M.MHIM(**full).cuda().eval()
        sd = {k: v.detach().cpu() for k, v in m.state_dict().items()}
        cfg = O.MHIMConfig(**{k: v for k, v in full.items() if k in O.MHIMConfig.__dataclass_fields__})
        with torch.no_grad():
            got = m.forward_test(x.cuda())
            cls, score = m.forward_teacher(x.cuda())
            ref = O.mhim_forward_test(cfg, sd, x)
            rc, rs = O.mhim_forward_teacher(cfg, sd, x)
        assert cases.rel_err(got, ref) < TOL, tag
        assert cases.rel_err(cls, rc) < TOL and cases.rel_err(score, rs) < TOL, tag
