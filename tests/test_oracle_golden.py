"""Oracle (oracle/mil_oracle.py) against the committed golden vectors made from the live reference
(tests/golden/make_golden.py).  CPU only; runs anywhere, incl. the GPU box that has no /root/reference."""
import os

import pytest
import torch
import torch.nn.functional as F

import cases
from oracle import mil_oracle as O

G = torch.load(os.path.join(os.path.dirname(__file__), "golden", "golden_v1.pt"), weights_only=False)
LABEL = torch.tensor([1])
TOL = 2e-6   # fp32 CPU vs fp32 CPU; GEMM blocking may differ between the two formulations


def leaf(sd):
    return {k: v.clone().requires_grad_(v.is_floating_point()) for k, v in sd.items()}


def check_grads(sd, golden, tol=2e-5):
    for k, g in golden.items():
        got = sd[k].grad
        assert got is not None, k
        if g["norm"] < 1e-6:      # e.g. attention.2.bias: d/dbc = sum_n g_s = 0 exactly; only rounding noise
            assert got.double().norm().item() < 1e-5, k
            continue
        assert abs(got.double().norm().item() - g["norm"]) <= tol * max(g["norm"], 1e-12), k
        assert cases.rel_err(got.flatten()[:8], g["head"]) <= 50 * tol or g["head"].abs().max() < 1e-9, k


@pytest.mark.parametrize("name", list(G["abmil"]))
def test_abmil(name):
    act, n, kind = name.split("_")
    i = list(G["abmil"]).index(name)
    sd = leaf(cases.abmil_state(11 + i))
    x = cases.make_bag(11 + i + 1000, int(n), 1024, kind)
    g = G["abmil"][name]
    assert cases.fingerprint(sd, x) == pytest.approx(g["fp"], rel=1e-12), "seeded RNG stream differs from golden"
    (out, pooled), attn, actv = O.abmil_dattention(sd, x, act, return_attn=True, return_act=True, return_img_feat=True)
    assert cases.rel_err(out, g["logits"]) <= TOL
    assert cases.rel_err(pooled, g["pooled"]) <= TOL
    assert cases.rel_err(attn[0, :16], g["attn_head"]) <= TOL
    F.cross_entropy(out, LABEL).backward()
    check_grads(sd, g["grads"])


@pytest.mark.parametrize("name", list(G["gated"]))
def test_gated(name):
    act, n, kind = name.split("_")
    i = list(G["gated"]).index(name)
    sd = leaf(cases.gated_state(31 + i))
    x = cases.make_bag(31 + i + 1000, int(n), 1024, kind)
    g = G["gated"][name]
    assert cases.fingerprint(sd, x) == pytest.approx(g["fp"], rel=1e-12)
    out = O.abmil_gated(sd, x, act)
    assert cases.rel_err(out, g["logits"]) <= TOL
    F.cross_entropy(out, LABEL).backward()
    check_grads(sd, g["grads"])


MHIM_CASES = {"attn_2000": ("attn", 2000, 1024, 51), "attn_33": ("attn", 33, 1024, 52),
              "dsmil_1000": ("dsmil", 1000, 1536, 61), "selfattn_600": ("selfattn", 600, 1024, 71)}


@pytest.mark.parametrize("name", list(MHIM_CASES))
def test_mhim(name):
    base, n, d, seed = MHIM_CASES[name]
    g = G["mhim"][name]
    cfg = O.MHIMConfig(**dict(cases.MHIM_KW, baseline=base, input_dim=d))
    sd_s, sd_t = leaf(cases.mhim_state(seed, base, D=d)), cases.mhim_state(seed + 1, base, D=d)
    x = cases.make_bag(seed + 1000, n, d)
    assert cases.fingerprint(sd_s, x) + cases.fingerprint(sd_t, x) == pytest.approx(g["fp"], rel=1e-12)
    tol = 5e-6 if base == "selfattn" else TOL
    with torch.no_grad():
        cls_tea, score = O.mhim_forward_teacher(cfg, sd_t, x)
    assert cases.rel_err(cls_tea, g["cls_tea"]) <= tol
    assert cases.rel_err(score, g["score"]) <= tol
    score = g["score"]        # from here on use the reference's own fp32 scores: index parity is defined on equal inputs
    tcf = cls_tea[0] if base == "dsmil" else cls_tea
    torch.manual_seed(seed + 7)
    logits, loss, ps, len_keep, new_q, ids = O.mhim_forward(cfg, sd_s, x, score, tcf, i=0, training=True)
    assert (ps, len_keep) == (g["ps"], g["len_keep"])
    assert cases.tensor_digest(ids) == g["mask_ids_digest"]              # bit-exact mask indices
    if base == "dsmil":
        for a, b in zip(logits, g["logits"]):
            assert cases.rel_err(a, b) <= tol
        lt = 0.5 * logits[0].view(1, -1) + 0.5 * logits[1].view(1, -1)
    else:
        assert cases.rel_err(logits, g["logits"]) <= tol
        lt = logits
    assert cases.rel_err(loss, g["loss"]) <= tol
    assert cases.rel_err(new_q[0, :, :8], g["new_global_q_head"]) <= tol
    (F.cross_entropy(lt, LABEL) + 0.5 * loss).backward()
    check_grads(sd_s, g["grads"], tol=2e-4 if base == "selfattn" else 1e-4)
    with torch.no_grad():
        ft = O.mhim_forward_test(cfg, sd_s, x)
        pu = O.mhim_pure(cfg, sd_s, x)
    if base == "dsmil":
        for a, b in zip(ft[0], g["forward_test"]):
            assert cases.rel_err(a, b) <= tol
        for a, b in zip(pu, g["pure_eval"]):
            assert cases.rel_err(a, b) <= tol
    else:
        assert cases.rel_err(ft, g["forward_test"]) <= tol
        assert cases.rel_err(pu, g["pure_eval"]) <= tol


@pytest.mark.parametrize("i", range(len(G["select_cases"])))
def test_select_mask(i):
    ps, r, hr, lg, h = G["select_cases"][i]
    g = G["select"][f"{ps}_{r}_{hr}_{int(lg)}_{h}"]
    gen = torch.Generator().manual_seed(90 + i)
    attn = torch.rand(1, h, ps, generator=gen) if h else torch.rand(1, ps, generator=gen)
    torch.manual_seed(90 + i + 3)
    lk, ids = O.select_mask(ps, attn, lg, r, len_keep_other=ps, random_ratio=hr)
    assert lk == g["len_keep"]
    assert g["kept_sorted"]                       # the reference's python-set complement was ascending here
    assert cases.tensor_digest(ids) == g["ids_digest"]


def test_transmil_milnet():
    g = G["transmil"]["700"]
    sd, x = cases.transmil_state(81), cases.make_bag(1081, 700, 1024)
    assert cases.fingerprint(sd, x) == pytest.approx(g["fp"], rel=1e-12)
    logits, attn, v = O.transmil_forward(sd, x, "relu", return_attn=True, return_act=True)
    assert cases.rel_err(logits, g["logits"]) <= 5e-6
    assert cases.rel_err(attn[0][0, :, :8], g["attn0_head"]) <= 5e-6
    assert cases.rel_err(attn[1][0, :, :8], g["attn1_head"]) <= 5e-6
    assert cases.rel_err(v[0, :, :2, :4], g["v_head"]) <= 5e-6
    g = G["milnet"]["500"]
    sd, x = cases.milnet_state(85), cases.make_bag(1085, 500, 1536)
    pred, classes, _, _ = O.milnet_forward(sd, x, "relu")
    assert cases.rel_err(pred, g["pred"]) <= TOL and cases.rel_err(classes, g["classes"]) <= TOL


def test_analytic_backward_matches_autograd():
    """the streaming-backward formulas the CUDA kernel implements == autograd (fp64)."""
    torch.manual_seed(3)
    N, D, Hh, Da = 97, 64, 48, 16
    for gated in (False, True):
        for act in ("relu", "gelu"):
            x = torch.randn(N, D, dtype=torch.float64)
            W1 = torch.randn(Hh, D, dtype=torch.float64, requires_grad=True)
            b1 = torch.randn(Hh, dtype=torch.float64, requires_grad=True)
            Wa = torch.randn(Da, Hh, dtype=torch.float64, requires_grad=True) * 0.2
            Wa.retain_grad()
            Wb = (torch.randn(Da, Hh, dtype=torch.float64) * 0.2).requires_grad_(True) if gated else None
            wc = torch.randn(Da, dtype=torch.float64, requires_grad=True)
            g_p = torch.randn(Hh, dtype=torch.float64)
            h = O.apply_act(O.affine(x, W1, b1), act)
            gate = torch.tanh(O.affine(h, Wa)) * (torch.sigmoid(O.affine(h, Wb)) if gated else 1.0)
            p, _ = O.softmax_pool(gate @ wc, h)
            (p @ g_p).backward()
            an = O.abmil_backward_analytic(x, W1.detach(), b1.detach(), Wa.detach(), None, wc.detach(), act, g_p,
                                           Wb.detach() if gated else None)
            assert cases.rel_err(an["W1"], W1.grad) < 1e-10
            assert cases.rel_err(an["b1"], b1.grad) < 1e-10
            assert cases.rel_err(an["Wa"], Wa.grad) < 1e-10
            assert cases.rel_err(an["wc"], wc.grad) < 1e-10
            if gated:
                assert cases.rel_err(an["Wb"], Wb.grad) < 1e-10


def test_shard_merge_is_associative():
    """merging per-shard (m, l, P) equals the unsharded softmax pool, for ragged splits."""
    torch.manual_seed(4)
    s, h = torch.randn(1000, dtype=torch.float64) * 3, torch.randn(1000, 32, dtype=torch.float64)
    p_ref, _ = O.softmax_pool(s, h)
    for cuts in ([1000], [500, 500], [1, 999], [250, 250, 250, 250], [3, 17, 480, 100, 100, 100, 100, 100]):
        parts, o = [], 0
        for c in cuts:
            parts.append(O.pool_partial(s[o:o + c], h[o:o + c]))
            o += c
        _, _, p = O.merge_partials(*zip(*parts))
        assert cases.rel_err(p, p_ref) < 1e-12


@pytest.mark.parametrize("mm", [0.9999, 0.99, 0.5, 0.0, 1.0])
def test_ema_update_matches_the_literal_reference_loop(mm):
    """Pins O.ema_update bit for bit to the reference's own statement (engines/base_engine.py:166-167) executed here on CPU
    tensors: same zip order, same two roundings per element."""
    sd_q, sd_k = cases.mhim_state(1, "attn"), cases.mhim_state(2, "attn")
    q = [torch.nn.Parameter(v.clone()) for v in sd_q.values()]
    k = [torch.nn.Parameter(v.clone()) for v in sd_k.values()]
    got = O.ema_update(q, k, mm)
    for param_q, param_k in zip(q, k):                                  # the reference's loop, verbatim in effect
        param_k.data.mul_(mm).add_(param_q.data, alpha=1. - mm)
    for a, b in zip(got, k):
        assert torch.equal(a, b.data)
    with pytest.raises(AssertionError):
        O.ema_update(q, k, 1.5)


GT = torch.load(os.path.join(os.path.dirname(__file__), "golden", "golden_trainer_v1.pt"), weights_only=False)


def _avg(logits):
    """engines/common_mil.py:27-28, 66-67: dsmil returns [bag, instance] logits, the engine averages them."""
    return 0.5 * logits[0].view(1, -1) + 0.5 * logits[1].view(1, -1) if isinstance(logits, (list, tuple)) else logits


@pytest.mark.parametrize("name", ["attn", "dsmil", "selfattn"])
def test_training_trajectory_matches_reference(name):
    """A few iterations of the reference's own training loop (CommonMIL.forward_func -> CE + aux_alpha * aux -> SGD step -> EMA
    teacher update through `.data`), recorded from the live reference by tests/golden/make_golden_trainer.py, replayed on the
    oracle: per-iteration teacher scores / cls_tea / logits / losses / keep_num and the final eval logits and weight norms."""
    g_all = GT if name == "attn" else GT["more"][name]
    T = g_all["cfg"]
    base = T["base"]
    cfg = O.MHIMConfig(**dict(cases.MHIM_KW, baseline=base, input_dim=T["D"]))
    sd_s, sd_t = leaf(cases.mhim_state(T["seed"], base, D=T["D"])), cases.mhim_state(T["seed"] + 1, base, D=T["D"])
    bags = [cases.make_bag(T["seed"] + 1000 + j, T["N"], T["D"]) for j in range(2)]
    assert cases.fingerprint(cases.mhim_state(T["seed"], base, D=T["D"]), bags[0]) == pytest.approx(g_all["fp"], rel=1e-12)
    tol = 1e-4 if base == "selfattn" else 2e-5   # fp32 vs fp32, a few optimiser steps apart (Nystrom: iterative pinv)
    for it, g in enumerate(g_all["steps"]):
        x = bags[it % 2]
        with torch.no_grad():
            cls_tea, score = O.mhim_forward_teacher(cfg, sd_t, x)
        tcf = cls_tea[0] if base == "dsmil" else cls_tea
        assert cases.rel_err(tcf, g["cls_tea"]) <= tol and cases.rel_err(score, g["score"]) <= tol
        torch.manual_seed(T["seed"] + 7 + it)
        logits, aux, ps, keep, new_q, _ = O.mhim_forward(cfg, sd_s, x, g["score"], tcf, i=it, training=True)
        logits = _avg(logits)
        assert (ps, keep) == (g["patch_num"], g["keep_num"])
        loss = F.cross_entropy(logits, LABEL) + T["aux_alpha"] * aux
        assert cases.rel_err(logits, g["logits"]) <= tol and cases.rel_err(aux, g["aux_loss"]) <= tol and cases.rel_err(loss, g["loss"]) <= tol
        loss.backward()
        with torch.no_grad():
            sd_s["merge.global_q_mm"].copy_(new_q)                       # merge.py:127-129 (train-mode side effect; no gradient)
            sd_s["merge.global_q"].copy_(new_q)                          # the same Parameter under its second name
            for k, p in sd_s.items():                                    # SGD step + zero_grad
                if p.grad is not None and not k.startswith("merge.global_q"):
                    p -= T["lr"] * p.grad
                p.grad = None
            keys = list(sd_t)
            new_t = O.ema_update([sd_s[k] for k in keys], [sd_t[k] for k in keys], T["mm"])      # base_engine.py:166-167
            sd_t = dict(zip(keys, new_t))
    with torch.no_grad():
        ev_s, ev_t = O.mhim_forward_test(cfg, sd_s, bags[0]), O.mhim_forward_test(cfg, sd_t, bags[0])
        if base == "dsmil":                                              # forward_test -> ([bag, inst], B); validate_func takes [0]
            ev_s, ev_t = ev_s[0], ev_t[0]
        assert cases.rel_err(_avg(ev_s), g_all["stu_eval"]) <= tol
        assert cases.rel_err(_avg(ev_t), g_all["tea_eval"]) <= tol
    for k, n in g_all["stu_norms"].items():
        assert abs(sd_s[k].double().norm().item() - n) <= tol * max(n, 1e-12), k
    for k, n in g_all["tea_norms"].items():
        assert abs(sd_t[k].double().norm().item() - n) <= tol * max(n, 1e-12), k
