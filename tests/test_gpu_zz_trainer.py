"""GPU replay of a few iterations of the reference's MHIM training loop against the trajectory recorded from the live reference
(tests/golden/golden_trainer_v1.pt, made by tests/golden/make_golden_trainer.py): teacher pass -> masked student pass -> CE + aux
loss -> backward -> SGD step -> EMA teacher update.  The EMA update alternates between the reference's literal loop (writes through
`.data`, engines/base_engine.py:166-167) and mhimk.engines.ema_update: both must leave the drop-in modules on the same trajectory."""
import json
import os

import pytest
import torch
import torch.nn.functional as F

import cases

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GT = torch.load(os.path.join(os.path.dirname(__file__), "golden", "golden_trainer_v1.pt"), weights_only=False)
TOL_FIRST, TOL_LATER = 1e-4, 1e-4        # the north-star gate on every iteration


def _avg(logits):
    """engines/common_mil.py:27-28, 66-67: dsmil returns [bag, instance] logits, the engine averages them."""
    return 0.5 * logits[0].view(1, -1) + 0.5 * logits[1].view(1, -1) if isinstance(logits, (list, tuple)) else logits


@pytest.mark.parametrize("name", ["attn", "dsmil", "selfattn"])
def test_training_trajectory_matches_reference(name):
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import mhimk
    from mhimk import modules as M
    from mhimk.engines import ema_update
    G = GT if name == "attn" else GT["more"][name]
    T = G["cfg"]
    # measured on B200 (profiles/round1e_trainer_trajectory_errs_*.json): attn <= 6e-6, dsmil <= 1e-5, selfattn <= 4.1e-5 (the iterative
    # pseudo-inverse amplifies rounding a few optimiser steps downstream): gate 1e-4 on the first iteration everywhere, 3e-4 later for Nystrom
    tol_first, tol_later = (TOL_FIRST, 3e-4) if name == "selfattn" else (TOL_FIRST, TOL_LATER)
    kw = dict(cases.MHIM_KW, baseline=T["base"], input_dim=T["D"], dropout=0.0)
    stu, tea = M.MHIM(**kw).cuda(), M.MHIM(**kw).cuda()
    stu.load_state_dict({k: v.cuda() for k, v in cases.mhim_state(T["seed"], T["base"], D=T["D"]).items()}, strict=True)
    tea.load_state_dict({k: v.cuda() for k, v in cases.mhim_state(T["seed"] + 1, T["base"], D=T["D"]).items()}, strict=True)
    for m in list(stu.modules()) + list(tea.modules()):
        if isinstance(m, torch.nn.Dropout):
            m.p = 0.0
    for p in tea.parameters():
        p.requires_grad = False
    stu.train(), tea.train()                                            # the reference keeps the teacher in train mode
    stu.merge._noise = lambda L, dev: torch.rand(L).to(dev)            # the CPU random stream the reference consumed
    bags = [cases.make_bag(T["seed"] + 1000 + j, T["N"], T["D"]).cuda() for j in range(2)]
    label = torch.tensor([1]).cuda()
    opt = torch.optim.SGD(stu.parameters(), lr=T["lr"])
    errs = {}
    for it, g in enumerate(G["steps"]):
        x = bags[it % 2]
        cls_tea, score = tea.forward_teacher(x)
        if T["base"] == "dsmil":
            cls_tea = cls_tea[0]                                         # engines/common_mil.py:26
        errs[f"{it}.cls_tea"] = cases.rel_err(cls_tea, g["cls_tea"])
        errs[f"{it}.score"] = cases.rel_err(score, g["score"])
        torch.manual_seed(T["seed"] + 7 + it)
        # index parity is defined on equal scores: the mask is taken from the reference's own fp32 scores of this iteration
        logits, aux, ps, keep = stu(x, g["score"].cuda(), cls_tea, i=it)
        logits = _avg(logits)
        assert (ps, keep) == (g["patch_num"], g["keep_num"])
        loss = F.cross_entropy(logits, label) + T["aux_alpha"] * aux
        errs[f"{it}.logits"] = cases.rel_err(logits, g["logits"])
        errs[f"{it}.aux_loss"] = cases.rel_err(aux, g["aux_loss"])
        errs[f"{it}.loss"] = cases.rel_err(loss, g["loss"])
        loss.backward()
        opt.step()
        opt.zero_grad()
        if it % 2 == 0:
            for param_q, param_k in zip(stu.parameters(), tea.parameters()):
                param_k.data.mul_(T["mm"]).add_(param_q.data, alpha=1. - T["mm"])
        else:
            ema_update(stu, tea, T["mm"])
    stu.eval(), tea.eval()
    ev_s, ev_t = stu.forward_test(bags[0]), tea.forward_test(bags[0])
    if T["base"] == "dsmil":                                             # forward_test -> ([bag, inst], B); validate_func takes [0]
        ev_s, ev_t = ev_s[0], ev_t[0]
    errs["stu_eval"] = cases.rel_err(_avg(ev_s), G["stu_eval"])
    errs["tea_eval"] = cases.rel_err(_avg(ev_t), G["tea_eval"])
    for who, model, norms in (("stu", stu, G["stu_norms"]), ("tea", tea, G["tea_norms"])):
        sd = model.state_dict()
        errs[f"{who}_norms"] = max(abs(sd[k].double().norm().item() - n) / max(n, 1e-12) for k, n in norms.items())
    if os.environ.get("MHIMK_DUMP_ERRS"):
        os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
        json.dump(errs, open(os.path.join(ROOT, "gpurun_out", f"trainer_errs_{name}.json"), "w"), indent=1)
    for k, e in errs.items():
        assert e < (tol_first if k.startswith("0.") else tol_later), (k, e, errs)
