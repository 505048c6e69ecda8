"""Generate tests/golden/golden_v1.pt from the LIVE reference (/root/reference, CPU).

Run in the build container only:  python tests/golden/make_golden.py
The reference classes are instantiated unmodified, loaded (strict=True) with the seeded weights of
tests/cases.py, every nn.Dropout p is set to 0 on the instance, and their outputs on seeded bags are
stored as small tensors / digests.  The bags and weights themselves are NOT stored: tests rebuild
them from the seeds and check `fingerprint` first.
"""
import os
import sys

import torch
import torch.nn.functional as F

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

import cases  # noqa: E402
from _refload import load_reference, zero_dropout  # noqa: E402

R = load_reference()
LABEL = torch.tensor([1])


def grads_summary(model):
    out = {}
    for k, p in model.named_parameters():
        if p.grad is not None:
            out[k] = {"norm": p.grad.double().norm().item(), "head": p.grad.flatten()[:8].clone()}
    return out


def run_abmil(act, N, kind, seed):
    sd = cases.abmil_state(seed)
    x = cases.make_bag(seed + 1000, N, 1024, kind)
    m = zero_dropout(R.abmil.DAttention(1024, 2, dropout=0.0, act=act))
    m.load_state_dict(sd, strict=True)
    m.train()
    logits, attn, actv = m(x.clone(), return_attn=True, return_act=True)
    pooled = m(x.clone(), return_img_feat=True)[1]
    F.cross_entropy(logits, LABEL).backward()
    return {"fp": cases.fingerprint(sd, x), "logits": logits.detach(), "pooled": pooled.detach(),
            "attn_head": attn[0, :16].detach(), "attn_digest": cases.tensor_digest(attn),
            "grads": grads_summary(m)}


def run_gated(act, N, kind, seed):
    sd = cases.gated_state(seed)
    x = cases.make_bag(seed + 1000, N, 1024, kind)
    m = zero_dropout(R.abmil.AttentionGated(1024, 2, act=act, dropout=0.0))
    m.load_state_dict(sd, strict=True)
    m.train()
    logits = m(x.clone())
    F.cross_entropy(logits, LABEL).backward()
    return {"fp": cases.fingerprint(sd, x), "logits": logits.detach(), "grads": grads_summary(m)}


def run_mhim(baseline, N, D, seed):
    kw = dict(cases.MHIM_KW, baseline=baseline, input_dim=D, dropout=0.0)
    sd_s = cases.mhim_state(seed, baseline, D=D)
    sd_t = cases.mhim_state(seed + 1, baseline, D=D)
    x = cases.make_bag(seed + 1000, N, D)
    stu = zero_dropout(R.mhim.MHIM(**kw))
    tea = zero_dropout(R.mhim.MHIM(**kw))
    stu.load_state_dict(sd_s, strict=True)
    tea.load_state_dict(sd_t, strict=True)
    stu.train()
    tea.train()
    cls_tea, score = tea.forward_teacher(x)
    tcf = cls_tea[0] if baseline == "dsmil" else cls_tea
    torch.manual_seed(seed + 7)
    logits, loss, ps, len_keep = stu(x, score, tcf, i=0)
    # the mask the student used (deterministic for hr=1: no randperm is consumed)
    lk, ids = stu.get_mask(N, 0, score)
    logit_t = 0.5 * logits[0].view(1, -1) + 0.5 * logits[1].view(1, -1) if baseline == "dsmil" else logits
    (F.cross_entropy(logit_t, LABEL) + 0.5 * loss).backward()
    out = {"fp": cases.fingerprint(sd_s, x) + cases.fingerprint(sd_t, x),
           "cls_tea": cls_tea.detach(), "score": score.detach(),
           "score_distinct": int(score.unique().numel()),
           "logits": [l.detach() for l in logits] if baseline == "dsmil" else logits.detach(),
           "loss": loss.detach(), "ps": ps, "len_keep": len_keep, "mask_len_keep": lk,
           "mask_ids_digest": cases.tensor_digest(ids), "mask_tail": ids[0, lk:].clone(),
           "new_global_q_head": stu.merge.global_q_mm.data[0, :, :8].clone(),
           "grads": grads_summary(stu)}
    stu.eval()
    ft = stu.forward_test(x)
    pu = stu.pure(x)
    # dsmil: forward_test returns ([bag, inst], B); pure (eval) returns [bag, inst]
    out["forward_test"] = [t.detach() for t in ft[0]] if baseline == "dsmil" else ft.detach()
    out["pure_eval"] = [t.detach() for t in pu] if baseline == "dsmil" else pu.detach()
    return out


def run_select(ps, ratio, hr, largest, seed, heads=0):
    g = torch.Generator().manual_seed(seed)
    if heads:
        attn = torch.rand(1, heads, ps, generator=g)
    else:
        attn = torch.rand(1, ps, generator=g)
    torch.manual_seed(seed + 3)
    lk, ids = R.masking.select_mask_fn(ps, attn, largest, ratio, len_keep_other=ps, random_ratio=hr)
    return {"len_keep": lk, "ids_digest": cases.tensor_digest(ids), "ids_tail": ids[0, lk:][:64].clone(),
            "kept_sorted": bool((ids[0, :lk][1:] > ids[0, :lk][:-1]).all()) if lk > 1 else True}


def run_transmil(N, seed):
    sd = cases.transmil_state(seed)
    x = cases.make_bag(seed + 1000, N, 1024)
    m = zero_dropout(R.transmil.TransMIL(1024, 2, dropout=0.0, act="relu")).eval()
    m.load_state_dict(sd, strict=True)
    with torch.no_grad():
        logits, attn, v = m(x, return_attn=True, return_act=True)
    return {"fp": cases.fingerprint(sd, x), "logits": logits, "attn0_head": attn[0][0, :, :8].clone(),
            "attn1_head": attn[1][0, :, :8].clone(), "v_head": v[0, :, :2, :4].clone()}


def run_milnet(N, seed):
    sd = cases.milnet_state(seed)
    x = cases.make_bag(seed + 1000, N, 1536)
    m = R.dsmil.MILNet(2, 0.0, "relu", input_dim=1536).eval()
    m.load_state_dict(sd, strict=True)
    with torch.no_grad():
        pred, classes = m(x)
    return {"fp": cases.fingerprint(sd, x), "pred": pred, "classes": classes}


def main():
    torch.set_num_threads(8)
    G = {"torch": torch.__version__, "reference_commit": "9d0c91abeb1b39f91c161b47cf14bc44f5a89da3"}
    G["abmil"] = {f"{a}_{n}_{k}": run_abmil(a, n, k, 11 + i) for i, (a, n, k) in
                  enumerate([("relu", 1024, "randn"), ("gelu", 333, "randn"), ("relu", 257, "relu"), ("gelu", 2, "randn")])}
    G["gated"] = {f"{a}_{n}_{k}": run_gated(a, n, k, 31 + i) for i, (a, n, k) in
                  enumerate([("relu", 1024, "randn"), ("gelu", 257, "relu")])}
    G["mhim"] = {"attn_2000": run_mhim("attn", 2000, 1024, 51), "attn_33": run_mhim("attn", 33, 1024, 52),
                 "dsmil_1000": run_mhim("dsmil", 1000, 1536, 61), "selfattn_600": run_mhim("selfattn", 600, 1024, 71)}
    sel = [(1000, 0.03, 1.0, True, 0), (1000, 0.03, 0.5, True, 0), (4099, 0.05, 1.0, True, 0), (255, 0.01, 1.0, True, 0),
           (1000, 0.5, 0.2, True, 0), (600, 0.1, 1.0, False, 0), (600, 0.03, 1.0, True, 8), (5, 0.03, 1.0, True, 0)]
    G["select"] = {f"{ps}_{r}_{hr}_{int(lg)}_{h}": run_select(ps, r, hr, lg, 90 + i, h) for i, (ps, r, hr, lg, h) in enumerate(sel)}
    G["select_cases"] = sel
    G["transmil"] = {"700": run_transmil(700, 81)}
    G["milnet"] = {"500": run_milnet(500, 85)}
    def compact(o):      # views would drag their whole storage into the file
        if isinstance(o, torch.Tensor):
            return o.detach().clone().contiguous()
        if isinstance(o, dict):
            return {k: compact(v) for k, v in o.items()}
        if isinstance(o, (list, tuple)):
            return type(o)(compact(v) for v in o)
        return o
    torch.save(compact(G), os.path.join(HERE, "golden_v1.pt"))
    print("wrote", os.path.join(HERE, "golden_v1.pt"), os.path.getsize(os.path.join(HERE, "golden_v1.pt")), "bytes")


if __name__ == "__main__":
    main()
