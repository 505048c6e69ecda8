"""Generate tests/golden/golden_trainer_v1.pt: a few iterations of the reference's MHIM training loop on the LIVE reference (CPU).

Run in the build container only:  python tests/golden/make_golden_trainer.py
What is replayed (unmodified reference classes, dropout neutralised, seeded weights of tests/cases.py):
  engines/common_mil.py:14-48   CommonMIL.forward_func  (teacher pass -> student pass -> tuple)
  engines/base_engine.py:97-151 train_loss = CE(logits, label) + aux_alpha * aux_loss; backward; optimizer.step(); zero_grad()
  engines/base_engine.py:155-167 the EMA teacher update written through `.data`
with plain SGD (the trajectory stays well-conditioned; Adam divides by the tiny second moments of near-zero gradients).
Stored per iteration: the teacher's scores (the GPU test feeds them to its student: index parity is defined on equal scores),
cls_tea, logits, aux loss, total loss, keep_num; after the loop: eval logits of student and teacher and a few weight norms.
"""
import os
import sys
from types import SimpleNamespace

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

import cases  # noqa: E402
from _refload import load_reference, zero_dropout  # noqa: E402

TRAINER = dict(base="attn", N=2000, D=1024, seed=81, iters=4, lr=0.02, mm=0.99, aux_alpha=0.5)
# further trajectories (stored under "more"; replayed by the CPU oracle test): the other two MHIM baselines
MORE = {"dsmil": dict(base="dsmil", N=600, D=1536, seed=83, iters=3, lr=0.02, mm=0.99, aux_alpha=0.5),
        "selfattn": dict(base="selfattn", N=300, D=1024, seed=85, iters=3, lr=0.02, mm=0.99, aux_alpha=0.5)}


def run(R, T):
    kw = dict(cases.MHIM_KW, baseline=T["base"], input_dim=T["D"], dropout=0.0)
    stu, tea = zero_dropout(R.mhim.MHIM(**kw)), zero_dropout(R.mhim.MHIM(**kw))
    stu.load_state_dict(cases.mhim_state(T["seed"], T["base"], D=T["D"]), strict=True)
    tea.load_state_dict(cases.mhim_state(T["seed"] + 1, T["base"], D=T["D"]), strict=True)
    for p in tea.parameters():
        p.requires_grad = False
    stu.train(), tea.train()                                            # base_engine.py:36-38
    bags = [cases.make_bag(T["seed"] + 1000 + j, T["N"], T["D"]) for j in range(2)]
    label = torch.tensor([1])
    args = SimpleNamespace(model="mhim", baseline=T["base"], aux_alpha=T["aux_alpha"])
    engine = R.common_mil.CommonMIL(None)
    crit = torch.nn.CrossEntropyLoss()
    opt = torch.optim.SGD(stu.parameters(), lr=T["lr"])
    steps = []
    for it in range(T["iters"]):
        x = bags[it % 2]
        with torch.no_grad():
            cls_tea, score = tea.forward_teacher(x)                     # what forward_func computes first (deterministic)
        torch.manual_seed(T["seed"] + 7 + it)                            # the Merge keep order draws from the CPU generator
        logits, _, aux_loss, patch_num, keep_num, _, _ = engine.forward_func(args, stu, tea, x, label, crit, 1, it, 0, it, None)
        loss = crit(logits, label) + T["aux_alpha"] * aux_loss
        loss.backward()
        opt.step()
        opt.zero_grad()
        for param_q, param_k in zip(stu.parameters(), tea.parameters()):  # base_engine.py:166-167
            param_k.data.mul_(T["mm"]).add_(param_q.data, alpha=1. - T["mm"])
        steps.append({"score": score.detach().clone(), "cls_tea": (cls_tea[0] if T["base"] == "dsmil" else cls_tea).detach().clone(),
                      "logits": logits.detach().clone(),
                      "aux_loss": aux_loss.detach().clone(), "loss": loss.detach().clone(), "patch_num": patch_num, "keep_num": keep_num})
    stu.eval(), tea.eval()
    val = R.common_mil.CommonMIL(None).validate_func                    # engines/common_mil.py:56-69 (dsmil: 0.5 bag + 0.5 instance logits)
    with torch.no_grad():
        out = {"cfg": T, "steps": steps, "stu_eval": val(args, stu, bags[0], label, None, 1, 0, None)[0].clone(),
               "tea_eval": val(args, tea, bags[0], label, None, 1, 0, None)[0].clone(),
               "stu_norms": {k: v.double().norm().item() for k, v in stu.state_dict().items()},
               "tea_norms": {k: v.double().norm().item() for k, v in tea.state_dict().items()},
               "fp": cases.fingerprint(cases.mhim_state(T["seed"], T["base"], D=T["D"]), bags[0])}
    print(T["base"], [float(s["loss"]) for s in steps])
    return out


def main():
    R = load_reference()
    out = run(R, TRAINER)
    out["more"] = {name: run(R, T) for name, T in MORE.items()}
    path = os.path.join(HERE, "golden_trainer_v1.pt")
    torch.save(out, path)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
