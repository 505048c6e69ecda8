"""Stress of the pair pipeline's tail split: random bag sizes (split and unsplit shapes interleaved on ONE workspace), every launch twice
(bit-identical) and against the single pipeline (different kernel, no split) within 2e-5."""
import os, sys, random
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import cases, mhimk
from mhimk import ops
sd = {k: v.cuda() for k, v in cases.abmil_state(3).items()}
rng = random.Random(int(os.environ.get("STRESS_SEED", 1)))
X = torch.randn(40000, 1024, device="cuda")
keep_all = (torch.rand(40000, device="cuda") > 0.3).to(torch.uint8)
worst = 0.0
n_iter = int(os.environ.get("STRESS_ITERS", 300))
for it in range(n_iter):
    N = rng.choice([rng.randint(1, 600), rng.randint(1, 9472), rng.randint(9473, 40000)])
    x = X[:N]
    keep = keep_all[:N] if it % 3 == 0 else None
    outs = []
    for pipe in ("pair", "pair", "single"):
        o = ops.abmil_fused_forward(x, sd["feature.0.weight"], sd["feature.0.bias"], "gelu", sd["attention.0.weight"], sd["attention.0.bias"],
                                    sd["attention.2.weight"], sd["attention.2.bias"], "tanh", keep=keep, want_scores=True, pipeline=pipe,
                                    Wcls=sd["classifier.weight"], bcls=sd["classifier.bias"])
        outs.append((o["pooled"].clone(), o["s"].clone(), o["logits"].clone()))
    torch.cuda.synchronize()
    assert all(torch.equal(a, b) for a, b in zip(outs[0], outs[1])), f"N={N}: repeated launch differs"
    fin = torch.isfinite(outs[2][1])
    e = max(float((outs[0][0] - outs[2][0]).abs().max() / outs[2][0].abs().max()),
            float((outs[0][1][fin] - outs[2][1][fin]).abs().max() / outs[2][1][fin].abs().max()) if bool(fin.any()) else 0.0)
    assert e < 2e-5 and torch.equal(torch.isfinite(outs[0][1]), fin), f"N={N}: pair vs single {e}"
    worst = max(worst, e)
print(f"{n_iter} random bag sizes: repeated launches bit-identical, pair (tail split) vs single pipeline worst rel diff {worst:.2e}")
