import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import cases, mhimk
pipe = os.environ.get("PIPE", "single"); prec = os.environ.get("PREC", "bf16x3"); N = int(os.environ.get("N", 4099))
wh = os.environ.get("WANT_H", "1") == "1"
sd = {k: v.cuda() for k, v in cases.abmil_state(3).items()}
x = cases.make_bag(5, N, 1024)[0].cuda()
for r in range(int(os.environ.get("REPS", 6))):
    out = mhimk.ops.abmil_fused_forward(x, sd["feature.0.weight"], sd["feature.0.bias"], "relu", sd["attention.0.weight"], sd["attention.0.bias"],
                                        sd["attention.2.weight"], sd["attention.2.bias"], "tanh", want_scores=True, want_h=wh, precision=prec, pipeline=pipe)
    torch.cuda.synchronize()
    print("rep", r, "ok pooled[0]", float(out["pooled"][0]), flush=True)
