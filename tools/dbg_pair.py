"""Quick parity probe of the pair pipeline against an fp64 torch reference on the GPU (debugging aid)."""
import os
import sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import cases  # noqa: E402
import mhimk  # noqa: E402

pipe = os.environ.get("PIPE", "pair")
for N in [int(v) for v in os.environ.get("NS", "64,128,200,1000,20000").split(",")]:
    for prec in os.environ.get("PRECS", "fp16,bf16x3").split(","):
        sd = {k: v.cuda() for k, v in cases.abmil_state(3).items()}
        x = cases.make_bag(5, N, 1024)[0].cuda()
        out = mhimk.ops.abmil_fused_forward(x, sd["feature.0.weight"], sd["feature.0.bias"], "relu", sd["attention.0.weight"], sd["attention.0.bias"],
                                            sd["attention.2.weight"], sd["attention.2.bias"], "tanh", want_scores=True, want_h=N <= 20000,
                                            precision=prec, pipeline=pipe)
        torch.cuda.synchronize()
        xd = x.double()
        h = torch.relu(xd @ sd["feature.0.weight"].double().t() + sd["feature.0.bias"].double())
        u = torch.tanh(h @ sd["attention.0.weight"].double().t() + sd["attention.0.bias"].double())
        s = (u @ sd["attention.2.weight"].double().t() + sd["attention.2.bias"].double())[:, 0]
        pr = torch.softmax(s, 0) @ h
        e = lambda a, b: float((a.double() - b).abs().max() / b.abs().max())
        msg = f"N={N} {prec} {pipe}: pooled {e(out['pooled'], pr):.2e} s {e(out['s'], s):.2e}"
        if out["h"] is not None:
            msg += f" h {e(out['h'], h):.2e}"
            bad = ((out["h"].double() - h).abs() > 1e-2 * h.abs().max()).nonzero()
            if len(bad):
                rows = torch.unique(bad[:, 0]).tolist()
                msg += f" n_bad {len(bad)} bad rows {rows[:12]} (mod 128: {[r % 128 for r in rows[:12]]})"
        print(msg, flush=True)
