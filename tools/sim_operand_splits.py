"""CPU emulation of the operand arithmetic of the fused pass: which split of the fp32 operands into 16-bit tensor-core operands meets
the 1e-4 parity gate?  (No GPU: operands are rounded with torch, products/accumulation in float64 = an upper bound on what fp32
TMEM accumulation achieves.)  Schemes: products listed as (X part, W part).
    python tools/sim_operand_splits.py"""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import cases
from oracle import mil_oracle as O


def split(t, dt, parts):
    hi = t.to(dt).to(torch.float32)
    if parts == 1:
        return [hi.double()]
    lo = (t - hi).to(dt).to(torch.float32)
    return [hi.double(), lo.double()]


SCHEMES = {
    "bf16x3  (Xh.Wh + Xl.Wh + Xh.Wl)  [shipped]": (torch.bfloat16, 2, 2, [(0, 0), (1, 0), (0, 1)]),
    "fp16x3  (Xh.Wh + Xl.Wh + Xh.Wl)": (torch.float16, 2, 2, [(0, 0), (1, 0), (0, 1)]),
    "fp16x2a (X.Wh + X.Wl)  X single": (torch.float16, 1, 2, [(0, 0), (0, 1)]),
    "fp16x2b (Xh.W + Xl.W)  W single": (torch.float16, 2, 1, [(0, 0), (1, 0)]),
    "bf16x2  (Xh.Wh + Xl.Wh) W single": (torch.bfloat16, 2, 1, [(0, 0), (1, 0)]),
    "fp16x1": (torch.float16, 1, 1, [(0, 0)]),
}


def gemm(x, w, scheme):
    dt, nx, nw, prods = scheme
    xs, ws = split(x, dt, nx), split(w, dt, nw)
    return sum(xs[i] @ ws[j].t() for i, j in prods)


def forward(sd, x, act, scheme):
    pre = (gemm(x, sd["feature.0.weight"], scheme) if scheme else x.double() @ sd["feature.0.weight"].double().t()) + sd["feature.0.bias"].double()
    h = O.apply_act(pre, act)
    h32 = h.float()                                                       # the epilogue hands fp32 h to GEMM2's operand split
    u = (gemm(h32, sd["attention.0.weight"], scheme) if scheme else h @ sd["attention.0.weight"].double().t()) + sd["attention.0.bias"].double()
    s = torch.tanh(u) @ sd["attention.2.weight"].double()[0] + sd["attention.2.bias"].double()
    a = torch.softmax(s, 0)
    p = a @ h
    return h, s, p, p @ sd["classifier.weight"].double().t() + sd["classifier.bias"].double()


print(f"{'scheme':46s} {'N':>6s} {'act':>5s}  rel_err(h)  rel_err(s)  rel_err(pooled)  rel_err(logits)   (gate 1e-4; s, t: 3e-4)")
for name, scheme in SCHEMES.items():
    for N, act, kind in [(1, "relu", "randn"), (129, "relu", "relu"), (1024, "gelu", "randn"), (4099, "relu", "randn")]:
        sd, x = cases.abmil_state(100 + N), cases.make_bag(200 + N, N, 1024, kind)[0]
        ref, got = forward(sd, x, act, None), forward(sd, x, act, scheme)
        e = [cases.rel_err(g, r) for g, r in zip(got, ref)]
        print(f"{name:46s} {N:6d} {act:>5s}  {e[0]:10.2e}  {e[1]:10.2e}  {e[2]:15.2e}  {e[3]:15.2e}   {'ok' if e[0] < 1e-4 and e[2] < 1e-4 and e[3] < 1e-4 and e[1] < 3e-4 else 'FAILS'}")
