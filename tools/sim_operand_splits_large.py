"""Round-2 re-examination of the operand arithmetic (Next #2d): the same CPU emulation as sim_operand_splits.py but at
BASELINE's bag sizes (N = 10 000 / 50 000) and gated exactly as north_star states it -- pooled vector and logits at 1e-4 (the
per-instance tensors h, s are reported, not gated: they only matter to calls whose per-instance outputs leave the kernel).
Two feature distributions: randn and relu(randn) (R50 features are non-negative -> W-rounding errors do not average out over N).
    python tools/sim_operand_splits_large.py [N ...]"""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import cases
from oracle import mil_oracle as O


def split(t, dt, parts):
    hi = t.to(dt).to(torch.float32)
    if parts == 1:
        return [hi]
    return [hi, (t - hi).to(dt).to(torch.float32)]


SCHEMES = {
    "bf16x3  Xh.Wh+Xl.Wh+Xh.Wl [shipped]": (torch.bfloat16, 2, 2, [(0, 0), (1, 0), (0, 1)]),
    "fp16x2a X.Wh+X.Wl   (X single)": (torch.float16, 1, 2, [(0, 0), (0, 1)]),
    "fp16x2b Xh.W+Xl.W   (W single)": (torch.float16, 2, 1, [(0, 0), (1, 0)]),
    "fp16x1": (torch.float16, 1, 1, [(0, 0)]),
}


def gemm(x, w, scheme, chunk=8192):
    """operands rounded to 16 bit, products and accumulation in float64 (upper bound on fp32 TMEM accumulation)"""
    dt, nx, nw, prods = scheme
    ws = [p.double() for p in split(w, dt, nw)]
    out = []
    for r0 in range(0, x.shape[0], chunk):
        xs = [p.double() for p in split(x[r0:r0 + chunk], dt, nx)]
        out.append(sum(xs[i] @ ws[j].t() for i, j in prods))
    return torch.cat(out)


def exact(x, w, chunk=8192):
    w = w.double()
    return torch.cat([x[r0:r0 + chunk].double() @ w.t() for r0 in range(0, x.shape[0], chunk)])


def forward(sd, x, act, scheme):
    pre = (gemm(x, sd["feature.0.weight"], scheme) if scheme else exact(x, sd["feature.0.weight"])) + sd["feature.0.bias"].double()
    h = O.apply_act(pre, act)
    u = (gemm(h.float(), sd["attention.0.weight"], scheme) if scheme else exact(h, sd["attention.0.weight"])) + sd["attention.0.bias"].double()
    s = torch.tanh(u) @ sd["attention.2.weight"].double()[0] + sd["attention.2.bias"].double()
    p = torch.softmax(s, 0) @ h
    return h, s, p, p @ sd["classifier.weight"].double().t() + sd["classifier.bias"].double()


if __name__ == "__main__":
    torch.set_num_threads(os.cpu_count())
    sizes = [int(a) for a in sys.argv[1:]] or [1024, 10000, 50000]
    print(f"{'scheme':38s} {'N':>6s} {'act':>5s} {'X':>6s}  rel_err(h)  rel_err(s)  rel_err(pooled)  rel_err(logits)   north_star gate: pooled, logits <= 1e-4")
    for name, scheme in SCHEMES.items():
        for N in sizes:
            for act, kind, seed in [("relu", "randn", 0), ("relu", "relu", 1), ("gelu", "randn", 2), ("gelu", "relu", 3)]:
                sd, x = cases.abmil_state(100 + N + seed), cases.make_bag(200 + N + seed, N, 1024, kind)[0]
                ref, got = forward(sd, x, act, None), forward(sd, x, act, scheme)
                e = [cases.rel_err(g, r) for g, r in zip(got, ref)]
                print(f"{name:38s} {N:6d} {act:>5s} {kind:>6s}  {e[0]:10.2e}  {e[1]:10.2e}  {e[2]:15.2e}  {e[3]:15.2e}   "
                      f"{'ok' if e[2] < 1e-4 and e[3] < 1e-4 else 'FAILS'}", flush=True)
