"""Stage-by-stage gradient comparison of the DSMIL bag head (baseline.py:131-152) at N ~ 7 770: our CUDA primitives (exact-fp32 and
tensor-core variants) against torch CPU fp64 on the same inputs -- where does the q.0 / v.1 gradient discrepancy enter?"""
import math, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import cases
from oracle import mil_oracle as O
import mhimk
from mhimk import ops

d, seed, n = 1536, 161, int(os.environ.get("DIAG_N", 7770))
sd = cases.mhim_state(seed, "dsmil", D=d)
x = cases.make_bag(seed + 1000, n, d)
with torch.no_grad():
    h0 = O.mhim_embed(sd, x, "gelu")
p = "online_encoder.b_classifier."
gB = torch.randn(2, 512, generator=torch.Generator().manual_seed(1)) * 1e-2


def run(dt, dev, lin, pool):
    W = {k: v.to(dt).to(dev).clone().requires_grad_(True) for k, v in sd.items() if k.startswith(p) or k.startswith("online_encoder.i_")}
    h = h0.to(dt).to(dev).clone().requires_grad_(True)
    st = {}

    def keep(name, t):
        t.retain_grad(); st[name] = t; return t
    classes = lin(h, W["online_encoder.i_classifier.0.weight"], W["online_encoder.i_classifier.0.bias"], "none")
    V = keep("V", lin(h, W[p + "v.1.weight"], W[p + "v.1.bias"], "relu"))
    r = keep("r0", lin(h, W[p + "q.0.weight"], W[p + "q.0.bias"], "relu"))
    Q = keep("Q", lin(r, W[p + "q.2.weight"], W[p + "q.2.bias"], "tanh"))
    crit = torch.sort(classes.detach(), 0, descending=True).indices[0]
    rm = keep("r0_crit", lin(h.index_select(0, crit), W[p + "q.0.weight"], W[p + "q.0.bias"], "relu"))
    qm = keep("q_max", lin(rm, W[p + "q.2.weight"], W[p + "q.2.bias"], "tanh"))
    logit = keep("logit", lin(Q, qm, None, "none") / math.sqrt(128))
    B = torch.stack([pool(logit[:, j], V) for j in range(2)])
    (B * gB.to(dt).to(dev)).sum().backward()
    out = {k: v.grad.detach().double().cpu() for k, v in st.items()}
    out.update({k.replace(p, ""): v.grad.detach().double().cpu() for k, v in W.items() if v.grad is not None})
    out["h"] = h.grad.detach().double().cpu()
    out["crit"] = crit.cpu()
    return out


ref = run(torch.float64, "cpu", lambda x_, W_, b_, a: O.apply_act(O.affine(x_, W_, b_), a), lambda s, V: torch.softmax(s, 0) @ V)
cpu32 = run(torch.float32, "cpu", lambda x_, W_, b_, a: O.apply_act(O.affine(x_, W_, b_), a), lambda s, V: torch.softmax(s, 0) @ V)
gpu_lin = lambda x_, W_, b_, a: ops.linear_act(x_, W_, b_, a)
gpu_pool = lambda s, V: ops.softmax_pool(s, V)[0]


def report(tag, got):
    assert torch.equal(got["crit"], ref["crit"])
    print(tag + ": " + ", ".join(f"{k} {cases.rel_err(got[k], ref[k]):.1e}" for k in ref if k != "crit"), flush=True)


report("cpu fp32 (torch)          ", cpu32)
report("gpu tensor cores (default)", run(torch.float32, "cuda", gpu_lin, gpu_pool))
orig, ops.WGRAD_TC = ops._tc_supported, False
ops._tc_supported = lambda *a, **k: False
report("gpu exact-fp32 CUDA cores ", run(torch.float32, "cuda", gpu_lin, gpu_pool))
report("gpu exact GEMMs, torch pool", run(torch.float32, "cuda", gpu_lin, lambda s, V: torch.softmax(s, 0) @ V))
ops._tc_supported, ops.WGRAD_TC = orig, True
