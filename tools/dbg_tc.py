import sys, torch
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import mhimk, cases
from oracle import mil_oracle as O
K = mhimk.ops
g = torch.Generator().manual_seed(4)
M, N, Kd = 1500, 512, 1024
x, W, b = torch.randn(M, Kd, generator=g), torch.randn(N, Kd, generator=g) * 0.03, torch.randn(N, generator=g) * 0.1
go = torch.randn(M, N, generator=g)
for act in ("gelu", "relu"):
    Wr, br = W.double().requires_grad_(True), b.double().requires_grad_(True)
    pre_ref = x.double() @ Wr.t() + br
    (O.apply_act(pre_ref, act) * go.double()).sum().backward()
    Wd, bd = W.cuda().requires_grad_(True), b.cuda().requires_grad_(True)
    y = K.linear_act(x.cuda(), Wd, bd, act)
    print(act, "fwd err", cases.rel_err(y, O.apply_act(pre_ref, act)))
    (y * go.cuda()).sum().backward()
    e = (Wd.grad.cpu().double() - Wr.grad).abs()
    print(act, "gW err", cases.rel_err(Wd.grad, Wr.grad), "gb err", cases.rel_err(bd.grad, br.grad), "worst at", divmod(int(e.argmax()), Kd), "rows with err>1e-3:", int((e.max(dim=1).values > 1e-3 * Wr.grad.abs().max()).sum()), "cols:", int((e.max(dim=0).values > 1e-3 * Wr.grad.abs().max()).sum()))
    # compare with exact SIMT forward then backward
    Ws = W.cuda().requires_grad_(True)
    pre = torch.empty(M, N, device="cuda")
    ys = K.sgemm(x.cuda(), Kd, 1, Ws, Kd, 1, M, N, Kd, bias=b.cuda(), act=act, pre_out=pre)
    print(act, "simt fwd err", cases.rel_err(ys, O.apply_act(pre_ref, act)), "pre", cases.rel_err(pre, pre_ref))
