"""Where the HOST time of an eager MHIM training step goes (cProfile over 30 steps; the GPU is faster than the host issues work)."""
import cProfile, pstats, os, sys, io
import torch, torch.nn.functional as F
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import cases, mhimk
from mhimk import modules as M
dev = torch.device("cuda"); LABEL = torch.tensor([1], device=dev)
base, N, D = os.environ.get("T_BASE", "attn"), int(os.environ.get("T_N", 10000)), int(os.environ.get("T_D", 1024))
kw = dict(cases.MHIM_KW, baseline=base, input_dim=D, dropout=0.25)
stu, tea = M.MHIM(**kw).to(dev).train(), M.MHIM(**kw).to(dev).train()
xb = cases.make_bag(3, N, D).to(dev)
def full():
    stu.zero_grad(set_to_none=True)
    ct, sc = tea.forward_teacher(xb)
    t_ = ct[0] if base == "dsmil" else ct
    lg, loss, _, _ = stu(xb, sc, t_, i=0)
    lt = 0.5 * lg[0].view(1, -1) + 0.5 * lg[1].view(1, -1) if base == "dsmil" else lg
    (F.cross_entropy(lt, LABEL) + 0.5 * loss).backward()
for _ in range(5): full()
torch.cuda.synchronize()
pr = cProfile.Profile(); pr.enable()
for _ in range(30): full()
pr.disable(); torch.cuda.synchronize()
for key in ("tottime", "cumulative"):
    s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats(key).print_stats(22); print(s.getvalue()[:5200])
