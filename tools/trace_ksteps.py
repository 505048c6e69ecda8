"""Per-k-step stamps of CTA 0 of the pair pipeline (MHIMK_TRACE=1): who waits for whom in the operand ring.
Needs a library built with the stamps compiled in:  KSTAMP=1 bash mhim-mil_b200/csrc/build.sh  (touch mil_fused2_sm100.cu first)."""
import os
import sys
import torch
os.environ["MHIMK_TRACE"] = "1"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import cases  # noqa: E402
import mhimk  # noqa: E402
from mhimk import ops  # noqa: E402

N = int(os.environ.get("PROF_N", 50000))
prec = os.environ.get("PROF_PREC", "fp16")
sd = {k: v.cuda() for k, v in cases.abmil_state(1).items()}
x = torch.randn(N, 1024, device="cuda")
for rep in range(2):
    ops.abmil_fused_forward(x, sd["feature.0.weight"], sd["feature.0.bias"], "relu", sd["attention.0.weight"], sd["attention.0.bias"],
                            sd["attention.2.weight"], sd["attention.2.bias"], "tanh", precision=prec)
torch.cuda.synchronize()
ws, _ = ops._fused_workspace(sd["feature.0.weight"], sd["attention.0.weight"], prec, ops._pipeline(None, prec))
base = ws.data_ptr()
a = (base + 255) & ~255
err = a + 512 * 1024 * 4 + 128 * 512 * 4
tr = ((err + 16) + 63) & ~63
off = tr - base
t = ws[off + 1024 * 8: off + (1024 + 7 * 64) * 8].view(torch.int64).view(7, 64).cpu()
names = ["W:empty/st", "I:full/st", "I:top/st", "C:empty", "C:arrived", "C:xfull", "X:xempty"]
t0 = int(t[2, 0])
print("k-step  " + "  ".join(f"{n:>10s}" for n in names))
for it in range(64):
    print(f"{it:6d}  " + "  ".join(f"{int(t[s, it]) - t0:10d}" for s in range(7)))
