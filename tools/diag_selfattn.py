"""Where does the selfattn / TransMIL error (3e-4 .. 1e-3 in round 1) come from?  The fp32 CPU oracle is within 1e-7 of its own fp64
evaluation, so it is GPU-side arithmetic: candidates are cuDNN TF32 convolutions (res_conv, PPEG; torch.backends.cudnn.allow_tf32
defaults to True), the bf16x3 projections (5e-6 each) amplified by the pseudo-inverse, or cuBLAS.  Prints errors per switch."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import cases
from oracle import mil_oracle as O
import mhimk
from mhimk import modules as M, ops

n, d, seed = int(os.environ.get("DIAG_N", 600)), 1024, 71
cfg = O.MHIMConfig(**dict(cases.MHIM_KW, baseline="selfattn", input_dim=d))
sd, x = cases.mhim_state(seed, "selfattn", D=d), cases.make_bag(seed + 1000, n, d)
m = M.MHIM(**dict(cases.MHIM_KW, baseline="selfattn", input_dim=d, dropout=0.0)).cuda().eval()
m.load_state_dict({k: v.cuda() for k, v in sd.items()}, strict=True)
for mod in m.modules():
    if isinstance(mod, torch.nn.Dropout):
        mod.p = 0.0
with torch.no_grad():
    ref_t, (ref_c, ref_s) = O.mhim_forward_test(cfg, sd, x), O.mhim_forward_teacher(cfg, sd, x)


def run(tag):
    with torch.no_grad():
        got, (c, s) = m.forward_test(x.cuda()), m.forward_teacher(x.cuda())
    print(f"{tag:60s} forward_test {cases.rel_err(got, ref_t):.2e}  cls {cases.rel_err(c, ref_c):.2e}  score {cases.rel_err(s, ref_s):.2e}", flush=True)


print("N =", n)
run("default (cudnn.allow_tf32=%s, matmul.allow_tf32=%s)" % (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32))
torch.backends.cudnn.allow_tf32 = False
run("cudnn.allow_tf32=False")
torch.backends.cuda.matmul.allow_tf32 = False
torch.set_float32_matmul_precision("highest")
run("+ matmul highest")
orig = ops._tc_supported
ops._tc_supported = lambda *a, **k: False
run("+ every Linear through the exact-fp32 CUDA-core GEMM")
ops._tc_supported = orig
