"""Evidence for the ReLU-gate explanation of the N = 10 000 gradient exceptions (tests/test_gpu_baseline_sizes.py): the same
teacher -> student -> backward pass with (a) the default tensor-core arithmetic and (b) every contraction on the exact-fp32
CUDA-core GEMM, full-tensor gradient errors against the CPU oracle."""
import os, sys
import torch, torch.nn.functional as F
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import cases
from oracle import mil_oracle as O
import mhimk
from mhimk import modules as M, ops
from test_gpu_baseline_sizes import build, tie_free, full_grad_errors


def run(base, n, d, seed, tag):
    cfg = O.MHIMConfig(**dict(cases.MHIM_KW, baseline=base, input_dim=d))
    (stu, sd_s), (tea, sd_t) = build(M, base, d, seed), build(M, base, d, seed + 1)
    stu.train(), tea.train()
    x = cases.make_bag(seed + 1000, n, d)
    with torch.no_grad():
        rc, rs = O.mhim_forward_teacher(cfg, sd_t, x)
    rs = tie_free(rs)
    tcf = rc[0] if base == "dsmil" else rc
    torch.manual_seed(seed + 7)
    stu.merge._noise = lambda L, dev: torch.rand(L).to(dev)
    logits, loss, _, _ = stu(x.cuda(), rs.cuda(), tcf.cuda(), i=0)
    sd_ref = {k: v.clone().requires_grad_(v.is_floating_point()) for k, v in sd_s.items()}
    torch.manual_seed(seed + 7)
    olg, oloss, *_ = O.mhim_forward(cfg, sd_ref, x, rs, tcf, i=0, training=True)
    lt, olt = (0.5 * logits[0].view(1, -1) + 0.5 * logits[1].view(1, -1), 0.5 * olg[0].view(1, -1) + 0.5 * olg[1].view(1, -1)) if base == "dsmil" else (logits, olg)
    (F.cross_entropy(lt, torch.tensor([1]).cuda()) + 0.5 * loss).backward()
    (F.cross_entropy(olt, torch.tensor([1])) + 0.5 * oloss).backward()
    errs = full_grad_errors(stu, sd_ref)
    worst = sorted(errs.items(), key=lambda kv: -kv[1])[:5]
    print(f"{base} N={n} {tag}: " + ", ".join(f"{k.split('online_encoder.')[-1]} {v:.1e}" for k, v in worst), flush=True)


for base, n, d, seed in (("attn", 10000, 1024, 151), ("dsmil", 10000, 1536, 161)):
    run(base, n, d, seed, "tensor cores (default)        ")
    orig, ops.WGRAD_TC = ops._tc_supported, False
    ops._tc_supported = lambda *a, **k: False
    run(base, n, d, seed, "exact-fp32 CUDA-core GEMM only")
    ops._tc_supported, ops.WGRAD_TC = orig, True
