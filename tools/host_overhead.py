"""CPU cost of one public-API step (DAttention eval forward), measured without waiting for the GPU."""
import os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import cases, mhimk
from mhimk.modules import DAttention
m = DAttention(1024, 2, dropout=0.0, act="relu").cuda().eval()
x_small = torch.randn(1, 256, 1024, device="cuda")      # tiny bag: GPU time ~ 0, so wall time = host time
with torch.no_grad():
    for _ in range(20): m(x_small)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(500): m(x_small)
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    print(f"host time per step (module call): {(t1 - t0) / 500 * 1e6:.1f} us")
    f0, a0, a2 = m.feature[0], m.attention[0], m.attention[2]
    t0 = time.perf_counter()
    for _ in range(500):
        mhimk.ops.abmil_fused_forward(x_small[0], f0.weight, f0.bias, "relu", a0.weight, a0.bias, a2.weight, a2.bias, "tanh")
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    print(f"host time per fused op call:     {(t1 - t0) / 500 * 1e6:.1f} us")
    p = torch.randn(1, 512, device="cuda")
    t0 = time.perf_counter()
    for _ in range(500):
        mhimk.ops.linear_act(p, m.classifier.weight, m.classifier.bias, "none")
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    print(f"host time per classifier GEMM:   {(t1 - t0) / 500 * 1e6:.1f} us")
    import cProfile, pstats
    pr = cProfile.Profile(); pr.enable()
    for _ in range(300): m(x_small)
    pr.disable(); torch.cuda.synchronize()
    pstats.Stats(pr).sort_stats("cumulative").print_stats(18)
