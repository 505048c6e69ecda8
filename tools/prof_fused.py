"""Timing probe used under gpurun / ncu: GEMM-store mode vs fused mode of the tcgen05 pipeline at the headline size."""
import os
import sys
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import cases  # noqa: E402
import mhimk  # noqa: E402

N = int(os.environ.get("PROF_N", 50000))
precs = os.environ.get("PROF_PREC", "bf16x3,fp16").split(",")
reps = int(os.environ.get("PROF_REPS", 5))
modes = os.environ.get("PROF_MODES", "store,fused").split(",")
sd = {k: v.cuda() for k, v in cases.abmil_state(1).items()}
xs = [torch.randn(N, 1024, device="cuda") for _ in range(3)]
act = os.environ.get("PROF_ACT", "relu")


def timeit(fn):
    fn(0)
    torch.cuda.synchronize()
    mhimk.ops.profile_fused(True)
    for i in range(reps):
        fn(i)
    torch.cuda.synchronize()
    n, tot = mhimk.ops.profile_collect()
    mhimk.ops.profile_fused(False)
    print(f"   kernel-only mean {tot / max(n, 1) * 1e3:8.1f} us over {n} launches", end="  |  ")
    ts = []
    for i in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn(i)
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return min(ts), sorted(ts)[len(ts) // 2]


for p in precs:
    if "store" in modes:
        t = timeit(lambda i: mhimk.ops.umma_selftest(xs[i % 3], sd["feature.0.weight"], p))
        print(f"store  {p:7s} N={N}: best {t[0]*1e3:8.1f} us  median {t[1]*1e3:8.1f} us  -> {N*4096/t[0]/1e6:7.1f} GB/s", flush=True)
    if "fused" in modes:
        t = timeit(lambda i: mhimk.ops.abmil_fused_forward(xs[i % 3], sd["feature.0.weight"], sd["feature.0.bias"], act, sd["attention.0.weight"],
                                                          sd["attention.0.bias"], sd["attention.2.weight"], sd["attention.2.bias"], "tanh", precision=p))
        print(f"fused  {p:7s} N={N}: best {t[0]*1e3:8.1f} us  median {t[1]*1e3:8.1f} us  -> {N*4096/t[0]/1e6:7.1f} GB/s", flush=True)
