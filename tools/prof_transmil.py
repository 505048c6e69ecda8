"""Kernel table of one TransMIL eval forward at N = 50 000 (torch.profiler / CUPTI) + event timing."""
import os, sys
import torch
from torch.profiler import profile, ProfilerActivity
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import cases, mhimk
from mhimk import modules as M
N = int(os.environ.get("T_N", 50000))
t = M.TransMIL(1024, 2, dropout=0.0, act="relu").cuda().eval()
t.load_state_dict({k: v.cuda() for k, v in cases.transmil_state(81).items()}, strict=True)
x = cases.make_bag(1, N, 1024).cuda()
with torch.no_grad():
    for _ in range(3):
        t(x)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        t(x)
    e1.record(); torch.cuda.synchronize()
    print(f"TransMIL eval forward N={N}: {e0.elapsed_time(e1) / 5:.3f} ms")
    with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
        for _ in range(2):
            t(x)
        torch.cuda.synchronize()
ka = prof.key_averages()
busy = sum(getattr(k, "self_device_time_total", 0) for k in ka if k.device_type.name != "CPU")
print(f"GPU busy per forward {busy / 2 / 1e3:.3f} ms, launches {sum(k.count for k in ka if getattr(k, 'self_device_time_total', 0) > 0 and k.device_type.name != 'CPU') / 2:.0f}")
print(ka.table(sort_by="self_cuda_time_total", row_limit=28, max_name_column_width=60))
