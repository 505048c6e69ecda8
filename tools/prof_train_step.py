"""Where one MHIM training step (teacher + student forward + backward) spends its time: GPU-busy time per kernel
(torch.profiler / CUPTI), host issue time (wall clock without waiting for the GPU) and the event-timed step.
T_BASE=attn|dsmil|selfattn, T_N, T_D; T_MODE=step|teacher|test profiles the whole step / the teacher pass / `forward_test` alone."""
import os, sys, time
import torch, torch.nn.functional as F
from torch.profiler import profile, ProfilerActivity
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import cases, mhimk
from mhimk import modules as M

dev = torch.device("cuda"); LABEL = torch.tensor([1], device=dev)
base, N, D = os.environ.get("T_BASE", "attn"), int(os.environ.get("T_N", 10000)), int(os.environ.get("T_D", 1024))
kw = dict(cases.MHIM_KW, baseline=base, input_dim=D, dropout=0.0)
stu, tea = M.MHIM(**kw).to(dev).train(), M.MHIM(**kw).to(dev).train()
xb = cases.make_bag(3, N, D).to(dev)


def full():
    stu.zero_grad(set_to_none=True)
    ct, sc = tea.forward_teacher(xb)
    t_ = ct[0] if base == "dsmil" else ct
    lg, loss, _, _ = stu(xb, sc, t_, i=0)
    lt = 0.5 * lg[0].view(1, -1) + 0.5 * lg[1].view(1, -1) if base == "dsmil" else lg
    (F.cross_entropy(lt, LABEL) + 0.5 * loss).backward()


mode = os.environ.get("T_MODE", "step")
if mode != "step":
    step_fn = full
    stu.eval()

    def full():                                            # noqa: F811 - the profiled callable
        with torch.no_grad():
            return tea.forward_teacher(xb) if mode == "teacher" else stu.forward_test(xb)

for _ in range(5):
    full()
torch.cuda.synchronize()
R = 20
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0 = time.perf_counter(); e0.record()
for _ in range(R):
    full()
t1 = time.perf_counter(); e1.record(); torch.cuda.synchronize()
print(f"{base} N={N} D={D}: event-timed step {e0.elapsed_time(e1) / R:.3f} ms, host issue time {(t1 - t0) / R * 1e3:.3f} ms")
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    for _ in range(5):
        full()
    torch.cuda.synchronize()
ka = prof.key_averages()
gpu_us = sum(getattr(k, "self_device_time_total", 0) for k in ka)
print(f"GPU-busy time per step (sum of kernel durations): {gpu_us / 5 / 1e3:.3f} ms; launches per step: "
      f"{sum(k.count for k in ka if getattr(k, 'self_device_time_total', 0) > 0 and k.device_type.name != 'CPU') / 5:.0f}")
print(ka.table(sort_by="self_cuda_time_total", row_limit=40, max_name_column_width=70))
print(ka.table(sort_by="self_cpu_time_total", row_limit=25, max_name_column_width=70))
