"""The MHIM training step of BASELINE.json's configs (teacher forward -> top-k mask -> student forward -> losses -> backward), as the
reference's trainer runs it (dropout 0.25, teacher in train mode), eager vs CUDA-graph replay (mhimk.engines.GraphedStep).
    T_BASE=attn|dsmil|selfattn T_N=10000 T_D=1024 python tools/bench_train_step.py"""
import json, os, sys, time, types
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import cases, mhimk
from mhimk import modules as M
from mhimk.engines import CommonMIL, GraphedStep

dev = torch.device("cuda")
base, N, D = os.environ.get("T_BASE", "attn"), int(os.environ.get("T_N", 10000)), int(os.environ.get("T_D", 1024))
drop = float(os.environ.get("T_DROPOUT", 0.25))
kw = dict(cases.MHIM_KW, baseline=base, input_dim=D, dropout=drop)
stu, tea = M.MHIM(**kw).to(dev).train(), M.MHIM(**kw).to(dev).train()
args = types.SimpleNamespace(model="mhim", baseline=base, aux_alpha=0.5)
eng, ce, label = CommonMIL(args), torch.nn.CrossEntropyLoss(), torch.tensor([1], device=dev)
bag = cases.make_bag(3, N, D).to(dev)


def step(x):
    stu.zero_grad(set_to_none=True)
    logits, lab, aux, *_ = eng.forward_func(args, stu, tea, x, label, ce, 1, 0, 0, 0, None)
    loss = ce(logits.view(1, -1), lab) + 0.5 * aux
    loss.backward()
    return loss.detach()


def timed(fn, reps=30):
    for _ in range(5):
        fn(bag)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); e0.record()
    for _ in range(reps):
        fn(bag)
    t1 = time.perf_counter(); e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps, (t1 - t0) / reps * 1e3


if os.environ.get("T_NCU"):                       # under ncu: two eager steps (the launch list), nothing else
    for _ in range(2):
        step(bag)
    torch.cuda.synchronize()
    sys.exit(0)
eager_ms, eager_host = timed(step)
g = GraphedStep(step)
graph_ms, graph_host = timed(g)
out = {"config": f"MHIM({base}) training step, N={N} x D={D}, dropout={drop}, teacher in train mode", "eager_ms_per_step": eager_ms,
       "eager_host_issue_ms": eager_host, "graphed_ms_per_step": graph_ms, "graphed_host_issue_ms": graph_host,
       "instances_per_s_graphed": N / (graph_ms * 1e-3)}
try:
    from torch.profiler import profile, ProfilerActivity
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        for _ in range(3):
            step(bag)
        torch.cuda.synchronize()
    ka = prof.key_averages()
    out["kernel_launches_per_step"] = round(sum(k.count for k in ka if getattr(k, "self_device_time_total", 0) > 0 and k.device_type.name != "CPU") / 3)
    out["gpu_busy_ms_per_step"] = sum(getattr(k, "self_device_time_total", 0) for k in ka if k.device_type.name != "CPU") / 3 / 1e3
except Exception as e:
    out["profile_error"] = repr(e)
print(json.dumps(out))
