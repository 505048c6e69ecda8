"""EMA teacher update: the reference's per-parameter loop (engines/base_engine.py:166-167) vs mhimk.engines.ema_update (one launch)."""
import os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import cases, mhimk
from mhimk import modules as M
from mhimk.engines import ema_update

for base in ("attn", "selfattn"):
    kw = dict(cases.MHIM_KW, baseline=base, input_dim=1024)
    stu, tea = M.MHIM(**kw).cuda(), M.MHIM(**kw).cuda()
    n_par, n_el = len(list(tea.parameters())), sum(p.numel() for p in tea.parameters())

    def loop(mm=0.9999):
        for param_q, param_k in zip(stu.parameters(), tea.parameters()):
            param_k.data.mul_(mm).add_(param_q.data, alpha=1. - mm)

    def timeit(fn, reps=200):
        for _ in range(10):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter(); e0.record()
        for _ in range(reps):
            fn()
        t1 = time.perf_counter(); e1.record(); torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps * 1e3, (t1 - t0) / reps * 1e6

    a, b = timeit(loop), timeit(lambda: ema_update(stu, tea, 0.9999))
    print(f"MHIM({base}): {n_par} parameters, {n_el} elements | reference loop {a[0]:.1f} us/step (host {a[1]:.1f} us, {2 * n_par} launches) | "
          f"ema_update {b[0]:.1f} us/step (host {b[1]:.1f} us, 1 launch)")
