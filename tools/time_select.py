"""Device time of the masked hard-instance selection (top-k + mask_ids) vs the reference's torch.topk + python-set path."""
import os, sys, time, math
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import mhimk
from oracle import mil_oracle as O
K = mhimk.ops
for N, ratio in [(10000, 0.03), (50000, 0.03), (200000, 0.03)]:
    k = int(math.ceil(N * ratio))
    s = (0.5 + torch.randint(0, 900, (N,)).float() * 5.96e-8).cuda()
    def ours():
        idx = K.topk(s, k, True)
        return K.mask_from_indices(idx, N)
    for _ in range(3): ours()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): ours()
    e1.record(); torch.cuda.synchronize()
    sc = s.cpu()[None]
    t0 = time.perf_counter(); O.select_mask(N, sc, True, ratio, len_keep_other=N, random_ratio=1.0); t1 = time.perf_counter()
    print(f"N={N:7d} k={k:5d}: device top-k + mask_ids {e0.elapsed_time(e1) / 20 * 1e3:8.1f} us   | oracle (torch.topk + complement) on CPU {1e3 * (t1 - t0):7.2f} ms")
