"""Device time of mil_topk_f32 alone over (N, k) -- which phase (passes over N vs the sort of the k winners) carries the cost."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import mhimk
K = mhimk.ops
for N in (1000, 10000, 50000):
    s = torch.rand(N, device="cuda")
    for k in (8, 300, 3000):
        if k > N:
            continue
        for largest in (True,):
            for _ in range(3): K.topk(s, k, largest)
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()                      # 20 calls in one graph: device time without the host's issue time
            with torch.cuda.graph(g):
                for _ in range(20): K.topk(s, k, largest)
            g.replay(); torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(5): g.replay()
            e1.record(); torch.cuda.synchronize()
            print(f"N={N:6d} k={k:5d}: {e0.elapsed_time(e1) / 100 * 1e3:7.1f} us per call (graph replay of 20 calls)")


# phase timeline of one launch (clock64 stamps the kernel leaves in the workspace's spare bytes): cache | select | compact | 4 sort passes | out
import ctypes
L = mhimk._lib.lib()
for N, k in ((1000, 8), (10000, 300), (50000, 3000), (200000, 6000)):
    s = torch.rand(N, device="cuda")
    ws = torch.empty(L.mil_topk_workspace_bytes(N), dtype=torch.uint8, device="cuda")
    idx = torch.empty(k, dtype=torch.int64, device="cuda")
    for _ in range(3):
        rc = L.mil_topk_f32(ctypes.c_void_p(s.data_ptr()), N, k, 1, ctypes.c_void_p(idx.data_ptr()), ctypes.c_void_p(ws.data_ptr()), ws.numel(),
                            ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
        assert rc == 0
    torch.cuda.synchronize()
    base = ws.data_ptr()
    off = ((((base + 15) & ~15) + 24 * N + 7) & ~7) - base
    st = ws[off:off + 72].view(torch.int64).cpu().tolist()
    d = [st[i + 1] - st[i] for i in range(8)]
    print(f"N={N} k={k}: cycles  cache {d[0]} | select {d[1]} | compact {d[2]} | sort passes {d[3]} {d[4]} {d[5]} {d[6]} | write-out {d[7]} | total {st[8] - st[0]}")
