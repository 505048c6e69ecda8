"""Phase timeline of CTA 0 of the fused kernel (MHIMK_TRACE=1): clock64 stamps per tile, printed as deltas in cycles."""
import os
import sys
import torch
os.environ["MHIMK_TRACE"] = "1"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import cases  # noqa: E402
import mhimk  # noqa: E402
from mhimk import ops  # noqa: E402

N = int(os.environ.get("PROF_N", 50000))
sd = {k: v.cuda() for k, v in cases.abmil_state(1).items()}
x = torch.randn(N, 1024, device="cuda")
names = ["g1_start", "g1_issued", "g2_start", "g2_issued", "acc_full", "e1_done", "e2_done", "u_full", "e3_done", "e4_done"]
for prec in os.environ.get("PROF_PREC", "bf16x3,fp16").split(","):
    for rep in range(2):
        ops.abmil_fused_forward(x, sd["feature.0.weight"], sd["feature.0.bias"], os.environ.get("PROF_ACT", "relu"), sd["attention.0.weight"],
                                sd["attention.0.bias"], sd["attention.2.weight"], sd["attention.2.bias"], "tanh", precision=prec)
    torch.cuda.synchronize()
    ws = ops._fused_workspace(sd["feature.0.weight"], sd["attention.0.weight"], prec, ops._pipeline(None, prec))[0]
    L = mhimk._lib.lib()
    total = L.mil_fused_workspace_bytes(1024, 512, 128, 0)
    base = ws.data_ptr()
    a = (base + 255) & ~255
    err = a + 512 * 1024 * 4 + 128 * 512 * 4
    tr = ((err + 16) + 63) & ~63
    off = tr - base
    t = ws[off:off + 16 * 16 * 8].view(torch.int64).view(16, 16).cpu()
    g = ws[off + 16 * 16 * 8: off + 16 * 16 * 8 + 148 * 2 * 8].view(torch.int64).view(148, 2).cpu()
    t_first, t_last = int(g[:, 0].min()), int(g[:, 1].max())
    dur = (g[:, 1] - g[:, 0]).double() / 1e3
    print(f"== {prec}: kernel span {(t_last - t_first) / 1e3:.1f} us; CTA start spread {(int(g[:, 0].max()) - t_first) / 1e3:.1f} us; "
          f"per-CTA duration min {dur.min():.1f} / median {dur.median():.1f} / max {dur.max():.1f} us; "
          f"3-tile CTAs (0..94) median {dur[:95].median():.1f}, 2-tile CTAs median {dur[95:].median():.1f}")
    fl = t.view(-1)
    print(f"   CTA 0: {int(fl[239]) - int(fl[238])} cycles in {(int(g[0, 1]) - int(g[0, 0])) / 1e3:.1f} us -> {(int(fl[239]) - int(fl[238])) / max(int(g[0, 1]) - int(g[0, 0]), 1):.3f} GHz")
    print("   grid_finalize stamps (cycles since entry): " + " ".join(str(int(fl[224 + k]) - int(fl[224])) for k in range(5)))
    order = torch.argsort(g[:, 1], descending=True)[:6]
    print("   last CTAs to finish (block, start us, end us): " + ", ".join(f"({int(i)}, {(int(g[i, 0]) - t_first) / 1e3:.1f}, {(int(g[i, 1]) - t_first) / 1e3:.1f})" for i in order))
    ends = torch.sort((g[:, 1] - t_first).double() / 1e3).values
    print("   end-time percentiles us: " + ", ".join(f"p{q}={float(ends[int(q / 100 * (len(ends) - 1))]):.1f}" for q in (0, 10, 50, 90, 100)))
    print(f"== {prec}: per-tile phase stamps of CTA 0, cycles relative to g1_start of tile 0")
    t0 = int(t[0, 0])
    print(f"   CTA 0: kernel entry -> g1_start of tile 0: {t0 - int(fl[238])} cycles; e4_done of the last tile -> kernel exit: {int(fl[239]) - max(int(t[i, 9]) for i in range(16))} cycles")
    for tl in range(int(os.environ.get("PROF_TILES", 4))):
        if int(t[tl, 0]) == 0:
            break
        row = {n: int(t[tl, i]) - t0 for i, n in enumerate(names)}
        print(f" tile {tl}: " + "  ".join(f"{n}={v}" for n, v in row.items()))
        print(f"          GEMM1 window {row['acc_full'] - row['g1_start']}  | E1 {row['e1_done'] - row['acc_full']}  E2 {row['e2_done'] - row['e1_done']}"
              f"  wait-u {row['u_full'] - row['e2_done']}  E3 {row['e3_done'] - row['u_full']}  E4 {row['e4_done'] - row['e3_done']}"
              f"  | epilogue total {row['e4_done'] - row['acc_full']}")
