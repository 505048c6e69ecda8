"""Repeat the fused forward many times and count runs whose h / pooled deviate (race hunting aid)."""
import os
import sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import cases  # noqa: E402
import mhimk  # noqa: E402
if os.environ.get("MHIMK_OLD_LIB"):      # A/B against another build of the library (path to its libmhimk.so; no pipeline bits)
    mhimk._lib.LIB_PATH = os.environ["MHIMK_OLD_LIB"]
    mhimk.ops.PIPELINES = {"single": 0, "pair": 0}

pipe = os.environ.get("PIPE", "single")
reps = int(os.environ.get("REPS", 30))
for N in [int(v) for v in os.environ.get("NS", "4099,10000").split(",")]:
    for prec in os.environ.get("PRECS", "bf16x3,fp16").split(","):
        sd = {k: v.cuda() for k, v in cases.abmil_state(3).items()}
        x = cases.make_bag(5, N, 1024)[0].cuda()
        ref = None
        nbad, rows_seen = 0, set()
        for r in range(reps):
            out = mhimk.ops.abmil_fused_forward(x, sd["feature.0.weight"], sd["feature.0.bias"], "relu", sd["attention.0.weight"], sd["attention.0.bias"],
                                                sd["attention.2.weight"], sd["attention.2.bias"], "tanh", want_scores=True, want_h=True, precision=prec, pipeline=pipe)
            torch.cuda.synchronize()
            ws, _ = mhimk.ops._fused_workspace(sd["feature.0.weight"], sd["attention.0.weight"], prec, pipe)
            a = (ws.data_ptr() + 255) & ~255
            off = a + 512 * 1024 * 4 + 128 * 512 * 4 - ws.data_ptr()
            code = int(ws[off:off + 4].view(torch.int32)[0])
            if code:
                print(f"   rep {r}: wait timed out: code {code & 255} block {(code >> 8) & 4095} warp {code >> 20}", flush=True)
                ws[off:off + 16].zero_()
            if ref is None:
                xd = x.double()
                ref = torch.relu(xd @ sd["feature.0.weight"].double().t() + sd["feature.0.bias"].double())
            bad = ((out["h"].double() - ref).abs() > 1e-2 * ref.abs().max()).nonzero()
            if len(bad):
                nbad += 1
                rows_seen |= set((torch.unique(bad[:, 0]) % 128).tolist())
        print(f"N={N} {prec} {pipe} dbg={os.environ.get('MHIMK_DEBUG', '0')}: {nbad}/{reps} bad runs; bad rows mod 128: {sorted(rows_seen)[:16]}", flush=True)
