"""Launch the fused forward REPS times on the same bag and report every launch that is not bit-identical to the last one
(pipeline-synchronisation regression probe; PIPE=single|pair PREC=bf16x3|fp16|bf16 N=... REPS=...)."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import cases, mhimk
if os.environ.get('MHIMK_OLD_LIB'):      # A/B against another build of the library
    mhimk._lib.LIB_PATH = os.environ['MHIMK_OLD_LIB']; mhimk.ops.PIPELINES = {'single': 0, 'pair': 0}
pipe = os.environ.get("PIPE", "single"); prec = os.environ.get("PREC", "bf16x3"); N = int(os.environ.get("N", 50000))
c = {k: v.cuda() for k, v in cases.abmil_state(5).items()}
x = cases.make_bag(9, N, 1024)[0].cuda()
outs = []
for r in range(int(os.environ.get("REPS", 12))):
    o = mhimk.ops.abmil_fused_forward(x, c["feature.0.weight"], c["feature.0.bias"], "relu", c["attention.0.weight"], c["attention.0.bias"],
                                      c["attention.2.weight"], c["attention.2.bias"], "tanh", want_scores=True, precision=prec, pipeline=pipe)
    outs.append((o["pooled"].clone(), o["s"].clone(), o["stats"].clone(), o["part"].clone()))
torch.cuda.synchronize()
ref = outs[-1]
for r, o in enumerate(outs):
    dp = (o[0] - ref[0]).abs().max().item(); ds = (o[1] - ref[1]).abs()
    bad = ds.nonzero().flatten()
    dpart = (o[3][:148] - ref[3][:148]).abs().amax(dim=1)
    print(f"rep {r}: pooled maxdiff {dp:.3e}  s ndiff {len(bad)} first {bad[:6].tolist()} maxdiff {ds.max().item():.3e}  stats {o[2].tolist()}  partial rows differing {dpart.nonzero().flatten()[:8].tolist()}")
