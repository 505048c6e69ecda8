for p in fp16 bf16x3; do
for d in 0 7 23 39 71 87 119; do echo "== prec $p dbg $d"; MHIMK_DEBUG=$d PROF_PREC=$p PROF_MODES=fused PROF_REPS=3 python tools/prof_fused.py; done
done
