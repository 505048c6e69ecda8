for p in fp16 bf16x3; do
for d in 0 1 2 4 5 7 15; do echo "== prec $p dbg $d"; MHIMK_DEBUG=$d PROF_PREC=$p PROF_MODES=fused PROF_REPS=3 python tools/prof_fused.py; done
echo "== prec $p grid 74"; MHIMK_GRID=74 PROF_PREC=$p PROF_MODES=fused PROF_REPS=3 python tools/prof_fused.py
done
