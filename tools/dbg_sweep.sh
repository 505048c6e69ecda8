#!/bin/bash
# timing attribution of the fused kernel: MHIMK_DEBUG bits skip parts of the pipeline (results are then wrong on purpose)
for d in ${DBGS:-0 1 2 4 8 128 3 7 15 32}; do
  echo "== MHIMK_DEBUG=$d"
  MHIMK_DEBUG=$d PROF_MODES=fused PROF_REPS=5 PROF_PREC=${PROF_PREC:-fp16} timeout 120 python tools/prof_fused.py 2>&1 | tail -2
done
