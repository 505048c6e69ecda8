# What sets the ~1000-cycle operand-ring stage of the pair kernel (768 GEMM1 MMA cycles)?  Needs a probe build (touch csrc/mil_fused2_sm100.cu; PROBE=1 csrc/build.sh):
# the no-conversion and prefetch-distance switches are compiled out of the shipped kernel.  Run on a B200; output = profiles/round2_ring_attribution.txt.
# Part 1: fewer CTAs (MHIMK_GRID) -> same GEMM1 window  => not chip-wide L2 / HBM contention.
for g in 148 74 38 2; do
  n=$(( g / 2 * 4 * 128 ))
  echo "GRID=$g N=$n"
  MHIMK_GRID=$g PROF_N=$n PROF_PREC=bf16x3 python tools/trace_fused.py 2>&1 | grep -E "GEMM1 window" | cut -c1-140
done
# Part 2: switch the ring's servers off one by one (MHIMK_DEBUG: 1 no W1 TMA, 2 no bag TMA, 4 no GEMM1 MMAs, 8 no conversion work; hand-shakes stay).
for d in 0 1 2 8 10 11 4 5 6 7 12 15; do
  echo "MHIMK_DEBUG=$d"
  MHIMK_DEBUG=$d PROF_N=37888 PROF_PREC=bf16x3 python tools/trace_fused.py 2>&1 | grep -E "GEMM1 window" | head -3 | cut -c1-60
done
