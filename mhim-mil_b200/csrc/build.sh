#!/bin/bash
# Builds libmhimk.so in-tree for sm_100a (no GPU needed: nvcc cross-compiles).
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
ROOT="$(cd "$HERE/../.." && pwd)"
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
OUT="$HERE/../libmhimk.so"
FLAGS=(-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -I"$ROOT/include" -I"$HERE")
mkdir -p "$HERE/build"
pids=()
SRCS="mil_api mil_simt mil_topk mil_fused_sm100 mil_fused2_sm100 mil_wgrad_sm100 mil_skinny mil_rows mil_nystrom"
for f in $SRCS; do
  if [ ! -f "$HERE/build/$f.o" ] || [ "$HERE/$f.cu" -nt "$HERE/build/$f.o" ] || [ "$HERE/mil_common.cuh" -nt "$HERE/build/$f.o" ] || [ "$HERE/mil_umma.cuh" -nt "$HERE/build/$f.o" ] || [ "$ROOT/include/mhimk.h" -nt "$HERE/build/$f.o" ]; then
    "$NVCC" "${FLAGS[@]}" ${PTXAS_V:+-Xptxas -v} ${KSTAMP:+-DMIL_KSTAMP} ${PROBE:+-DMIL_PROBE} -c "$HERE/$f.cu" -o "$HERE/build/$f.o" &
    pids+=($!)
  fi
done
for p in "${pids[@]:-}"; do [ -n "$p" ] && wait "$p"; done
OBJS=()
for f in $SRCS; do OBJS+=("$HERE/build/$f.o"); done
"$NVCC" -shared -o "$OUT" "${OBJS[@]}" -lcudart
echo "built $OUT"
