#!/bin/bash
# round 2, call F: graphed training step, faster skinny kernels
set -u
mkdir -p gpurun_out
( timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log )
grep -E "passed|failed|^FAILED|^E  " gpurun_out/pytest_gpu.log | cut -c1-600 | tail -20
for b in attn dsmil; do
  d=1024; [ $b = dsmil ] && d=1536
  T_BASE=$b T_D=$d timeout 300 python tools/bench_train_step.py > gpurun_out/train_step_$b.json 2> gpurun_out/train_step_$b.err; tail -1 gpurun_out/train_step_$b.json; tail -3 gpurun_out/train_step_$b.err
done
T_BASE=selfattn T_N=50000 timeout 300 python tools/bench_train_step.py > gpurun_out/train_step_selfattn.json 2> gpurun_out/train_step_selfattn.err; tail -1 gpurun_out/train_step_selfattn.json; tail -3 gpurun_out/train_step_selfattn.err
timeout 120 python tools/prof_train_step.py > gpurun_out/train_prof_attn.txt 2>&1; head -40 gpurun_out/train_prof_attn.txt | cut -c1-200
