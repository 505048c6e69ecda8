#!/bin/bash
set -u
mkdir -p gpurun_out
( timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log )
grep -E "passed|failed|^FAILED|^E  " gpurun_out/pytest_gpu.log | cut -c1-600 | tail -20
for b in attn dsmil; do
  d=1024; [ $b = dsmil ] && d=1536
  T_BASE=$b T_D=$d timeout 300 python tools/bench_train_step.py > gpurun_out/train_step_$b.json 2> gpurun_out/train_step_$b.err; tail -1 gpurun_out/train_step_$b.json; tail -2 gpurun_out/train_step_$b.err | cut -c1-300
done
timeout 120 python tools/prof_train_step.py > gpurun_out/train_prof_attn.txt 2>&1; head -36 gpurun_out/train_prof_attn.txt | cut -c1-75,150-215
