#!/bin/bash
# round 2, call B: full GPU suite (tensor-core weight gradient added), training-step profile
set -u
mkdir -p gpurun_out
( timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider -s > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log )
grep -E "full-tensor gradient errors|passed|failed|^FAILED|^E  " gpurun_out/pytest_gpu.log | cut -c1-1500 | tail -40
timeout 120 python tools/prof_train_step.py > gpurun_out/train_prof_attn.txt 2>&1; head -40 gpurun_out/train_prof_attn.txt
T_BASE=dsmil T_D=1536 timeout 120 python tools/prof_train_step.py > gpurun_out/train_prof_dsmil.txt 2>&1; head -3 gpurun_out/train_prof_dsmil.txt
