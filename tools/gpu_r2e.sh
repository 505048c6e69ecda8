#!/bin/bash
# round 2, call E: skinny kernels, oracle fix (merge.norm), DSMIL gradient diag, train-step profile
set -u
mkdir -p gpurun_out
( timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider -s > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log )
grep -E "full-tensor gradient errors|passed|failed|^FAILED|^E  " gpurun_out/pytest_gpu.log | cut -c1-900 | tail -30
timeout 300 python tools/diag_dsmil_grads.py > gpurun_out/diag_dsmil_grads.txt 2>&1; cat gpurun_out/diag_dsmil_grads.txt | cut -c1-1500
timeout 120 python tools/prof_train_step.py > gpurun_out/train_prof_attn.txt 2>&1; head -45 gpurun_out/train_prof_attn.txt | cut -c1-230
T_BASE=dsmil T_D=1536 timeout 120 python tools/prof_train_step.py > gpurun_out/train_prof_dsmil.txt 2>&1; head -3 gpurun_out/train_prof_dsmil.txt
