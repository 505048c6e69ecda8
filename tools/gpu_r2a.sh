#!/bin/bash
# round 2, call A: full GPU suite (new dropout / BASELINE-size tests included), selfattn error diagnosis, bench sanity
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
( timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider --durations=15 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log )
tail -40 gpurun_out/pytest_gpu.log
timeout 300 python tools/diag_selfattn.py > gpurun_out/diag_selfattn.txt 2>&1; cat gpurun_out/diag_selfattn.txt
DIAG_N=50000 timeout 300 python tools/diag_selfattn.py > gpurun_out/diag_selfattn_50k.txt 2>&1; cat gpurun_out/diag_selfattn_50k.txt
timeout 600 python bench.py --steps 50 --warmup 5 > gpurun_out/bench_bf16x3.json 2> gpurun_out/bench_bf16x3.err; tail -1 gpurun_out/bench_bf16x3.json
timeout 60 python tools/prof_train_step.py > gpurun_out/train_prof_attn.txt 2>&1; head -3 gpurun_out/train_prof_attn.txt
