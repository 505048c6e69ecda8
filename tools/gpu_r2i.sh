#!/bin/bash
set -u
mkdir -p gpurun_out
( timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log )
grep -E "passed|failed|^FAILED|^E  " gpurun_out/pytest_gpu.log | cut -c1-500 | tail -40
CFG_REPS=5 timeout 300 python tools/bench_configs.py > gpurun_out/configs.json 2> gpurun_out/configs.err; tail -c 2500 gpurun_out/configs.json; tail -3 gpurun_out/configs.err
