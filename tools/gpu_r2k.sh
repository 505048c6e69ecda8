#!/bin/bash
set -u
mkdir -p gpurun_out
( timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log )
grep -E "passed|failed|^FAILED|^E  " gpurun_out/pytest_gpu.log | cut -c1-400 | tail -20
timeout 300 python tools/prof_transmil.py > gpurun_out/prof_transmil.txt 2>&1; head -24 gpurun_out/prof_transmil.txt | cut -c1-62,140-215
CFG_REPS=6 timeout 600 python tools/bench_configs.py > gpurun_out/configs.json 2> gpurun_out/configs.err; tail -c 3500 gpurun_out/configs.json; tail -3 gpurun_out/configs.err
