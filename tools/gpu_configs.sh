#!/bin/bash
# One short gpurun call: GPU parity suite, then step times of the BASELINE.json configs (tools/bench_configs.py).
set -u
mkdir -p gpurun_out
( timeout 120 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_e.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu_e.log )
tail -3 gpurun_out/pytest_gpu_e.log
CFG_REPS=5 timeout 200 python tools/bench_configs.py > gpurun_out/configs.json 2> gpurun_out/configs.err
echo "configs exit $?"; tail -c 3000 gpurun_out/configs.json; tail -5 gpurun_out/configs.err
