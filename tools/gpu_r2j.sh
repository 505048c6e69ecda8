#!/bin/bash
set -u
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_gpu_engines.py -m gpu -q -p no:cacheprovider > gpurun_out/pytest_engines.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_engines.log )
grep -E "passed|failed|^FAILED|^E  " gpurun_out/pytest_engines.log | cut -c1-300 | tail -8
timeout 300 python tools/prof_transmil.py > gpurun_out/prof_transmil.txt 2>&1; head -40 gpurun_out/prof_transmil.txt | cut -c1-70,150-215
