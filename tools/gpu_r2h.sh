#!/bin/bash
set -u
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_engines.py -m gpu -q -p no:cacheprovider > gpurun_out/pytest_engines.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_engines.log )
grep -E "passed|failed|^FAILED|^E  " gpurun_out/pytest_engines.log | cut -c1-700 | tail -30
