#!/bin/bash
# round 2, call C: fp16x3 arithmetic + shard record + everything so far
set -u
mkdir -p gpurun_out
( timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider -s > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log )
grep -E "full-tensor gradient errors|umma fp16x3|passed|failed|^FAILED|^E  " gpurun_out/pytest_gpu.log | cut -c1-1200 | tail -50
for p in fp16x3 bf16x3; do
  timeout 600 python bench.py --steps 50 --warmup 5 --precision $p --no-cpu-baseline > gpurun_out/bench_$p.json 2> gpurun_out/bench_$p.err; tail -1 gpurun_out/bench_$p.json | cut -c1-900
done
timeout 120 python tools/prof_train_step.py > gpurun_out/train_prof_attn.txt 2>&1; head -3 gpurun_out/train_prof_attn.txt
