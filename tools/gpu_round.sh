#!/bin/bash
# One gpurun call: GPU parity tests, bench (both arms), ncu launch list + full capture of the fused kernel.
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
( timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log ) 
tail -5 gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; tail -2 gpurun_out/smoke.log
for p in bf16x3 fp16; do
  timeout 600 python bench.py --steps 50 --warmup 5 --precision $p > gpurun_out/bench_$p.json 2> gpurun_out/bench_$p.err; tail -1 gpurun_out/bench_$p.json
done
timeout 600 python bench.py --steps 50 --warmup 5 --precision fp16 --pipeline pair --no-cpu-baseline > gpurun_out/bench_fp16_pair.json 2> gpurun_out/bench_fp16_pair.err; tail -1 gpurun_out/bench_fp16_pair.json
timeout 600 python bench.py --steps 50 --warmup 5 --precision bf16x3 --pipeline single --no-cpu-baseline > gpurun_out/bench_bf16x3_single.json 2> gpurun_out/bench_bf16x3_single.err; tail -1 gpurun_out/bench_bf16x3_single.json
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_reference.json 2>&1; tail -1 gpurun_out/bench_reference.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
   python bench.py --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:mil_fused -s 2 -c 2 -o gpurun_out/fused_full -f \
   env PROF_MODES=fused PROF_REPS=1 python tools/prof_fused.py > gpurun_out/ncu_full.log 2>&1
timeout 300 env PROF_REPS=10 python tools/prof_fused.py > gpurun_out/prof_fused.log 2>&1; cat gpurun_out/prof_fused.log
timeout 300 python tools/trace_fused.py > gpurun_out/trace_fused.log 2>&1; tail -30 gpurun_out/trace_fused.log
MHIMK_PIPELINE=2 PROF_PREC=fp16 timeout 300 python tools/trace_fused.py > gpurun_out/trace_fused_pair.log 2>&1; tail -12 gpurun_out/trace_fused_pair.log
MHIMK_PIPELINE=1 PROF_PREC=bf16x3 timeout 300 python tools/trace_fused.py > gpurun_out/trace_fused_single_bf16x3.log 2>&1; tail -8 gpurun_out/trace_fused_single_bf16x3.log
# the other BASELINE.json configs, the training-step profile and the EMA update (round 1, run E)
CFG_REPS=5 timeout 200 python tools/bench_configs.py > gpurun_out/configs.json 2> gpurun_out/configs.err; tail -c 1500 gpurun_out/configs.json
timeout 60 python tools/prof_train_step.py > gpurun_out/train_prof_attn.txt 2>&1; head -1 gpurun_out/train_prof_attn.txt
T_BASE=dsmil T_D=1536 timeout 60 python tools/prof_train_step.py > gpurun_out/train_prof_dsmil.txt 2>&1; head -1 gpurun_out/train_prof_dsmil.txt
timeout 40 python tools/time_ema.py > gpurun_out/time_ema.txt 2>&1; tail -2 gpurun_out/time_ema.txt
ls -la gpurun_out
