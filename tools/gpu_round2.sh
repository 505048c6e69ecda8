#!/bin/bash
# Round-2 measurement set for ONE gpurun call on one B200: GPU parity suite, smoke, bench (both arms), the other BASELINE configs (ours /
# graphed / eager torch CUDA / CPU), training-step profiles, ncu launch lists and one `ncu --set full` capture of the dominant kernels.
set -u
mkdir -p gpurun_out
O=gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/smi.txt 2>&1
( timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider > $O/pytest_gpu.log 2>&1; echo "pytest exit $?" >> $O/pytest_gpu.log ); tail -3 $O/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > $O/smoke.log 2>&1; tail -1 $O/smoke.log
timeout 600 python bench.py --steps 50 --warmup 5 > $O/bench_bf16x3.json 2> $O/bench_bf16x3.err; tail -1 $O/bench_bf16x3.json | cut -c1-400
timeout 600 python bench.py --steps 50 --warmup 5 --precision fp16x3 --no-cpu-baseline > $O/bench_fp16x3.json 2> $O/bench_fp16x3.err
timeout 600 python bench.py --steps 50 --warmup 5 --precision fp16 --no-cpu-baseline > $O/bench_fp16.json 2> $O/bench_fp16.err
timeout 600 python bench.py --impl reference --steps 20 --warmup 3 > $O/bench_reference.json 2> $O/bench_reference.err; tail -1 $O/bench_reference.json | cut -c1-300
CFG_REPS=8 timeout 600 python tools/bench_configs.py > $O/configs_stream.json 2> $O/configs.err; tail -n +2 $O/configs_stream.json | tail -n 60 > /dev/null
for b in attn dsmil; do d=1024; [ $b = dsmil ] && d=1536
  T_BASE=$b T_D=$d timeout 300 python tools/bench_train_step.py > $O/train_step_$b.json 2> $O/train_step_$b.err; tail -1 $O/train_step_$b.json
done
timeout 120 python tools/prof_train_step.py > $O/train_prof_attn.txt 2>&1
timeout 300 python tools/prof_transmil.py > $O/prof_transmil.txt 2>&1; head -1 $O/prof_transmil.txt
# ncu: launch lists (serialised, cold cache: shares, not absolutes) ...
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_bench.csv python bench.py --steps 4 --warmup 3 --no-cpu-baseline > $O/bench_under_ncu.log 2>&1
T_REPS=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file $O/launches_train_step_attn.csv env T_NCU=1 python tools/bench_train_step.py > $O/train_under_ncu.log 2>&1
# ... and one full capture each of the fused forward, the tensor-core weight gradient and the streaming softmax-over-N pool
timeout 900 ncu --set full --clock-control none --import-source on -k regex:mil_fused2 -s 2 -c 1 -o $O/fused2_full -f env PROF_MODES=fused PROF_PREC=bf16x3 PROF_REPS=1 python tools/prof_fused.py > $O/ncu_fused2.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:^wgrad_kernel" -s 1 -c 1 -o $O/wgrad_full -f env T_NCU=1 python tools/bench_train_step.py > $O/ncu_wgrad.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:colsoftmax_pool_kernel -s 2 -c 1 -o $O/colpool_full -f env T_N=50000 python tools/prof_transmil.py > $O/ncu_colpool.log 2>&1
for r in fused2_full wgrad_full colpool_full; do
  [ -f $O/$r.ncu-rep ] && ncu -i $O/$r.ncu-rep --page raw --csv > $O/$r.raw.csv 2>/dev/null
done
ls -la $O | head -60
