#!/bin/bash
# Round-2 multi-GPU measurements on ONE 8-GPU node: NCCL parity of the instance-sharded path at 8 ranks, bench.py at N = 2 / 4 / 8
# (bag-parallel headline + the `sharded` giant-bag object).
set -u
mkdir -p gpurun_out
O=gpurun_out
nvidia-smi --query-gpu=index,name --format=csv > $O/smi_multi.txt 2>&1
timeout 600 python -m pytest tests/test_gpu_dist.py -m gpu -q -p no:cacheprovider -s > $O/pytest_dist_8gpu.log 2>&1; tail -3 $O/pytest_dist_8gpu.log
timeout 300 python tools/prof_transmil.py > $O/prof_transmil.txt 2>&1; head -12 $O/prof_transmil.txt | cut -c1-62,140-215
for n in 8 4 2; do
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29600 + n)) bench.py --gpus $n --steps 20 --warmup 5 \
     > $O/bench_${n}gpu.json 2> $O/bench_${n}gpu.err
  tail -1 $O/bench_${n}gpu.json | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print({k: d[k] for k in ('n_gpus','value','ms_per_step')}, 'e2e', round(d['e2e']['value']/1e6,2), 'Minst/s', round(d['e2e']['h2d_gbs_aggregate'],1), 'GB/s agg')
print(' sharded:', {k: (round(v,3) if isinstance(v,float) else v) for k,v in d['sharded'].items() if k in ('ms_per_bag','local_fused_kernel_ms','exchange_overhead_us','single_gpu_ms_per_bag','speedup_vs_single_gpu')})"
done
