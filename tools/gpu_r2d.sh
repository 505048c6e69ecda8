#!/bin/bash
# round 2, call D (2 GPUs): instance-sharded parity (NCCL), bench with the `sharded` object, reference arm
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_dist.py tests/test_gpu_dropout.py tests/test_gpu_baseline_sizes.py tests/test_gpu_umma.py -m gpu -q -p no:cacheprovider > gpurun_out/pytest_gpu_2gpu.log 2>&1; tail -5 gpurun_out/pytest_gpu_2gpu.log | cut -c1-400
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/bench_2gpu.json 2> gpurun_out/bench_2gpu.err; tail -1 gpurun_out/bench_2gpu.json | cut -c1-3000; tail -5 gpurun_out/bench_2gpu.err
timeout 300 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_reference.json 2>&1; tail -1 gpurun_out/bench_reference.json | cut -c1-700
timeout 300 python tools/diag_gate_flips.py > gpurun_out/diag_gate_flips.txt 2>&1; cat gpurun_out/diag_gate_flips.txt
