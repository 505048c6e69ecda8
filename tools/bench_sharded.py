"""BASELINE config 4: giant bag N=200 000 x D=1024 instance-sharded over the ranks of one node (torchrun).
Each rank streams its rows through the fused pass; ONE all-gather of (m, l, P[512]) merges the shards.
    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 tools/bench_sharded.py"""
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import cases  # noqa: E402
import mhimk  # noqa: E402
from mhimk import dist as D  # noqa: E402
from mhimk.modules import DAttention  # noqa: E402

N, DIM, STEPS = int(os.environ.get("SHARD_N", 200000)), 1024, int(os.environ.get("SHARD_STEPS", 30))
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
m = DAttention(DIM, 2, dropout=0.0, act="relu").to(dev).eval()
m.load_state_dict({k: v.to(dev) for k, v in cases.abmil_state(2021).items()}, strict=True)
m.precision = os.environ.get("SHARD_PREC", "bf16x3")
lo, hi = D.row_slices(N, world)[rank]
xs = [torch.randn(hi - lo, DIM, device=dev, generator=torch.Generator(device=dev).manual_seed(7 + rank + 100 * i)) for i in range(3)]


def step(i):
    with torch.no_grad():                       # inference path on every rank count (without it the 1-GPU arm ran the autograd path)
        if world > 1:
            return D.sharded_abmil_forward(m, xs[i % 3])[0]
        return m(xs[i % 3][None])


for i in range(5):
    step(i)
if world > 1:
    dist.barrier()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for i in range(STEPS):
    out = step(i)
e1.record()
if world > 1:
    dist.barrier()
torch.cuda.synchronize()
t = torch.tensor([e0.elapsed_time(e1) / STEPS], device=dev, dtype=torch.float64)
if world > 1:
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
if rank == 0:
    ms = float(t.item())
    print(json.dumps({"workload": f"abmil.DAttention eval fwd, giant bag N={N} x D={DIM}, instance-sharded x{world}", "precision": m.precision,
                      "ms_per_bag": ms, "instances_per_s": N / (ms * 1e-3), "scaling": "strong", "n_gpus": world,
                      "exchange": "1 all-gather of (m, l, P[512]) = 2056 B per rank"}))
if world > 1:
    dist.destroy_process_group()
