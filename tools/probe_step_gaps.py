"""Per-step GPU timeline of the bench loop under torchrun: where does the fixed start-up cost of a short timed region come from?"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch, torch.distributed as dist
import cases, mhimk
from mhimk.modules import DAttention
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local); dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
m = DAttention(1024, 2, dropout=0.0, act="relu").to(dev).eval()
bags = [torch.randn(1, 50000, 1024, device=dev) for _ in range(4)]
mode = os.environ.get("PROBE_SYNC", "barrier")
with torch.no_grad():
    for i in range(5):
        m(bags[i % 4])
    for trial in range(3):
        if world > 1 and mode == "barrier":
            dist.barrier()
        torch.cuda.synchronize()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(21)]
        ev[0].record()
        for i in range(20):
            m(bags[i % 4])
            ev[i + 1].record()
        torch.cuda.synchronize()
        d = [ev[i].elapsed_time(ev[i + 1]) * 1e3 for i in range(20)]
        if rank == 0:
            print(f"world={world} sync={mode} trial {trial}: total {sum(d):.0f} us; per step us:", " ".join(f"{x:.0f}" for x in d), flush=True)
if world > 1:
    dist.destroy_process_group()
