"""GPU (>= 2 devices): instance-sharded forward == single-GPU forward; global top-k over sharded scores == single-GPU top-k.
Skipped on a single-GPU box; run with `gpurun --gpus 2 -- python -m pytest tests/test_gpu_dist.py -m gpu`."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, ret):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        import cases
        import mhimk
        from mhimk import dist as D
        from mhimk.modules import DAttention
        N = 10007
        m = DAttention(1024, 2, dropout=0.0, act="relu").cuda().eval()
        m.load_state_dict({k: v.cuda() for k, v in cases.abmil_state(3).items()}, strict=True)
        x = cases.make_bag(4, N, 1024)[0].cuda()
        with torch.no_grad():
            ref_logits, ref_attn = m(x[None].clone(), return_attn=True)
        lo, hi = D.row_slices(N, world)[rank]
        logits, stats, s = D.sharded_abmil_forward(m, x[lo:hi].contiguous(), want_scores=True)
        assert cases.rel_err(logits, ref_logits) < 1e-5, cases.rel_err(logits, ref_logits)
        attn_local = torch.exp(s - stats[0]) / stats[1]
        assert cases.rel_err(attn_local, ref_attn[0, lo:hi]) < 1e-4
        full = torch.exp(torch.randn(N, generator=torch.Generator().manual_seed(1))).cuda()
        k = 300
        got = D.global_topk(full[lo:hi].contiguous(), k, lo, N, True)
        assert torch.equal(got, mhimk.ops.topk(full, k, True))
        ret[rank] = "ok"
    except Exception:
        import traceback
        ret[rank] = traceback.format_exc()
    finally:
        dist.destroy_process_group()


def test_instance_sharded_matches_single_gpu():
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 CUDA devices")
    world = min(torch.cuda.device_count(), 8)
    with mp.Manager() as mgr:
        ret = mgr.dict()
        mp.spawn(_worker, args=(world, 29731, ret), nprocs=world, join=True)
        assert all(ret.get(r) == "ok" for r in range(world)), dict(ret)
