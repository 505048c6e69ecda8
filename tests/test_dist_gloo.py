"""world_size-2 (and 3, ragged) gloo runs on CPU of the multi-GPU protocol: shard partial exchange + merge, global top-k
over instance-sharded scores, flat gradient all-reduce, bag/row partitioning helpers."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _lowest_index_topk(score, k, largest):
    """CPU stand-in for mil_topk_f32's total order (value, then lowest index) -- host logic under test is the protocol."""
    key = score.double() if largest else -score.double()
    order = sorted(range(score.numel()), key=lambda i: (-key[i].item(), i))[:k]
    return torch.tensor(order, dtype=torch.int64)


def _worker(rank, world, port, n_rows, ret):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import mhimk
        from mhimk import dist as D
        from oracle import mil_oracle as O
        g = torch.Generator().manual_seed(5)
        s, h = torch.randn(n_rows, generator=g, dtype=torch.float64) * 3, torch.randn(n_rows, 64, generator=g, dtype=torch.float64)
        score = (torch.randint(0, 50, (n_rows,), generator=g).float() / 50)              # heavy ties
        lo, hi = D.row_slices(n_rows, world, multiple=128)[rank]
        # --- partial exchange + merge
        if hi > lo:
            m, l, P = O.pool_partial(s[lo:hi], h[lo:hi])
            part = torch.cat([m[None], l[None], P])
        else:
            part = torch.zeros(66, dtype=torch.float64)
        with pytest.raises(RuntimeError):
            D.exchange_and_merge(part)                                                   # CPU partials are refused by default
        stats, pooled = D.exchange_and_merge(part, allow_host_merge=True)
        ref, _ = O.softmax_pool(s, h)
        assert float((pooled - ref).abs().max()) < 1e-12
        # --- global top-k == single-device top-k incl. tie order
        k = 37
        got = D.global_topk(score[lo:hi], k, lo, n_rows, True, topk_fn=_lowest_index_topk)
        assert got.tolist() == _lowest_index_topk(score, k, True).tolist()
        got = D.global_topk(score[lo:hi], k, lo, n_rows, False, topk_fn=_lowest_index_topk)
        assert got.tolist() == _lowest_index_topk(score, k, False).tolist()
        # --- gradient all-reduce
        p1, p2 = torch.nn.Parameter(torch.zeros(3, 4)), torch.nn.Parameter(torch.zeros(5))
        p1.grad, p2.grad = torch.full((3, 4), float(rank + 1)), torch.arange(5.0) * (rank + 1)
        D.allreduce_grads([p1, p2])
        mean = sum(range(1, world + 1)) / world
        assert torch.allclose(p1.grad, torch.full((3, 4), mean)) and torch.allclose(p2.grad, torch.arange(5.0) * mean)
        ret[rank] = "ok"
    except Exception as e:  # surfaced by the parent
        import traceback
        ret[rank] = traceback.format_exc()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,n_rows", [(2, 1000), (3, 300), (2, 100)])
def test_protocol_over_gloo(world, n_rows):
    port = 29500 + (os.getpid() + world * 7 + n_rows) % 2000
    with mp.get_context("spawn").Manager() as mgr:          # not fork(): the parent already runs torch threads
        ret = mgr.dict()
        mp.spawn(_worker, args=(world, port, n_rows, ret), nprocs=world, join=True)
        assert all(ret.get(r) == "ok" for r in range(world)), dict(ret)


def test_partition_helpers():
    sys.path.insert(0, ROOT)
    import mhimk
    from mhimk import dist as D
    for n, w in [(8, 8), (10, 8), (3, 8), (1000, 7)]:
        parts = [D.bag_slice(n, r, w) for r in range(w)]
        assert sorted(i for p in parts for i in p) == list(range(n))
        assert max(len(p) for p in parts) - min(len(p) for p in parts) <= 1
    for n, w in [(200000, 8), (1000, 8), (129, 2), (5, 4)]:
        sl = D.row_slices(n, w)
        assert sl[0][0] == 0 and sl[-1][1] == n and all(a[1] == b[0] for a, b in zip(sl, sl[1:]))
        assert all((hi - lo) % 128 == 0 for lo, hi in sl[:-1] if hi < n)
