"""Multi-GPU partitioning of the path.  One process per GPU, torch.distributed for the plumbing.

bag-parallel      : bags are independent units -> shard the bag list, NO data-path collective (bag_slice).
instance-sharded  : rank g holds rows [off_g, off_g + n_g) of one giant bag.  Each rank runs the fused pass on its rows and
                    contributes its softmax statistics (m_g, l_g, P_g[H]) = 2 KB; ONE all-gather, then the same
                    log-sum-exp merge on every rank (mil_pool_merge_f32) -> identical pooled vector everywhere.
                    Masked selection needs one more small exchange: local top-k candidates -> global top-k with the same
                    total order (value, then lowest GLOBAL index).
data-parallel grads: one flat all-reduce of the weight gradients per step (absent upstream; optional).

The collectives are latency-bound (2 KB .. 400 KB): they are issued on the compute stream right after the partial kernel.
"""
from typing import Callable, List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist


def bag_slice(n_bags: int, rank: int, world: int) -> range:
    """Bags [lo, hi) owned by `rank` (contiguous, sizes differ by at most one)."""
    base, rem = divmod(n_bags, world)
    lo = rank * base + min(rank, rem)
    return range(lo, lo + base + (1 if rank < rem else 0))


def row_slices(n_rows: int, world: int, multiple: int = 128) -> List[Tuple[int, int]]:
    """Row ranges of an instance-sharded bag: contiguous, rounded to `multiple` rows (the kernel's tile height)."""
    per = -(-n_rows // world)
    per = -(-per // multiple) * multiple
    return [(min(n_rows, r * per), min(n_rows, (r + 1) * per)) for r in range(world)]


def make_partial(stats: torch.Tensor, pooled: torch.Tensor) -> torch.Tensor:
    """(m, l, pooled = P / l) of one shard -> the exchange record [m, l, P[H]]."""
    return torch.cat([stats, pooled * stats[1]])


def _merge_host(parts: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """Reference merge on a [world, 2+H] tensor in torch ops -- protocol tests on CPU (gloo) only."""
    m_i, l_i, P_i = parts[:, 0], parts[:, 1], parts[:, 2:]
    live = l_i > 0
    m = m_i[live].max()
    w = torch.where(live, torch.exp(m_i - m), torch.zeros_like(m_i))
    l = (l_i * w).sum()
    return torch.stack([m, l]), (P_i * w[:, None]).sum(0) / l


def exchange_and_merge(partial: torch.Tensor, group=None, allow_host_merge: bool = False) -> Tuple[torch.Tensor, torch.Tensor]:
    """All-gather this rank's [2+H] partial and merge -> (stats [2], pooled [H]), bit-identical on every rank."""
    world = dist.get_world_size(group)
    flat = torch.empty(world * partial.numel(), dtype=partial.dtype, device=partial.device)
    dist.all_gather_into_tensor(flat, partial.contiguous().reshape(-1), group=group)
    gathered = flat.view(world, partial.numel())
    if gathered.is_cuda:
        from . import ops
        return ops.pool_merge(gathered)
    if not allow_host_merge:
        raise RuntimeError("mhimk.dist: partials are on the CPU; the product path merges on the GPU (mil_pool_merge_f32)")
    return _merge_host(gathered)


def global_topk(score_local: torch.Tensor, k: int, row_offset: int, n_global: int, largest: bool = True, group=None,
                topk_fn: Optional[Callable] = None) -> torch.Tensor:
    """Global top-k over an instance-sharded score vector -> int64 [k] GLOBAL indices, same on every rank and equal to the
    single-GPU result on the concatenated scores (value order, ties lowest global index first).

    Each rank contributes its local top-min(k, n_local) (score, global index) pairs; candidates are scattered into a dense
    n_global vector filled with the worst possible value and selected once more.
    """
    if topk_fn is None:
        from . import ops
        topk_fn = ops.topk
    world = dist.get_world_size(group)
    n_local = score_local.numel()
    kl = min(k, n_local)
    cand_idx = topk_fn(score_local, kl, largest) if kl > 0 else torch.empty(0, dtype=torch.int64, device=score_local.device)
    rec = torch.full((k, 2), -1.0, dtype=torch.float64, device=score_local.device)       # (global index, score); -1 = empty slot
    if kl > 0:
        rec[:kl, 0] = (cand_idx + row_offset).double()
        rec[:kl, 1] = score_local[cand_idx].double()
    flat = torch.empty(world * k * 2, dtype=torch.float64, device=score_local.device)
    dist.all_gather_into_tensor(flat, rec.reshape(-1), group=group)
    allrec = flat.view(world * k, 2)
    valid = allrec[:, 0] >= 0
    fill = float("-inf") if largest else float("inf")
    dense = torch.full((n_global,), fill, dtype=score_local.dtype, device=score_local.device)
    dense[allrec[valid, 0].long()] = allrec[valid, 1].to(score_local.dtype)
    return topk_fn(dense, k, largest)


def allreduce_grads(params: Sequence[torch.nn.Parameter], group=None, average: bool = True) -> None:
    """One flat all-reduce of the weight gradients (data-parallel training over bags)."""
    grads = [p.grad for p in params if p.grad is not None]
    if not grads:
        return
    flat = torch.cat([g.reshape(-1) for g in grads])
    dist.all_reduce(flat, group=group)
    if average:
        flat /= dist.get_world_size(group)
    o = 0
    for g in grads:
        g.copy_(flat[o:o + g.numel()].view_as(g))
        o += g.numel()


@torch.no_grad()
def sharded_abmil_forward(model, x_local: torch.Tensor, group=None, want_scores: bool = False):
    """Instance-sharded forward of a mhimk DAttention (BASELINE config 5: giant bag split along N).

    x_local: this rank's rows [n_g, D] (may be empty on trailing ranks).  Returns (logits [1,C], stats, s_local or None).
    Three launches per rank: the fused pass (its tail writes the exchange record (m, l, P[H]) itself), ONE all-gather of
    2056 B per rank (NCCL, on the compute stream), one merge + classifier kernel (mil_shard_merge_cls_f32).
    """
    from . import ops
    f0, a0, a2 = model.feature[0], model.attention[0], model.attention[2]
    H = f0.out_features
    world = dist.get_world_size(group)
    if x_local.shape[0] > 0:
        out = ops.abmil_fused_forward(x_local, f0.weight, f0.bias, model.act, a0.weight, a0.bias, a2.weight, a2.bias, "tanh",
                                      want_scores=want_scores, precision=model.precision, volatile=model.training, want_record=True)
        rec, s = out["record"], out["s"]
    else:
        rec, s = torch.zeros(2 + H, device=x_local.device), None
    gathered = torch.empty((world, 2 + H), dtype=torch.float32, device=x_local.device)
    dist.all_gather_into_tensor(gathered.view(-1), rec, group=group)
    stats, _, logits = ops.shard_merge_cls(gathered, model.classifier.weight, model.classifier.bias)
    return logits, stats, s
