"""Masked hard-instance selection on the device (reference: modules/mhim_modules/masking.py:9-110).

Same signatures and return values as the reference -- (len_keep: python int, mask_ids: int64 [1, ps] = kept ids
ascending followed by the masked ids) -- but top-k, complement and concatenation run in mil_topk_f32 /
mil_mask_from_indices: no `.tolist()` / python-set round trip (masking.py:78-80), no host sync on the common path.
Tie rule (the reference inherits torch.topk's unspecified one): value, then LOWEST index first.
"""
import math

import torch

from ... import ops


def _k_of(ps: int, ratio: float) -> int:
    return int(math.ceil(ps * ratio))                       # python float product + ceil, as masking.py:61


def select_mask_fn(ps, attn, largest, mask_ratio, mask_ids_other=None, len_keep_other=None, cls_attn_topk_idx_other=None,
                   random_ratio=1.0, select_inv=False, msa_fusion="vote"):
    ps_eff, ratio0 = ps, mask_ratio
    mask_ratio = mask_ratio / random_ratio
    if mask_ratio > 1:                                       # clamp branch (masking.py:33-35)
        random_ratio, mask_ratio = ratio0, 1.0
    if mask_ids_other is not None and cls_attn_topk_idx_other is None:
        cls_attn_topk_idx_other = mask_ids_other[:, len_keep_other:].squeeze()
        ps_eff = ps - cls_attn_topk_idx_other.size(0)

    if attn.dim() > 2:                                       # per-head attention [1, h, N]
        heads = attn.size(1)
        if msa_fusion == "mean":
            k = int(math.ceil(ps_eff * mask_ratio) // heads)
            idx = torch.unique(torch.cat([ops.topk(attn[0, h], k, largest) for h in range(heads)]))
        else:                                                # 'vote' (masking.py:49-59)
            k = _k_of(ps_eff, mask_ratio)
            votes = torch.zeros(ps, dtype=torch.float32, device=attn.device)
            for h in range(heads):
                votes.index_add_(0, ops.topk(attn[0, h], k, largest), torch.ones(k, device=attn.device))
            idx = ops.topk(votes, k, True)
    else:
        idx = ops.topk(attn.reshape(-1), _k_of(ps_eff, mask_ratio), bool(largest))

    if random_ratio < 1.0:                                   # random subset of the top-k (masking.py:66-71)
        n = idx.size(0)
        perm = torch.randperm(n, device=idx.device)
        idx = idx[perm[: int(math.ceil(n * random_ratio))]]
    if mask_ids_other is not None:                           # union with an earlier mask (v1 paths only)
        idx = torch.cat([idx, cls_attn_topk_idx_other.reshape(-1)]).unique()

    mask_ids, keep, _ = ops.mask_from_indices(idx, ps)
    len_keep = ps - idx.size(0)                              # sizes are known on the host: no sync
    if select_inv:
        return ps - len_keep, torch.cat([idx, mask_ids[0, :len_keep]]).unsqueeze(0)
    return len_keep, mask_ids


def mask_fn(x, ids_shuffle=None, len_keep=None):
    """Rows of x [1, L, D] at the first len_keep ids (masking.py:91-110)."""
    assert ids_shuffle is not None
    if x.is_cuda and x.dim() == 3 and x.shape[0] == 1 and x.dtype == torch.float32 and x.shape[2] % 4 == 0 and ids_shuffle.shape[1] == x.shape[1]:
        # mask_ids is a permutation of the rows: own gather, and a backward that writes every row once (no index sort)
        return ops.take_rows(x[0], ids_shuffle[0], len_keep)[None]
    return x[:, ids_shuffle[0, :len_keep]]
