"""Merge / MCA: recycle randomly dropped instances into k tokens (reference: modules/mhim_modules/merge.py:14-203).

The large contraction (to_kv over the dropped rows, L_m x 512 -> 1024) runs in the CUDA GEMM; the k x L_m attention is tiny.
"""
import math

import torch
import torch.nn.functional as F
from torch import nn

from ... import ops
from .. import _common as C
from .masking import select_mask_fn


class MCA(C.MilModule):
    def __init__(self, dim, heads=8, dim_head=64, dropout=0.0):
        super().__init__()
        inner = dim_head * heads
        self.heads, self.scale = heads, dim_head ** -0.5
        self.attend, self.dropout = nn.Softmax(dim=-1), nn.Dropout(dropout)
        self.to_kv = nn.Linear(dim, inner * 2, bias=False)
        self.to_q = nn.Linear(dim, inner, bias=False)
        self.to_out = nn.Sequential(nn.Linear(inner, dim), nn.Dropout(dropout)) if not (heads == 1 and dim_head == dim) else nn.Identity()

    def forward(self, x, _q):
        """x [1,n,d], _q [1,m,d] -> [1,m,d]"""
        kv = C.lin(self.to_kv, x[0])
        q = C.lin(self.to_q, _q[0])
        inner = q.shape[-1]
        if q.shape[0] <= 8 and inner // self.heads <= 64 and kv.shape[1] == 2 * inner:
            # own kernels: scores, softmax over the n instances, weighted sum (and their backward); the attention dropout
            # (merge.py:60) enters as a pre-scaled keep mask drawn by torch's generator
            pmask = None
            if self.training and self.dropout.p > 0:
                pmask = F.dropout(torch.ones((self.heads, q.shape[0], kv.shape[0]), dtype=torch.float32, device=q.device), self.dropout.p, True)
            out = ops.mca_attend(q, kv, self.heads, self.scale, pmask)
        else:
            split = lambda t: t.reshape(t.shape[0], self.heads, -1).permute(1, 0, 2)
            qh, kh, vh = split(q), split(kv[:, :inner]), split(kv[:, inner:])
            attn = self.dropout(self.attend(qh @ kh.transpose(-1, -2) * self.scale))
            out = (attn @ vh).permute(1, 0, 2).reshape(q.shape[0], inner)
        if isinstance(self.to_out, nn.Identity):
            return out[None]
        return self.to_out[1](C.lin(self.to_out[0], out))[None]


class Merge(C.MilModule):
    def __init__(self, dim, heads=8, merge_h_dim=64, dropout=0.1, k=10, g_q_mm=1.0, merge_ratio=0.2, global_q_enable=True, no_merge=False,
                 g_q_grad=False, mask_type="random", **kwargs):
        super().__init__()
        self.norm = nn.LayerNorm(dim)
        self.attn = MCA(dim, heads, merge_h_dim, dropout)
        self.merge_k = self.k = k
        self.no_merge, self.mask_type = no_merge, mask_type
        self.g_q_mm, self.g_q_grad, self.merge_ratio = g_q_mm, g_q_grad, merge_ratio
        self.global_q = None
        if global_q_enable:
            val = math.sqrt(6.0 / float(3 * 16 * 16 + dim))          # VPT-style uniform range (merge.py:106,114)
            if g_q_grad:
                self.global_q_grad = nn.Parameter(torch.empty(1, k, dim).uniform_(-val, val))
            if g_q_mm != 1.0:
                self.global_q_mm = nn.Parameter(torch.empty(1, k, dim).uniform_(-val, val), requires_grad=False)
            if g_q_grad and g_q_mm == 1.0:
                self.global_q = self.global_q_grad
            elif not g_q_grad and g_q_mm != 1.0:
                self.global_q = self.global_q_mm                     # same Parameter under two state_dict keys, as upstream

    @staticmethod
    def _noise(L, device):
        """U(0,1) keys whose argsort is the random keep order (merge.py:164); tests substitute a CPU-seeded stream."""
        return torch.rand(L, device=device)

    def update_q_ema(self, new):
        self.global_q_mm.data.mul_(self.g_q_mm).add_(new, alpha=1.0 - self.g_q_mm)

    def merge(self, x):
        z = self.attn(self.norm(x), self.norm(self.global_q))
        if self.training and self.global_q is not None and self.g_q_mm != 1.0:
            self.update_q_ema(z[:, : self.k].detach())
        return z

    def masking(self, x, attn):
        L = x.shape[1]
        if self.mask_type == "random":
            n_keep = int(L * self.merge_ratio)
            order = torch.argsort(self._noise(L, x.device), dim=0)
        else:                                                        # 'low'
            n_keep, order = select_mask_fn(L, attn, False, 1 - self.merge_ratio)
            order = order.squeeze(0)
        if x.is_cuda and x.shape[0] == 1 and x.dtype == torch.float32 and x.shape[2] % 4 == 0 and order.numel() == L and 0 < n_keep < L:
            keep, drop = ops.split_rows(x[0], order, n_keep)             # one gather each; the backward is a single scatter
            return keep[None], drop[None]
        return x[:, order[:n_keep]], x[:, order[n_keep:]]

    def forward(self, x, attn=None):
        if self.training:
            x_keep, x_drop = self.masking(x, attn)
            if self.no_merge:
                return torch.cat((x_keep, self.global_q), dim=1) if self.global_q is not None else x_keep
            return torch.cat((x_keep, self.merge(x_drop)), dim=1)
        if not self.no_merge:
            return torch.cat((x, self.merge(x)), dim=1)
        return torch.cat((x, self.global_q), dim=1) if self.global_q is not None else x
