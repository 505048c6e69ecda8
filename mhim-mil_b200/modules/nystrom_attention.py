"""Nystrom self-attention with the reference's interface (modules/nystrom_attention.py:12-27, 31-152).

Without autograd (inference, MHIM's teacher pass) the whole layer runs on the library's own kernels in the streaming form of SURVEY 9.7
(ops.nystrom_attention_forward): tensor cores for to_qkv, q k_l^T, k q_l^T, softmax . Z and to_out; CUDA-core kernels for the landmark
means, the softmaxes, the softmax-over-N aggregation, the 256 x 256 pseudo-inverse products (batched fp32 GEMM) and the 33-tap
residual convolution -- no cuBLAS / cuDNN call.  With autograd (the selfattn student) the N-row projections run in the library's
GEMMs and the landmark algebra is expressed with differentiable torch ops.
"""
from math import ceil

import torch
import torch.nn.functional as F
from torch import nn

from .. import ops
from . import _common as C


def moore_penrose_iter_pinv(x, iters=6):
    """6-step iterative pseudo-inverse; ONE global scalar normaliser over all heads (nystrom_attention.py:18)."""
    ax = x.abs()
    z = x.transpose(-1, -2) / (ax.sum(dim=-1).max() * ax.sum(dim=-2).max())
    eye = torch.eye(x.shape[-1], device=x.device, dtype=x.dtype)[None]
    for _ in range(iters):
        xz = x @ z
        z = 0.25 * z @ (13 * eye - xz @ (15 * eye - xz @ (7 * eye - xz)))
    return z


class NystromAttention(C.MilModule):
    def __init__(self, dim, dim_head=64, heads=8, num_landmarks=256, pinv_iterations=6, residual=True, residual_conv_kernel=33, eps=1e-8,
                 dropout=0.0):
        super().__init__()
        inner = heads * dim_head
        self.eps, self.num_landmarks, self.pinv_iterations, self.heads, self.scale = eps, num_landmarks, pinv_iterations, heads, dim_head ** -0.5
        self.to_qkv = nn.Linear(dim, inner * 3, bias=False)
        self.to_out = nn.Sequential(nn.Linear(inner, dim), nn.Dropout(dropout))
        self.residual = residual
        if residual:
            self.res_conv = nn.Conv2d(heads, heads, (residual_conv_kernel, 1), padding=(residual_conv_kernel // 2, 0), groups=heads, bias=False)

    def forward(self, x, attn_mask=None, return_attn=False, no_norm=False):
        if attn_mask is not None:
            raise NotImplementedError("attn_mask is dead code upstream (nystrom_attention.py:122 references an undefined name)")
        C.require_cuda(x, "NystromAttention")
        b, n, dim = x.shape
        if b != 1:
            raise RuntimeError("mhimk NystromAttention: batch must be 1 bag")
        h, m, iters = self.heads, self.num_landmarks, self.pinv_iterations
        if (not C.grad_needed(self, x) and m <= 256 and (self.to_qkv.weight.shape[0] // 3) // h <= 64 and x.dtype == torch.float32 and n > 1
                and (not self.residual or self.res_conv.weight.shape[2] <= 33)):
            res = ops.nystrom_attention_forward(x[0], self.to_qkv.weight, self.to_out[0].weight, self.to_out[0].bias,
                                                self.res_conv.weight if self.residual else None, h, m, iters, self.scale, return_attn=return_attn,
                                                no_norm=no_norm, volatile=self.training)
            if not return_attn:
                return self.to_out[1](res)[None]
            return self.to_out[1](res[0])[None], res[1][None], res[2][None]
        t = x[0]
        pad = (m - n % m) % m                                   # FRONT zero padding to a multiple of m (:70-73)
        if pad:
            t = torch.cat([t.new_zeros(pad, dim), t], dim=0)
        npad = t.shape[0]
        qkv = C.lin(self.to_qkv, t)
        inner = qkv.shape[-1] // 3
        split = lambda u: u.reshape(npad, h, -1).permute(1, 0, 2)
        q, k, v = split(qkv[:, :inner]) * self.scale, split(qkv[:, inner:2 * inner]), split(qkv[:, 2 * inner:])
        l = ceil(n / m)
        q_l = q.reshape(h, m, l, -1).sum(dim=2) / l             # landmark = segment sum / l (:93-109)
        k_l = k.reshape(h, m, l, -1).sum(dim=2) / l
        s1, s2, s3 = q @ k_l.transpose(-1, -2), q_l @ k_l.transpose(-1, -2), q_l @ k.transpose(-1, -2)
        a1, a2, a3 = s1.softmax(dim=-1), s2.softmax(dim=-1), s3.softmax(dim=-1)
        a2i = moore_penrose_iter_pinv(a2, iters)
        out = (a1 @ a2i) @ (a3 @ v)
        if self.residual:
            out = out + self.res_conv(v[None])[0]
        out = out.permute(1, 0, 2).reshape(npad, inner)
        out = self.to_out[1](C.lin(self.to_out[0], out))[-n:][None]
        if not return_attn:
            return out
        if no_norm:
            r = (s1[:, -n][:, None] @ moore_penrose_iter_pinv(s2, iters)) @ s3
        else:
            r = (a1[:, -n][:, None] @ a2i) @ a3
        return out, r[None, :, 0, -n + 1:], v[None, :, -n + 1:]
