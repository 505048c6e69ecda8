"""PPEG positional encoding used by MHIM's SAttention (reference: modules/emb_position.py:85-120)."""
import math

import torch
from torch import nn

from .. import ops


class SINCOS(nn.Module):
    """2-D sin/cos positional embedding added to the embedded instances (reference: modules/emb_position.py:5-84, used by
    abmil.DAttention(pos='sincos'), abmil.py:162-163, 214-215).  `pos` [1, N+1, 2] or [N+1, 2]: row 0 = (W, H) of the slide's patch grid,
    rows 1.. = (x, y) of every instance.  The reference builds the full H x W table and gathers row y * W + x; the entry only depends on
    (x, y), so it is evaluated directly: [sin(x w), cos(x w), sin(y w), cos(y w)], w_k = 10000^(-k / (C/4)), k < C/4."""

    def forward(self, x, pos=None):
        B, N, C = x.shape
        if pos is None:
            raise RuntimeError("SINCOS needs the patch coordinates `pos`")
        if pos.dim() == 3:
            if pos.size(0) != 1:
                raise RuntimeError("mhimk SINCOS: batch must be 1 bag")
            pos = pos[0]
        xy = pos[1:].to(device=x.device, dtype=torch.float32)
        quarter = C // 4
        omega = 1.0 / (10000 ** (torch.arange(quarter, dtype=torch.float32, device=x.device) / quarter))
        ax, ay = xy[:, 0:1] * omega, xy[:, 1:2] * omega
        emb = torch.cat([torch.sin(ax), torch.cos(ax), torch.sin(ay), torch.cos(ay)], dim=1)
        return x + emb[None]


class PPEG(nn.Module):
    def __init__(self, dim=512, k=7, conv_1d=False, bias=True):
        super().__init__()
        mk = lambda ks: nn.Conv2d(dim, dim, (ks, 1) if conv_1d else ks, 1, (ks // 2, 0) if conv_1d else ks // 2, groups=dim, bias=bias)
        self.proj, self.proj1, self.proj2 = mk(k), mk(5), mk(3)

    def forward(self, x):
        if x.dim() == 2:
            x = x.unsqueeze(0)
        B, N, Cc = x.shape
        H = W = int(math.ceil(math.sqrt(N)))
        add = H * W - N
        x = torch.cat([x, x[:, :add]], dim=1)                   # wrap-pad to a square
        if H < 7:                                                # minimum 7x7 grid, zero filled
            H = W = 7
            zp = H * W - (N + add)
            x = torch.cat([x, x.new_zeros(B, zp, Cc)], dim=1)
            add += zp
        convs = (self.proj, self.proj1, self.proj2)
        own = (B == 1 and x.is_cuda and x.dtype == torch.float32 and all(c.kernel_size[0] == c.kernel_size[1] and c.kernel_size[0] <= 7 for c in convs)
               and not (torch.is_grad_enabled() and (x.requires_grad or self.proj.weight.requires_grad)))
        if own:                                                  # one depth-wise 7x7 kernel of the library (summed kernels + identity)
            y = ops.ppeg_forward(x[0].contiguous(), H, W, convs)[None]
            return y[:, :-add] if add > 0 else y
        g = x.transpose(1, 2).reshape(B, Cc, H, W)
        y = (self.proj(g) + g + self.proj1(g) + self.proj2(g)).flatten(2).transpose(1, 2)
        return y[:, :-add] if add > 0 else y
