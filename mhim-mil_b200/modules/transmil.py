"""TransMIL with the reference's interface (modules/transmil.py:23-175)."""
import math

import torch
from torch import nn

from .. import ops
from . import _common as C
from .nystrom_attention import NystromAttention


class TransLayer(C.MilModule):
    def __init__(self, norm_layer=nn.LayerNorm, dim=512, n_heads=8):
        super().__init__()
        self.norm = norm_layer(dim)
        self.attn = NystromAttention(dim=dim, dim_head=dim // n_heads, heads=n_heads, num_landmarks=dim // 2, pinv_iterations=6, residual=True,
                                     dropout=0.1)

    def forward(self, x, need_attn=False, need_v=False, no_norm=False):
        xn = C.layer_norm(self.norm, x)
        if need_attn:
            z, attn, v = self.attn(xn, return_attn=True, no_norm=no_norm)
            return (x + z, attn, v) if need_v else (x + z, attn)
        return x + self.attn(xn)


class PPEG(C.MilModule):
    def __init__(self, dim=512):
        super().__init__()
        self.proj = nn.Conv2d(dim, dim, 7, 1, 7 // 2, groups=dim)
        self.proj1 = nn.Conv2d(dim, dim, 5, 1, 5 // 2, groups=dim)
        self.proj2 = nn.Conv2d(dim, dim, 3, 1, 3 // 2, groups=dim)

    def forward(self, x, H, W):
        B, _, Cc = x.shape
        cls, tok = x[:, :1], x[:, 1:]
        if B == 1 and x.is_cuda and not C.grad_needed(self, x) and x.dtype == torch.float32:
            return torch.cat((cls, ops.ppeg_forward(tok[0].contiguous(), H, W, (self.proj, self.proj1, self.proj2))[None]), dim=1)
        g = tok.transpose(1, 2).reshape(B, Cc, H, W)
        y = (self.proj(g) + g + self.proj1(g) + self.proj2(g)).flatten(2).transpose(1, 2)
        return torch.cat((cls, y), dim=1)


def _init_transmil(module):
    for m in module.modules():
        if isinstance(m, (nn.Conv2d, nn.Linear)):
            nn.init.xavier_normal_(m.weight)
            if m.bias is not None:
                nn.init.zeros_(m.bias)
        elif isinstance(m, nn.LayerNorm):
            nn.init.ones_(m.weight)
            if m.bias is not None:
                nn.init.zeros_(m.bias)


class TransMIL(C.MilModule):
    def __init__(self, input_dim, n_classes, dropout, act, mil_norm=None, mil_bias=True, inner_dim=512, embed_feat=True, pos="ppeg", n_heads=8,
                 **kwargs):
        super().__init__()
        if mil_norm not in (None, "none"):
            raise NotImplementedError("mhimk TransMIL: mil_norm='bn'/'ln' is outside the accelerated path")
        self.pos, self.mil_norm = pos, None
        self.pos_layer = nn.Identity() if pos == "none" else PPEG(dim=inner_dim)
        self.act = act.lower() if act.lower() in ("relu", "gelu") else "none"
        self.embed_feat, self.p_drop = embed_feat, 0.25 if dropout else 0.0
        feat = []
        if embed_feat:
            feat += [nn.Linear(input_dim, inner_dim, bias=mil_bias)]
            if self.act != "none":
                feat += [C.act_module(self.act)]
            if dropout:
                feat += [nn.Dropout(0.25)]
        self.feature = nn.Sequential(*feat) if feat else nn.Identity()
        self.norm1 = nn.Identity()
        self.cls_token = nn.Parameter(torch.randn(1, 1, inner_dim) * 1e-6)
        self.n_classes = n_classes
        self.layer1, self.layer2 = TransLayer(dim=inner_dim, n_heads=n_heads), TransLayer(dim=inner_dim, n_heads=n_heads)
        self.norm = nn.LayerNorm(inner_dim)
        self.classifier = nn.Linear(inner_dim, n_classes, bias=mil_bias)
        _init_transmil(self)

    def forward(self, x, return_attn=False, return_act=False, **kwargs):
        C.require_cuda(x, "TransMIL")
        if x.dim() == 2:
            x = x.unsqueeze(0)
        h = x
        if self.embed_feat:
            h = C.lin(self.feature[0], x[0], self.act)[None]
            if self.training and self.p_drop > 0:
                h = torch.nn.functional.dropout(h, self.p_drop, True)
        n0 = h.shape[1]
        side = int(math.ceil(math.sqrt(n0)))
        add = side * side - n0
        h = torch.cat([h, h[:, :add]], dim=1)                            # wrap-pad to a square (:124-127)
        h = torch.cat((self.cls_token.expand(h.shape[0], -1, -1), h), dim=1)
        attn, v = [], None
        if return_attn:
            h, a, v = self.layer1(h, need_attn=True, need_v=True)
            attn.append((a[:, :, :-add] if add > 0 else a).clone())
        else:
            h = self.layer1(h)
        if self.pos != "none":
            h = self.pos_layer(h, side, side)
        if return_attn:
            h, a = self.layer2(h, need_attn=True)
            attn.append((a[:, :, :-add] if add > 0 else a).clone())
        else:
            h = self.layer2(h)
        logits = C.lin(self.classifier, C.layer_norm(self.norm, h[:, :1])[:, 0])          # only the cls row is read (transmil.py:164-165)
        if return_attn:
            out = [logits, attn]
            if return_act:
                out.append(v)
            return out
        return logits
