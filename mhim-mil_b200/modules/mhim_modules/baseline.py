"""MHIM's encoders on pre-embedded instances h [1, L, 512] (reference: modules/mhim_modules/baseline.py:8-288).

DAttention / Attention / AttentionGated : attention pooling           -> CUDA GEMMs + mil_softmax_pool
BClassifier / DSMIL                     : critical-instance attention -> CUDA GEMMs + one mil_softmax_pool per class
TransLayer / SAttention                 : 2 x Nystrom + PPEG          -> see nystrom_attention.py
"""
import math

import torch
from einops import repeat
from torch import nn

from ... import ops
from .. import _common as C
from ..emb_position import PPEG
from ..nystrom_attention import NystromAttention


def _act_name(act):
    return act if act in ("gelu", "relu", "tanh") else "none"


class Attention(C.MilModule):
    def __init__(self, input_dim=512, act="relu", bias=False, dropout=False):
        super().__init__()
        self.L, self.D, self.K = input_dim, 128, 1
        self.act, self.p_drop = _act_name(act), 0.25 if dropout else 0.0
        layers = [nn.Linear(self.L, self.D, bias=bias)]
        if self.act != "none":
            layers += [C.act_module(self.act)]
        if dropout:
            layers += [nn.Dropout(0.25)]
        layers += [nn.Linear(self.D, self.K, bias=bias)]
        self.attention = nn.Sequential(*layers)

    def logits(self, h):
        u = C.lin(self.attention[0], h, self.act)
        if self.training and self.p_drop > 0:
            u = torch.nn.functional.dropout(u, self.p_drop, True)
        return C.lin(self.attention[-1], u)[:, 0]

    def forward(self, x, no_norm=False):
        h = x[0]
        s = self.logits(h)
        pooled, attn = ops.softmax_pool(s, h)
        return pooled[None, None], (s if no_norm else attn)[None, None]


class AttentionGated(C.MilModule):
    def __init__(self, input_dim=512, act="relu", bias=False, dropout=False):
        super().__init__()
        self.L, self.D, self.K = input_dim, 128, 1
        self.act, self.p_drop = _act_name(act), 0.25 if dropout else 0.0
        a = [nn.Linear(self.L, self.D, bias=bias)] + ([C.act_module(self.act)] if self.act != "none" else [])
        b = [nn.Linear(self.L, self.D, bias=bias), nn.Sigmoid()]
        if dropout:
            a += [nn.Dropout(0.25)]
            b += [nn.Dropout(0.25)]
        self.attention_a, self.attention_b = nn.Sequential(*a), nn.Sequential(*b)
        self.attention_c = nn.Linear(self.D, self.K, bias=bias)

    def forward(self, x, no_norm=False):
        h = x[0]
        ga, gb = C.lin(self.attention_a[0], h, self.act), C.lin(self.attention_b[0], h, "sigmoid")
        if self.training and self.p_drop > 0:
            ga, gb = torch.nn.functional.dropout(ga, self.p_drop, True), torch.nn.functional.dropout(gb, self.p_drop, True)
        s = C.lin(self.attention_c, ga * gb)[:, 0]
        pooled, attn = ops.softmax_pool(s, h)
        return pooled[None, None], (s if no_norm else attn)[None, None]


class DAttention(C.MilModule):
    def __init__(self, input_dim=512, act="relu", gated=False, bias=False, dropout=False):
        super().__init__()
        self.gated = gated
        self.attention = AttentionGated(input_dim, act, bias, dropout) if gated else Attention(input_dim, act, bias, dropout)

    def forward(self, x, return_attn=False, no_norm=False, return_act=False, **kwargs):
        C.require_cuda(x, "DAttention")
        pooled, attn = self.attention(x, no_norm)
        if return_attn:
            out = [pooled.squeeze(1), attn.squeeze(1)]
            if return_act:
                out.append(x)
            return out
        return pooled.squeeze(1)


class BClassifier(C.MilModule):
    def __init__(self, input_size, output_class, dropout_v=0.0, nonlinear=True, passing_v=True):
        super().__init__()
        self.q = (nn.Sequential(nn.Linear(input_size, 128), nn.ReLU(), nn.Linear(128, 128), nn.Tanh()) if nonlinear
                  else nn.Linear(input_size, 128))
        self.v = nn.Sequential(nn.Dropout(dropout_v), nn.Linear(input_size, input_size), nn.ReLU()) if passing_v else nn.Identity()
        self.fcc = nn.Conv1d(output_class, output_class, kernel_size=input_size)
        self.nonlinear, self.passing_v = nonlinear, passing_v

    def _q(self, t):
        if not self.nonlinear:
            return C.lin(self.q, t)
        return C.lin(self.q[2], C.lin(self.q[0], t, "relu"), "tanh")

    def forward(self, feats, c, no_norm=False):
        """feats [N,K], c [N,C] -> (pred [1,C], A [N,C], B [1,C,K])"""
        V = feats
        if self.passing_v:
            V = C.lin(self.v[1], self.v[0](feats), "relu")
        Q = self._q(feats)
        crit, _ = ops.col_argmax(c)                                    # critical instance per class (the reference sorts all N rows, :137)
        q_max = self._q(feats.index_select(0, crit))
        logit = ops.linear_act(Q, q_max, None, "none") / math.sqrt(Q.shape[1])
        Bs, As = [], []
        for j in range(logit.shape[1]):
            pooled, a = ops.softmax_pool(logit[:, j], V)
            Bs.append(pooled)
            As.append(a)
        B = torch.stack(Bs)[None]                                      # [1,C,K]
        A = logit if no_norm else torch.stack(As, dim=1)
        pred = C.conv1d_full(self.fcc, B).view(1, -1)
        return pred, A, B


class DSMIL(C.MilModule):
    def __init__(self, n_classes=2, mask_ratio=0.0, mlp_dim=512, cls_attn=True, attn_index="max"):
        super().__init__()
        self.i_classifier = nn.Sequential(nn.Linear(mlp_dim, n_classes))
        self.b_classifier = BClassifier(mlp_dim, n_classes)
        self.cls_attn, self.attn_index, self.mask_ratio = cls_attn, attn_index, mask_ratio

    def attention(self, x, no_norm=False, return_attn=False, return_cam=False, **kwargs):
        feats = x.squeeze(0)
        classes = C.lin(self.i_classifier[0], feats)
        pred, A, B = self.b_classifier(feats, classes, no_norm)
        crit, _ = ops.col_argmax(classes)                              # max over instances = gather of the critical rows (differentiable)
        inst = classes[crit, torch.arange(classes.shape[1], device=crit.device)]
        attn = None
        if return_attn:
            src = classes if self.cls_attn else A
            attn = (src.max(dim=-1).values if self.attn_index == "max" else src[:, int(self.attn_index)]).unsqueeze(0)
            if return_cam:
                attn = [attn, classes.unsqueeze(0)]
        return pred, inst.unsqueeze(0), attn, B

    def forward(self, x, return_attn=False, no_norm=False, **kwargs):
        C.require_cuda(x, "DSMIL")
        logits, inst, attn, B = self.attention(x, no_norm, return_attn=return_attn, **kwargs)
        return ([logits, inst], B, attn) if return_attn else ([logits, inst], B)


class TransLayer(C.MilModule):
    def __init__(self, norm_layer=nn.LayerNorm, dim=512, head=8):
        super().__init__()
        self.norm = norm_layer(dim)
        self.attn = NystromAttention(dim=dim, dim_head=dim // 8, heads=head, num_landmarks=dim // 2, pinv_iterations=6, residual=True, dropout=0.1)

    def forward(self, x, need_attn=False, need_v=False, no_norm=False):
        xn = C.layer_norm(self.norm, x)
        if need_attn:
            z, attn, v = self.attn(xn, return_attn=True, no_norm=no_norm)
            return (x + z, attn, v) if need_v else (x + z, attn)
        return x + self.attn(xn)


class SAttention(C.MilModule):
    def __init__(self, mlp_dim=512, pos_pos=0, pos="ppeg", peg_k=7, head=8):
        super().__init__()
        self.norm = nn.LayerNorm(mlp_dim)
        self.cls_token = nn.Parameter(torch.randn(1, 1, mlp_dim))
        self.layer1, self.layer2 = TransLayer(dim=mlp_dim, head=head), TransLayer(dim=mlp_dim, head=head)
        if pos != "ppeg":
            raise NotImplementedError("mhimk SAttention: only pos='ppeg' (the MHIM default) is provided")
        self.pos_embedding = PPEG(dim=mlp_dim, k=peg_k)
        self.pos_pos = pos_pos

    def forward(self, x, return_attn=False, return_act=False, no_norm=False, **kwargs):
        C.require_cuda(x, "SAttention")
        attn, v = [], None
        if self.pos_pos == -2:
            x = self.pos_embedding(x)
        x = torch.cat((repeat(self.cls_token, "1 n d -> b n d", b=x.shape[0]), x), dim=1)
        if self.pos_pos == -1:
            x = self.pos_embedding(x)
        if return_attn:
            x, a, v = self.layer1(x, need_attn=True, need_v=True, no_norm=no_norm)
            attn.append(a.clone())
        else:
            x = self.layer1(x)
        if self.pos_pos == 0:
            x = torch.cat((x[:, :1], self.pos_embedding(x[:, 1:])), dim=1)
        if return_attn:
            x, a, _ = self.layer2(x, need_attn=True, need_v=True, no_norm=no_norm)
            attn.append(a.clone())
        else:
            x = self.layer2(x)
        cls = C.layer_norm(self.norm, x[:, :1])[:, 0, :]               # only the cls row is read (baseline.py:281)
        if return_attn:
            out = [cls, attn]
            if return_act:
                out.append(v)
            return out
        return cls
