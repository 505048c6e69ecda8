"""ABMIL heads with the reference's interface (modules/abmil.py:51-143 AttentionGated, :145-251 DAttention).

state_dict keys are identical to the reference (`feature.0.*`, `attention.{0,2}.*`, `classifier.*`; gated:
`attention_{a,b}.0.*`, `attention_c.*`, `classifier.0.*`) so its checkpoints load with strict=True.  The nn containers
exist only to own the parameters; forward() calls the CUDA kernels directly:
  * no autograd needed  -> ONE fused tcgen05 pass over the bag (ops.abmil_fused_forward), + classifier
  * training            -> linear_act / softmax_pool primitives with CUDA backward
"""
import torch
import torch.nn.functional as F
from torch import nn

from .. import ops
from ._common import MilModule, act_module, grad_needed, init_linear_layers, lin, require_cuda
from .emb_position import SINCOS


class DAttention(MilModule):
    def __init__(self, input_dim, n_classes, dropout, act, mil_norm=None, mil_bias=True, mil_cls_bias=True, inner_dim=512,
                 embed_feat=True, embed_norm_pos=0, pos=None, **kwargs):
        super().__init__()
        assert pos in ("sincos", "none", None) and embed_norm_pos in (0, 1)                   # abmil.py:159-160
        self.L, self.D, self.K = inner_dim, 128, 1
        self.mil_norm = mil_norm if mil_norm in ("bn", "ln") else None
        self.embed_norm_pos, self.pos = embed_norm_pos, pos
        self.act = act.lower() if act.lower() == "gelu" else "relu"       # anything but gelu falls back to ReLU (abmil.py:183-186)
        self.p_drop = 0.25 if dropout else 0.0                            # hard-coded 0.25 when truthy (abmil.py:188-189)
        if mil_bias:
            mil_cls_bias = True
        self.pos_embed = SINCOS() if pos == "sincos" else nn.Identity()
        layers = []
        # the norm layers own their parameters under the reference's keys (abmil.py:167-178): `norm.*` / `norm1.*`, or `feature.0.*` for
        # the input LayerNorm, which shifts the Linear to `feature.1.*`
        self.norm1 = self.norm = nn.Identity()
        if self.mil_norm == "bn":
            self.norm = nn.BatchNorm1d(input_dim if embed_norm_pos == 0 else inner_dim)
            self.norm1 = nn.BatchNorm1d(self.L * self.K)
        elif self.mil_norm == "ln":
            if embed_norm_pos == 0:
                layers += [nn.LayerNorm(input_dim, bias=mil_bias)]
            else:
                self.norm = nn.LayerNorm(inner_dim, bias=mil_bias)
            self.norm1 = nn.LayerNorm(self.L * self.K, bias=mil_bias)
        self._lin = len(layers)                                            # index of the Linear inside `feature`
        if embed_feat:
            layers += [nn.Linear(input_dim, inner_dim, bias=mil_bias), act_module(self.act)]
            if dropout:
                layers += [nn.Dropout(0.25)]
        self.feature = nn.Sequential(*layers) if layers else nn.Identity()
        self.attention = nn.Sequential(nn.Linear(self.L, self.D, bias=mil_bias), nn.Tanh(), nn.Linear(self.D, self.K, bias=mil_bias))
        self.classifier = nn.Linear(self.L * self.K, n_classes, bias=mil_cls_bias)
        self.embed_feat = embed_feat
        self.precision = ops.DEFAULT_PRECISION
        init_linear_layers(self)

    def _fused_ok(self, x):
        return (self.embed_feat and self.mil_norm is None and self.pos != "sincos" and self.feature[0].bias is not None and self.L == 512
                and x.shape[-1] % 32 == 0 and self.classifier.out_features <= 64)

    def _dropout(self, rows, device):
        if self.training and self.p_drop > 0 and self.embed_feat:
            return ops.next_dropout(self.p_drop, rows, self.L, device)
        return None

    def forward(self, x, return_attn=False, no_norm=False, return_act=False, pos=None, return_img_feat=False, **kwargs):
        require_cuda(x, "DAttention")
        if x.dim() == 2:
            x.unsqueeze_(0)                                                # in place, like abmil.py:204-205
        if x.shape[0] != 1:
            raise RuntimeError("mhimk DAttention: batch must be 1 bag (as everywhere in the reference)")
        x2 = x[0]
        att0, att2 = self.attention[0], self.attention[2]
        if not grad_needed(self, x) and self._fused_ok(x):
            f0 = self.feature[0]
            out = ops.abmil_fused_forward(x2, f0.weight, f0.bias, self.act, att0.weight, att0.bias, att2.weight, att2.bias, "tanh",
                                          want_scores=return_attn, want_h=return_attn and return_act, precision=self.precision,
                                          Wcls=self.classifier.weight, bcls=self.classifier.bias, volatile=self.training,
                                          dropout=self._dropout(x2.shape[0], x2.device))
            pooled, fused_logits = out["pooled"], out["logits"]
            attn = torch.exp(out["s"] - out["stats"][0]) / out["stats"][1] if return_attn else None
            h = out["h"]
        else:
            # composed path (autograd, or a configuration outside the fused kernel: mil_norm bn / ln, pos = sincos)
            if self.mil_norm == "bn" and self.embed_norm_pos == 0:
                x2 = self.norm(x2)                                         # BatchNorm1d over the instances (abmil.py:207-211: the transposes only move the channel axis)
            if self.embed_feat:                                            # dropout inside the GEMM's epilogue (abmil.py:188-189)
                if self._lin == 1:
                    x2 = self.feature[0](x2)                               # input LayerNorm (mil_norm = 'ln', embed_norm_pos = 0)
                f0 = self.feature[self._lin]
                h = ops.linear_act(x2, f0.weight, f0.bias, self.act, volatile=f0.training, dropout=self._dropout(x2.shape[0], x2.device))
            else:
                h = self.feature(x2) if self._lin == 1 else x2
            if self.pos == "sincos":
                h = self.pos_embed(h[None], pos=pos)[0]
            if self.embed_norm_pos == 1 and self.mil_norm is not None:
                h = self.norm(h)
            u = lin(att0, h, "tanh")
            s = lin(att2, u)[:, 0]
            pooled, attn = ops.softmax_pool(s, h)
            fused_logits = None
        img_feat = pooled[None]
        if fused_logits is not None:
            logits = fused_logits
        else:
            logits = lin(self.classifier, self.norm1(img_feat))
        if return_img_feat:
            logits = [logits, img_feat]
        if return_attn:
            res = [logits, attn[None]]
            if return_act:
                res.append(h[None])
            return res
        return logits


class AttentionGated(MilModule):
    def __init__(self, input_dim, n_classes, act="relu", dropout=0.0, mil_norm=None, mil_bias=True, mil_cls_bias=True, inner_dim=512,
                 embed_feat=True, embed_norm_pos=0, pos=None, **kwargs):
        super().__init__()
        if mil_norm not in (None, "none"):
            raise NotImplementedError("mhimk AttentionGated: mil_norm='bn'/'ln' is outside the accelerated path")
        self.L, self.D, self.K = inner_dim, 384, 1                       # D = 384 as hard-coded at abmil.py:55
        self.mil_norm, self.embed_norm_pos, self.pos = None, embed_norm_pos, pos
        self.act = act if act in ("gelu", "relu") else "none"
        self.p_feat = float(dropout)
        self.p_att = 0.25 if dropout else 0.0
        feat = [nn.Linear(input_dim, inner_dim, bias=mil_bias)]
        if act in ("gelu", "relu"):
            feat += [act_module(act)]
        feat += [nn.Dropout(dropout)]
        self.feature = nn.Sequential(*feat)
        a, b = [nn.Linear(self.L, self.D, bias=mil_bias), nn.Tanh()], [nn.Linear(self.L, self.D, bias=mil_bias), nn.Sigmoid()]
        if dropout:
            a += [nn.Dropout(0.25)]
            b += [nn.Dropout(0.25)]
        self.attention_a, self.attention_b = nn.Sequential(*a), nn.Sequential(*b)
        self.attention_c = nn.Linear(self.D, self.K, bias=mil_bias)
        self.classifier = nn.Sequential(nn.Linear(self.L * self.K, n_classes, bias=mil_bias))
        self.norm = self.norm1 = nn.Identity()
        init_linear_layers(self)

    def forward(self, x, **kwargs):
        require_cuda(x, "AttentionGated")
        if x.dim() == 2:
            x.unsqueeze_(0)
        if x.shape[0] != 1:
            raise RuntimeError("mhimk AttentionGated: batch must be 1 bag")
        h = lin(self.feature[0], x[0], self.act)
        if self.training and self.p_feat > 0:
            h = F.dropout(h, self.p_feat, True)
        ga = lin(self.attention_a[0], h, "tanh")
        gb = lin(self.attention_b[0], h, "sigmoid")
        if self.training and self.p_att > 0:
            ga, gb = F.dropout(ga, self.p_att, True), F.dropout(gb, self.p_att, True)
        s = lin(self.attention_c, ga * gb)[:, 0]
        pooled, _ = ops.softmax_pool(s, h)
        return lin(self.classifier[0], pooled[None])
