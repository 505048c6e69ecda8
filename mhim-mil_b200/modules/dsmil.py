"""DSMIL with the reference's interface (modules/dsmil.py:59-172): MILNet = feature -> instance classifier ->
critical-instance attention (BClassifier) -> Conv1d bag head."""
import math

import torch
import torch.nn.functional as F
from torch import nn

from .. import ops
from . import _common as C


class BClassifier(C.MilModule):
    def __init__(self, input_size, output_class, dropout_v=0.0, nonlinear=True, passing_v=True, bias=True, norm=False):
        super().__init__()
        self.q = (nn.Sequential(nn.Linear(input_size, 128, bias=bias), nn.ReLU(), nn.Linear(128, 128), nn.Tanh()) if nonlinear
                  else nn.Linear(input_size, 128, bias=bias))
        self.v = nn.Sequential(nn.Dropout(dropout_v), nn.Linear(input_size, input_size, bias=bias), nn.ReLU()) if passing_v else nn.Identity()
        if norm in ("bn", "ln"):
            raise NotImplementedError("mhimk dsmil.BClassifier: mil_norm is outside the accelerated path")
        self.mil_norm, self.norm = norm, nn.Identity()
        self.fcc = nn.Conv1d(output_class, output_class, kernel_size=input_size, bias=bias)
        self.nonlinear, self.passing_v = nonlinear, passing_v

    def _q(self, t):
        if not self.nonlinear:
            return C.lin(self.q, t)
        return C.lin(self.q[2], C.lin(self.q[0], t, "relu"), "tanh")

    def forward(self, feats, c):
        """feats [1,N,K], c [1,N,C] -> (pred [1,C], A [1,N,C], B [1,C,K])  (dsmil.py:85-109)"""
        f, cl = feats[0], c[0]
        V = C.lin(self.v[1], self.v[0](f), "relu") if self.passing_v else f
        Q = self._q(f)
        crit, _ = ops.col_argmax(cl)                                   # the reference sorts all N rows to read row 0 (dsmil.py:91-92)
        q_max = self._q(f.index_select(0, crit))
        logit = ops.linear_act(Q, q_max, None, "none") / math.sqrt(Q.shape[-1])
        Bs, As = [], []
        for j in range(logit.shape[1]):
            pooled, a = ops.softmax_pool(logit[:, j], V)
            Bs.append(pooled)
            As.append(a)
        B = torch.stack(Bs)[None]
        pred = C.conv1d_full(self.fcc, B)                              # Conv1d(C, C, kernel = K) on [1, C, K] == one Linear over C * K inputs
        return pred, torch.stack(As, dim=1)[None], B


class MILNet(C.MilModule):
    def __init__(self, n_classes, dropout, act, input_dim=1024, mil_norm=None, mil_bias=True, inner_dim=512, **kwargs):
        super().__init__()
        if mil_norm not in (None, "none"):
            raise NotImplementedError("mhimk MILNet: mil_norm='bn'/'ln' is outside the accelerated path")
        self.mil_norm = None
        self.act = act.lower() if act.lower() in ("relu", "gelu") else "none"
        feat = [nn.Linear(input_dim, inner_dim, bias=mil_bias)]
        if self.act != "none":
            feat += [C.act_module(self.act)]
        self.feature = nn.Sequential(*feat)
        self.dp = nn.Dropout(dropout) if dropout > 0.0 else nn.Identity()
        self.norm1 = self.norm = nn.Identity()
        self.i_classifier = nn.Linear(inner_dim, n_classes, bias=mil_bias)
        self.b_classifier = BClassifier(inner_dim, n_classes, bias=mil_bias, norm=mil_norm)
        C.init_linear_layers(self)

    def forward(self, x, label=None, loss=None, pos=None, **kwargs):
        C.require_cuda(x, "MILNet")
        ps, bs = x.size(1), x.size(0)
        if bs != 1:
            raise RuntimeError("mhimk MILNet: batch must be 1 bag")
        feats = self.dp(C.lin(self.feature[0], x[0], self.act))[None]
        classes = C.lin(self.i_classifier, feats[0])[None]
        pred, A, B = self.b_classifier(feats, classes)
        crit, _ = ops.col_argmax(classes[0])                           # max-pooling of the instance logits (dsmil.py:160) as a gather of the
        inst = classes[0][crit, torch.arange(classes.shape[-1], device=crit.device)][None]    # critical rows (differentiable)
        if self.training:
            if isinstance(loss, nn.CrossEntropyLoss):
                max_loss = loss(inst.view(bs, -1), label)
            elif isinstance(loss, nn.BCEWithLogitsLoss):
                max_loss = loss(inst.view(bs, -1), label.view(bs, -1).float())
            else:
                max_loss = loss(logits=inst.view(bs, -1), Y=label[0], c=label[1])
            return pred, max_loss, ps
        return pred, inst
