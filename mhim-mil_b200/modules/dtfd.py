"""DTFD-MIL (double-tier feature distillation) with the reference's interface (modules/dtfd.py:32-272), on the kernels of the path
("pure reuse of a1-a6").

Tier 1: the bag is cut into `group` pseudo-bags; each is pooled by a GATED attention head (tanh x sigmoid, Da = 128; :107-146) and, for the
MaxS / MaxMinS distillations, its instances are ranked by a CAM score (:29-32, 200-207).  Tier 2: the pseudo-bag features go through a second gated
attention + classifier (:95-105).  Kernel mapping:
  dimReduction.fc1 (+act, +Dropout .25)   tensor-core Linear with the dropout in its epilogue            (ops.linear_act)
  attention V / U projections             ONE pass over ALL instances (the per-row logits do not depend on the pseudo-bag; the reference
                                          recomputes them per group, :191), tensor-core Linear x 2 + GEMV-shaped Linear
  per pseudo-bag softmax + pooling        mil_softmax_pool on the group's row slice (train: contiguous chunks, :176-178; test: shuffled ids, :232-235)
  CAM ranking (MaxS / MaxMinS)            GEMV-shaped Linear + mil_col_argmax (the reference sorts all n instances to take the first / last)
  tier 2 and the classifiers              GEMV-shaped Linear kernels (5 rows)
state_dict keys are the reference's (`classifier.fc.*`, `attention.attention_{V,U}.0.*`, `attention.attention_weights.*`, `dimReduction.fc1.weight`,
`UClassifier.{attention,classifier}.*`).  Only numLayer_Res = 0 / swin = False (the DTFD defaults, :149-153) are provided.
"""
import random

import numpy as np
import torch
from torch import nn

from .. import ops
from . import _common as C


class Classifier_1fc(C.MilModule):
    def __init__(self, n_channels, n_classes, droprate=0.0, n_robust=0):
        super().__init__()
        self.fc = nn.Linear(n_channels, n_classes)
        self.droprate = droprate
        if droprate != 0.0:
            self.dropout = nn.Dropout(p=droprate)
        C.init_linear_layers(self)

    def forward(self, x):
        if self.droprate != 0.0:
            x = self.dropout(x)
        return C.lin(self.fc, x)


class DimReduction(C.MilModule):
    def __init__(self, n_channels, m_dim=512, numLayer_Res=0, dropout=False, act="relu", n_robust=0, swin=False, **kwargs):
        super().__init__()
        if numLayer_Res or swin:
            raise NotImplementedError("mhimk DimReduction: residual blocks / the swin encoder are outside the accelerated path (DTFD uses neither)")
        self.fc1 = nn.Linear(n_channels, m_dim, bias=False)
        self.act = "relu" if act.lower() == "relu" else "gelu"
        self.relu1 = C.act_module(self.act)
        self.drop, self.dropout, self.numRes = nn.Dropout(0.25), dropout, 0
        C.init_linear_layers(self)

    def forward(self, x):
        drop = None
        if self.dropout and self.training and self.fc1.out_features % 32 == 0:
            drop = ops.next_dropout(self.drop.p, x.shape[0], self.fc1.out_features, x.device)
        y = ops.linear_act(x, self.fc1.weight, None, self.act, volatile=self.fc1.training, dropout=drop)
        if self.dropout and self.training and drop is None:
            y = self.drop(y)
        return y


class Attention(C.MilModule):
    """Gated attention logits (dtfd.py:107-146): A = w . (tanh(V x) * sigmoid(U x)) + b, [K, N]; softmax over N when isNorm."""

    def __init__(self, L=512, D=128, K=1, n_robust=0):
        super().__init__()
        self.L, self.D, self.K = L, D, K
        self.attention_V = nn.Sequential(nn.Linear(L, D), nn.Tanh())
        self.attention_U = nn.Sequential(nn.Linear(L, D), nn.Sigmoid())
        self.attention_weights = nn.Linear(D, K)
        C.init_linear_layers(self)

    def logits(self, x):
        """x [N, L] -> raw attention logits [N] (K = 1)"""
        gate = C.lin(self.attention_V[0], x, "tanh") * C.lin(self.attention_U[0], x, "sigmoid")
        return C.lin(self.attention_weights, gate)[:, 0]

    def forward(self, x, isNorm=True):
        s = self.logits(x)
        if isNorm:
            _, a = ops.softmax_pool(s, x)
            return a[None]
        return s[None]


class Attention_with_Classifier(C.MilModule):
    def __init__(self, L=512, D=128, K=1, num_cls=2, droprate=0, n_robust=0):
        super().__init__()
        self.attention = Attention(L, D, K)
        self.classifier = Classifier_1fc(L, num_cls, droprate)

    def forward(self, x):
        pooled, _ = ops.softmax_pool(self.attention.logits(x), x)          # AA @ x (dtfd.py:101-103)
        return self.classifier(pooled[None])


class DTFD(C.MilModule):
    def __init__(self, device, lr, weight_decay, steps, input_dim=1024, inner_dim=512, n_classes=2, group=5, distill="AFS", **kwargs):
        super().__init__()
        self.classifier = Classifier_1fc(inner_dim, n_classes, 0.25)
        self.attention = Attention(inner_dim)
        self.dimReduction = DimReduction(input_dim, inner_dim, dropout=0.25)
        self.UClassifier = Attention_with_Classifier(L=inner_dim, num_cls=n_classes, droprate=0.25)
        self.group, self.distill = group, distill
        self.ce_cri = nn.CrossEntropyLoss(reduction="none").to(device)
        trainable = list(self.classifier.parameters()) + list(self.attention.parameters()) + list(self.dimReduction.parameters())
        self.optimizer0 = torch.optim.Adam(trainable, lr=lr, weight_decay=weight_decay)          # as upstream (:159-166); unused by forward
        self.scheduler0 = torch.optim.lr_scheduler.CosineAnnealingLR(self.optimizer0, steps, 0)

    def _tier1(self, mid, s_all, chunks):
        """Per pseudo-bag: softmax over its rows, pooled feature, optional CAM-ranked instances -> the tier-2 input."""
        feats = []
        wcls = self.classifier.fc.weight
        for idx in chunks:
            if isinstance(idx, slice):
                h, s = mid[idx], s_all[idx]
            else:
                h, s = mid.index_select(0, idx), s_all.index_select(0, idx)
            pooled, a = ops.softmax_pool(s.contiguous(), h.contiguous())
            if self.distill == "AFS":
                feats.append(pooled[None])
                continue
            # CAM: softmax_c(a_n h_n . W_c)[-1] (get_cam_1d has no bias, :29-32); the reference sorts all n rows for the first / last one
            cam = ops.linear_act(h.detach().contiguous(), wcls.detach(), None, "none") * a[:, None]
            pos = torch.softmax(cam, dim=1)[:, -1:].contiguous()
            top, _ = ops.col_argmax(pos)
            if self.distill == "MaxS":
                sel = top
            elif self.distill == "MaxMinS":
                low, _ = ops.col_argmax(-pos)
                sel = torch.cat([top, low])
            else:
                raise ValueError(self.distill)
            feats.append(h.index_select(0, sel))
        return torch.cat(feats, dim=0)

    def train_forward(self, x, label):
        n = x.shape[0]
        bounds = np.cumsum([0] + [len(c) for c in np.array_split(np.arange(n), self.group)])    # contiguous chunks (dtfd.py:176-178)
        mid = self.dimReduction(x)
        s_all = self.attention.logits(mid)
        pseudo = self._tier1(mid, s_all, [slice(int(a), int(b)) for a, b in zip(bounds[:-1], bounds[1:]) if b > a])
        return self.UClassifier(pseudo)

    def test_forward(self, x):
        n = x.shape[0]
        mid = self.dimReduction(x)
        s_all = self.attention.logits(mid)
        ids = list(range(n))
        random.shuffle(ids)                                                  # python's RNG, as upstream (:232-233)
        chunks = [torch.as_tensor(c, dtype=torch.int64, device=x.device) for c in np.array_split(np.array(ids), self.group) if len(c)]
        return self.UClassifier(self._tier1(mid, s_all, chunks))

    def forward(self, x, label=None, **kwargs):
        C.require_cuda(x, "DTFD")
        x = x.squeeze(0)
        return self.train_forward(x, label) if self.training else self.test_forward(x)
