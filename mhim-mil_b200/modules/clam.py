"""CLAM_SB / CLAM_MB with the reference's interface (modules/clam.py:93-331; upstream: mahmoodlab/CLAM), on the kernels of the path
(the head is a gated ABMIL -- rows a1/a2 -- plus an instance-level branch on 2 x k_sample selected rows).

Kernel mapping for one bag x [N, D]:
  attention_net.0 (+ReLU/GELU, +Dropout)      tensor-core Linear, dropout in its epilogue                         (ops.linear_act)
  Attn_Net_Gated a / b (+Dropout .25) / c     tensor-core Linear x 2 (tanh, sigmoid), GEMV-shaped Linear to the K attention columns
  softmax over N + M = A h                    mil_softmax_pool per attention column (K = 1 single branch, K = n_classes multi branch)
  instance branch (clam.py:137-167)           mil_topk on the raw attention column (softmax is monotone: same ids as topk(softmax(A)), and no ties
                                              from underflow), 2 x k_sample row gather, GEMV-shaped instance classifier, smooth top-1 SVM loss
  bag classifier(s)                           GEMV-shaped Linear

Differences from upstream that do not change results: the in-/out-of-class choice per instance classifier is a device-side select on the label
(upstream reads `inst_labels[i].item()`, one host sync per class and step, clam.py:192), the hard / smooth split of the SVM loss is a per-row
`where` (upstream branches on `smooth.data.sum()`, svm.py:96-105), and the never-returned `results_dict` (numpy copies of the instance
predictions, :224-229) is not built -- so a training step has no host synchronisation and can be captured in a CUDA graph.
`mil_norm='ln'` is accepted and ignored exactly like upstream (its LayerNorm is dropped by the next assignment, clam.py:102-105).
state_dict keys are the reference's, `instance_loss_fn.labels` included.
"""
import math

import torch
from torch import nn

from .. import ops
from . import _common as C


class SmoothTop1SVM(nn.Module):
    """modules/topk/svm.py:84-108 for the binary instance task (alpha = 1, tau = 1, thresh = 1e3): rows whose top-2 gap is at least
    tau log(thresh) take the max-margin form (functional.py:9-17), the others tau logsumexp((x + delta - x_y) / tau) (:35-42); mean over rows."""

    def __init__(self, n_classes=2, alpha=None, tau=1.0):
        super().__init__()
        self.alpha, self.tau, self.thresh, self.n_classes = (1 if alpha is None else alpha), float(tau), 1e3, n_classes
        self.register_buffer("labels", torch.arange(n_classes))

    def rows(self, x, y):
        """per-row loss [n]"""
        top = x.topk(2, dim=1).values
        hard = (top[:, 0] - top[:, 1]) >= self.tau * math.log(self.thresh)
        z = x + self.alpha * (y[:, None] != self.labels[None, :]).to(x.dtype) - x.gather(1, y[:, None])
        smooth_loss = self.tau * torch.logsumexp(z / self.tau, dim=1)
        return torch.where(hard, z.max(dim=1).values, smooth_loss)

    def forward(self, x, y):
        return self.rows(x, y).sum() / x.shape[0]


class Attn_Net(C.MilModule):
    """clam.py:31-48: Linear -> Tanh (-> Dropout .25) -> Linear to the attention columns."""

    def __init__(self, L=1024, D=256, dropout=False, n_classes=1):
        super().__init__()
        mods = [nn.Linear(L, D), nn.Tanh()]
        if dropout:
            mods.append(nn.Dropout(0.25))
        mods.append(nn.Linear(D, n_classes))
        self.module = nn.Sequential(*mods)
        self.dropout = bool(dropout)

    def forward(self, x):
        a = _lin_drop(self.module[0], x, "tanh", self.module[2] if self.dropout else None, self.training)
        return C.lin(self.module[-1], a), x


class Attn_Net_Gated(C.MilModule):
    """clam.py:58-80: A = c(tanh(a x) * sigmoid(b x))."""

    def __init__(self, L=1024, D=256, dropout=False, n_classes=1, bias=True):
        super().__init__()
        a, b = [nn.Linear(L, D, bias=bias), nn.Tanh()], [nn.Linear(L, D, bias=bias), nn.Sigmoid()]
        if dropout:
            a.append(nn.Dropout(0.25))
            b.append(nn.Dropout(0.25))
        self.attention_a, self.attention_b = nn.Sequential(*a), nn.Sequential(*b)
        self.attention_c = nn.Linear(D, n_classes, bias=bias)
        self.dropout = bool(dropout)

    def forward(self, x):
        a = _lin_drop(self.attention_a[0], x, "tanh", self.attention_a[2] if self.dropout else None, self.training)
        b = _lin_drop(self.attention_b[0], x, "sigmoid", self.attention_b[2] if self.dropout else None, self.training)
        return C.lin(self.attention_c, a * b), x


def _lin_drop(layer, x, act, drop, training):
    """act(layer(x)) followed by `drop` (an nn.Dropout or None): the mask is applied in the Linear's epilogue when the kernel can."""
    if drop is None or not training or drop.p == 0.0:
        return C.lin(layer, x, act)
    if layer.out_features % 32 == 0:
        spec = ops.next_dropout(drop.p, x.shape[0], layer.out_features, x.device)
        return ops.linear_act(x, layer.weight, layer.bias, act, volatile=layer.training, dropout=spec)
    return drop(C.lin(layer, x, act))


class CLAM_SB(C.MilModule):
    def __init__(self, input_dim=1024, gate=True, size_arg="small", dropout=0.0, k_sample=8, n_classes=2, instance_loss_fn=None, subtyping=False,
                 test=False, act="relu", n_robust=0, mil_bias=True, mil_norm=None, inner_dim=512, **kwargs):
        super().__init__()
        self.size_dict = {"small": [input_dim, 512, 256], "big": [input_dim, 512, 384], "hipt": [192, 512, 256]}
        size = self.size_dict[size_arg]
        self.act = "gelu" if act.lower() == "gelu" else "relu"
        fc = [nn.Linear(size[0], inner_dim, bias=mil_bias), C.act_module(self.act)]
        if dropout != 0.0:
            fc.append(nn.Dropout(dropout))
        net = Attn_Net_Gated if gate else Attn_Net
        fc.append(net(L=inner_dim, D=size[2], dropout=dropout, n_classes=1))
        self.attention_net = nn.Sequential(*fc)
        self.classifiers = nn.Linear(inner_dim, n_classes, bias=mil_bias)
        self.instance_classifiers = nn.ModuleList([nn.Linear(inner_dim, 2) for _ in range(n_classes)])
        self.k_sample, self.n_classes, self.subtyping = k_sample, n_classes, subtyping
        self.instance_loss_fn = SmoothTop1SVM(2)
        self.fc_dropout = dropout != 0.0
        C.init_linear_layers(self)

    def relocate(self):
        device = torch.device("cuda" if torch.cuda.is_available() else "cpu")
        self.to(device)

    @staticmethod
    def create_positive_targets(length, device):
        return torch.full((length,), 1, device=device).long()

    @staticmethod
    def create_negative_targets(length, device):
        return torch.full((length,), 0, device=device).long()

    # ---- pieces shared with CLAM_MB ----
    def _embed(self, x):
        """x [N, D] -> (h [N, 512], raw attention logits [N, K])"""
        fc0 = self.attention_net[0]
        h = _lin_drop(fc0, x, self.act, self.attention_net[2] if self.fc_dropout else None, self.training)
        a_raw, _ = self.attention_net[-1](h)
        return h, a_raw

    def _instance_losses(self, a_raw, h, label, multi_branch):
        """Sum over the instance classifiers (clam.py:186-209 / :293-311) for one bag, selected on the device by `label` [] int64: classifier
        i == label sees the k_sample top rows of its attention column (target 1) and the k_sample bottom rows (target 0, :137-154); the others
        contribute only with subtyping: their top rows with target 0 (:157-167).  One top-k pair per attention column (the single-branch
        model ranks ONE column for all classifiers) and ONE Linear launch for all instance classifiers (their weights stacked)."""
        k, nc, dev = self.k_sample, self.n_classes, h.device
        cols = range(nc) if multi_branch else (0,)
        picks = []
        for c in cols:
            col = a_raw[:, c].detach().contiguous()
            picks.append(torch.cat([ops.topk(col, k, largest=True), ops.topk(col, k, largest=False)]))
        inst = h.index_select(0, torch.cat(picks))                                    # [len(cols) * 2k, 512]
        W = torch.cat([c.weight for c in self.instance_classifiers], dim=0)           # [2 nc, 512]
        b = torch.cat([c.bias for c in self.instance_classifiers], dim=0)
        logits_all = ops.linear_act(inst, W, b, "none", volatile=self.training)      # [len(cols) * 2k, 2 nc]
        tgt_in = torch.cat([self.create_positive_targets(k, dev), self.create_negative_targets(k, dev)])
        tgt_out = self.create_negative_targets(k, dev)
        total = h.new_zeros(())
        for i in range(nc):
            r0 = (i if multi_branch else 0) * 2 * k
            lg = logits_all[r0:r0 + 2 * k, 2 * i:2 * i + 2]
            loss_in = self.instance_loss_fn(lg, tgt_in)
            loss_other = self.instance_loss_fn(lg[:k], tgt_out) if self.subtyping else torch.zeros_like(loss_in)
            total = total + torch.where(label == i, loss_in, loss_other)
        return total / nc if self.subtyping else total

    def _bag(self, x, label, instance_eval, multi_branch):
        h, a_raw = self._embed(x)                                                # [N, 512], [N, K]
        total = self._instance_losses(a_raw, h, label, multi_branch) if instance_eval else None
        pooled = torch.stack([ops.softmax_pool(a_raw[:, kcol], h)[0] for kcol in range(a_raw.shape[1])], dim=0)     # M [K, 512]
        return pooled, a_raw, total

    def forward(self, h, label=None, instance_eval=True, return_features=False, attention_only=False, **kwargs):
        instance_eval = instance_eval if label is not None else False
        if h.dim() == 2:
            h = h.unsqueeze(0)
        if type(label) == list:
            instance_eval = False
        bs, ps = h.size(0), h.size(1)
        if attention_only:
            return torch.stack([self._embed(h[b])[1].t() for b in range(bs)], dim=0)             # [bs, 1, N]
        logits, total = [], 0.0
        for b in range(bs):
            lab = label[b].reshape(()).to(h.device) if instance_eval else None
            pooled, _, inst = self._bag(h[b], lab, instance_eval, False)
            logits.append(C.lin(self.classifiers, pooled).max(dim=0).values)                    # [K = 1, C].max over K
            if instance_eval:
                total = total + inst
        bag_logits = torch.stack(logits, dim=0)
        if instance_eval:
            return bag_logits, total, ps
        if type(label) == list:
            return bag_logits, 0, ps
        return bag_logits


class CLAM_MB(CLAM_SB):
    def __init__(self, input_dim=1024, gate=True, size_arg="small", dropout=0.0, k_sample=8, n_classes=2, instance_loss_fn=None, subtyping=False,
                 act="relu", **kwargs):
        nn.Module.__init__(self)
        self.size_dict = {"small": [input_dim, 512, 256], "big": [input_dim, 512, 384]}
        size = self.size_dict[size_arg]
        self.act = "gelu" if act.lower() == "gelu" else "relu"
        fc = [nn.Linear(size[0], size[1]), C.act_module(self.act)]
        if dropout != 0.0:
            fc.append(nn.Dropout(dropout))
        net = Attn_Net_Gated if gate else Attn_Net
        fc.append(net(L=size[1], D=size[2], dropout=dropout, n_classes=n_classes))
        self.attention_net = nn.Sequential(*fc)
        self.classifiers = nn.ModuleList([nn.Linear(size[1], 1) for _ in range(n_classes)])       # one bag classifier per class
        self.instance_classifiers = nn.ModuleList([nn.Linear(size[1], 2) for _ in range(n_classes)])
        self.k_sample, self.n_classes, self.subtyping = k_sample, n_classes, subtyping
        self.instance_loss_fn = SmoothTop1SVM(2)
        self.fc_dropout = dropout != 0.0
        C.init_linear_layers(self)

    def forward(self, h, label=None, instance_eval=True, return_features=False, attention_only=False, **kwargs):
        instance_eval = instance_eval if label is not None else False
        if type(label) == list:
            instance_eval = False
        ps = h.size(1) if h.dim() == 3 else h.size(0)
        x = h.reshape(-1, h.shape[-1])                                          # h.squeeze() of a [1, N, D] bag
        if attention_only:
            return self._embed(x)[1].t()                                        # [K, N]
        lab = label.reshape(-1)[0].to(x.device) if instance_eval else None
        pooled, _, total = self._bag(x, lab, instance_eval, True)              # M [C, 512]
        W = torch.cat([c.weight for c in self.classifiers], dim=0)              # [C, 512]: logits[0, c] = classifiers[c](M[c])
        bias = torch.cat([c.bias for c in self.classifiers], dim=0)
        logits = ((pooled * W).sum(dim=1) + bias)[None]
        if instance_eval:
            return logits, total, ps
        if type(label) == list:
            return logits, 0, ps
        return logits
