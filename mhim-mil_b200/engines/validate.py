"""Validation without a host round trip per bag.

Reference: engines/base_engine.py:262-316 -- per bag: validate_func, `torch.cat` of the logits, a loss update and
`torch.cuda.synchronize()`; then torchmetrics on the concatenated logits (engines/metrics.py).  Here the bags are enqueued back to
back: logits land in a preallocated device matrix, the loss accumulates on the device, nothing synchronises until the metrics are
read.  Metrics (accuracy, macro precision / recall / F1, binary AUROC with the usual midrank tie handling) are computed on the
device; the ranking for the AUROC uses the library's own radix sort (mil_topk_f32 with k = n).
"""
from typing import Dict, Iterable, Optional, Tuple

import torch

from .. import ops
from .common_mil import CommonMIL


@torch.no_grad()
def collect_logits(args, model, bags: Iterable, n_bags: int, n_classes: int, criterion=None, device=None):
    """Runs `CommonMIL.validate_func` over (bag, label) pairs -> (logits [n_bags, C], labels [n_bags], mean loss (device scalar) or None).
    No synchronisation, no per-bag allocation of the result."""
    eng = CommonMIL(args)
    logits_all = labels_all = loss_sum = None
    i = 0
    for bag, label in bags:
        out, lab = eng.validate_func(args, model, bag, label, criterion, 1, i, None)
        if isinstance(out, (list, tuple)):
            out = out[0]
        if logits_all is None:
            dev = out.device if device is None else device
            logits_all = torch.empty((n_bags, n_classes), dtype=torch.float32, device=dev)
            labels_all = torch.empty(n_bags, dtype=torch.int64, device=dev)
            loss_sum = torch.zeros((), dtype=torch.float32, device=dev)
        logits_all[i].copy_(out.reshape(-1))
        labels_all[i].copy_(lab.reshape(-1)[0])
        if criterion is not None:
            loss_sum += criterion(out.view(1, -1), lab.view(1))
        i += 1
    if i != n_bags:
        raise RuntimeError(f"mhimk validate: expected {n_bags} bags, got {i}")
    return logits_all, labels_all, (loss_sum / n_bags if criterion is not None else None)


def binary_auroc(score: torch.Tensor, target: torch.Tensor) -> torch.Tensor:
    """Area under the ROC curve from ranks (Mann-Whitney U, ties share their mean rank) -- equals sklearn / torchmetrics' trapezoidal
    AUROC.  score [n] float32 CUDA, target [n] in {0, 1}.  Device scalar; nan if one class is absent."""
    n = score.numel()
    order = ops.topk(score, n, largest=False)                          # ascending by value, ties by index (the library's own sort)
    s = score[order]
    rank = torch.arange(1, n + 1, device=score.device, dtype=torch.float64)
    # mean rank within every run of equal scores
    new = torch.ones(n, dtype=torch.bool, device=score.device)
    new[1:] = s[1:] != s[:-1]
    gid = torch.cumsum(new.to(torch.int64), 0) - 1
    gsum = torch.zeros(int(n), dtype=torch.float64, device=score.device).index_add_(0, gid, rank)
    gcnt = torch.zeros(int(n), dtype=torch.float64, device=score.device).index_add_(0, gid, torch.ones_like(rank))
    mid = (gsum / gcnt.clamp_min(1))[gid]
    pos = target[order] == 1
    n_pos = pos.sum().double()
    n_neg = n - n_pos
    u = mid[pos].sum() - n_pos * (n_pos + 1) / 2
    return u / (n_pos * n_neg)


def classification_metrics(logits: torch.Tensor, labels: torch.Tensor) -> Dict[str, torch.Tensor]:
    """acc, macro precision / recall / f1, auc (binary: AUROC of softmax[:, 1]; multi-class: macro one-vs-rest) as device scalars."""
    n, C = logits.shape
    prob = torch.softmax(logits.float(), dim=-1)
    pred = prob.argmax(-1)
    conf = torch.zeros((C, C), dtype=torch.float64, device=logits.device).index_put_((labels, pred), torch.ones(n, dtype=torch.float64, device=logits.device),
                                                                                   accumulate=True)
    tp = conf.diag()
    prec = tp / conf.sum(0).clamp_min(1)
    rec = tp / conf.sum(1).clamp_min(1)
    f1 = 2 * prec * rec / (prec + rec).clamp_min(1e-12)
    if C == 2:
        auc = binary_auroc(prob[:, 1].contiguous(), labels)
    else:
        auc = torch.stack([binary_auroc(prob[:, c].contiguous(), (labels == c).long()) for c in range(C)]).mean()
    return {"acc": tp.sum() / n, "precision": prec.mean(), "recall": rec.mean(), "f1": f1.mean(), "auc": auc}


@torch.no_grad()
def validate(args, model, bags: Iterable, n_bags: int, n_classes: int, criterion=None) -> Tuple[Dict[str, float], torch.Tensor, torch.Tensor]:
    """One validation epoch: ONE device->host transfer at the end.  Returns (metrics as python floats, logits, labels)."""
    was_training = model.training
    model.eval()
    logits, labels, loss = collect_logits(args, model, bags, n_bags, n_classes, criterion)
    m = classification_metrics(logits, labels)
    if loss is not None:
        m["loss"] = loss.double()
    keys = sorted(m)
    vals = torch.stack([m[k].double() for k in keys]).cpu().tolist()       # the only synchronisation of the epoch
    model.train(was_training)
    return dict(zip(keys, vals)), logits, labels
