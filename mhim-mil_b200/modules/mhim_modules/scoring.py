"""Attention -> instance score (reference: modules/mhim_modules/scoring.py:9-58)."""
import torch

from ... import ops


def _cam_max(cam: torch.Tensor) -> torch.Tensor:
    return torch.softmax(cam, dim=1).max(dim=1).values


def get_pseudo_score(classifier, feat, attention):
    """score_n = max_c softmax_c(a_n (h_n . W_c) + b_0); feat [1,n,d], attention [1,n] -> [1,n]."""
    w, b = list(classifier.parameters())[-2], list(classifier.parameters())[-1]
    h, a = feat[0], attention.reshape(-1)
    t = ops.sgemm(h.contiguous(), h.shape[1], 1, w, w.shape[1], 1, h.shape[0], w.shape[0], h.shape[1])
    return _cam_max(t * a[:, None] + b.data[0]).unsqueeze(0)


def get_pseudo_score_trans(classifier, feat, attention, to_out):
    """Per-head v [1,h,n,d] * attn [1,h,n] -> [n, h d] -> to_out -> CAM (scoring.py:9-34)."""
    w, b = list(classifier.parameters())[-2], list(classifier.parameters())[-1]
    v, a = feat[0], attention[0]
    f = (v * a[:, :, None]).permute(1, 0, 2).reshape(v.shape[1], -1)
    f = to_out(f)
    return _cam_max(f @ w.t() + b.data[0]).unsqueeze(0)
