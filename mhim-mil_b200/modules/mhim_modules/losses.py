"""Distillation loss (reference: modules/mhim_modules/losses.py:10-44)."""
import torch
import torch.nn.functional as F
from torch import nn


class SoftTargetCrossEntropy(nn.Module):
    def __init__(self, temp_t=1.0, temp_s=1.0):
        super().__init__()
        self.temp_t, self.temp_s = temp_t, temp_s

    def forward(self, x: torch.Tensor, target: torch.Tensor, mean: bool = True) -> torch.Tensor:
        per_row = -(F.softmax(target / self.temp_t, dim=-1) * F.log_softmax(x / self.temp_s, dim=-1)).sum(dim=-1)
        return per_row.mean() if mean else per_row
