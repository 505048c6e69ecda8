"""Helpers shared by the drop-in modules."""
import torch
from torch import nn

from .. import ops


def init_linear_layers(module: nn.Module) -> None:
    """Reference init policy (abmil.py:8-21, mhim_modules/utils.py:8-22): Xavier-normal Linear weights, zero biases,
    LayerNorm weight 1 / bias 0."""
    for m in module.modules():
        if isinstance(m, nn.Linear):
            nn.init.xavier_normal_(m.weight)
            if m.bias is not None:
                nn.init.zeros_(m.bias)
        elif isinstance(m, nn.LayerNorm):
            if m.bias is not None:
                nn.init.zeros_(m.bias)
            nn.init.ones_(m.weight)


def act_module(name: str) -> nn.Module:
    name = name.lower()
    return {"relu": nn.ReLU, "gelu": nn.GELU, "tanh": nn.Tanh}[name]()


def grad_needed(module: nn.Module, *tensors) -> bool:
    if not torch.is_grad_enabled():
        return False
    if any(t is not None and t.requires_grad for t in tensors):
        return True
    return any(p.requires_grad for p in module.parameters())


def lin(layer: nn.Linear, x: torch.Tensor, act: str = "none") -> torch.Tensor:
    """act(layer(x)) through the CUDA GEMM of libmhimk (x is [M, in]).  In train mode the cached weight image is not trusted
    (`volatile`): the reference's EMA teacher update writes through `.data` (engines/base_engine.py:166-167), which autograd's
    version counter does not see."""
    return ops.linear_act(x, layer.weight, layer.bias, act, volatile=layer.training)


def layer_norm(norm: nn.LayerNorm, x: torch.Tensor) -> torch.Tensor:
    """norm(x) for x [1, n, dim]: the library's own kernel when no autograd is needed, torch's differentiable one otherwise."""
    if (x.is_cuda and x.dtype == torch.float32 and x.dim() == 3 and x.shape[0] == 1 and norm.weight is not None
            and not (torch.is_grad_enabled() and (x.requires_grad or norm.weight.requires_grad))):
        return ops.layernorm(x[0], norm.weight, norm.bias, norm.eps)[None]
    return norm(x)


def conv1d_full(conv: nn.Conv1d, B: torch.Tensor) -> torch.Tensor:
    """DSMIL's bag head `Conv1d(C, C, kernel_size=K)` applied to B [1, C, K] (dsmil.py:98-99, baseline.py:149-150): the kernel spans the
    whole length, so it is the Linear pred[o] = sum_{c,k} w[o,c,k] B[c,k] + b[o] over the flattened C*K inputs -> [1, C].  Runs in the
    library's own (GEMV-shaped) kernel with its CUDA backward instead of cuDNN."""
    Cc, Kk = B.shape[1], B.shape[2]
    W = conv.weight.reshape(conv.out_channels, Cc * Kk)
    return ops.linear_act(B.reshape(1, Cc * Kk), W, conv.bias, "none", volatile=conv.training)


class MilModule(nn.Module):
    """Base of the drop-in modules.  Behaves exactly like nn.Module (no parameters, no state_dict keys, no hooks); it only tells
    the weight-image caches of `ops` that a train()/eval() switch happened, so the first forward after the switch rebuilds its
    images: the last `.data` update of a training epoch happens after the last train-mode forward."""

    def train(self, mode: bool = True):
        ops.weights_touched()
        return super().train(mode)


def require_cuda(x: torch.Tensor, who: str) -> None:
    if not x.is_cuda:
        raise RuntimeError(f"{who}: input is on {x.device}; mhimk modules run only on a CUDA (sm_100) device -- no CPU fallback")
