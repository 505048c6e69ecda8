"""Adam / AdamW with the whole-model step in ONE kernel launch.

Reference: train_utils.py:55-65 builds `torch.optim.Adam(params)` (default `--opt adam`, options.py:56) or `AdamW(params)` over one
param group {params, lr, weight_decay}; engines/base_engine.py:110-120 calls `optimizer.step()` after every bag.  At MIL sizes
(1.6-3.7 M parameters in 13-32 tensors) that step is launch-bound.  `FusedAdam` is a torch.optim.Optimizer with the same
constructor arguments, hyper-parameter defaults, param-group semantics (schedulers work unchanged) and state_dict layout
({"step", "exp_avg", "exp_avg_sq"} per parameter) as torch's, so checkpoints interchange; `step()` runs mil_adam_step_f32 over a
cached device table of (param, grad, exp_avg, exp_avg_sq) segments -- one table and one launch per param group.
"""
import ctypes
import math

import torch

from .. import _lib, ops

SEG = 32768                      # elements per segment = per CTA


class FusedAdam(torch.optim.Optimizer):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0, adamw=False, capturable=False):
        if lr < 0.0 or eps < 0.0 or not 0.0 <= betas[0] < 1.0 or not 0.0 <= betas[1] < 1.0 or weight_decay < 0.0:
            raise ValueError("FusedAdam: invalid hyper-parameters")
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay, adamw=adamw, capturable=capturable))
        self._tables = {}

    @staticmethod
    def adamw(params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2, **kw):
        """torch.optim.AdamW's defaults (decoupled weight decay 1e-2)."""
        return FusedAdam(params, lr, betas, eps, weight_decay, adamw=True, **kw)

    def _table(self, gi, plist, state):
        key = tuple((p.data_ptr(), p.grad.data_ptr(), p.numel()) for p in plist)
        hit = self._tables.get(gi)
        if hit is not None and hit[0] == key:
            return hit[1]
        rows = []
        for p in plist:
            st = state[p]
            ptrs = (p.data_ptr(), p.grad.data_ptr(), st["exp_avg"].data_ptr(), st["exp_avg_sq"].data_ptr())
            for o in range(0, p.numel(), SEG):
                rows.append(tuple(a + 4 * o for a in ptrs) + (min(SEG, p.numel() - o),))
        table = torch.tensor(rows, dtype=torch.int64).reshape(-1, 5).to(plist[0].device)
        self._tables[gi] = (key, table)
        return table

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        L = _lib.lib()
        for gi, group in enumerate(self.param_groups):
            plist = [p for p in group["params"] if p.grad is not None]
            if not plist:
                continue
            for p in plist:
                if not p.is_cuda or p.dtype != torch.float32 or p.grad.dtype != torch.float32 or p.grad.is_sparse:
                    raise RuntimeError("mhimk FusedAdam: parameters and gradients must be dense float32 CUDA tensors -- no CPU path")
                if not (p.is_contiguous() and p.grad.is_contiguous()):
                    raise RuntimeError("mhimk FusedAdam: parameters and gradients must be contiguous")
                st = self.state[p]
                if len(st) == 0:
                    st["step"] = torch.zeros((), dtype=torch.float32, device=p.device if group["capturable"] else "cpu")
                    st["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                    st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
            b1, b2 = group["betas"]
            steps = {id(self.state[p]["step"]): self.state[p]["step"] for p in plist}
            for s in steps.values():
                s += 1
            table = self._table(gi, plist, self.state)
            first = self.state[plist[0]]["step"]
            if group["capturable"]:
                # every parameter of the group has stepped the same number of times: one device-resident counter feeds the kernel
                bc1 = bc2s = 1.0
                step_dev = _lib.ptr(first)
            else:
                t = float(first)
                bc1, bc2s, step_dev = 1.0 - b1 ** t, math.sqrt(1.0 - b2 ** t), None
            _lib.check(L.mil_adam_step_f32(_lib.ptr(table), table.shape[0], _lib.c_float(group["lr"]), ctypes.c_double(b1), ctypes.c_double(b2),
                                           _lib.c_float(group["eps"]), _lib.c_float(group["weight_decay"]), 1 if group["adamw"] else 0,
                                           _lib.c_float(bc1), _lib.c_float(bc2s), step_dev, _lib.stream_ptr()), "mil_adam_step_f32")
        ops.weights_touched()                              # the fused pass's cached weight images follow the update
        return loss
