"""EMA teacher update in ONE kernel launch.

Reference: engines/base_engine.py:155-167 (classification) and :478-489 (survival) --

    for param_q, param_k in zip(model.parameters(), model_ema.parameters()):
        param_k.data.mul_(mm).add_(param_q.data, alpha=1. - mm)

i.e. two launches per parameter (~60-120 per training step) through `.data`, which autograd's version counter does not see.
`ema_update` does the same arithmetic (same two roundings per element) in one launch of mil_ema_update_f32 over a cached device
table of parameter segments, and tells the weight-image caches about it (ops.weights_touched).
"""
import torch

from .. import _lib, ops

SEG = 32768                      # elements per segment = per CTA (128 KB read+write each way)
_TABLES = {}                     # (ptrs, sizes) -> device int64 [n_seg, 3] = mil_ema_seg_t records


def _segments(pairs):
    rows = []
    for k, q in pairs:
        n, pk, pq = k.numel(), k.data_ptr(), q.data_ptr()
        for o in range(0, n, SEG):
            rows.append((pk + 4 * o, pq + 4 * o, min(SEG, n - o)))
    return rows


def ema_update(model: torch.nn.Module, model_ema: torch.nn.Module, mm: float) -> None:
    """model_ema <- mm * model_ema + (1 - mm) * model, parameter by parameter, in place (buffers are not touched, as upstream)."""
    assert 0.0 <= mm <= 1.0, "Momentum needs to be between 0.0 and 1.0, got %.5f" % mm          # base_engine.py:164
    pairs = list(zip(model_ema.parameters(), model.parameters()))       # (param_k, param_q), the reference's zip order
    if not pairs:
        return
    # the segment table depends only on addresses, sizes and dtypes: validate and build it once per such layout
    key = tuple((k.data_ptr(), q.data_ptr(), k.numel(), k.dtype, q.dtype) for k, q in pairs)
    table = _TABLES.get(key)
    if table is None:
        dev = pairs[0][0].device
        for k, q in pairs:
            if not (k.is_cuda and q.is_cuda) or k.device != dev or q.device != dev:
                raise RuntimeError("mhimk ema_update: every parameter of both models must live on the same CUDA device -- no CPU path")
            if k.dtype != torch.float32 or q.dtype != torch.float32 or k.shape != q.shape:
                raise RuntimeError("mhimk ema_update: parameters must be float32 and pairwise of equal shape")
            if not (k.is_contiguous() and q.is_contiguous()):
                raise RuntimeError("mhimk ema_update: parameters must be contiguous")
        if len(_TABLES) > 16:
            _TABLES.clear()
        table = torch.tensor(_segments(pairs), dtype=torch.int64).reshape(-1, 3).to(dev)
        _TABLES[key] = table
    L = _lib.lib()
    _lib.check(L.mil_ema_update_f32(_lib.ptr(table), table.shape[0], _lib.c_float(mm), _lib.c_float(1.0 - mm), _lib.stream_ptr()),
               "mil_ema_update_f32")
    ops.weights_touched()
