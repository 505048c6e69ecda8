"""Tensor-level ops over the C ABI (include/mhimk.h).  Every op runs a hand-written CUDA kernel from libmhimk.so on
the current CUDA stream; there is no CPU / eager-PyTorch fallback (CPU tensors raise).

Differentiable primitives (torch.autograd.Function with CUDA backward):
  linear_act      y = act(x W^T + b)                    (every nn.Linear(+activation) on the path)
  softmax_pool    p = softmax_L(s) @ h                  (abmil.py:231-234, baseline.py:33-36, dsmil.py:94-96)
No-grad fused paths:
  abmil_fused_forward   one streaming tcgen05 pass X -> (scores, pooled)  (teacher / inference)
  topk, mask_from_indices, cam_score                    (masked hard-instance selection, scoring.py:37-58)
"""
from ctypes import c_float, c_int, c_int64, c_size_t, c_void_p
from typing import Optional, Tuple

import torch

from . import _lib
from ._lib import ACT, PREC, check, ptr, stream_ptr

# Arithmetic of the tensor-core contractions.  "bf16x3" (bf16 hi + lo, 3 products; fp32's exponent range) is the parity arithmetic.
# "fp16x3" (fp16 hi + lo, operands must satisfy |x| <= 65504) has a 32x smaller operand roundoff at the same cost, but measured on a
# B200 it buys little where it would matter: the fp32 accumulation inside tcgen05.mma is not round-to-nearest and its error grows with
# the contraction length -- 3.5e-7 at K = 32, 7.7e-7 at K = 128, 2.3e-6 at K = 512, 4.2e-6 at K = 1024 (fp16x3) against 6.0e-6 / 4.6e-6 /
# 4.8e-6 / 6.9e-6 (bf16x3), tests/test_gpu_umma.py -- so at the path's K = 512 .. 1536 both sit on the accumulator's floor.
DEFAULT_PRECISION = "bf16x3"
GRAD_PRECISION = "bf16x3"


# --------------------------------------------------------------------------------------------------------------
# dropout inside the kernels (mil_dropout_t)
# --------------------------------------------------------------------------------------------------------------
class DropSpec:
    """Host description of a dropout fused into a kernel: probability `p` and one of
      * the in-kernel Philox stream (seed, offset) given as host integers (mode 2),
      * caller-supplied keep bits (int32 [rows, ncols/32], bit i of word (r, c) = keep flag of column 32c+i; mode 1),
      * `seed_dev`: an int64 CUDA tensor [2] = (seed, offset) the kernel reads at run time (mode 3; CUDA-graph replays)."""
    __slots__ = ("p", "seed", "offset", "keep_bits", "seed_dev", "_c")

    def __init__(self, p, seed=0, offset=0, keep_bits=None, seed_dev=None):
        if not 0.0 <= p < 1.0:
            raise ValueError(f"mhimk: dropout p={p} must be in [0, 1)")
        if keep_bits is not None:
            keep_bits = _need(keep_bits, "keep_bits", torch.int32)
        if seed_dev is not None:
            seed_dev = _need(seed_dev, "seed_dev", torch.int64)
        self.p, self.seed, self.offset = float(p), int(seed) & (2 ** 64 - 1), int(offset) & (2 ** 64 - 1)
        self.keep_bits, self.seed_dev = keep_bits, seed_dev
        mode, pointer = (1, keep_bits.data_ptr()) if keep_bits is not None else ((3, seed_dev.data_ptr()) if seed_dev is not None else (2, None))
        self._c = _lib.DropoutT(mode, self.p, self.seed, self.offset, pointer)

    def c(self):
        import ctypes
        return ctypes.byref(self._c)


DROPOUT_HOOK = None          # tests: callable (rows, ncols, p, device) -> DropSpec, e.g. keep bits packed from the reference's own mask


def next_dropout(p, rows=None, ncols=None, device=None):
    """A fresh DropSpec drawn from torch's CUDA generator of the device, the way torch's own CUDA dropout does it: seed = the
    generator's seed, offset = its Philox offset, which this call advances -- so torch.manual_seed() makes runs reproducible and
    successive calls get independent streams.  DROPOUT_HOOK, when set, supplies the spec instead (mask-in parity tests)."""
    if p <= 0.0:
        return None
    if DROPOUT_HOOK is not None:
        return DROPOUT_HOOK(rows, ncols, p, device)
    if torch.cuda.is_current_stream_capturing():
        # inside a CUDA-graph capture host integers would be frozen into the graph: draw (seed, offset) on the DEVICE with torch's
        # graph-safe generator (the drawing kernel is part of the graph, so every replay gets a fresh pair) -- mil_dropout_t mode 3
        return DropSpec(p, seed_dev=torch.randint(0, 2 ** 62, (2,), dtype=torch.int64, device=device if device is not None else "cuda"))
    idx = device.index if isinstance(device, torch.device) and device.index is not None else torch.cuda.current_device()
    gen = torch.cuda.default_generators[idx]
    off = gen.get_offset()
    gen.set_offset(off + 4)                            # torch requires multiples of 4 (one Philox4x32 call)
    return DropSpec(p, gen.initial_seed(), off // 4)


def pack_keep_bits(keep: torch.Tensor) -> torch.Tensor:
    """bool/0-1 mask [rows, ncols] -> int32 keep words [rows, ncols/32] (torch ops on the mask's device; test / integration helper)."""
    rows, ncols = keep.shape
    w = (keep.reshape(rows, ncols // 32, 32) != 0).to(torch.int64) << torch.arange(32, device=keep.device, dtype=torch.int64)
    w = w.sum(-1)
    return torch.where(w >= 2 ** 31, w - 2 ** 32, w).to(torch.int32).contiguous()


def dropout_bits(rows: int, ncols: int, spec: DropSpec, device) -> torch.Tensor:
    """The keep words the in-kernel Philox stream of `spec` yields for a [rows, ncols] tensor (mil_dropout_bits)."""
    out = torch.empty((rows, ncols // 32), dtype=torch.int32, device=device)
    check(_lib.lib().mil_dropout_bits(rows, ncols, spec.c(), ptr(out), stream_ptr()), "mil_dropout_bits")
    return out


def apply_dropout(y: torch.Tensor, spec: DropSpec) -> torch.Tensor:
    """y * keep / (1 - p) for a contiguous [rows, ncols] tensor (no autograd; the primitive under linear_act's non-tensor-core path)."""
    out = torch.empty_like(y)
    check(_lib.lib().mil_act_bwd_drop_f32(ptr(y), ptr(y), y.shape[0], y.shape[1], ACT["none"], spec.c(), ptr(out), stream_ptr()), "mil_act_bwd_drop_f32")
    return out


def _need(t: torch.Tensor, name: str, dtype=torch.float32) -> torch.Tensor:
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise RuntimeError(f"mhimk: `{name}` must be a CUDA tensor -- there is no CPU path (got {getattr(t, 'device', type(t))})")
    if t.dtype != dtype:
        if dtype == torch.float32 and t.dtype in (torch.float16, torch.bfloat16):
            # `--amp` (engines/base_engine.py:78: torch.autocast(float16)) hands the kernels half-precision activations produced by
            # autocast'ed torch ops: compute in fp32 like torch.amp.custom_fwd(cast_inputs=torch.float32) would
            t = t.float()
        else:
            raise RuntimeError(f"mhimk: `{name}` must be {dtype} (got {t.dtype})")
    return t if t.is_contiguous() else t.contiguous()


def _ws(nbytes: int, device) -> torch.Tensor:
    return torch.empty(max(int(nbytes), 16), dtype=torch.uint8, device=device)


# --------------------------------------------------------------------------------------------------------------
# raw GEMM
# --------------------------------------------------------------------------------------------------------------
def sgemm(A, sAm, sAk, B, sBn, sBk, M, N, K, bias=None, act="none", pre_out=None, row_ids=None, splitk=1, out=None):
    """C[M,N] = act(sum_k A(m,k) B(n,k) + bias) on fp32 CUDA cores (mil_sgemm_f32)."""
    L = _lib.lib()
    C = out if out is not None else torch.empty((M, N), dtype=torch.float32, device=A.device)
    ws = _ws(splitk * M * N * 4 if splitk > 1 else 16, A.device)
    check(L.mil_sgemm_f32(ptr(A), sAm, sAk, ptr(row_ids), ptr(B), sBn, sBk, ptr(bias), ptr(C), N, ptr(pre_out), M, N, K, ACT[act], splitk,
                          ptr(ws), ws.numel(), stream_ptr()), "mil_sgemm_f32")
    return C


def _splitk_for(M, N, K):
    tiles = ((M + 127) // 128) * ((N + 127) // 128)
    want = max(1, (2 * 148 + tiles - 1) // tiles)
    return int(max(1, min(want, (K + 511) // 512, 64)))


# weight images of the tensor-core Linear, cached like the fused pass's (same validity rule)
_TC_CACHE = {}
TC_MIN_ROWS = 256            # below this the exact-fp32 CUDA-core GEMM is as fast and has no image to build


def _tc_supported(M, N, K, W):
    return (M >= TC_MIN_ROWS and K % 32 == 0 and K >= 32 and N % 64 == 0 and N >= 64 and W.is_contiguous() and W.data_ptr() % 16 == 0
            and _lib.lib().mil_device_supported())


SKINNY = True                # tests can switch the GEMV-shaped kernels off (the 128 x 128-tile fp32 GEMM then takes those shapes)


def _skinny(M, N, K):
    return SKINNY and (N <= 8 or M <= 8) and _lib.lib().mil_skinny_supported(M, N, K) != 0


def _drop_dead(cache, limit):
    """Forget the entries whose owner tensor is gone (their workspaces would otherwise pin HBM until the cache is cleared)."""
    dead = [k for k, v in cache.items() if v[0]() is None]
    for k in dead:
        del cache[k]
    if len(cache) > limit:
        cache.clear()


# Validity of the cached 16-bit weight images.  An image is reused only while (a) the weight is the very same live tensor at the
# same address and shape, (b) its autograd version counter is unchanged (optimizer.step / load_state_dict / copy_ bump it), (c) no
# one has called weights_touched() since, and (d) the caller does not say `volatile`.  (d) exists because writes through `.data` --
# the reference's EMA teacher update, engines/base_engine.py:166-167 -- do NOT bump the version counter: the drop-in modules pass
# volatile=self.training, so in train mode every forward rebuilds its images (a few microseconds into the same buffer), and
# MilModule.train()/eval() call weights_touched() so that the first eval-mode forward after training rebuilds once.
_EPOCH = 0


def weights_touched() -> None:
    """Declare that weights may have been modified behind autograd's back (e.g. through `.data` while in eval mode):
    every cached weight image is rebuilt at its next use."""
    global _EPOCH
    _EPOCH += 1


# W^T copies for the tensor-core dX (the NT kernel wants [K, N] rows).  One persistent buffer per live weight: refreshed in place
# (copy_ bumps the buffer's version, which in turn refreshes the image _tc_block keys on it).
_WT_CACHE = {}


def _transposed(W, volatile=False):
    import weakref
    ver = (W._version, _EPOCH)
    hit = _WT_CACHE.get(id(W))
    if hit is not None and hit[0]() is W and hit[2].shape == (W.shape[1], W.shape[0]) and hit[2].device == W.device:
        if volatile or hit[1] != ver:
            hit[2].copy_(W.detach().t())
            _WT_CACHE[id(W)] = (hit[0], ver, hit[2])
        return hit[2]
    Wt = W.detach().t().contiguous()
    _drop_dead(_WT_CACHE, 256)
    _WT_CACHE[id(W)] = (weakref.ref(W), ver, Wt)
    return Wt


def _tc_block(x, Wb, bb, act, pre_b, y_b, key_obj, blk, precision, volatile=False, dropout=None):
    """One column block (<= 512 wide, contiguous rows of the weight) through mil_linear_act_tc_f32."""
    import weakref
    L = _lib.lib()
    M, K = x.shape
    N = Wb.shape[0]
    key = (id(key_obj), blk, N, precision)
    ver = (key_obj._version, _EPOCH)
    where = (key_obj.data_ptr(), tuple(key_obj.shape))
    hit = _TC_CACHE.get(key)
    if hit is not None and hit[0]() is key_obj and hit[3] == where:
        ws = hit[2]                                   # same live weight: keep the buffer, rebuild the image in place if stale
        ready = 0 if (volatile or hit[1] != ver) else 1
    else:
        ws, ready = _ws(L.mil_linear_tc_workspace_bytes(N, K), x.device), 0
        _drop_dead(_TC_CACHE, 256)
    if not ready:                                     # not valid until the call below has rebuilt the image (a failed
        _TC_CACHE[key] = (weakref.ref(key_obj), None, ws, where)   # call must not leave a "ready" entry behind)
    # x / y may be column blocks of wider row-major buffers: pass their row strides as leading dimensions
    check(L.mil_linear_act_tc_ld_f32(c_void_p(x.data_ptr()), x.stride(0), M, K, ptr(Wb), ptr(bb), N, ACT[act], ptr(pre_b), c_void_p(y_b.data_ptr()),
                                     y_b.stride(0), dropout.c() if dropout else None, ptr(ws), ws.numel(), ready, PREC[precision], stream_ptr()),
          "mil_linear_act_tc_f32")
    if not ready:
        _TC_CACHE[key] = (weakref.ref(key_obj), ver, ws, where)


def linear_forward(x, W, b, act, pre=None, precision=None, volatile=False, dropout=None):
    """y = drop(act(x W^T + b)): tcgen05 path when the shape allows (column blocks of <= 512), else the fp32 CUDA-core GEMM."""
    M, K = x.shape
    N = W.shape[0]
    precision = precision or DEFAULT_PRECISION
    if not _tc_supported(M, N, K, W):
        if _skinny(M, N, K) and W.is_contiguous():                  # GEMV-shaped: streaming kernel instead of 128 x 128 GEMM tiles
            y = torch.empty((M, N), dtype=torch.float32, device=x.device)
            check(_lib.lib().mil_skinny_fwd_f32(ptr(x), K, M, K, ptr(W), ptr(b), N, ACT[act], ptr(pre), ptr(y), stream_ptr()), "mil_skinny_fwd_f32")
        else:
            y = sgemm(x, K, 1, W, K, 1, M, N, K, bias=b, act=act, pre_out=pre)
        return apply_dropout(y, dropout) if dropout else y
    widths, left = [], N
    while left > 0:                                   # 512-wide blocks, then one of 256 / 192 / 128 / 64
        wdt = 512 if left >= 512 else (256 if left >= 256 else left)
        widths.append(wdt)
        left -= wdt
    if len(widths) == 1:
        y = torch.empty((M, N), dtype=torch.float32, device=x.device)
        _tc_block(x, W, b, act, pre, y, W, 0, precision, volatile, dropout)
        return y
    if dropout:
        raise RuntimeError("mhimk: fused dropout on a Linear wider than 512 columns is not provided")
    # wider than 512 columns: column blocks of ONE output buffer (the kernel takes a leading dimension), no concatenation afterwards
    y = torch.empty((M, N), dtype=torch.float32, device=x.device)
    o = 0
    for i, wdt in enumerate(widths):
        _tc_block(x, W[o:o + wdt], None if b is None else b[o:o + wdt], act, None if pre is None else pre[:, o:o + wdt], y[:, o:o + wdt], W, i, precision,
                  volatile)
        o += wdt
    return y


class _LinearAct(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, W, b, act, volatile=False, dropout=None):
        x, W = _need(x, "x"), _need(W, "weight")
        b = _need(b, "bias") if b is not None else None
        M, K = x.shape
        N = W.shape[0]
        grad = x.requires_grad or W.requires_grad
        # saved for the backward: relu -> the output (dropped or not); gelu -> the pre-activation; through a dropout tanh / sigmoid
        # also keep the pre-activation (their derivative needs the undropped value)
        need_pre = grad and (act == "gelu" or (dropout is not None and act in ("tanh", "sigmoid")))
        pre = torch.empty((M, N), dtype=torch.float32, device=x.device) if need_pre else None
        y = linear_forward(x, W, b, act, pre, volatile=volatile, dropout=dropout)
        ctx.act = act
        ctx.volatile = volatile
        ctx.dropout = dropout
        ctx.has_bias = b is not None
        ctx.save_for_backward(x, W, pre if need_pre else y)
        return y

    @staticmethod
    def backward(ctx, g_y):
        x, W, saved = ctx.saved_tensors
        L = _lib.lib()
        g_y = _need(g_y, "grad")
        M, K = x.shape
        N = W.shape[0]
        if ctx.dropout is not None:
            g_pre = torch.empty_like(g_y)
            check(L.mil_act_bwd_drop_f32(ptr(g_y), ptr(saved), M, N, ACT[ctx.act], ctx.dropout.c(), ptr(g_pre), stream_ptr()), "mil_act_bwd_drop_f32")
        elif ctx.act in ("none", None):
            g_pre = g_y
        else:
            g_pre = torch.empty_like(g_y)
            check(L.mil_act_bwd_f32(ptr(g_y), ptr(saved), g_y.numel(), ACT[ctx.act], ptr(g_pre), stream_ptr()), "mil_act_bwd_f32")
        gx = gW = gb = None
        want_b = ctx.has_bias and ctx.needs_input_grad[2]
        if _skinny(M, N, K) and W.is_contiguous() and x.is_contiguous():
            g_pre = g_pre.contiguous()
            if ctx.needs_input_grad[1]:
                gW = torch.empty((N, K), dtype=torch.float32, device=x.device)
                gb = torch.empty(N, dtype=torch.float32, device=x.device) if want_b else None
            elif want_b:
                gb = g_pre.sum(0)
            if ctx.needs_input_grad[0]:
                gx = torch.empty((M, K), dtype=torch.float32, device=x.device)
            ws = _ws(L.mil_skinny_workspace_bytes(M, N, K), x.device)
            check(L.mil_skinny_bwd_f32(ptr(g_pre), ptr(x), K, ptr(W), M, N, K, ptr(gW), ptr(gb) if gW is not None else None, ptr(gx), ptr(ws), ws.numel(),
                                       stream_ptr()), "mil_skinny_bwd_f32")
            return gx, gW, gb, None, None, None
        if ctx.needs_input_grad[1]:
            gW, gb = weight_grad(g_pre, x, want_b)
        if want_b and gb is None:
            gb = torch.empty(N, dtype=torch.float32, device=x.device)
            ws = _ws(64 * N * 4, x.device)
            check(L.mil_colsum_f32(ptr(g_pre), M, N, ptr(gb), ptr(ws), ws.numel(), stream_ptr()), "mil_colsum_f32")
        if ctx.needs_input_grad[0]:
            # gx[m,k] = sum_n g_pre[m,n] W[n,k].  With W^T materialised ([K,N], a few hundred KB) this is the NT form again and
            # runs on the tensor cores; otherwise A(m,n) = g_pre[m*N + n], B(k,n) = W[n*K + k] on the CUDA cores.
            if _tc_supported(M, K, N, W) and N % 32 == 0 and K % 64 == 0:
                gx = linear_forward(g_pre.contiguous(), _transposed(W, ctx.volatile), None, "none", precision=GRAD_PRECISION)
            else:
                gx = sgemm(g_pre, N, 1, W, 1, K, M, K, N)
        return gx, gW, gb, None, None, None


WGRAD_TC = True              # tests / profiling can switch the tensor-core weight gradient off (exact fp32 CUDA-core split-K GEMM instead)


def weight_grad(g_pre, x, want_bias=False):
    """(gW[n,k] = sum_m g_pre[m,n] x[m,k], gb[n] = sum_m g_pre[m,n] or None): the contraction over the instances.
    Tensor cores (mil_wgrad_tc_f32: bf16 hi+lo, 3 products, deterministic slice reduction, bias gradient from the same pass) when
    the shape allows -- N % 128 == 0, K % 256 == 0, M >= 256 -- else the fp32 CUDA-core split-K GEMM
    (A(n,m) = g_pre[m*N + n], B(k,m) = x[m*K + k]; the caller adds mil_colsum_f32 for the bias)."""
    M, K = x.shape
    N = g_pre.shape[1]
    L = _lib.lib()
    if (WGRAD_TC and M >= TC_MIN_ROWS and N % 128 == 0 and K % 256 == 0 and g_pre.is_contiguous() and x.is_contiguous()
            and g_pre.data_ptr() % 16 == 0 and x.data_ptr() % 16 == 0 and L.mil_device_supported()):
        gW = torch.empty((N, K), dtype=torch.float32, device=x.device)
        gb = torch.empty(N, dtype=torch.float32, device=x.device) if want_bias else None
        ws = _ws(L.mil_wgrad_tc_workspace_bytes(M, N, K), x.device)
        check(L.mil_wgrad_tc_f32(ptr(g_pre), N, ptr(x), K, M, N, K, ptr(gW), ptr(gb), ptr(ws), ws.numel(), stream_ptr()), "mil_wgrad_tc_f32")
        return gW, gb
    return sgemm(g_pre, 1, N, x, 1, K, N, K, M, splitk=_splitk_for(N, K, M)), None


def linear_act(x: torch.Tensor, W: torch.Tensor, b: Optional[torch.Tensor] = None, act: str = "none",
               volatile: bool = False, dropout: Optional[DropSpec] = None) -> torch.Tensor:
    """drop(act(x @ W.T + b)) for x [M,K]; differentiable (CUDA backward).  volatile=True: do not trust a cached weight image
    (see weights_touched).  dropout: a DropSpec (next_dropout(p)) applied inside the GEMM's epilogue; the backward regenerates
    the same mask."""
    if not (torch.is_grad_enabled() and (x.requires_grad or W.requires_grad or (b is not None and b.requires_grad))):
        # no autograd wanted: skip the autograd.Function machinery (≈ 10 us of host time per call; the eval paths are launch-bound)
        return linear_forward(_need(x, "x"), _need(W, "weight"), _need(b, "bias") if b is not None else None, act, None, volatile=volatile, dropout=dropout)
    return _LinearAct.apply(x, W, b, act, volatile, dropout)


def linear_act_rows(x, W, b, act, row_ids):
    """No-grad variant that reads only the rows `row_ids` (int64) of x: fuses masking.py:108's gather into the GEMM."""
    x, W = _need(x, "x"), _need(W, "weight")
    row_ids = _need(row_ids, "row_ids", torch.int64)
    M, K, N = row_ids.numel(), x.shape[1], W.shape[0]
    return sgemm(x, K, 1, W, K, 1, M, N, K, bias=b, act=act, row_ids=row_ids)


# --------------------------------------------------------------------------------------------------------------
# softmax over instances + pooling
# --------------------------------------------------------------------------------------------------------------
def _pool_fwd(s, h, keep, want_attn):
    L = _lib.lib()
    n, H = h.shape
    npart = L.mil_pool_num_partials(n)
    part = torch.empty((npart, 2 + H), dtype=torch.float32, device=h.device)
    stats = torch.empty(2, dtype=torch.float32, device=h.device)
    pooled = torch.empty(H, dtype=torch.float32, device=h.device)
    attn = torch.empty(n, dtype=torch.float32, device=h.device) if want_attn else None
    check(L.mil_softmax_pool_fwd_f32(ptr(s), s.stride(0), ptr(h), n, H, ptr(keep), ptr(part), ptr(stats), ptr(pooled), ptr(attn), stream_ptr()),
          "mil_softmax_pool_fwd_f32")
    return pooled, stats, attn


class _SoftmaxPool(torch.autograd.Function):
    @staticmethod
    def forward(ctx, s, h, keep):
        h = _need(h, "h")
        if not s.is_cuda or s.dtype != torch.float32 or s.dim() != 1:
            raise RuntimeError("mhimk: `s` must be a 1-D float32 CUDA tensor")
        pooled, stats, attn = _pool_fwd(s, h, keep, True)
        ctx.save_for_backward(s, h, stats, pooled)
        ctx.keep = keep
        ctx.mark_non_differentiable(attn)
        return pooled, attn

    @staticmethod
    def backward(ctx, g_p, _g_attn):
        s, h, stats, pooled = ctx.saved_tensors
        L = _lib.lib()
        g_p = _need(g_p, "grad")
        n, H = h.shape
        g_s = torch.empty(n, dtype=torch.float32, device=h.device)
        g_h = torch.empty_like(h) if ctx.needs_input_grad[1] else None
        check(L.mil_softmax_pool_bwd_f32(ptr(s), s.stride(0), ptr(h), n, H, ptr(ctx.keep), ptr(stats), ptr(pooled), ptr(g_p), ptr(g_s), 1,
                                         ptr(g_h), 0, stream_ptr()), "mil_softmax_pool_bwd_f32")
        return g_s, g_h, None


def softmax_pool(s: torch.Tensor, h: torch.Tensor, keep: Optional[torch.Tensor] = None) -> Tuple[torch.Tensor, torch.Tensor]:
    """(pooled[H], attn[L]) = (softmax(s) @ h, softmax(s)); s [L] may be a strided column view.  attn is not differentiable."""
    if not (torch.is_grad_enabled() and (s.requires_grad or h.requires_grad)):
        if not s.is_cuda or s.dtype != torch.float32 or s.dim() != 1:
            raise RuntimeError("mhimk: `s` must be a 1-D float32 CUDA tensor")
        pooled, _, attn = _pool_fwd(s, _need(h, "h"), keep, True)
        return pooled, attn
    return _SoftmaxPool.apply(s, h, keep)


def pool_merge(part: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """Merge partials [(m, l, P[H])] -> (stats[2], pooled[H]); also the instance-shard merge of SURVEY §9.3."""
    L = _lib.lib()
    part = _need(part, "part")
    n, w = part.shape
    stats = torch.empty(2, dtype=torch.float32, device=part.device)
    pooled = torch.empty(w - 2, dtype=torch.float32, device=part.device)
    check(L.mil_pool_merge_f32(ptr(part), n, w - 2, ptr(stats), ptr(pooled), stream_ptr()), "mil_pool_merge_f32")
    return stats, pooled


def cam_score(s, t, stats, bias0):
    """score_n = max_c softmax_c(a_n t_nc + bias0) with a_n = exp(s_n - m)/l  (scoring.py:49-58).
    bias0: a python float, or a CUDA fp32 tensor whose first element is read on the device (no host sync)."""
    L = _lib.lib()
    s, t = _need(s, "s"), _need(t, "t")
    n, C = t.shape
    out = torch.empty(n, dtype=torch.float32, device=s.device)
    if isinstance(bias0, torch.Tensor):
        b = _need(bias0.detach().reshape(-1), "bias0")
        check(L.mil_cam_score_dev_f32(ptr(s), ptr(t), n, C, ptr(stats), ptr(b), ptr(out), stream_ptr()), "mil_cam_score_dev_f32")
    else:
        check(L.mil_cam_score_f32(ptr(s), ptr(t), n, C, ptr(stats), c_float(bias0), ptr(out), stream_ptr()), "mil_cam_score_f32")
    return out


# --------------------------------------------------------------------------------------------------------------
# masked hard-instance selection
# --------------------------------------------------------------------------------------------------------------
def topk(score: torch.Tensor, k: int, largest: bool = True) -> torch.Tensor:
    """Indices (int64 [k]) of the k largest/smallest scores ordered by value, ties lowest-index-first (masking.py:62)."""
    L = _lib.lib()
    score = _need(score.reshape(-1), "score")
    n = score.numel()
    idx = torch.empty(k, dtype=torch.int64, device=score.device)
    if k == 0:
        return idx
    ws = _ws(L.mil_topk_workspace_bytes(n), score.device)
    check(L.mil_topk_f32(ptr(score), n, k, 1 if largest else 0, ptr(idx), ptr(ws), ws.numel(), stream_ptr()), "mil_topk_f32")
    return idx


def col_argmax(a: torch.Tensor):
    """(idx int64 [C], val [C]): the row of the maximum of every column of a [M, C], lowest row first among ties (the critical
    instance of DSMIL, dsmil.py:91-92; no gradient)."""
    a = _need(a.detach(), "a")
    M, C = a.shape
    idx = torch.empty(C, dtype=torch.int64, device=a.device)
    val = torch.empty(C, dtype=torch.float32, device=a.device)
    check(_lib.lib().mil_col_argmax_f32(ptr(a), M, C, ptr(idx), ptr(val), stream_ptr()), "mil_col_argmax_f32")
    return idx, val


def mask_from_indices(idx: torch.Tensor, n: int):
    """(mask_ids int64 [1,n] = [kept ascending || idx], keep uint8 [n], len_keep int64 device scalar)  (masking.py:77-86)."""
    L = _lib.lib()
    idx = _need(idx.reshape(-1), "idx", torch.int64)
    mask_ids = torch.empty((1, n), dtype=torch.int64, device=idx.device)
    keep = torch.empty(n, dtype=torch.uint8, device=idx.device)
    len_keep = torch.empty(1, dtype=torch.int64, device=idx.device)
    check(L.mil_mask_from_indices(ptr(idx), idx.numel(), n, ptr(mask_ids), ptr(keep), ptr(len_keep), None, 0, stream_ptr()), "mil_mask_from_indices")
    return mask_ids, keep, len_keep


# --------------------------------------------------------------------------------------------------------------
# Nystrom / TransMIL forward (no autograd): every contraction on the library's own kernels
# --------------------------------------------------------------------------------------------------------------
NYSTROM_PRECISION = "fp16x3"     # K = 64 / 256 contractions of O(1) operands: fp16 hi+lo sits well below bf16 hi+lo there (tests/test_gpu_umma.py)


def layernorm(x: torch.Tensor, w: torch.Tensor, b: Optional[torch.Tensor], eps: float = 1e-5) -> torch.Tensor:
    """LayerNorm over the last axis of x [rows, cols] (no autograd; mil_layernorm_fwd_f32)."""
    x = _need(x, "x")
    y = torch.empty_like(x)
    check(_lib.lib().mil_layernorm_fwd_f32(ptr(x), x.shape[0], x.shape[1], ptr(w), ptr(b), c_float(eps), ptr(y), stream_ptr()), "mil_layernorm_fwd_f32")
    return y


def bmm(a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    """[B, M, K] @ [B, K, N] in exact fp32 on the CUDA cores (mil_sgemm_batched_f32); contiguous operands, no autograd."""
    a, b = _need(a, "a"), _need(b, "b")
    B, M, K = a.shape
    N = b.shape[2]
    c = torch.empty((B, M, N), dtype=torch.float32, device=a.device)
    check(_lib.lib().mil_sgemm_batched_f32(ptr(a), K, 1, M * K, ptr(b), 1, N, K * N, ptr(c), N, M * N, M, N, K, B, stream_ptr()), "mil_sgemm_batched_f32")
    return c


def bmm_nt(a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    """[B, M, K] @ [B, N, K]^T (contiguous) -> [B, M, N], exact fp32."""
    a, b = _need(a, "a"), _need(b, "b")
    B, M, K = a.shape
    N = b.shape[1]
    c = torch.empty((B, M, N), dtype=torch.float32, device=a.device)
    check(_lib.lib().mil_sgemm_batched_f32(ptr(a), K, 1, M * K, ptr(b), K, 1, N * K, ptr(c), N, M * N, M, N, K, B, stream_ptr()), "mil_sgemm_batched_f32")
    return c


def _tc_nt_view(x_view, W, y_view, ws, precision):
    """y_view[M, N] = x_view[M, K] @ W[N, K]^T on the tensor cores for operands that are column blocks of wider buffers (row stride =
    leading dimension); W contiguous, its image is rebuilt on every call (it is an activation, not a parameter)."""
    M, K = x_view.shape
    N = W.shape[0]
    check(_lib.lib().mil_linear_act_tc_ld_f32(c_void_p(x_view.data_ptr()), x_view.stride(0), M, K, ptr(W), None, N, ACT["none"], None,
                                              c_void_p(y_view.data_ptr()), y_view.stride(0), None, ptr(ws), ws.numel(), 0, PREC[precision], stream_ptr()),
          "mil_linear_act_tc_ld_f32")


def pinv_iter(x: torch.Tensor, iters: int = 6) -> torch.Tensor:
    """Moore-Penrose iteration of nystrom_attention.py:12-27 on x [heads, m, m]; ONE global scalar normaliser over all heads (:18).
    The 4 x iters batched products run in mil_sgemm_batched_f32 (exact fp32); the affine combinations are elementwise."""
    ax = x.abs()
    z = (x.transpose(-1, -2) / (ax.sum(dim=-1).max() * ax.sum(dim=-2).max())).contiguous()
    eye = torch.eye(x.shape[-1], device=x.device, dtype=x.dtype)[None]
    for _ in range(iters):
        xz = bmm(x, z)
        z = 0.25 * bmm(z, 13 * eye - bmm(xz, 15 * eye - bmm(xz, 7 * eye - xz)))
    return z


@torch.no_grad()
def nystrom_attention_forward(xn, Wqkv, Wout, bout, conv_w, heads, m, iters, scale, return_attn=False, no_norm=False, volatile=False):
    """NystromAttention.forward (nystrom_attention.py:65-152) for one bag without autograd, in the streaming form of SURVEY 9.7.

    xn [n, dim] (already normalised by the caller's LayerNorm).  Returns out [n, dim] and, with return_attn, the cls-row attention
    over the other tokens [heads, n-1] and v [heads, n-1, dh].
      to_qkv                     tensor cores, written as three column blocks of ONE [n_pad, 3*inner] buffer (no cat)
      landmarks                  segment means of q and k (mil_segment_mean_f32)
      pass A, per head           S = k q_l^T [n_pad, m] (tensor cores) -> kv = softmax_N(S)^T v with column max / sum
                                 (mil_colsoftmax_pool_f32): attn3 @ v without a transposed / second similarity tensor
      landmark block             A2 = softmax(q_l k_l^T) (batched fp32 GEMM + row softmax), A2^+ (pinv_iter), Z = A2^+ kv
      pass B, per head           P = softmax_m(q k_l^T) [n_pad, m] (tensor cores + row softmax) -> out_h = P Z_h (tensor cores, written
                                 into its 64 columns of the merged-heads buffer); reassociated: (attn1 @ attn2^+) @ (attn3 @ v)
                                 = attn1 @ (attn2^+ @ (attn3 @ v)) saves the [n, m] x [m, m] product
      residual conv              33 taps over the token axis, depth-wise per head, accumulated into the same buffer
      to_out                     tensor cores
    """
    L = _lib.lib()
    xn = _need(xn, "x")
    n, dim = xn.shape
    inner = Wqkv.shape[0] // 3
    dh = inner // heads
    dev = xn.device
    pad = (m - n % m) % m
    npad = n + pad
    seg = npad // m
    xp = torch.cat([xn.new_zeros(pad, dim), xn], dim=0) if pad else xn                       # FRONT zero padding (:70-73)
    qkv = torch.empty((npad, 3 * inner), dtype=torch.float32, device=dev)
    for blk in range(3):                                                                      # q | k | v column blocks, no bias (:52)
        Wb = Wqkv[blk * inner:(blk + 1) * inner]
        _tc_block(xp, Wb, None, "none", None, qkv[:, blk * inner:(blk + 1) * inner], Wqkv, blk, DEFAULT_PRECISION, volatile)
    ql, kl = torch.empty((heads, m, dh), dtype=torch.float32, device=dev), torch.empty((heads, m, dh), dtype=torch.float32, device=dev)
    check(L.mil_segment_mean_f32(ptr(qkv), 3 * inner, m, seg, 0, heads, dh, c_float(scale), ptr(ql), stream_ptr()), "mil_segment_mean_f32")   # q is scaled (:89)
    check(L.mil_segment_mean_f32(ptr(qkv), 3 * inner, m, seg, inner, heads, dh, c_float(1.0), ptr(kl), stream_ptr()), "mil_segment_mean_f32")
    kl_s = kl * scale                                                                         # (scale q) k_l^T == q (scale k_l)^T
    s2 = bmm_nt(ql, kl)                                                                       # [heads, m, m]
    a2 = s2.clone()
    check(L.mil_row_softmax_f32(ptr(a2), heads * m, m, stream_ptr()), "mil_row_softmax_f32")
    S = torch.empty((npad, m), dtype=torch.float32, device=dev)
    ws_t = _ws(L.mil_linear_tc_workspace_bytes(max(m, dh), max(m, dh)), dev)
    ws_p = _ws(L.mil_colsoftmax_pool_workspace_bytes(npad, m), dev)
    kv = torch.empty((heads, m, dh), dtype=torch.float32, device=dev)
    M3, L3 = torch.empty((heads, m), dtype=torch.float32, device=dev), torch.empty((heads, m), dtype=torch.float32, device=dev)
    for h in range(heads):                                                                    # pass A
        _tc_nt_view(qkv[:, inner + h * dh: inner + (h + 1) * dh], ql[h], S, ws_t, NYSTROM_PRECISION)
        check(L.mil_colsoftmax_pool_f32(ptr(S), c_void_p(qkv.data_ptr() + 4 * (2 * inner + h * dh)), 3 * inner, npad, m, dh, ptr(kv[h]), ptr(M3[h]), ptr(L3[h]),
                                        ptr(ws_p), ws_p.numel(), stream_ptr()), "mil_colsoftmax_pool_f32")
    a2i = pinv_iter(a2, iters)
    Zt = bmm(a2i, kv).transpose(1, 2).contiguous()                                            # [heads, dh, m]: the "weight" of the second product
    out = torch.empty((npad, inner), dtype=torch.float32, device=dev)
    cls_rows = torch.empty((heads, m), dtype=torch.float32, device=dev) if return_attn else None
    for h in range(heads):                                                                    # pass B
        _tc_nt_view(qkv[:, h * dh:(h + 1) * dh], kl_s[h], S, ws_t, NYSTROM_PRECISION)
        if return_attn and no_norm:
            cls_rows[h].copy_(S[pad])                                                         # raw similarities of the cls query (:146-147)
        check(L.mil_row_softmax_f32(ptr(S), npad, m, stream_ptr()), "mil_row_softmax_f32")
        if return_attn and not no_norm:
            cls_rows[h].copy_(S[pad])
        _tc_nt_view(S, Zt[h], out[:, h * dh:(h + 1) * dh], ws_t, NYSTROM_PRECISION)
    if conv_w is not None:                                                                    # out += res_conv(v) (:135-136)
        taps = conv_w.shape[2]
        check(L.mil_dwconv_tokens_f32(c_void_p(qkv.data_ptr() + 4 * 2 * inner), 3 * inner, npad, heads, dh, ptr(conv_w.reshape(heads, taps).contiguous()), taps, ptr(out),
                                      inner, 1, stream_ptr()), "mil_dwconv_tokens_f32")
    y = linear_forward(out, Wout, bout, "none", volatile=volatile)[pad:]
    if not return_attn:
        return y
    # cls-row attention over the last n-1 keys (:143-150): r = attn1[cls] @ attn2^+ (no_norm: raw sims and the pinv of the raw block)
    r = bmm(cls_rows[:, None, :].contiguous(), pinv_iter(s2, iters) if no_norm else a2i)[:, 0]            # [heads, m]
    attn = torch.empty((heads, n - 1), dtype=torch.float32, device=dev)
    rows0 = pad + 1
    for h in range(heads):
        Sk = S[: n - 1]
        _tc_nt_view(qkv[rows0:, inner + h * dh: inner + (h + 1) * dh], ql[h], Sk, ws_t, NYSTROM_PRECISION)          # sim3^T for the real keys
        if no_norm:
            check(L.mil_skinny_fwd_f32(ptr(Sk), m, n - 1, m, ptr(r[h].contiguous()), None, 1, ACT["none"], None, ptr(attn[h]), stream_ptr()), "mil_skinny_fwd_f32")
        else:
            check(L.mil_expdot_rows_f32(ptr(Sk), n - 1, m, ptr(M3[h]), ptr((r[h] / L3[h]).contiguous()), ptr(attn[h]), stream_ptr()), "mil_expdot_rows_f32")
    v = qkv[rows0:, 2 * inner:].reshape(n - 1, heads, dh).permute(1, 0, 2)
    return y, attn, v


@torch.no_grad()
def ppeg_forward(tokens, H, W, convs):
    """PPEG on the [H*W, C] token grid (transmil.py:50-64, emb_position.py:85-120): proj(x) + x + proj1(x) + proj2(x) as ONE depth-wise 7x7
    convolution with the summed kernel (5x5 and 3x3 zero-padded to 7x7, identity at the centre) and summed bias."""
    tokens = _need(tokens, "tokens")
    C = tokens.shape[1]
    w = torch.zeros((C, 7, 7), dtype=torch.float32, device=tokens.device)
    bias = torch.zeros(C, dtype=torch.float32, device=tokens.device)
    for conv in convs:
        k = conv.weight.shape[-1]
        o = (7 - k) // 2
        w[:, o:o + k, o:o + k] += conv.weight[:, 0]
        if conv.bias is not None:
            bias += conv.bias
    w[:, 3, 3] += 1.0
    y = torch.empty_like(tokens)
    check(_lib.lib().mil_ppeg_f32(ptr(tokens), H, W, C, ptr(w.reshape(C, 49).t().contiguous()), ptr(bias), ptr(y), stream_ptr()), "mil_ppeg_f32")
    return y


# --------------------------------------------------------------------------------------------------------------
# row selection by a permutation; Merge's cross-attention
# --------------------------------------------------------------------------------------------------------------
class _SplitRows(torch.autograd.Function):
    """(x[perm[:n_a]], x[perm[n_a:n_a+n_b]]) for perm = (a prefix of) a permutation of the rows; rows of perm beyond n_a + n_b get a zero
    gradient.  The backward is one scatter in which every row of gx is written once (no index sort, no atomics)."""

    @staticmethod
    def forward(ctx, x, perm, n_a, n_b):
        x, perm = _need(x, "x"), _need(perm.reshape(-1), "perm", torch.int64)
        rows, cols = x.shape
        L = _lib.lib()
        a = torch.empty((n_a, cols), dtype=torch.float32, device=x.device)
        b = torch.empty((n_b, cols), dtype=torch.float32, device=x.device) if n_b else None
        check(L.mil_take_rows_f32(ptr(x), ptr(perm), n_a, cols, ptr(a), stream_ptr()), "mil_take_rows_f32")
        if n_b:
            check(L.mil_take_rows_f32(ptr(x), c_void_p(perm.data_ptr() + 8 * n_a), n_b, cols, ptr(b), stream_ptr()), "mil_take_rows_f32")
        ctx.save_for_backward(perm)
        ctx.dims = (rows, cols, n_a, n_b)
        if perm.numel() != rows:
            raise RuntimeError("mhimk split_rows: `perm` must be a permutation of all rows of x")
        return (a, b) if n_b else (a, a.new_empty(0, cols))

    @staticmethod
    def backward(ctx, ga, gb):
        (perm,) = ctx.saved_tensors
        rows, cols, n_a, n_b = ctx.dims
        gx = torch.empty((rows, cols), dtype=torch.float32, device=perm.device)
        ga = _need(ga, "grad") if ga is not None else torch.zeros((n_a, cols), dtype=torch.float32, device=perm.device)
        gbp = ptr(_need(gb, "grad")) if (n_b and gb is not None) else None
        # rows perm[n_a + n_b:] (masked instances) receive zeros: the kernel writes 0 wherever gb is absent, so pass gb only when it spans them all
        if n_b and n_a + n_b != rows:
            raise RuntimeError("mhimk split_rows: the two parts must cover the permutation")
        check(_lib.lib().mil_scatter_rows_f32(ptr(ga), gbp, ptr(perm), n_a, rows, cols, ptr(gx), stream_ptr()), "mil_scatter_rows_f32")
        return gx, None, None, None


def take_rows(x: torch.Tensor, perm: torch.Tensor, n_take: int) -> torch.Tensor:
    """x[perm[:n_take]] for a full permutation `perm` of the rows of x [rows, cols] (mask_fn, masking.py:108); differentiable."""
    return _SplitRows.apply(x, perm, n_take, 0)[0]


def split_rows(x: torch.Tensor, perm: torch.Tensor, n_keep: int):
    """(x[perm[:n_keep]], x[perm[n_keep:]]) (Merge's random keep / drop split, merge.py:171-174); differentiable."""
    return _SplitRows.apply(x, perm, n_keep, x.shape[0] - n_keep)


class _MCA(torch.autograd.Function):
    @staticmethod
    def forward(ctx, q, kv, heads, scale, pmask):
        q, kv = _need(q, "q"), _need(kv, "kv")
        kq, inner = q.shape
        L_, dh = kv.shape[0], inner // heads
        lib = _lib.lib()
        P = torch.empty((heads, kq, L_), dtype=torch.float32, device=q.device)
        out = torch.empty((kq, inner), dtype=torch.float32, device=q.device)
        ws = _ws(lib.mil_mca_workspace_bytes(L_, kq, heads, dh), q.device)
        check(lib.mil_mca_fwd_f32(ptr(q), ptr(kv), L_, kq, heads, dh, c_float(scale), ptr(pmask), ptr(P), ptr(out), ptr(ws), ws.numel(), stream_ptr()),
              "mil_mca_fwd_f32")
        ctx.save_for_backward(q, kv, P, pmask if pmask is not None else q.new_empty(0))
        ctx.cfg = (heads, scale, pmask is not None)
        return out

    @staticmethod
    def backward(ctx, g):
        q, kv, P, pmask = ctx.saved_tensors
        heads, scale, has_mask = ctx.cfg
        kq, inner = q.shape
        L_, dh = kv.shape[0], inner // heads
        lib = _lib.lib()
        g = _need(g, "grad")
        dS = torch.empty_like(P)
        dq, dkv = torch.empty_like(q), torch.empty_like(kv)
        ws = _ws(lib.mil_mca_workspace_bytes(L_, kq, heads, dh), q.device)
        check(lib.mil_mca_bwd_f32(ptr(g), ptr(q), ptr(kv), ptr(P), ptr(pmask) if has_mask else None, L_, kq, heads, dh, c_float(scale), ptr(dS), ptr(dq),
                                  ptr(dkv), ptr(ws), ws.numel(), stream_ptr()), "mil_mca_bwd_f32")
        return dq, dkv, None, None, None


def mca_attend(q: torch.Tensor, kv: torch.Tensor, heads: int, scale: float, pmask: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Merge's cross-attention core (merge.py:52-65): q [k, heads*dh] (k <= 8), kv [L, 2*heads*dh] = [K | V] -> [k, heads*dh] =
    concat_h( dropout(softmax_L(scale q_h K_h^T)) V_h ); pmask [heads, k, L] is the dropout mask scaled by 1/(1-p) or None."""
    return _MCA.apply(q, kv, heads, float(scale), pmask)


# --------------------------------------------------------------------------------------------------------------
# fused tcgen05 forward
# --------------------------------------------------------------------------------------------------------------
# 16-bit hi/lo weight images for the fused pass, cached per (weight tensor OBJECTS, precision).  An entry is valid only
# while both tensors are the very same live objects (weakrefs; a recycled address is not enough) and their autograd
# version counters are unchanged (an optimizer step / load_state_dict / copy_ bumps `_version`).
_WS_CACHE = {}


PIPELINES = {"single": 1, "pair": 2}


def _pipeline(name, precision=None):
    """'single' (one CTA per 128-row tile) or 'pair' (two CTAs share every MMA, double-buffered TMEM accumulators, the epilogue
    overlaps the next tile's GEMM1).  Default: 'pair' (measured 5-7 % faster in the 3-product parity arithmetic, 3-4 % in the 1-product
    modes); MHIMK_PIPELINE=1|2 overrides the default."""
    import os
    if name in (None, "auto"):
        name = {"1": "single", "2": "pair"}.get(os.environ.get("MHIMK_PIPELINE", ""), "pair")
    if name not in PIPELINES:
        raise ValueError(f"mhimk: unknown fused pipeline {name!r}")
    return name


def _fused_workspace(W1, Wa, precision, pipeline="pair", volatile=False):
    """(workspace, ready, commit): the caller-owned buffer holding the 16-bit images of W1 / Wa, whether the kernel may use them
    as they are (1) or must rebuild them first (0), and the callback that marks the rebuilt images valid -- called only after
    the C call has returned 0 (a failed call must not leave a "ready" entry behind)."""
    import weakref
    L = _lib.lib()
    key = (id(W1), id(Wa), precision, pipeline)
    ver = (W1._version, Wa._version, _EPOCH)
    where = (W1.data_ptr(), Wa.data_ptr(), tuple(W1.shape), tuple(Wa.shape))
    hit = _WS_CACHE.get(key)
    if hit is not None and hit[0]() is W1 and hit[1]() is Wa and hit[4] == where:
        ws = hit[3]
        ready = 0 if (volatile or hit[2] != ver) else 1
    else:
        ws, ready = _ws(L.mil_fused_workspace_bytes(W1.shape[1], W1.shape[0], Wa.shape[0], 0), W1.device), 0
        _drop_dead(_WS_CACHE, 64)
    if ready:
        return ws, 1, lambda: None
    refs = (weakref.ref(W1), weakref.ref(Wa))
    _WS_CACHE[key] = (refs[0], refs[1], None, ws, where)

    def commit():
        _WS_CACHE[key] = (refs[0], refs[1], ver, ws, where)
    return ws, 0, commit


@torch.no_grad()
def abmil_fused_forward(x, W1, b1, act, Wa, ba, wc, bc, att_act="tanh", keep=None, Wp=None, want_scores=False, want_h=False,
                        precision: str = DEFAULT_PRECISION, Wcls=None, bcls=None, pipeline=None, volatile: bool = False,
                        dropout: Optional[DropSpec] = None, want_record: bool = False):
    """One streaming pass over x [N,D]: returns dict(pooled[H], stats[2] = (m, l), s[N]?, t[N,C]?, h[N,H]?, part).

    h = act(x W1^T + b1); s = wc . att_act(Wa h + ba) + bc; pooled = softmax_N(s) @ h; logits = Wcls pooled + bcls when a
    classifier is given (same kernel).  dropout: a DropSpec applied to h right after the activation (train-mode teacher,
    mhim.py:193-194).  (mil_abmil_fused_fwd_f32)
    """
    L = _lib.lib()
    x, W1, b1, Wa, wc = _need(x, "x"), _need(W1, "W1"), _need(b1, "b1"), _need(Wa, "Wa"), _need(wc.reshape(-1), "wc")
    N, D = x.shape
    H, Da = W1.shape[0], Wa.shape[0]
    dev = x.device
    npart = L.mil_fused_num_partials()
    ncls = Wcls.shape[0] if Wcls is not None else 0
    # one allocation for every small output (partials, stats, pooled, logits, exchange record): the host side of a bag is a handful of
    # python calls, and at 8 ranks per node each torch.empty is a contended allocator round trip (7 % at N > 1)
    n_part = npart * (2 + H)
    n_small = n_part + (-n_part) % 4 + 4 + H + (-H) % 4 + ((ncls + 3) // 4) * 4 + (2 + H if want_record else 0)
    small = torch.empty(n_small, dtype=torch.float32, device=dev)            # every sub-buffer starts 16-byte aligned
    o = n_part + (-n_part) % 4
    part, stats = small[:n_part].view(npart, 2 + H), small[o:o + 2]
    o += 4
    pooled = small[o:o + H]
    o += H + (-H) % 4
    logits = small[o:o + ncls].view(1, ncls) if Wcls is not None else None
    o += ((ncls + 3) // 4) * 4
    rec = small[o:o + 2 + H] if want_record else None
    s = torch.empty(N, dtype=torch.float32, device=dev) if want_scores else None
    C = Wp.shape[0] if Wp is not None else 0
    t = torch.empty((N, C), dtype=torch.float32, device=dev) if Wp is not None else None
    h = torch.empty((N, H), dtype=torch.float32, device=dev) if want_h else None
    pipeline = _pipeline(pipeline, precision)
    ws, ready, commit = _fused_workspace(W1, Wa, precision, pipeline, volatile)
    check(L.mil_abmil_fused_fwd_f32(ptr(x), N, D, H, ptr(W1), ptr(b1), ACT[act], ptr(Wa), ptr(ba), None, None, Da, ACT[att_act], ptr(wc),
                                    ptr(bc), ptr(keep), ptr(Wp), C, ptr(s), ptr(t), ptr(h), ptr(part), ptr(stats), ptr(pooled),
                                    ptr(rec), ptr(Wcls), ptr(bcls), ncls, ptr(logits), dropout.c() if dropout else None, ptr(ws), ws.numel(), ready,
                                    PREC[precision] | (PIPELINES[pipeline] << 8), stream_ptr()),
          "mil_abmil_fused_fwd_f32")
    commit()
    return {"pooled": pooled, "stats": stats, "s": s, "t": t, "h": h, "part": part, "logits": logits, "record": rec}


def shard_merge_cls(rec: torch.Tensor, Wcls=None, bcls=None):
    """Merge the gathered per-rank records [(m, l, P[H])] and apply the classifier in one launch -> (stats[2], pooled[H], logits[1,C]?)
    (mil_shard_merge_cls_f32; the tail of an instance-sharded forward)."""
    L = _lib.lib()
    rec = _need(rec, "rec")
    n, w = rec.shape
    stats = torch.empty(2, dtype=torch.float32, device=rec.device)
    pooled = torch.empty(w - 2, dtype=torch.float32, device=rec.device)
    ncls = Wcls.shape[0] if Wcls is not None else 0
    logits = torch.empty((1, ncls), dtype=torch.float32, device=rec.device) if Wcls is not None else None
    check(L.mil_shard_merge_cls_f32(ptr(rec), n, w - 2, ptr(Wcls), ptr(bcls), ncls, ptr(stats), ptr(pooled), ptr(logits), stream_ptr()),
          "mil_shard_merge_cls_f32")
    return stats, pooled, logits


def profile_fused(enable: bool):
    """Bracket every fused-kernel launch with CUDA events (kernel-only time for bench.py's roofline)."""
    _lib.lib().mil_profile_enable(1 if enable else 0)


def profile_collect():
    """-> (number of timed fused launches, their summed duration in ms); synchronises the recorded events."""
    import ctypes
    tot = ctypes.c_double(0.0)
    n = _lib.lib().mil_profile_collect(ctypes.byref(tot))
    return n, tot.value


@torch.no_grad()
def umma_selftest(A, B, precision: str = DEFAULT_PRECISION):
    """C = A @ B.T through the fused pass's TMA -> split -> tcgen05 -> TMEM pipeline (tests only)."""
    L = _lib.lib()
    A, B = _need(A, "A"), _need(B, "B")
    M, K = A.shape
    N = B.shape[0]
    C = torch.zeros((M, N), dtype=torch.float32, device=A.device)
    ws = _ws(N * K * 4 + 2048, A.device)
    check(L.mil_umma_selftest_f32(ptr(A), ptr(B), ptr(C), M, N, K, PREC[precision], ptr(ws), ws.numel(), stream_ptr()), "mil_umma_selftest_f32")
    return C
