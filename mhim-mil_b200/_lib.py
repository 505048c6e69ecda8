"""ctypes binding of libmhimk.so (the C ABI declared in include/mhimk.h).

The product path has NO fallback: if the shared library is missing or the device is not sm_100,
every op raises.  Build with `python __graft_entry__.py build` (or mhim-mil_b200/csrc/build.sh).
"""
import ctypes
import os
from ctypes import c_char_p, c_float, c_int, c_int64, c_size_t, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libmhimk.so")
ABI_VERSION = 2

ACT = {"none": 0, None: 0, "relu": 1, "gelu": 2, "tanh": 3, "sigmoid": 4}
PREC = {"bf16x3": 0, "fp16": 1, "bf16": 2, "fp16x3": 3}



class DropoutT(ctypes.Structure):
    """mil_dropout_t of include/mhimk.h (a HOST struct passed by pointer)."""
    _fields_ = [("mode", c_int), ("p", c_float), ("seed", ctypes.c_uint64), ("offset", ctypes.c_uint64), ("keep_bits", c_void_p)]


_lib = None

_SIGS = {
    "mil_abi_version": (c_int, []),
    "mil_last_error": (c_char_p, []),
    "mil_device_supported": (c_int, []),
    "mil_abmil_fused_fwd_f32": (c_int, [c_void_p, c_int64, c_int, c_int, c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p,
                                        c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p,
                                        c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_size_t, c_int, c_int, c_void_p]),
    "mil_dropout_bits": (c_int, [c_int64, c_int, c_void_p, c_void_p, c_void_p]),
    "mil_profile_enable": (None, [c_int]),
    "mil_profile_collect": (c_int, [ctypes.POINTER(ctypes.c_double)]),
    "mil_fused_num_partials": (c_int, []),
    "mil_fused_workspace_bytes": (c_size_t, [c_int, c_int, c_int, c_int]),
    "mil_sgemm_f32": (c_int, [c_void_p, c_int64, c_int64, c_void_p, c_void_p, c_int64, c_int64, c_void_p, c_void_p, c_int64, c_void_p,
                              c_int64, c_int64, c_int64, c_int, c_int, c_void_p, c_size_t, c_void_p]),
    "mil_linear_act_tc_f32": (c_int, [c_void_p, c_int64, c_int, c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_size_t,
                                      c_int, c_int, c_void_p]),
    "mil_linear_act_tc_ld_f32": (c_int, [c_void_p, c_int64, c_int64, c_int, c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p, c_int64, c_void_p, c_void_p,
                                         c_size_t, c_int, c_int, c_void_p]),
    "mil_sgemm_batched_f32": (c_int, [c_void_p, c_int64, c_int64, c_int64, c_void_p, c_int64, c_int64, c_int64, c_void_p, c_int64, c_int64, c_int64, c_int64,
                                      c_int64, c_int, c_void_p]),
    "mil_layernorm_fwd_f32": (c_int, [c_void_p, c_int64, c_int, c_void_p, c_void_p, c_float, c_void_p, c_void_p]),
    "mil_segment_mean_f32": (c_int, [c_void_p, c_int64, c_int, c_int, c_int, c_int, c_int, c_float, c_void_p, c_void_p]),
    "mil_row_softmax_f32": (c_int, [c_void_p, c_int64, c_int, c_void_p]),
    "mil_colsoftmax_pool_f32": (c_int, [c_void_p, c_void_p, c_int64, c_int64, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "mil_colsoftmax_pool_workspace_bytes": (c_size_t, [c_int64, c_int]),
    "mil_expdot_rows_f32": (c_int, [c_void_p, c_int64, c_int, c_void_p, c_void_p, c_void_p, c_void_p]),
    "mil_dwconv_tokens_f32": (c_int, [c_void_p, c_int64, c_int64, c_int, c_int, c_void_p, c_int, c_void_p, c_int64, c_int, c_void_p]),
    "mil_ppeg_f32": (c_int, [c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p]),
    "mil_linear_tc_workspace_bytes": (c_size_t, [c_int, c_int]),
    "mil_skinny_supported": (c_int, [c_int64, c_int, c_int]),
    "mil_skinny_fwd_f32": (c_int, [c_void_p, c_int64, c_int64, c_int, c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p]),
    "mil_skinny_bwd_f32": (c_int, [c_void_p, c_void_p, c_int64, c_void_p, c_int64, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "mil_skinny_workspace_bytes": (c_size_t, [c_int64, c_int, c_int]),
    "mil_wgrad_tc_f32": (c_int, [c_void_p, c_int64, c_void_p, c_int64, c_int64, c_int, c_int, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "mil_wgrad_tc_workspace_bytes": (c_size_t, [c_int64, c_int, c_int]),
    "mil_act_bwd_f32": (c_int, [c_void_p, c_void_p, c_int64, c_int, c_void_p, c_void_p]),
    "mil_act_bwd_drop_f32": (c_int, [c_void_p, c_void_p, c_int64, c_int, c_int, c_void_p, c_void_p, c_void_p]),
    "mil_colsum_f32": (c_int, [c_void_p, c_int64, c_int64, c_void_p, c_void_p, c_size_t, c_void_p]),
    "mil_softmax_pool_fwd_f32": (c_int, [c_void_p, c_int64, c_void_p, c_int64, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "mil_pool_num_partials": (c_int, [c_int64]),
    "mil_softmax_pool_bwd_f32": (c_int, [c_void_p, c_int64, c_void_p, c_int64, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int64,
                                         c_void_p, c_int, c_void_p]),
    "mil_pool_merge_f32": (c_int, [c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p]),
    "mil_shard_merge_cls_f32": (c_int, [c_void_p, c_int, c_int, c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p]),
    "mil_cam_score_f32": (c_int, [c_void_p, c_void_p, c_int64, c_int, c_void_p, c_float, c_void_p, c_void_p]),
    "mil_cam_score_dev_f32": (c_int, [c_void_p, c_void_p, c_int64, c_int, c_void_p, c_void_p, c_void_p, c_void_p]),
    "mil_take_rows_f32": (c_int, [c_void_p, c_void_p, c_int64, c_int, c_void_p, c_void_p]),
    "mil_scatter_rows_f32": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_int, c_void_p, c_void_p]),
    "mil_mca_fwd_f32": (c_int, [c_void_p, c_void_p, c_int64, c_int, c_int, c_int, c_float, c_void_p, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "mil_mca_bwd_f32": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int, c_int, c_int, c_float, c_void_p, c_void_p, c_void_p,
                                c_void_p, c_size_t, c_void_p]),
    "mil_mca_workspace_bytes": (c_size_t, [c_int64, c_int, c_int, c_int]),
    "mil_adam_step_f32": (c_int, [c_void_p, c_int, c_float, ctypes.c_double, ctypes.c_double, c_float, c_float, c_int, c_float, c_float, c_void_p, c_void_p]),
    "mil_ema_update_f32": (c_int, [c_void_p, c_int, c_float, c_float, c_void_p]),
    "mil_topk_f32": (c_int, [c_void_p, c_int64, c_int64, c_int, c_void_p, c_void_p, c_size_t, c_void_p]),
    "mil_mask_from_indices": (c_int, [c_void_p, c_int64, c_int64, c_void_p, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "mil_col_argmax_f32": (c_int, [c_void_p, c_int64, c_int, c_void_p, c_void_p, c_void_p]),
    "mil_topk_workspace_bytes": (c_size_t, [c_int64]),
    "mil_umma_selftest_f32": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_size_t, c_void_p]),
    "mil_pair_plan_item": (c_int, [c_int64, c_int, c_int, c_int, c_int, c_void_p]),
}

EXPORTS = tuple(_SIGS)


def lib():
    """Load (once) and return the ctypes handle; raises if the extension was not built."""
    global _lib
    if _lib is None:
        if not os.path.isfile(LIB_PATH):
            raise RuntimeError(f"{LIB_PATH} not found: the CUDA extension is not built (run `python __graft_entry__.py build`). "
                               "mhimk has no CPU or PyTorch fallback.")
        h = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in _SIGS.items():
            fn = getattr(h, name)
            fn.restype = res
            fn.argtypes = args
        if h.mil_abi_version() != ABI_VERSION:
            raise RuntimeError(f"libmhimk.so ABI {h.mil_abi_version()} != expected {ABI_VERSION}")
        _lib = h
    return _lib


def check(rc: int, what: str):
    if rc != 0:
        msg = lib().mil_last_error().decode("utf-8", "replace")
        raise RuntimeError(f"{what} failed (rc={rc}): {msg}")


def ptr(t):
    """Device pointer of a tensor (None -> NULL)."""
    return None if t is None else c_void_p(t.data_ptr())


def stream_ptr():
    import torch
    return c_void_p(torch._C._cuda_getCurrentRawStream(torch.cuda.current_device()))
