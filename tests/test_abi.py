"""CPU-only: the C-ABI library builds/loads and exports every symbol include/mhimk.h declares."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "mhimk.h")
LIB = os.path.join(ROOT, "mhim-mil_b200", "libmhimk.so")


def declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(mil_[a-z0-9_]+)\s*\(", src)))


def test_header_declares_expected_entry_points():
    syms = declared_symbols()
    for s in ("mil_abmil_fused_fwd_f32", "mil_sgemm_f32", "mil_softmax_pool_fwd_f32", "mil_softmax_pool_bwd_f32", "mil_topk_f32",
              "mil_mask_from_indices", "mil_pool_merge_f32", "mil_cam_score_f32", "mil_last_error", "mil_abi_version"):
        assert s in syms


def test_library_exports_every_declared_symbol():
    if not os.path.isfile(LIB):
        import __graft_entry__
        __graft_entry__.build()
    h = ctypes.CDLL(LIB)
    for s in declared_symbols():
        assert hasattr(h, s), f"{s} declared in mhimk.h but not exported"
    h.mil_abi_version.restype = ctypes.c_int
    assert h.mil_abi_version() == 1


def test_python_binding_covers_the_header():
    import mhimk
    assert sorted(mhimk._lib.EXPORTS) == declared_symbols()


def test_cpu_tensors_are_rejected_loudly():
    import torch
    import mhimk
    with pytest.raises(RuntimeError, match="no CPU path"):
        mhimk.ops.linear_act(torch.zeros(4, 8), torch.zeros(2, 8), None, "relu")


def test_fused_pipeline_selection(monkeypatch):
    """Host logic of the pipeline switch (no GPU): default, environment override, explicit names, bad names."""
    import mhimk
    monkeypatch.delenv("MHIMK_PIPELINE", raising=False)
    assert mhimk.ops._pipeline(None, "bf16x3") == "pair" and mhimk.ops._pipeline("auto", "fp16") == "pair"
    monkeypatch.setenv("MHIMK_PIPELINE", "1")
    assert mhimk.ops._pipeline(None, "bf16x3") == "single"
    assert mhimk.ops._pipeline("pair", "bf16x3") == "pair"          # an explicit choice wins over the environment
    with pytest.raises(ValueError):
        mhimk.ops._pipeline("triple")
    assert mhimk.ops.PIPELINES == {"single": 1, "pair": 2}          # = MIL_PIPE_SINGLE / MIL_PIPE_PAIR of include/mhimk.h
    hdr = open(HEADER).read()
    assert "MIL_PIPE_SINGLE  = 1" in hdr and "MIL_PIPE_PAIR    = 2" in hdr
