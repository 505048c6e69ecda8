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
    assert h.mil_abi_version() == int(re.search(r"#define MIL_ABI_VERSION (\d+)", open(HEADER).read()).group(1)) == 2


def test_python_binding_covers_the_header():
    import mhimk
    assert sorted(mhimk._lib.EXPORTS) == declared_symbols()


def test_cpu_tensors_are_rejected_loudly():
    import torch
    import mhimk
    with pytest.raises(RuntimeError, match="no CPU path"):
        mhimk.ops.linear_act(torch.zeros(4, 8), torch.zeros(2, 8), None, "relu")


def test_ema_update_rejects_cpu_models_loudly():
    import torch
    from mhimk.engines import ema_update
    with pytest.raises(RuntimeError, match="no CPU path"):
        ema_update(torch.nn.Linear(2, 2), torch.nn.Linear(2, 2), 0.5)
    with pytest.raises(AssertionError, match="Momentum"):
        ema_update(torch.nn.Linear(2, 2), torch.nn.Linear(2, 2), -0.1)


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


def test_weight_image_cache_policy(monkeypatch):
    """Host logic of the cached 16-bit weight images (no GPU): reuse while the weight is the same live tensor with an unchanged
    version counter; rebuild IN THE SAME BUFFER after an optimizer-style in-place update, when the caller says `volatile`
    (train mode), and after weights_touched() / a train()-eval() switch.  A write through `.data` -- the reference's EMA update,
    engines/base_engine.py:166-167 -- does not bump the version counter, which is why the last three exist."""
    import gc
    import torch
    import mhimk
    from mhimk import ops
    from mhimk.modules._common import MilModule

    class FakeLib:
        def mil_fused_workspace_bytes(self, *a):
            return 64

    monkeypatch.setattr(ops, "_ws", lambda nbytes, device: torch.empty(max(int(nbytes), 16), dtype=torch.uint8))
    monkeypatch.setattr(ops._lib, "lib", lambda: FakeLib())
    W1, Wa = torch.nn.Parameter(torch.randn(8, 4)), torch.nn.Parameter(torch.randn(2, 8))
    def use(**kw):                                                  # a successful kernel call: (ws, ready), images marked valid
        ws, ready, commit = ops._fused_workspace(W1, Wa, "bf16x3", "pair", **kw)
        commit()
        return ws, ready

    ws, ready, commit = ops._fused_workspace(W1, Wa, "bf16x3", "pair")
    assert ready == 0
    # the C call failed (commit never ran): the next call must rebuild, not trust the unbuilt buffer
    ws1, ready, commit = ops._fused_workspace(W1, Wa, "bf16x3", "pair")
    assert ready == 0 and ws1 is ws
    commit()
    ws2, ready = use()
    assert ready == 1 and ws2 is ws
    ws3, ready = use(volatile=True)
    assert ready == 0 and ws3 is ws
    assert use()[1] == 1
    W1.data.mul_(0.5)                                               # invisible to the version counter ...
    assert use()[1] == 1
    ops.weights_touched()                                           # ... hence the explicit notification
    assert use()[1] == 0
    assert use()[1] == 1
    MilModule().eval()                                              # a train()/eval() switch is such a notification
    assert use()[1] == 0
    with torch.no_grad():
        W1.mul_(2)                                                  # what optimizer.step() / load_state_dict do
    ws4, ready = use()
    assert ready == 0 and ws4 is ws                                 # rebuilt in place, no new allocation per step
    # W^T copy for the tensor-core dX: one persistent buffer per live weight, refreshed in place
    Wt = ops._transposed(W1)
    assert torch.equal(Wt, W1.detach().t())
    v = Wt._version
    assert ops._transposed(W1) is Wt and Wt._version == v
    W1.data.mul_(3.0)
    assert ops._transposed(W1, volatile=True) is Wt and Wt._version > v and torch.equal(Wt, W1.detach().t())
    n_ws, n_wt = len(ops._WS_CACHE), len(ops._WT_CACHE)
    del W1, Wa
    gc.collect()
    ops._drop_dead(ops._WS_CACHE, 64)
    ops._drop_dead(ops._WT_CACHE, 256)
    assert len(ops._WS_CACHE) == n_ws - 1 and len(ops._WT_CACHE) == n_wt - 1      # dead owners release their buffers


def test_modules_survive_the_plumbing_of_the_reference_main():
    """What the reference's factory / main do to a model before training (modules/__init__.py:179-214, main.py:224-236):
    deepcopy for the teacher, load_state_dict of the student's dict, .to(memory_format=channels_last), attribute pokes.
    Pure host logic: no kernel is called."""
    import copy
    import torch
    import cases
    from mhimk import modules as M
    for base, n_par in (("attn", 13), ("dsmil", 21), ("selfattn", 32)):
        stu = M.MHIM(**dict(cases.MHIM_KW, baseline=base, input_dim=1024))
        tea = copy.deepcopy(stu).to(memory_format=torch.channels_last)
        assert tea.load_state_dict(stu.state_dict(), strict=True).missing_keys == []
        assert len(list(tea.parameters())) == n_par == len(list(stu.parameters()))
        assert all(a is not b for a, b in zip(stu.parameters(), tea.parameters()))
        tea.merge_test = False
        assert tea.training and isinstance(tea, torch.nn.Module)
    for cls, args in ((M.DAttention, (1024, 2, 0.25, "relu")), (M.AttentionGated, (1024, 2)), (M.TransMIL, (1024, 2, 0.25, "relu")),
                      (M.MILNet, (2, 0.25, "relu"))):
        m = cls(*args)
        assert copy.deepcopy(m).load_state_dict(m.state_dict(), strict=True).unexpected_keys == []
