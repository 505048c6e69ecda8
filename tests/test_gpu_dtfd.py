"""DTFD-MIL drop-in (mhimk.modules.dtfd) on the path's kernels against the CPU oracle (pinned to the live reference class in
tests/test_oracle_vs_reference.py): both forwards, the three distillations, gradients, the engine adapter."""
import random
import types

import pytest
import torch
import torch.nn.functional as F

import cases
from oracle import mil_oracle as O

pytestmark = pytest.mark.gpu
TOL = 1e-4


@pytest.fixture(scope="module")
def M():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import mhimk
    from mhimk import modules
    return modules


def build(M, seed, distill):
    m = M.DTFD(torch.device("cuda"), 1e-4, 1e-5, 10, distill=distill).cuda()
    sd = cases.dtfd_state(seed)
    m.load_state_dict({k: v.cuda() for k, v in sd.items()}, strict=True)          # the reference's keys
    for mod in m.modules():
        if isinstance(mod, torch.nn.Dropout):
            mod.p = 0.0
    m.dimReduction.dropout = False
    return m, sd


@pytest.mark.parametrize("distill", ["AFS", "MaxS", "MaxMinS"])
@pytest.mark.parametrize("N", [7, 333, 10000])
def test_dtfd_forward_and_gradients(M, distill, N):
    m, sd = build(M, N, distill)
    x = cases.make_bag(N + 3, N, 1024)
    m.train()
    got = m(x.cuda())
    sd_ref = {k: v.clone().requires_grad_() for k, v in sd.items()}
    ref = O.dtfd_forward(sd_ref, x[0], True, distill=distill)
    assert cases.rel_err(got, ref) < TOL
    F.cross_entropy(got, torch.tensor([1]).cuda()).backward()
    F.cross_entropy(ref, torch.tensor([1])).backward()
    for k, p in m.named_parameters():
        g = sd_ref[k].grad
        if k.endswith("attention_weights.bias"):                         # softmax is shift-invariant: this gradient is mathematically zero
            assert p.grad is None or float(p.grad.abs().max()) < 1e-6, k
            continue
        if g is None or float(g.abs().max()) == 0.0:
            assert p.grad is None or float(p.grad.abs().max()) < 1e-7, k
            continue
        # dimReduction is followed by a ReLU: same gate-flip caveat as tests/test_gpu_baseline_sizes.py at N = 10 000
        tol = 2e-2 if (N >= 10000 and k == "dimReduction.fc1.weight") else TOL
        assert cases.rel_err(p.grad, g) < tol, (k, cases.rel_err(p.grad, g))
    m.eval()
    random.seed(5)
    with torch.no_grad():
        got_t = m(x.cuda())
    random.seed(5)
    ids = list(range(N))
    random.shuffle(ids)
    with torch.no_grad():
        assert cases.rel_err(got_t, O.dtfd_forward(sd, x[0], False, distill=distill, test_ids=ids)) < TOL


def test_dtfd_through_the_engine_adapter_with_dropout(M):
    """engines/common_mil.py's default branch (`model(bag, pos=pos)`), train mode with the reference's dropouts active: runs, finite,
    and different masks per call."""
    from mhimk.engines import CommonMIL
    m = M.DTFD(torch.device("cuda"), 1e-4, 1e-5, 10).cuda().train()
    args = types.SimpleNamespace(model="dtfd", baseline="attn", aux_alpha=0.0)
    bag, label = cases.make_bag(1, 3000, 1024).cuda(), torch.tensor([1]).cuda()
    out = CommonMIL(args).forward_func(args, m, None, bag, label, torch.nn.CrossEntropyLoss(), 1, 0, 0, 0, None)
    out2 = CommonMIL(args).forward_func(args, m, None, bag, label, torch.nn.CrossEntropyLoss(), 1, 0, 0, 0, None)
    assert tuple(out[0].shape) == (1, 2) and torch.isfinite(out[0]).all() and not torch.equal(out[0], out2[0])
    F.cross_entropy(out[0], label).backward()
    assert m.dimReduction.fc1.weight.grad is not None and torch.isfinite(m.dimReduction.fc1.weight.grad).all()
