"""GPU: the CLAM_SB / CLAM_MB drop-ins (mhimk.modules.clam) against the CPU oracle (pinned to the live classes in
tests/test_oracle_vs_reference.py): eval logits and attention, the training forward with the instance-level branch (smooth top-1 SVM loss),
every gradient tensor, dropout through caller-supplied keep bits, and a CUDA-graph replay of the whole step (no host synchronisation)."""
import pytest
import torch
import torch.nn.functional as F

import cases
from oracle import mil_oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def M():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import mhimk  # noqa: F401
    from mhimk import modules
    return modules


def build(M, mb, sd, **kw):
    cls = M.CLAM_MB if mb else M.CLAM_SB
    m = cls(input_dim=1024, **kw).cuda()
    m.load_state_dict({k: v.cuda() for k, v in sd.items()}, strict=True)
    return m


@pytest.mark.parametrize("mb", [False, True])
@pytest.mark.parametrize("gate,subtyping,n_cls", [(True, False, 2), (True, True, 3), (False, False, 2)])
@pytest.mark.parametrize("act,N", [("relu", 300), ("gelu", 5000)])
def test_clam_eval_train_forward_and_gradients(M, mb, gate, subtyping, n_cls, act, N):
    sd = cases.clam_state(N + n_cls, mb, C=n_cls, gate=gate)
    x = cases.make_bag(N + 7, N, 1024)[0]
    m = build(M, mb, sd, gate=gate, n_classes=n_cls, subtyping=subtyping, act=act, dropout=0.0)
    xg = x.cuda()
    m.eval()
    with torch.no_grad():
        lg = m(xg[None])
        att = m(xg[None], attention_only=True)
    lg_ref, _, a_ref = O.clam_forward(sd, x, mb, n_cls, gate, act, subtyping=subtyping)
    assert cases.rel_err(lg, lg_ref) < 1e-4 and cases.rel_err(att.reshape(a_ref.shape), a_ref) < 1e-4
    m.train()
    for lab in (1, 0):
        m.zero_grad(set_to_none=True)
        label = torch.tensor([lab], device="cuda")
        lg, il, ps = m(xg[None], label=label, instance_eval=True)
        (F.cross_entropy(lg, label) + il).backward()
        sdl = {k: v.clone().requires_grad_(v.is_floating_point()) for k, v in sd.items()}
        lr, ir, _ = O.clam_forward(sdl, x, mb, n_cls, gate, act, subtyping=subtyping, label=lab)
        (F.cross_entropy(lr, torch.tensor([lab])) + ir).backward()
        assert ps == N and cases.rel_err(lg, lr) < 1e-4 and cases.rel_err(il, ir) < 1e-4
        # ReLU gate flips (profiles/round2_gradient_parity.md) reach the first layer's gradient at N = 5 000; everything else is at 1e-4
        for k, p_ in m.named_parameters():
            ref = sdl[k].grad
            if ref is None or float(ref.abs().max()) < 1e-6:
                assert p_.grad is None or float(p_.grad.abs().max()) < 1e-5, k
                continue
            gate_flip = act == "relu" and k.startswith("attention_net.0.")
            assert cases.rel_err(p_.grad, ref) < (5e-3 if gate_flip else 1e-4), (k, lab)
    assert m(xg[None], label=[1])[1] == 0                                      # a list label switches the instance branch off (clam.py:177-178)


@pytest.mark.parametrize("mb", [False, True])
def test_clam_dropout_with_caller_masks(M, mb):
    """dropout = 0.25: the fc Dropout and the attention net's two Dropout(0.25)s run in the Linear epilogues; with the keep bits supplied
    by the test the step equals the oracle with the same masks."""
    from mhimk import ops
    N, C = 1500, 2
    sd = cases.clam_state(21, mb, C=C, fc_drop=True)
    x = cases.make_bag(22, N, 1024)[0]
    m = build(M, mb, sd, n_classes=C, dropout=0.25, act="gelu").train()
    g = torch.Generator().manual_seed(5)
    masks = [torch.rand(N, w, generator=g) > 0.25 for w in (512, 256, 256)]
    it = iter(masks)

    def hook(rows, ncols, p, device):
        k = next(it)
        assert tuple(k.shape) == (rows, ncols) and p == 0.25
        return ops.DropSpec(p=p, keep_bits=ops.pack_keep_bits(k.to(device)))
    ops.DROPOUT_HOOK = hook
    try:
        label = torch.tensor([1], device="cuda")
        lg, il, _ = m(x.cuda()[None], label=label, instance_eval=True)
        (F.cross_entropy(lg, label) + il).backward()
    finally:
        ops.DROPOUT_HOOK = None
    sdl = {k: v.clone().requires_grad_(v.is_floating_point()) for k, v in sd.items()}
    dm = [k.float() / 0.75 for k in masks]
    lr, ir, _ = O.clam_forward(sdl, x, mb, C, True, "gelu", label=1, fc_drop=True, drop_h=dm[0], drop_a=dm[1], drop_b=dm[2])
    (F.cross_entropy(lr, torch.tensor([1])) + ir).backward()
    assert cases.rel_err(lg, lr) < 1e-4 and cases.rel_err(il, ir) < 1e-4
    for k, p_ in m.named_parameters():
        ref = sdl[k].grad
        if ref is None or float(ref.abs().max()) < 1e-6:
            continue
        assert cases.rel_err(p_.grad, ref) < 1e-4, k


def test_clam_step_replays_from_a_cuda_graph(M):
    """No host synchronisation in the training forward (the class label is consumed on the device): the step captures and replays."""
    from mhimk.engines import GraphedStep
    N = 2000
    sd = cases.clam_state(31, False)
    m = build(M, False, sd, n_classes=2, dropout=0.0).train()
    x = cases.make_bag(32, N, 1024)[0].cuda()

    def step(bag, label):
        m.zero_grad(set_to_none=True)
        lg, il, _ = m(bag[None], label=label, instance_eval=True)
        loss = F.cross_entropy(lg, label) + il
        loss.backward()
        return loss

    gs = GraphedStep(step)
    for lab in (1, 0, 1):
        label = torch.tensor([lab], device="cuda")
        lg_e = step(x, label).detach().clone()
        ge = {k: p.grad.clone() for k, p in m.named_parameters() if p.grad is not None}
        lg_g = gs(x, label).clone()
        assert cases.rel_err(lg_g, lg_e) < 1e-6
        for k, p in m.named_parameters():
            if k in ge and float(ge[k].abs().max()) > 1e-6:
                assert cases.rel_err(p.grad, ge[k]) < 1e-5, k
    assert gs.n_graphs == 1
