"""CUDA-graph replay of the MHIM training step (mhimk.engines.GraphedStep): same numbers as the eager step, fresh dropout masks on
every replay, one graph per bag size."""
import types

import pytest
import torch
import torch.nn.functional as F

import cases

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def M():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import mhimk
    from mhimk import modules
    return modules


def build(M, base, d, seed, dropout):
    m = M.MHIM(**dict(cases.MHIM_KW, baseline=base, input_dim=d, dropout=dropout)).cuda()
    m.load_state_dict({k: v.cuda() for k, v in cases.mhim_state(seed, base, D=d).items()}, strict=True)
    for name, mod in m.named_modules():
        if isinstance(mod, torch.nn.Dropout) and name != "dp":
            mod.p = 0.0
    return m.train()


@pytest.mark.parametrize("base,d", [("attn", 1024), ("dsmil", 1536)])
def test_graphed_step_equals_eager_step(M, base, d):
    from mhimk.engines import CommonMIL, GraphedStep
    n = 3000
    stu, tea = build(M, base, d, 5, 0.0), build(M, base, d, 6, 0.0)
    noise = torch.rand(n - 90, generator=torch.Generator().manual_seed(1)).cuda()
    stu.merge._noise = lambda L, dev: noise[:L]                     # deterministic keep order: eager and graph see the same one
    args = types.SimpleNamespace(model="mhim", baseline=base, aux_alpha=0.5)
    eng, ce, label = CommonMIL(args), torch.nn.CrossEntropyLoss(), torch.tensor([1]).cuda()

    def step(bag):
        stu.zero_grad(set_to_none=True)
        logits, lab, aux, *_ = eng.forward_func(args, stu, tea, bag, label, ce, 1, 0, 0, 0, None)
        loss = ce(logits.view(1, -1), lab) + 0.5 * aux
        loss.backward()
        return loss.detach(), logits.detach()

    q0 = stu.merge.global_q_mm.data.clone()
    bag = cases.make_bag(3, n, d).cuda()
    loss_e, logits_e = (t.clone() for t in step(bag))
    grads_e = {k: p.grad.clone() for k, p in stu.named_parameters() if p.grad is not None}
    g = GraphedStep(step)
    for _ in range(2):                                                # capture (+ warm-ups), then a pure replay
        stu.merge.global_q_mm.data.copy_(q0)                          # the EMA side effect of Merge would otherwise drift the input
        loss_g, logits_g = g(bag)
    # warm-up / capture ran the step several times: global_q_mm moved by the EMA; restore and replay once more for the comparison
    stu.merge.global_q_mm.data.copy_(q0)
    loss_g, logits_g = g(bag)
    assert g.n_graphs == 1
    assert cases.rel_err(loss_g, loss_e) < 1e-6 and cases.rel_err(logits_g, logits_e) < 1e-6
    for k, p in stu.named_parameters():
        if k in grads_e:
            assert cases.rel_err(p.grad, grads_e[k]) < 1e-5, k
    bag2 = cases.make_bag(4, n + 128, d).cuda()                       # another bag size: a second graph, the first one still valid
    noise2 = torch.rand(n + 128 - 94, generator=torch.Generator().manual_seed(2)).cuda()
    stu.merge._noise = lambda L, dev: (noise2 if L == noise2.numel() else noise)[:L]
    g(bag2)
    assert g.n_graphs == 2
    stu.merge.global_q_mm.data.copy_(q0)
    loss_g2, _ = g(bag)
    assert cases.rel_err(loss_g2, loss_e) < 1e-6


def test_graphed_step_draws_fresh_dropout_masks(M):
    """With the reference's dropout = 0.25 every replay must see a new mask (device-resident Philox seed words refreshed inside the
    graph), and the same torch seed must reproduce the same sequence."""
    from mhimk.engines import GraphedStep
    tea = build(M, "attn", 1024, 6, 0.25)
    x = cases.make_bag(3, 2000, 1024).cuda()

    def teacher(bag):
        return tea.forward_teacher(bag)[1]

    g = GraphedStep(teacher)
    torch.manual_seed(11)
    a = g(x).clone()
    b = g(x).clone()
    assert not torch.equal(a, b)                                      # fresh mask per replay
    assert abs(float(a.mean()) - float(b.mean())) < 1e-3
    tea.eval()
    assert torch.equal(tea.forward_teacher(x)[1], tea.forward_teacher(x)[1])


@pytest.mark.parametrize("base,D", [("dsmil", 1536), ("attn", 1024)])
def test_graphed_inference_in_place_buffers(M, base, D):
    """Inference replayed from a graph: `buffers()` hands out the static input so a loader can fill it in place; the replay equals the
    eager call bit for bit (same kernels, same order) and follows new bag contents."""
    from mhimk.engines import GraphedStep
    tea = build(M, base, D, 7, 0.0).eval()
    x1, x2 = cases.make_bag(4, 1500, D).cuda(), cases.make_bag(5, 1500, D).cuda()
    with torch.no_grad():
        g = GraphedStep(lambda bag: tea.forward_test(bag))
        buf = g.buffers(x1)[0]
        assert buf.data_ptr() != x1.data_ptr() and g.n_graphs == 1
        for x in (x1, x2):
            buf.copy_(x)
            out = g(buf)
            ref = tea.forward_test(x)
            while isinstance(out, (tuple, list)):                     # dsmil returns nested logit lists
                out, ref = out[0], ref[0]
            assert torch.equal(out, ref)
    assert g.n_graphs == 1
