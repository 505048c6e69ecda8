"""The callers either side of the hot path: one-launch Adam / AdamW, the bag loader, validation on the device."""
import os

import pytest
import torch

import cases

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def E():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import mhimk
    from mhimk import engines
    return engines


class Net(torch.nn.Module):
    def __init__(self, seed):
        super().__init__()
        g = torch.Generator().manual_seed(seed)
        self.a = torch.nn.Parameter(torch.randn(70001, generator=g))          # spans CTAs, odd length
        self.b = torch.nn.Parameter(torch.randn(3, generator=g))
        self.c = torch.nn.Parameter(torch.randn(257, 129, generator=g))
        self.frozen = torch.nn.Parameter(torch.randn(5, generator=g), requires_grad=False)


@pytest.mark.parametrize("adamw,wd", [(False, 1e-5), (False, 0.0), (True, 1e-2)])
def test_fused_adam_matches_torch(E, adamw, wd):
    """Same trajectory as torch.optim.Adam / AdamW (train_utils.py:55-65: one param group {params, lr, weight_decay}) over 6 steps,
    interchangeable state_dict, lr changes by a scheduler honoured."""
    ref, mine = Net(1).cuda(), Net(1).cuda()
    params = lambda m: [{"params": filter(lambda p: p.requires_grad, m.parameters()), "lr": 2e-4, "weight_decay": wd}]
    o_ref = (torch.optim.AdamW if adamw else torch.optim.Adam)(params(ref))
    o_mine = E.FusedAdam(params(mine), adamw=adamw) if not adamw else E.FusedAdam.adamw(params(mine))
    g = torch.Generator().manual_seed(2)
    for step in range(6):
        for pr, pm in zip(ref.parameters(), mine.parameters()):
            if pr.requires_grad:
                gr = torch.randn(pr.shape, generator=g).cuda() * 0.1
                pr.grad, pm.grad = gr.clone(), gr.clone()
        if step == 3:                                                        # what StepLR / cosine schedulers do
            for o in (o_ref, o_mine):
                o.param_groups[0]["lr"] = 5e-5
        o_ref.step(), o_mine.step()
        for (k, pr), pm in zip(ref.named_parameters(), mine.parameters()):
            assert cases.rel_err(pm, pr) < 2e-6, (step, k)
    sd_r, sd_m = o_ref.state_dict(), o_mine.state_dict()
    assert set(sd_r["state"][0]) == set(sd_m["state"][0]) == {"step", "exp_avg", "exp_avg_sq"}
    assert float(sd_m["state"][0]["step"]) == 6.0
    assert cases.rel_err(sd_m["state"][0]["exp_avg_sq"], sd_r["state"][0]["exp_avg_sq"]) < 2e-6
    o2 = E.FusedAdam(params(mine), adamw=adamw)                               # resume from torch's own checkpoint
    import copy
    sd_load = copy.deepcopy(o_ref.state_dict())                               # (state_dict() hands out the live tensors; a checkpoint file would not alias them)
    sd_load["param_groups"][0].update(adamw=adamw, capturable=False)
    o2.load_state_dict(sd_load)
    for pr, pm in zip(ref.parameters(), mine.parameters()):
        if pr.requires_grad:
            pm.data.copy_(pr.data)
            gr = torch.randn(pr.shape, generator=g).cuda() * 0.1
            pr.grad, pm.grad = gr.clone(), gr.clone()
    o_ref.step(), o2.step()
    for pr, pm in zip(ref.parameters(), mine.parameters()):
        assert cases.rel_err(pm, pr) < 2e-6


def test_fused_adam_capturable_equals_host_counter(E):
    a, b = Net(3).cuda(), Net(3).cuda()
    oa, ob = E.FusedAdam([p for p in a.parameters() if p.requires_grad], lr=1e-3), E.FusedAdam([p for p in b.parameters() if p.requires_grad], lr=1e-3, capturable=True)
    g = torch.Generator().manual_seed(4)
    for _ in range(4):
        for pa, pb in zip(a.parameters(), b.parameters()):
            if pa.requires_grad:
                gr = torch.randn(pa.shape, generator=g).cuda()
                pa.grad, pb.grad = gr.clone(), gr.clone()
        oa.step(), ob.step()
    for pa, pb in zip(a.parameters(), b.parameters()):
        assert cases.rel_err(pb, pa) < 1e-6
    assert ob.state[next(iter(ob.state))]["step"].is_cuda


def test_fused_adam_rejects_cpu_parameters(E):
    o = E.FusedAdam([torch.nn.Parameter(torch.zeros(4))])
    o.param_groups[0]["params"][0].grad = torch.zeros(4)
    with pytest.raises(RuntimeError, match="no CPU path"):
        o.step()


def test_metrics_match_sklearn(E):
    from sklearn.metrics import accuracy_score, f1_score, precision_score, recall_score, roc_auc_score
    g = torch.Generator().manual_seed(7)
    for n, C in [(200, 2), (57, 2), (300, 3)]:
        logits = torch.randn(n, C, generator=g)
        logits[::7] = logits[3]                                               # exact score ties
        labels = torch.randint(0, C, (n,), generator=g)
        m = E.classification_metrics(logits.cuda(), labels.cuda())
        prob = torch.softmax(logits, -1).numpy()
        pred = prob.argmax(-1)
        y = labels.numpy()
        auc = roc_auc_score(y, prob[:, 1]) if C == 2 else roc_auc_score(y, prob, multi_class="ovr", average="macro")
        want = {"acc": accuracy_score(y, pred), "precision": precision_score(y, pred, average="macro", zero_division=0),
                "recall": recall_score(y, pred, average="macro", zero_division=0), "f1": f1_score(y, pred, average="macro", zero_division=0), "auc": auc}
        for k, v in want.items():
            assert abs(float(m[k]) - v) < 1e-6, (k, float(m[k]), v)


def test_validate_and_bag_loader(E, tmp_path):
    """`.pt` [N, D] files (the CLAM feature layout) -> BagLoader -> validate(): same logits as calling the model bag by bag, every bag
    exactly once over two ranks' slices, labels and names aligned."""
    import types
    from mhimk.modules import DAttention
    sizes, D = [300, 1, 1025, 64, 777], 1024
    paths, labels, bags = [], [], []
    for i, n in enumerate(sizes):
        x = cases.make_bag(100 + i, n, D)[0]
        p = os.path.join(tmp_path, f"slide_{i}.pt")
        torch.save(x, p)
        paths.append(p); labels.append(i % 2); bags.append(x)
    model = DAttention(D, 2, dropout=0.0, act="relu").cuda().eval()
    args = types.SimpleNamespace(model="abmil", baseline="attn", aux_alpha=0.0)
    loader = E.BagLoader(paths, labels)
    seen = []

    def pairs():
        for bag, label, name in loader:
            seen.append((name, tuple(bag.shape), int(label)))
            yield bag, label

    metrics, logits, labs = E.validate(args, model, pairs(), len(sizes), 2, torch.nn.CrossEntropyLoss())
    assert [s[0] for s in seen] == [f"slide_{i}.pt" for i in range(5)] and [s[1] for s in seen] == [(1, n, D) for n in sizes]
    assert labs.cpu().tolist() == labels and loader.h2d_bytes == sum(sizes) * D * 4
    with torch.no_grad():
        for i, x in enumerate(bags):
            assert cases.rel_err(logits[i], model(x[None].cuda())[0]) < 1e-6
    assert 0.0 <= metrics["acc"] <= 1.0 and metrics["loss"] > 0
    got = []
    for r in range(2):
        got += [name for _, _, name in E.BagLoader(paths, labels, rank=r, world=2)]
    assert sorted(got) == sorted(f"slide_{i}.pt" for i in range(5))
    # in-memory tensors and a seeded shuffle
    order = [name for _, _, name in E.BagLoader(bags, labels, shuffle_seed=1)]
    assert sorted(order) == [f"bag{i}" for i in range(5)]
