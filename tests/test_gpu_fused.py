"""GPU parity of the fused tcgen05 ABMIL forward against the CPU oracle (fp64 and fp32)."""
import pytest
import torch

import cases
from oracle import mil_oracle as O

pytestmark = pytest.mark.gpu

# relative tolerance (max|d| / max|ref|) per arithmetic; the north-star gate is 1e-4 for the parity mode
TOL = {"bf16x3": 1e-4, "fp16x3": 1e-4, "fp16": 2e-3, "bf16": 2e-2}


@pytest.fixture(scope="module")
def K():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import mhimk
    return mhimk.ops


def run_oracle(sd, x, act, keep=None):
    sd64 = {k: v.double() for k, v in sd.items()}
    h = O.apply_act(O.affine(x[0].double(), sd64["feature.0.weight"], sd64["feature.0.bias"]), act)
    u = torch.tanh(O.affine(h, sd64["attention.0.weight"], sd64["attention.0.bias"]))
    s = O.affine(u, sd64["attention.2.weight"], sd64["attention.2.bias"])[:, 0]
    if keep is not None:
        s = s.masked_fill(keep == 0, float("-inf"))
    a = torch.softmax(s, 0)
    return h, s, a @ h


@pytest.mark.parametrize("pipe", ["pair", "single"])
@pytest.mark.parametrize("prec", ["fp16x3", "bf16x3", "fp16", "bf16"])
@pytest.mark.parametrize("N,act,kind", [(1, "relu", "randn"), (100, "gelu", "randn"), (128, "relu", "randn"), (129, "relu", "relu"),
                                        (1024, "gelu", "randn"), (4099, "relu", "randn"), (50000, "relu", "randn")])
def test_fused_forward(K, pipe, prec, N, act, kind):
    sd = cases.abmil_state(100 + N)
    x = cases.make_bag(200 + N, N, 1024, kind)
    h_ref, s_ref, p_ref = run_oracle(sd, x, act)
    c = {k: v.cuda() for k, v in sd.items()}
    Wp = torch.randn(2, 512, generator=torch.Generator().manual_seed(1)) * 0.05
    out = K.abmil_fused_forward(x[0].cuda(), c["feature.0.weight"], c["feature.0.bias"], act, c["attention.0.weight"], c["attention.0.bias"],
                                c["attention.2.weight"], c["attention.2.bias"], "tanh", Wp=Wp.cuda(), want_scores=True, want_h=(N <= 4099),
                                precision=prec, Wcls=c["classifier.weight"], bcls=c["classifier.bias"], pipeline=pipe)
    torch.cuda.synchronize()
    assert cases.rel_err(out["pooled"], p_ref) < TOL[prec]
    assert cases.rel_err(out["s"], s_ref) < TOL[prec] * 3
    assert cases.rel_err(out["t"], h_ref @ Wp.double().t()) < TOL[prec] * 3
    if out["h"] is not None:
        assert cases.rel_err(out["h"], h_ref) < TOL[prec]
    logits = out["pooled"].cpu().double() @ sd["classifier.weight"].double().t() + sd["classifier.bias"].double()
    ref_logits = p_ref @ sd["classifier.weight"].double().t() + sd["classifier.bias"].double()
    assert cases.rel_err(logits, ref_logits) < TOL[prec]
    assert cases.rel_err(out["logits"][0], ref_logits) < TOL[prec]          # classifier fused into the kernel tail
    stats = out["stats"].cpu().double()
    assert abs(float(stats[0]) - float(s_ref.max())) < 3 * TOL[prec] * max(1.0, float(s_ref.abs().max()))
    assert cases.rel_err(stats[1], torch.exp(s_ref - s_ref.max()).sum()) < 30 * TOL[prec]


@pytest.mark.parametrize("pipe", ["pair", "single"])
@pytest.mark.parametrize("prec", ["fp16x3", "bf16x3", "fp16"])
@pytest.mark.parametrize("N,D", [(777, 1536), (300, 96), (150, 32), (2000, 2048)])
def test_fused_forward_other_feature_widths(K, pipe, prec, N, D):
    """D = 1536 (GigaPath, BASELINE config 3), widths that are multiples of 32 but not of 64 (the pair pipeline's 32-wide stages),
    a single k-step, and D = 2048."""
    sd = cases.abmil_state(300 + D, D=D)
    x = cases.make_bag(400 + D, N, D)
    h_ref, s_ref, p_ref = run_oracle(sd, x, "gelu")
    c = {k: v.cuda() for k, v in sd.items()}
    out = K.abmil_fused_forward(x[0].cuda(), c["feature.0.weight"], c["feature.0.bias"], "gelu", c["attention.0.weight"], c["attention.0.bias"],
                                c["attention.2.weight"], c["attention.2.bias"], "tanh", want_scores=True, want_h=True, precision=prec, pipeline=pipe,
                                Wcls=c["classifier.weight"], bcls=c["classifier.bias"])
    torch.cuda.synchronize()
    assert cases.rel_err(out["h"], h_ref) < TOL[prec]
    assert cases.rel_err(out["s"], s_ref) < TOL[prec] * 3
    assert cases.rel_err(out["pooled"], p_ref) < TOL[prec]
    ref_logits = p_ref @ sd["classifier.weight"].double().t() + sd["classifier.bias"].double()
    assert cases.rel_err(out["logits"][0], ref_logits) < TOL[prec]


@pytest.mark.parametrize("pipe", ["pair", "single"])
def test_fused_forward_with_keep_mask(K, pipe):
    N = 3000
    sd = cases.abmil_state(7)
    x = cases.make_bag(8, N, 1024)
    keep = (torch.rand(N, generator=torch.Generator().manual_seed(3)) > 0.2).to(torch.uint8)
    _, s_ref, p_ref = run_oracle(sd, x, "gelu", keep)
    c = {k: v.cuda() for k, v in sd.items()}
    out = K.abmil_fused_forward(x[0].cuda(), c["feature.0.weight"], c["feature.0.bias"], "gelu", c["attention.0.weight"], c["attention.0.bias"],
                                c["attention.2.weight"], c["attention.2.bias"], "tanh", keep=keep.cuda(), want_scores=True, pipeline=pipe)
    assert cases.rel_err(out["pooled"], p_ref) < 1e-4
    sk = out["s"].cpu()
    assert torch.isinf(sk[keep == 0]).all() and cases.rel_err(sk[keep == 1], s_ref[keep == 1]) < 3e-4


@pytest.mark.parametrize("pipe", ["pair", "single"])
@pytest.mark.parametrize("prec", ["fp16x3", "bf16x3", "fp16"])
def test_fused_forward_repeated_launches_are_identical(K, pipe, prec):
    """Pipeline-synchronisation regression: 3 tiles per CTA, cached weight images, 12 back-to-back launches must agree bit for bit
    (a skipped mbarrier phase in the converter groups used to corrupt a few rows of h or dead-lock about once in 20 launches)."""
    N = 50000
    sd = cases.abmil_state(5)
    c = {k: v.cuda() for k, v in sd.items()}
    x = cases.make_bag(9, N, 1024)[0].cuda()
    outs = []
    for _ in range(12):
        o = K.abmil_fused_forward(x, c["feature.0.weight"], c["feature.0.bias"], "relu", c["attention.0.weight"], c["attention.0.bias"],
                                  c["attention.2.weight"], c["attention.2.bias"], "tanh", want_scores=True, precision=prec, pipeline=pipe)
        outs.append((o["pooled"].clone(), o["s"].clone()))
    torch.cuda.synchronize()
    for pooled, s_ in outs[1:]:
        assert torch.equal(pooled, outs[0][0]) and torch.equal(s_, outs[0][1])
    xd = x.double()
    h = torch.relu(xd @ c["feature.0.weight"].double().t() + c["feature.0.bias"].double())
    s_ref = (torch.tanh(h @ c["attention.0.weight"].double().t() + c["attention.0.bias"].double()) @ c["attention.2.weight"].double().t()
             + c["attention.2.bias"].double())[:, 0]
    assert cases.rel_err(outs[0][1], s_ref) < TOL[prec] * 3
    assert cases.rel_err(outs[0][0], torch.softmax(s_ref, 0) @ h) < TOL[prec]


def test_fused_matches_golden_logits(K):
    """Committed golden vector made from the live reference: abmil relu N=1024 (BASELINE config 0 shape)."""
    import os
    G = torch.load(os.path.join(os.path.dirname(__file__), "golden", "golden_v1.pt"), weights_only=False)["abmil"]["relu_1024_randn"]
    sd, x = cases.abmil_state(11), cases.make_bag(1011, 1024, 1024)
    c = {k: v.cuda() for k, v in sd.items()}
    out = K.abmil_fused_forward(x[0].cuda(), c["feature.0.weight"], c["feature.0.bias"], "relu", c["attention.0.weight"], c["attention.0.bias"],
                                c["attention.2.weight"], c["attention.2.bias"], "tanh")
    assert cases.rel_err(out["pooled"], G["pooled"][0]) < 1e-4
    logits = out["pooled"].cpu() @ sd["classifier.weight"].t() + sd["classifier.bias"]
    assert cases.rel_err(logits, G["logits"][0]) < 1e-4


@pytest.mark.parametrize("prec", ["bf16x3", "fp16"])
def test_pair_tail_split_shapes_and_rearm(K, prec):
    """Tail split of the pair pipeline (the tiles of a last wave that fills at most half of the pairs are shared by K range between two pairs
    each; the owner adds the helper's dumped accumulator): split and unsplit shapes, with and without whole waves before the split one, on ONE
    workspace (same weights) in an interleaved order so the flags raised by one launch must have been re-armed for the next, each launch
    twice and bit-identical."""
    sd = cases.abmil_state(41)
    c = {k: v.cuda() for k, v in sd.items()}
    Wp = (torch.randn(2, 512, generator=torch.Generator().manual_seed(2)) * 0.05).cuda()
    # tiles: 1 | 19 | 37 (all split) | 38 (more than half of the 74 pairs: no split) | 74 + 1 | 74 + 30 | 2 x 74 + 21 (split after whole waves)
    sizes = [100, 19 * 128, 37 * 128 - 5, 38 * 128, 75 * 128 - 60, 104 * 128, 169 * 128 + 1, 100, 75 * 128 - 60]
    seen = {}
    for N in sizes:
        x = cases.make_bag(500 + N, N, 1024)
        h_ref, s_ref, p_ref = run_oracle(sd, x, "gelu")
        xg = x[0].cuda()
        outs = []
        for _ in range(2):
            o = K.abmil_fused_forward(xg, c["feature.0.weight"], c["feature.0.bias"], "gelu", c["attention.0.weight"], c["attention.0.bias"],
                                      c["attention.2.weight"], c["attention.2.bias"], "tanh", Wp=Wp, want_scores=True, want_h=True, precision=prec,
                                      pipeline="pair", Wcls=c["classifier.weight"], bcls=c["classifier.bias"])
            outs.append({k: o[k].clone() for k in ("pooled", "s", "h", "t", "logits")})
        torch.cuda.synchronize()
        for k in outs[0]:
            assert torch.equal(outs[0][k], outs[1][k]), (N, k)
        o = outs[0]
        assert cases.rel_err(o["h"], h_ref) < TOL[prec], N
        assert cases.rel_err(o["s"], s_ref) < TOL[prec] * 3, N
        assert cases.rel_err(o["t"], h_ref @ Wp.cpu().double().t()) < TOL[prec] * 3, N
        assert cases.rel_err(o["pooled"], p_ref) < TOL[prec], N
        if N in seen:                                                          # the same bag later in the sequence: same bits
            assert torch.equal(seen[N], o["pooled"])
        seen[N] = o["pooled"]
