"""Seeded synthetic bags and weights for the parity tests.

Weights are produced by an explicit recipe (NOT by the reference's constructors), keyed by the
reference's own state_dict names, so the same tensors can be rebuilt on the GPU box
where /root/reference does not exist.  tests/golden/make_golden.py loads them into the live
reference classes with strict=True, which also proves the key/shape lists below are exact.
"""
import hashlib
import math

import torch


def _g(seed):
    g = torch.Generator(device="cpu")
    g.manual_seed(seed)
    return g


def _mat(g, *shape):
    fan_out, fan_in = shape[0], math.prod(shape[1:])
    return torch.randn(*shape, generator=g) * math.sqrt(2.0 / (fan_in + fan_out))


def _vec(g, n, scale=0.02):
    return torch.randn(n, generator=g) * scale


def _lin(sd, g, key, out_f, in_f, bias=True):
    sd[key + ".weight"] = _mat(g, out_f, in_f)
    if bias:
        sd[key + ".bias"] = _vec(g, out_f)


def _ln(sd, g, key, n):
    sd[key + ".weight"] = 1.0 + 0.1 * torch.randn(n, generator=g)
    sd[key + ".bias"] = 0.1 * torch.randn(n, generator=g)


def make_bag(seed, N, D, kind="randn"):
    x = torch.randn(1, N, D, generator=_g(seed))
    if kind == "relu":          # R50 features are post-ReLU / avg-pool: non-negative
        x = torch.relu(x)
    return x


def abmil_state(seed, D=1024, H=512, Da=128, C=2):
    g, sd = _g(seed), {}
    _lin(sd, g, "feature.0", H, D)
    _lin(sd, g, "attention.0", Da, H)
    _lin(sd, g, "attention.2", 1, Da)
    _lin(sd, g, "classifier", C, H)
    return sd


def abmil_norm_state(seed, mil_norm, embed_norm_pos, D=1024, H=512, Da=128, C=2):
    """abmil.DAttention with mil_norm 'bn' / 'ln' (abmil.py:167-178): the reference's keys for each variant."""
    g, sd = _g(seed), {}
    lin = "feature.0"
    if mil_norm == "ln" and embed_norm_pos == 0:
        _ln(sd, g, "feature.0", D)
        lin = "feature.1"
    _lin(sd, g, lin, H, D)
    _lin(sd, g, "attention.0", Da, H)
    _lin(sd, g, "attention.2", 1, Da)
    _lin(sd, g, "classifier", C, H)
    if mil_norm == "ln":
        if embed_norm_pos == 1:
            _ln(sd, g, "norm", H)
        _ln(sd, g, "norm1", H)
    elif mil_norm == "bn":
        for key, n in (("norm", D if embed_norm_pos == 0 else H), ("norm1", H)):
            _ln(sd, g, key, n)
            sd[key + ".running_mean"] = 0.1 * torch.randn(n, generator=g)
            sd[key + ".running_var"] = 1.0 + 0.2 * torch.rand(n, generator=g)
            sd[key + ".num_batches_tracked"] = torch.tensor(3)
    return sd


def sincos_pos(seed, N, W=97, H=61):
    g = _g(seed)
    xy = torch.stack([torch.randint(0, W, (N,), generator=g), torch.randint(0, H, (N,), generator=g)], 1)
    return torch.cat([torch.tensor([[W, H]]), xy])[None]


def dtfd_state(seed, D=1024, H=512, Da=128, C=2):
    """modules/dtfd.py:149-153: the keys of DTFD(device, lr, wd, steps)."""
    g, sd = _g(seed), {}
    _lin(sd, g, "classifier.fc", C, H)
    for p in ("attention.", "UClassifier.attention."):
        _lin(sd, g, p + "attention_V.0", Da, H)
        _lin(sd, g, p + "attention_U.0", Da, H)
        _lin(sd, g, p + "attention_weights", 1, Da)
    sd["dimReduction.fc1.weight"] = _mat(g, H, D)
    _lin(sd, g, "UClassifier.classifier.fc", C, H)
    return sd


def clam_state(seed, multi_branch, D=1024, H=512, Da=256, C=2, gate=True, fc_drop=False):
    """modules/clam.py:106-132 (CLAM_SB) / :245-274 (CLAM_MB): the attention net sits at index 3 of `attention_net` when the model was built
    with dropout != 0 (a Dropout at index 2), else at 2; its own Dropout(0.25)s then shift the plain net's last Linear to `module.3`."""
    g, sd = _g(seed), {}
    K = C if multi_branch else 1
    a = f"attention_net.{3 if fc_drop else 2}."
    _lin(sd, g, "attention_net.0", H, D)
    if gate:
        _lin(sd, g, a + "attention_a.0", Da, H)
        _lin(sd, g, a + "attention_b.0", Da, H)
        _lin(sd, g, a + "attention_c", K, Da)
    else:
        _lin(sd, g, a + "module.0", Da, H)
        _lin(sd, g, a + f"module.{3 if fc_drop else 2}", K, Da)
    if multi_branch:
        for c in range(C):
            _lin(sd, g, f"classifiers.{c}", 1, H)
    else:
        _lin(sd, g, "classifiers", C, H)
    for c in range(C):
        _lin(sd, g, f"instance_classifiers.{c}", 2, H)
    sd["instance_loss_fn.labels"] = torch.arange(2)
    return sd


def gated_state(seed, D=1024, H=512, Da=384, C=2):
    g, sd = _g(seed), {}
    _lin(sd, g, "feature.0", H, D)
    _lin(sd, g, "attention_a.0", Da, H)
    _lin(sd, g, "attention_b.0", Da, H)
    _lin(sd, g, "attention_c", 1, Da)
    _lin(sd, g, "classifier.0", C, H)
    return sd


def _nystrom_layer(sd, g, p, dim=512, heads=8):
    _ln(sd, g, p + "norm", dim)
    sd[p + "attn.to_qkv.weight"] = _mat(g, 3 * dim, dim)
    _lin(sd, g, p + "attn.to_out.0", dim, dim)
    sd[p + "attn.res_conv.weight"] = torch.randn(heads, 1, 33, 1, generator=g) * 0.05


def _ppeg(sd, g, p, dim=512):
    for name, k in (("proj", 7), ("proj1", 5), ("proj2", 3)):
        sd[p + name + ".weight"] = torch.randn(dim, 1, k, k, generator=g) * (0.5 / k)
        sd[p + name + ".bias"] = _vec(g, dim)


def _dsmil_bag(sd, g, p, H=512, C=2):
    _lin(sd, g, p + "q.0", 128, H)
    _lin(sd, g, p + "q.2", 128, 128)
    _lin(sd, g, p + "v.1", H, H)
    sd[p + "fcc.weight"] = _mat(g, C, C, H)
    sd[p + "fcc.bias"] = _vec(g, C)


def mhim_state(seed, baseline, D=1024, H=512, C=2, merge_k=5):
    g, sd = _g(seed), {}
    _lin(sd, g, "feature.0", H, D)
    gq = (torch.rand(1, merge_k, H, generator=g) * 2 - 1) * 0.0877
    sd["merge.global_q_mm"] = gq
    sd["merge.global_q"] = gq.clone()
    _ln(sd, g, "merge.norm", H)
    sd["merge.attn.to_kv.weight"] = _mat(g, 2 * H, H)
    sd["merge.attn.to_q.weight"] = _mat(g, H, H)
    _lin(sd, g, "merge.attn.to_out.0", H, H)
    e = "online_encoder."
    if baseline == "attn":
        sd[e + "attention.attention.0.weight"] = _mat(g, 128, H)
        sd[e + "attention.attention.2.weight"] = _mat(g, 1, 128)
    elif baseline == "dsmil":
        _lin(sd, g, e + "i_classifier.0", C, H)
        _dsmil_bag(sd, g, e + "b_classifier.", H, C)
    elif baseline == "selfattn":
        _ln(sd, g, e + "norm", H)
        sd[e + "cls_token"] = torch.randn(1, 1, H, generator=g)
        _nystrom_layer(sd, g, e + "layer1.", H)
        _nystrom_layer(sd, g, e + "layer2.", H)
        _ppeg(sd, g, e + "pos_embedding.", H)
    else:
        raise ValueError(baseline)
    _lin(sd, g, "predictor", C, H)
    return sd


def transmil_state(seed, D=1024, H=512, C=2):
    g, sd = _g(seed), {}
    sd["cls_token"] = torch.randn(1, 1, H, generator=g) * 0.02
    _ppeg(sd, g, "pos_layer.", H)
    _lin(sd, g, "feature.0", H, D)
    _nystrom_layer(sd, g, "layer1.", H)
    _nystrom_layer(sd, g, "layer2.", H)
    _ln(sd, g, "norm", H)
    _lin(sd, g, "classifier", C, H)
    return sd


def milnet_state(seed, D=1536, H=512, C=2):
    g, sd = _g(seed), {}
    _lin(sd, g, "feature.0", H, D)
    _lin(sd, g, "i_classifier", C, H)
    _dsmil_bag(sd, g, "b_classifier.", H, C)
    return sd


MHIM_KW = dict(mlp_dim=512, n_classes=2, temp_t=0.1, act="gelu", mask_ratio_h=0.03, mask_ratio_hr=1.0,
               da_act="relu", attn2score=True, merge_enable=True, merge_k=5, merge_mm=0.9999, merge_ratio=0.8)


def tensor_digest(t: torch.Tensor) -> str:
    return hashlib.sha256(t.detach().cpu().contiguous().numpy().tobytes()).hexdigest()


def fingerprint(sd, x) -> float:
    """One number that changes if the seeded RNG stream ever differs between machines."""
    tot = float(x.double().sum())
    for k in sorted(sd):
        tot += float(sd[k].detach().double().abs().sum())
    return tot


def rel_err(a: torch.Tensor, b: torch.Tensor) -> float:
    """max|a-b| / max|b| -- the per-tensor relative error used throughout."""
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).abs().max() / (b.abs().max() + 1e-300))
