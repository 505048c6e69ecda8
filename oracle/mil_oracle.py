"""CPU oracle for the MHIM-MIL per-bag aggregation hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``mhim-mil_b200/`` may import this file; only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference``
legs use it, and only as the checker or the timed CPU baseline.

What it is: an independent restatement, as plain functions over tensors, of what the reference
(DearCaat/MHIM-MIL @ 9d0c91a) computes on the path score -> (masked hard-instance) select ->
attention-weighted pool.  Every function cites the reference file:line it follows.  It is dtype-
generic (fp32 like the reference, or fp64 to measure rounding) and differentiable through
``torch.autograd`` so gradients can be checked too.

Parity pin: the reference ships no tests / golden vectors.  The oracle is
pinned instead against OUTPUTS OF THE LIVE REFERENCE generated in the build container by
``tests/golden/make_golden.py`` (committed next to its output) and, when ``/root/reference``
exists, against the reference classes directly (``tests/test_oracle_vs_reference.py``).

Weights are passed as ``dict[str, Tensor]`` with the reference's own ``state_dict`` keys.
Batch is always 1 (as everywhere in the reference).
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F

Tensor = torch.Tensor
SD = Dict[str, Tensor]


# ----------------------------------------------------------------------------------------------
# small helpers
# ----------------------------------------------------------------------------------------------
def apply_act(x: Tensor, name: Optional[str]) -> Tensor:
    name = (name or "none").lower()
    if name == "relu":
        return torch.relu(x)
    if name == "gelu":
        return F.gelu(x)  # exact erf form, nn.GELU() default
    if name == "tanh":
        return torch.tanh(x)
    if name == "sigmoid":
        return torch.sigmoid(x)
    if name == "none":
        return x
    raise ValueError(name)


def affine(x: Tensor, w: Tensor, b: Optional[Tensor] = None) -> Tensor:
    y = x @ w.transpose(-1, -2)
    return y if b is None else y + b


def layer_norm(x: Tensor, w: Tensor, b: Optional[Tensor], eps: float = 1e-5) -> Tensor:
    mu = x.mean(-1, keepdim=True)
    var = ((x - mu) ** 2).mean(-1, keepdim=True)
    y = (x - mu) / torch.sqrt(var + eps) * w
    return y if b is None else y + b


def softmax_pool(s: Tensor, h: Tensor) -> Tuple[Tensor, Tensor]:
    """softmax over instances of s[L] and the weighted sum of h[L,H].  Returns (p[H], a[L])."""
    a = torch.softmax(s, dim=0)
    return a @ h, a


def pool_partial(s: Tensor, h: Tensor) -> Tuple[Tensor, Tensor, Tensor]:
    """One shard's softmax statistics: m=max s, l=sum e^(s-m), P=sum e^(s-m) h."""
    m = s.max()
    e = torch.exp(s - m)
    return m, e.sum(), e @ h


def merge_partials(ms: Sequence[Tensor], ls: Sequence[Tensor], Ps: Sequence[Tensor]) -> Tuple[Tensor, Tensor, Tensor]:
    """Log-sum-exp merge of shard partials -> (m, l, p) with p already normalised."""
    m = torch.stack(list(ms)).max()
    w = [torch.exp(mi - m) for mi in ms]
    l = sum(li * wi for li, wi in zip(ls, w))
    P = sum(Pi * wi for Pi, wi in zip(Ps, w))
    return m, l, P / l


# ----------------------------------------------------------------------------------------------
# a1 / a2 : plain ABMIL heads
# ----------------------------------------------------------------------------------------------
def sincos_embed(pos: Tensor, C: int) -> Tensor:
    """modules/emb_position.py:9-72: row y * W + x of the 2-D sin/cos table = [sin(x w), cos(x w), sin(y w), cos(y w)] with
    w_k = 10000^(-k / (C/4)).  pos [N+1, 2]: row 0 = (W, H), rows 1.. = (x, y) per instance.  -> [N, C]"""
    xy = pos[1:].to(torch.float32)
    q = C // 4
    omega = 1.0 / (10000 ** (torch.arange(q, dtype=torch.float32) / q))
    ax, ay = xy[:, 0:1] * omega, xy[:, 1:2] * omega
    return torch.cat([torch.sin(ax), torch.cos(ax), torch.sin(ay), torch.cos(ay)], dim=1)


def _batchnorm_rows(x: Tensor, sd: SD, key: str, training: bool, eps: float = 1e-5) -> Tensor:
    """nn.BatchNorm1d over the instances (abmil.py:207-211, 217-221: transposing [1,N,C] -> [1,C,N] only moves the channel axis)."""
    if training:
        mu, var = x.mean(0), x.var(0, unbiased=False)
    else:
        mu, var = sd[key + ".running_mean"], sd[key + ".running_var"]
    return (x - mu) / torch.sqrt(var + eps) * sd[key + ".weight"] + sd[key + ".bias"]


def abmil_dattention(sd: SD, x: Tensor, act: str = "relu", return_attn: bool = False,
                     return_act: bool = False, return_img_feat: bool = False, drop_mask: Optional[Tensor] = None,
                     mil_norm: Optional[str] = None, embed_norm_pos: int = 0, pos: Optional[Tensor] = None, training: bool = False):
    """modules/abmil.py:203-251 (DAttention.forward).  drop_mask [N,512] (keep/(1-p), i.e. what nn.Dropout(0.25) of `feature`
    (:188-189) multiplies by in train mode) or None for dropout off.  mil_norm in (None, 'bn', 'ln') with embed_norm_pos 0 / 1
    (:167-178, 207-223): 'ln' at position 0 is feature.0 (the Linear moves to feature.1); pos [N+1, 2] = sincos coordinates (:214-215).

    feature (:213) -> attention Linear/Tanh/Linear (:229) -> softmax over N (:231-232) ->
    weighted sum (:234) -> norm1 (:237) -> classifier (:238).
    """
    if x.dim() == 2:
        x = x.unsqueeze(0)
    h = x[0]
    lin = "feature.0"
    if mil_norm == "bn" and embed_norm_pos == 0:
        h = _batchnorm_rows(h, sd, "norm", training)
    if mil_norm == "ln" and embed_norm_pos == 0:
        h = layer_norm(h, sd["feature.0.weight"], sd.get("feature.0.bias"))
        lin = "feature.1"
    if lin + ".weight" in sd:
        h = apply_act(affine(h, sd[lin + ".weight"], sd.get(lin + ".bias")), act)
        if drop_mask is not None:
            h = h * drop_mask.to(h.dtype)
    if pos is not None:
        h = h + sincos_embed(pos[0] if pos.dim() == 3 else pos, h.shape[1]).to(h.dtype)
    if embed_norm_pos == 1 and mil_norm == "bn":
        h = _batchnorm_rows(h, sd, "norm", training)
    elif embed_norm_pos == 1 and mil_norm == "ln":
        h = layer_norm(h, sd["norm.weight"], sd.get("norm.bias"))
    u = torch.tanh(affine(h, sd["attention.0.weight"], sd.get("attention.0.bias")))
    s = affine(u, sd["attention.2.weight"], sd.get("attention.2.bias"))[:, 0]
    p, a = softmax_pool(s, h)
    z = p[None]
    if mil_norm == "ln":
        z = layer_norm(z, sd["norm1.weight"], sd.get("norm1.bias"))
    elif mil_norm == "bn":
        z = _batchnorm_rows(z, sd, "norm1", False)           # batch of one bag: only meaningful in eval mode (train mode raises upstream)
    logits = affine(z, sd["classifier.weight"], sd.get("classifier.bias"))
    out = [logits, p[None]] if return_img_feat else logits
    if return_attn:
        res = [out, a[None]]
        if return_act:
            res.append(h[None])
        return res
    return out


def abmil_gated(sd: SD, x: Tensor, act: str = "relu") -> Tensor:
    """modules/abmil.py:111-143 (AttentionGated.forward): tanh branch * sigmoid branch, Da=384."""
    if x.dim() == 2:
        x = x.unsqueeze(0)
    h = apply_act(affine(x[0], sd["feature.0.weight"], sd.get("feature.0.bias")), act)
    ga = torch.tanh(affine(h, sd["attention_a.0.weight"], sd.get("attention_a.0.bias")))
    gb = torch.sigmoid(affine(h, sd["attention_b.0.weight"], sd.get("attention_b.0.bias")))
    s = affine(ga * gb, sd["attention_c.weight"], sd.get("attention_c.bias"))[:, 0]
    p, _ = softmax_pool(s, h)
    return affine(p[None], sd["classifier.0.weight"], sd.get("classifier.0.bias"))


# ----------------------------------------------------------------------------------------------
# a3 : MHIM's pooling heads on pre-embedded h
# ----------------------------------------------------------------------------------------------
def mhim_attention_pool(sd: SD, h: Tensor, da_act: str = "gelu", prefix: str = "online_encoder.",
                        no_norm: bool = False) -> Tuple[Tensor, Tensor]:
    """modules/mhim_modules/baseline.py:8-41,88-110.  Bias-free Linear 512->128, act, Linear 128->1,
    softmax over L, A @ h.  h is [L,512].  Returns (p[512], attn[L]) (raw logits if no_norm, :38-41).
    """
    known = da_act in ("gelu", "relu", "tanh")                 # any other name builds NO activation module (baseline.py:17-22): the
    u = apply_act(affine(h, sd[prefix + "attention.attention.0.weight"]), da_act if known else "none")     # second Linear is then index 1
    s = affine(u, sd[prefix + ("attention.attention.2.weight" if known else "attention.attention.1.weight")])[:, 0]
    p, a = softmax_pool(s, h)
    return p, (s if no_norm else a)


# ----------------------------------------------------------------------------------------------
# a5 : attention -> score
# ----------------------------------------------------------------------------------------------
def pseudo_score(w_pred: Tensor, b_pred: Tensor, h: Tensor, attn: Tensor) -> Tensor:
    """modules/mhim_modules/scoring.py:37-58: score_n = max_c softmax_c(a_n * (h_n . W_c) + b_0).

    Note only bias[0] is added, to every class (:54), so it cancels in the softmax.
    """
    cam = (h * attn[:, None]) @ w_pred.t() + b_pred[0]          # [L,C]
    return torch.softmax(cam, dim=1).max(dim=1).values


def pseudo_score_trans(w_pred: Tensor, b_pred: Tensor, v: Tensor, attn: Tensor,
                       w_out: Tensor, b_out: Tensor) -> Tensor:
    """scoring.py:9-34: per-head v[8,n,64]*attn[8,n] -> [n,512] -> layer1.attn.to_out -> CAM."""
    hds, n, d = v.shape
    f = (v * attn[:, :, None]).permute(1, 0, 2).reshape(n, hds * d)
    f = affine(f, w_out, b_out)
    cam = f @ w_pred.t() + b_pred[0]
    return torch.softmax(cam, dim=1).max(dim=1).values


# ----------------------------------------------------------------------------------------------
# a6 / a7 : masked hard-instance selection
# ----------------------------------------------------------------------------------------------
def topk_count(ps: int, ratio: float) -> int:
    """k exactly as masking.py:61 computes it: python float product, numpy ceil, int()."""
    return int(math.ceil(ps * ratio))


def select_mask(ps: int, attn: Tensor, largest: bool, mask_ratio: float,
                mask_ids_other: Optional[Tensor] = None, len_keep_other: Optional[int] = None,
                topk_idx_other: Optional[Tensor] = None, random_ratio: float = 1.0,
                select_inv: bool = False, msa_fusion: str = "vote") -> Tuple[int, Tensor]:
    """modules/mhim_modules/masking.py:9-88.

    Consumes torch's global RNG exactly where the reference does (randperm, :67).  The kept ids
    are returned ascending (the reference builds them from a python set, :77-80, which iterates
    ascending for the table sizes that occur at MHIM's mask ratios).
    """
    ps_eff = ps
    ratio0 = mask_ratio
    mask_ratio = mask_ratio / random_ratio                                   # :32
    if mask_ratio > 1:                                                       # :33-35
        random_ratio = ratio0
        mask_ratio = 1.0
    if mask_ids_other is not None and topk_idx_other is None:                # :37-40
        topk_idx_other = mask_ids_other[:, len_keep_other:].squeeze()
        ps_eff = ps - topk_idx_other.size(0)

    if attn.dim() > 2:                                                       # multi-head [1,h,N]
        if msa_fusion == "mean":                                             # :44-48
            k = int(math.ceil(ps_eff * mask_ratio) // attn.size(1))
            idx = torch.topk(attn, k, largest=largest).indices
            idx = torch.unique(idx.flatten())
        else:                                                                # vote, :49-59
            k = topk_count(ps_eff, mask_ratio)
            hidx = torch.topk(attn, k=k, sorted=False, largest=largest).indices
            votes = torch.zeros_like(attn).scatter_(2, hidx, 1.0).sum(dim=1)
            idx = torch.topk(votes, k=k, sorted=False).indices[0]
    else:                                                                    # :60-63
        k = topk_count(ps_eff, mask_ratio)
        idx = torch.topk(attn, k, largest=largest).indices.squeeze(0)

    if random_ratio < 1.0:                                                   # :66-71
        perm = torch.randperm(idx.size(0), device=idx.device)
        idx = idx[perm[: int(math.ceil(idx.size(0) * random_ratio))]]

    if mask_ids_other is not None:                                           # :74-75
        idx = torch.cat([idx, topk_idx_other]).unique()

    len_keep = ps - idx.size(0)                                              # :77
    keep_flag = torch.ones(ps, dtype=torch.bool, device=attn.device)
    keep_flag[idx] = False
    kept = torch.nonzero(keep_flag).flatten()                                # ascending complement
    if select_inv:                                                           # :82-84
        return ps - len_keep, torch.cat([idx, kept]).unsqueeze(0)
    return len_keep, torch.cat([kept, idx]).unsqueeze(0)                     # :86


def mask_gather(x: Tensor, mask_ids: Tensor, len_keep: int) -> Tensor:
    """masking.py:91-110: rows of x[1,L,D] at the first len_keep ids."""
    return x[:, mask_ids[0, :len_keep]]


@dataclass
class MHIMConfig:
    """Constructor arguments of modules/mhim.py:22-27 that change the arithmetic."""
    input_dim: int = 1024
    mlp_dim: int = 512
    mask_ratio: float = 0.0
    n_classes: int = 2
    temp_t: float = 1.0
    act: str = "relu"
    mask_ratio_h: float = 0.0
    mrh_sche: Optional[Sequence[float]] = None
    mask_ratio_hr: float = 0.0
    mask_ratio_l: float = 0.0
    da_act: str = "gelu"
    baseline: str = "selfattn"
    head: int = 8
    attn2score: bool = True
    merge_enable: bool = True
    merge_k: int = 1
    merge_mm: float = 0.9998
    merge_ratio: float = 0.0
    merge_test: bool = False


def mhim_get_mask(cfg: MHIMConfig, ps: int, i: Optional[int], attn: Optional[Tensor],
                  mrh: Optional[float] = None) -> Tuple[int, Optional[Tensor]]:
    """modules/mhim.py:109-179 (select_inv=False, msa_fusion='vote' as set at :59-60)."""
    len_keep, ids = ps, None
    if attn is not None and cfg.mask_ratio > 0.0:                            # :124-128
        len_keep, ids = select_mask(ps, attn, False, cfg.mask_ratio, random_ratio=0.001)
    if attn is not None and cfg.mask_ratio_l > 0.0:                          # :133-149
        if ids is None:
            len_keep, ids = select_mask(ps, attn, False, cfg.mask_ratio_l)
        else:
            other = ids[:, len_keep:].squeeze()
            len_keep, ids = select_mask(ps, attn, False, cfg.mask_ratio_l, mask_ids_other=ids,
                                        len_keep_other=ps, topk_idx_other=other)
    r_h = cfg.mask_ratio_h                                                   # :152-156
    if cfg.mrh_sche is not None:
        r_h = cfg.mrh_sche[i]
    if mrh is not None:
        r_h = mrh
    if r_h > 0.0:                                                            # :158-177
        if ids is None:
            len_keep, ids = select_mask(ps, attn, True, r_h, len_keep_other=ps,
                                        random_ratio=cfg.mask_ratio_hr)
        else:
            other = ids[:, len_keep:].squeeze()
            len_keep, ids = select_mask(ps, attn, True, r_h, mask_ids_other=ids, len_keep_other=ps,
                                        topk_idx_other=other, random_ratio=cfg.mask_ratio_hr)
    return len_keep, ids


# ----------------------------------------------------------------------------------------------
# a8 : Merge / MCA
# ----------------------------------------------------------------------------------------------
def mca(sd: SD, x: Tensor, q_in: Tensor, prefix: str = "merge.attn.", heads: int = 8) -> Tensor:
    """modules/mhim_modules/merge.py:43-65 with dropout off.  x [n,512], q_in [k,512] -> [k,512]."""
    kv = affine(x, sd[prefix + "to_kv.weight"])
    inner = kv.shape[-1] // 2
    dh = inner // heads
    kk, vv = kv[:, :inner], kv[:, inner:]
    q = affine(q_in, sd[prefix + "to_q.weight"])
    split = lambda t: t.reshape(t.shape[0], heads, dh).permute(1, 0, 2)      # [h, n, d]
    qh, kh, vh = split(q), split(kk), split(vv)
    dots = qh @ kh.transpose(-1, -2) * dh ** -0.5
    out = torch.softmax(dots, dim=-1) @ vh                                   # [h, k, d]
    out = out.permute(1, 0, 2).reshape(q_in.shape[0], inner)
    return affine(out, sd[prefix + "to_out.0.weight"], sd[prefix + "to_out.0.bias"])


class _LayerNormInputReadAtBackward(torch.autograd.Function):
    """LayerNorm whose backward reads its INPUT at backward time (mean / rstd from the forward), which is what aten's layer_norm
    backward does with its saved input.  It matters for exactly one tensor of the path: the reference updates `global_q_mm` through
    `.data` (merge.py:127-129; same storage as `global_q`) inside forward(), AFTER LayerNorm(global_q) has saved it and BEFORE
    backward() runs -- autograd's version counter does not see a `.data` write -- so the reference's gradient of merge.norm.weight
    (and of global_q) is evaluated with the already-updated tokens: 2e-4 away from the mathematically clean value at mm = 0.9999.
    A drop-in follows the reference (it runs the same torch LayerNorm in the same order); so does this restatement."""

    @staticmethod
    def forward(ctx, x, w, b, read_input, eps):
        mu = x.mean(-1, keepdim=True)
        rstd = 1.0 / torch.sqrt(((x - mu) ** 2).mean(-1, keepdim=True) + eps)
        ctx.save_for_backward(w, mu, rstd)
        ctx.read_input, ctx.has_b = read_input, b is not None
        y = (x - mu) * rstd * w
        return y if b is None else y + b

    @staticmethod
    def backward(ctx, dy):
        w, mu, rstd = ctx.saved_tensors
        xh = (ctx.read_input().to(dy.dtype) - mu) * rstd
        g = dy * w
        dx = rstd * (g - g.mean(-1, keepdim=True) - xh * (g * xh).mean(-1, keepdim=True))
        return dx, (dy * xh).sum(0), (dy.sum(0) if ctx.has_b else None), None, None


def merge_tokens(sd: SD, x: Tensor, prefix: str = "merge.") -> Tensor:
    """merge.py:131-144: MCA(LN(x), LN(global_q)) -> [k,512].  The EMA side effect itself is applied by merge_forward."""
    nw, nb = sd[prefix + "norm.weight"], sd[prefix + "norm.bias"]
    holder = sd[prefix + "global_q"]
    gq = holder[0]
    if torch.is_grad_enabled() and (nw.requires_grad or holder.requires_grad):
        qn = _LayerNormInputReadAtBackward.apply(gq, nw, nb, lambda: holder.data[0], 1e-5)
    else:
        qn = layer_norm(gq, nw, nb)
    return mca(sd, layer_norm(x, nw, nb), qn, prefix + "attn.")


def merge_forward(sd: SD, x: Tensor, merge_ratio: float, training: bool, mm: float,
                  prefix: str = "merge.") -> Tuple[Tensor, Optional[Tensor]]:
    """merge.py:146-203 (mask_type='random').  x [L,512].

    training: ids = argsort(rand(L)) (:163-165); first int(L*merge_ratio) kept in that random order
    (:171-174), the rest merged into k tokens (:141); returns (cat(x_keep, z), new_global_q) where
    new_global_q is the EMA the reference writes into global_q_mm (:127-129).
    eval: cat(x, merge(x)) (:199).
    """
    if training:
        L = x.shape[0]
        n_keep = int(L * merge_ratio)
        ids = torch.argsort(torch.rand(L, device=x.device), dim=0)
        z = merge_tokens(sd, x[ids[n_keep:]], prefix)
        gq = sd[prefix + "global_q"]
        new_q = gq * mm + z.detach()[None].to(gq.dtype) * (1.0 - mm) if mm != 1.0 else None
        if new_q is not None and torch.is_grad_enabled() and (gq.requires_grad or sd[prefix + "norm.weight"].requires_grad):
            gq.data.copy_(new_q.detach())      # the reference's in-forward `.data` write (merge.py:127-129): backward sees the new tokens
        return torch.cat([x[ids[:n_keep]], z], dim=0), (new_q.detach() if new_q is not None and new_q.requires_grad else new_q)
    return torch.cat([x, merge_tokens(sd, x, prefix)], dim=0), None


# ----------------------------------------------------------------------------------------------
# a13 : DSMIL
# ----------------------------------------------------------------------------------------------
def dsmil_bag(sd: SD, feats: Tensor, classes: Tensor, prefix: str, no_norm: bool = False,
              v_has_dropout_slot: bool = True):
    """dsmil.py:85-109 / baseline.py:131-152.  feats [N,512], classes [N,C] -> (pred[1,C], A[N,C], B[C,512]).

    Critical instance per class = argmax over N (the reference sorts, :91 / :137, only to take row 0).
    """
    vkey = "v.1" if v_has_dropout_slot else "v.0"
    V = torch.relu(affine(feats, sd[prefix + vkey + ".weight"], sd.get(prefix + vkey + ".bias")))
    qnet = lambda t: torch.tanh(affine(torch.relu(affine(t, sd[prefix + "q.0.weight"], sd.get(prefix + "q.0.bias"))),
                                       sd[prefix + "q.2.weight"], sd[prefix + "q.2.bias"]))
    Q = qnet(feats)
    crit = torch.sort(classes, 0, descending=True).indices[0]               # [C]
    q_max = qnet(feats[crit])
    logit = (Q @ q_max.t()) / math.sqrt(Q.shape[1])
    A = torch.softmax(logit, dim=0)
    B = A.t() @ V                                                            # [C,512]
    w, b = sd[prefix + "fcc.weight"], sd.get(prefix + "fcc.bias")           # [C,C,512]
    pred = torch.einsum("ock,ck->o", w, B)
    if b is not None:
        pred = pred + b
    return pred[None], (logit if no_norm else A), B


def milnet_forward(sd: SD, x: Tensor, act: str = "relu"):
    """dsmil.py:142-172 (MILNet.forward), eval branch: (prediction_bag[1,C], max-instance logits[1,C])."""
    h = apply_act(affine(x[0], sd["feature.0.weight"], sd.get("feature.0.bias")), act)
    classes = affine(h, sd["i_classifier.weight"], sd.get("i_classifier.bias"))
    pred, A, B = dsmil_bag(sd, h, classes, "b_classifier.")
    return pred, classes.max(dim=0).values[None], A, B


def mhim_dsmil_encoder(sd: SD, h: Tensor, cls_attn: bool = True, return_attn: bool = False,
                       no_norm: bool = False, prefix: str = "online_encoder."):
    """baseline.py:166-194 (DSMIL.attention/forward) on h [L,512]."""
    classes = affine(h, sd[prefix + "i_classifier.0.weight"], sd[prefix + "i_classifier.0.bias"])
    pred, A, B = dsmil_bag(sd, h, classes, prefix + "b_classifier.", no_norm=no_norm)
    inst = classes.max(dim=0).values[None]
    attn = None
    if return_attn:
        attn = (classes.max(dim=-1).values if cls_attn else A.max(dim=-1).values)[None]
    return [pred, inst], B[None], attn


# ----------------------------------------------------------------------------------------------
# a11 / a12 : Nystrom attention, TransMIL, SAttention
# ----------------------------------------------------------------------------------------------
def pinv_iter(x: Tensor, iters: int = 6) -> Tensor:
    """nystrom_attention.py:12-27.  x [h,m,m]; ONE global scalar normaliser over all heads (:18)."""
    ax = x.abs()
    z = x.transpose(-1, -2) / (ax.sum(-1).max() * ax.sum(-2).max())
    eye = torch.eye(x.shape[-1], dtype=x.dtype, device=x.device)[None]
    for _ in range(iters):
        xz = x @ z
        z = 0.25 * z @ (13 * eye - xz @ (15 * eye - xz @ (7 * eye - xz)))
    return z


def nystrom_attention(sd: SD, x: Tensor, prefix: str, heads: int = 8, m: int = 256, iters: int = 6,
                      return_attn: bool = False, no_norm: bool = False):
    """nystrom_attention.py:65-152, dropout off, attn_mask=None.  x [n,dim] -> out [n,dim]
    (+ cls-row attention [h,n-1] and v[h,n-1,dh] when return_attn)."""
    n, dim = x.shape
    pad = (m - n % m) % m                                                    # :70-73 front zero-pad
    if pad:
        x = torch.cat([x.new_zeros(pad, dim), x], dim=0)
    npad = x.shape[0]
    qkv = affine(x, sd[prefix + "to_qkv.weight"])
    inner = qkv.shape[-1] // 3
    dh = inner // heads
    split = lambda t: t.reshape(npad, heads, dh).permute(1, 0, 2)
    q, k, v = (split(qkv[:, i * inner:(i + 1) * inner]) for i in range(3))
    q = q * dh ** -0.5                                                       # :90
    l = math.ceil(n / m)                                                     # :94
    q_l = q.reshape(heads, m, l, dh).sum(2) / l                              # :95-109
    k_l = k.reshape(heads, m, l, dh).sum(2) / l
    s1 = q @ k_l.transpose(-1, -2)                                           # :114-116
    s2 = q_l @ k_l.transpose(-1, -2)
    s3 = q_l @ k.transpose(-1, -2)
    a1, a2, a3 = s1.softmax(-1), s2.softmax(-1), s3.softmax(-1)              # :130
    a2i = pinv_iter(a2, iters)                                               # :131
    out = (a1 @ a2i) @ (a3 @ v)                                              # :132
    rw = sd.get(prefix + "res_conv.weight")                                  # [h,1,ks,1]
    if rw is not None:                                                       # :135-136
        ks = rw.shape[2]
        out = out + F.conv2d(v[None], rw, padding=(ks // 2, 0), groups=heads)[0]
    out = out.permute(1, 0, 2).reshape(npad, inner)                          # :140
    out = affine(out, sd[prefix + "to_out.0.weight"], sd[prefix + "to_out.0.bias"])[-n:]
    if not return_attn:
        return out
    if no_norm:                                                              # :127-129,146-148
        r = (s1[:, -n][:, None] @ pinv_iter(s2, iters)) @ s3
    else:                                                                    # :144-145
        r = (a1[:, -n][:, None] @ a2i) @ a3
    return out, r[:, 0, -n + 1:], v[:, -n + 1:]


def trans_layer(sd: SD, x: Tensor, prefix: str, heads: int = 8, need_attn: bool = False, no_norm: bool = False):
    """transmil.py:38-48 / baseline.py:210-220: x + Nystrom(LayerNorm(x)), m = dim//2, 6 pinv iters."""
    dim = x.shape[-1]
    y = layer_norm(x, sd[prefix + "norm.weight"], sd[prefix + "norm.bias"])
    r = nystrom_attention(sd, y, prefix + "attn.", heads=heads, m=dim // 2, return_attn=need_attn, no_norm=no_norm)
    if need_attn:
        return x + r[0], r[1], r[2]
    return x + r


def _ppeg_convs(sd: SD, grid: Tensor, prefix: str) -> Tensor:
    c = grid.shape[1]
    y = grid
    for name, ks in (("proj", sd[prefix + "proj.weight"].shape[-1]), ("proj1", 5), ("proj2", 3)):
        y = y + F.conv2d(grid, sd[prefix + name + ".weight"], sd.get(prefix + name + ".bias"),
                         padding=ks // 2, groups=c)
    return y


def ppeg_transmil(sd: SD, x: Tensor, H: int, W: int, prefix: str = "pos_layer.") -> Tensor:
    """transmil.py:57-64: x [1+H*W, C]; cls token passes through, the rest goes through 7/5/3 depthwise convs."""
    cls, tok = x[:1], x[1:]
    grid = tok.t().reshape(1, -1, H, W)
    y = _ppeg_convs(sd, grid, prefix)
    return torch.cat([cls, y.flatten(2)[0].t()], dim=0)


def ppeg_selfpad(sd: SD, x: Tensor, prefix: str) -> Tensor:
    """emb_position.py:92-120: wrap-pad to a square (min 7x7, zero-filled), convs, trim back. x [N,C]."""
    N, C = x.shape
    H = W = int(math.ceil(math.sqrt(N)))
    add = H * W - N
    x = torch.cat([x, x[:add]], dim=0)
    if H < 7:
        H = W = 7
        zp = H * W - (N + add)
        x = torch.cat([x, x.new_zeros(zp, C)], dim=0)
        add += zp
    y = _ppeg_convs(sd, x.t().reshape(1, C, H, W), prefix).flatten(2)[0].t()
    return y[:-add] if add > 0 else y


def transmil_forward(sd: SD, x: Tensor, act: str = "relu", heads: int = 8, return_attn: bool = False,
                     return_act: bool = False):
    """transmil.py:110-175 with dropout off, mil_norm=None, pos='ppeg'."""
    h = apply_act(affine(x[0], sd["feature.0.weight"], sd.get("feature.0.bias")), act)
    n0 = h.shape[0]
    side = int(math.ceil(math.sqrt(n0)))
    add = side * side - n0
    h = torch.cat([h, h[:add]], dim=0)                                       # :124-127
    h = torch.cat([sd["cls_token"][0], h], dim=0)                            # :130-132
    attn: List[Tensor] = []
    v = None
    if return_attn:
        h, a, v = trans_layer(sd, h, "layer1.", heads, need_attn=True)
        attn.append((a[:, :-add] if add > 0 else a)[None])
    else:
        h = trans_layer(sd, h, "layer1.", heads)
    h = ppeg_transmil(sd, h, side, side)
    if return_attn:
        h, a, _ = trans_layer(sd, h, "layer2.", heads, need_attn=True)
        attn.append((a[:, :-add] if add > 0 else a)[None])
    else:
        h = trans_layer(sd, h, "layer2.", heads)
    cls = layer_norm(h, sd["norm.weight"], sd["norm.bias"])[:1]
    logits = affine(cls, sd["classifier.weight"], sd.get("classifier.bias"))
    if return_attn:
        out = [logits, attn]
        if return_act:
            out.append(v[None])
        return out
    return logits


def sattention_forward(sd: SD, h: Tensor, heads: int = 8, return_attn: bool = False, return_act: bool = False,
                       no_norm: bool = False, prefix: str = "online_encoder."):
    """baseline.py:244-288 (pos_pos=0, pos='ppeg').  h [L,512] -> cls feature [1,512]."""
    x = torch.cat([sd[prefix + "cls_token"][0], h], dim=0)
    attn: List[Tensor] = []
    v = None
    if return_attn:
        x, a, v = trans_layer(sd, x, prefix + "layer1.", heads, need_attn=True, no_norm=no_norm)
        attn.append(a[None])
    else:
        x = trans_layer(sd, x, prefix + "layer1.", heads)
    x = torch.cat([x[:1], ppeg_selfpad(sd, x[1:], prefix + "pos_embedding.")], dim=0)   # :265-266
    if return_attn:
        x, a, _ = trans_layer(sd, x, prefix + "layer2.", heads, need_attn=True, no_norm=no_norm)
        attn.append(a[None])
    else:
        x = trans_layer(sd, x, prefix + "layer2.", heads)
    cls = layer_norm(x, sd[prefix + "norm.weight"], sd[prefix + "norm.bias"])[:1]
    if return_attn:
        out = [cls, attn]
        if return_act:
            out.append(v[None])
        return out
    return cls


# ----------------------------------------------------------------------------------------------
# a4 / a9 / a10 : MHIM entry points
# ----------------------------------------------------------------------------------------------
def soft_target_ce(student: Tensor, teacher: Tensor, temp_t: float, temp_s: float = 1.0) -> Tensor:
    """losses.py:26-44: mean over rows of -softmax(t/Tt) . log_softmax(s/Ts)."""
    return (-(torch.softmax(teacher / temp_t, -1) * torch.log_softmax(student / temp_s, -1)).sum(-1)).mean()


def mhim_embed(sd: SD, x: Tensor, act: str, drop_mask: Optional[Tensor] = None) -> Tensor:
    """mhim.py:193-194/244-245/285-286/335-336: feature = Linear(D->512)+act on every row, then `self.dp`: drop_mask [N,512]
    holds keep/(1-p) (what nn.Dropout multiplies by in train mode; pinned against the live reference under a fixed torch seed in
    tests/test_oracle_vs_reference.py), None = dropout off / eval."""
    h = apply_act(affine(x[0], sd["feature.0.weight"], sd["feature.0.bias"]), act)
    return h if drop_mask is None else h * drop_mask.to(h.dtype)


def _encode(cfg: MHIMConfig, sd: SD, h: Tensor, return_attn=False, return_act=False, no_norm=False):
    if cfg.baseline == "attn":
        p, a = mhim_attention_pool(sd, h, cfg.da_act, no_norm=no_norm)
        if return_attn:
            out = [p[None], a[None]]
            if return_act:
                out.append(h[None])
            return out
        return p[None]
    if cfg.baseline == "selfattn":
        return sattention_forward(sd, h, cfg.head, return_attn, return_act, no_norm)
    if cfg.baseline == "dsmil":
        lg, B, attn = mhim_dsmil_encoder(sd, h, cls_attn=cfg.attn2score, return_attn=return_attn, no_norm=no_norm)
        return (lg, B, attn) if return_attn else (lg, B)
    raise ValueError(cfg.baseline)


def mhim_forward_teacher(cfg: MHIMConfig, sd: SD, x: Tensor, drop_mask: Optional[Tensor] = None):
    """mhim.py:181-227 (merge_test=False): returns (cls_feat, score)."""
    h = mhim_embed(sd, x, cfg.act, drop_mask)
    if cfg.baseline == "dsmil":                                              # :202-205
        _, B, attn = _encode(cfg, sd, h, return_attn=True)
        return B, attn
    feat, attn, act = _encode(cfg, sd, h, return_attn=True, return_act=True)
    if cfg.attn2score:                                                       # :215-222
        wp, bp = sd["predictor.weight"], sd["predictor.bias"]
        if cfg.baseline == "selfattn":
            p = "online_encoder.layer1.attn.to_out.0."
            score = pseudo_score_trans(wp, bp, act[0], attn[0][0], sd[p + "weight"], sd[p + "bias"])[None]
        else:
            score = pseudo_score(wp, bp, act[0], attn[0])[None]
        return feat, score
    if isinstance(attn, (list, tuple)):                                      # :224-225
        attn = attn[0]
    return feat, attn


def mhim_forward_test(cfg: MHIMConfig, sd: SD, x: Tensor):
    """mhim.py:229-272 with return_attn=False."""
    h = mhim_embed(sd, x, cfg.act)
    if cfg.merge_test:
        h, _ = merge_forward(sd, h, cfg.merge_ratio, False, cfg.merge_mm)
    if cfg.baseline == "dsmil":
        return _encode(cfg, sd, h)
    return affine(_encode(cfg, sd, h), sd["predictor.weight"], sd["predictor.bias"])


def mhim_pure(cfg: MHIMConfig, sd: SD, x: Tensor):
    """mhim.py:274-298: no masking, no merging."""
    h = mhim_embed(sd, x, cfg.act)
    if cfg.baseline == "dsmil":
        return _encode(cfg, sd, h)[0]
    return affine(_encode(cfg, sd, h), sd["predictor.weight"], sd["predictor.bias"])


def mhim_forward(cfg: MHIMConfig, sd: SD, x: Tensor, attn: Tensor, teacher_cls_feat: Optional[Tensor],
                 i: Optional[int] = None, training: bool = True, drop_mask: Optional[Tensor] = None):
    """mhim.py:318-378: the student pass.  Returns (logits, cls_loss, ps, len_keep, new_global_q, mask_ids)."""
    h = mhim_embed(sd, x, cfg.act, drop_mask)
    ps = h.shape[0]
    len_keep, ids = mhim_get_mask(cfg, ps, i, attn)                          # :341
    h = h[ids[0, :len_keep]]                                                 # :342
    h, new_q = merge_forward(sd, h, cfg.merge_ratio, training, cfg.merge_mm)  # :351
    len_keep2 = h.shape[0]
    if cfg.baseline == "dsmil":                                              # :355-364
        logit, feat = _encode(cfg, sd, h)
    else:
        feat = _encode(cfg, sd, h)
        logit = affine(feat, sd["predictor.weight"], sd["predictor.bias"])
    loss = soft_target_ce(feat, teacher_cls_feat.detach(), cfg.temp_t) if teacher_cls_feat is not None else 0.0
    return logit, loss, ps, len_keep2, new_q, ids


# ----------------------------------------------------------------------------------------------
# f-3 : DTFD-MIL (modules/dtfd.py)
# ----------------------------------------------------------------------------------------------
def dtfd_gated_logits(sd: SD, x: Tensor, prefix: str) -> Tensor:
    """dtfd.py:135-139 (Attention.forward before the softmax): w . (tanh(V x) * sigmoid(U x)) + b -> [N]"""
    av = torch.tanh(affine(x, sd[prefix + "attention_V.0.weight"], sd[prefix + "attention_V.0.bias"]))
    au = torch.sigmoid(affine(x, sd[prefix + "attention_U.0.weight"], sd[prefix + "attention_U.0.bias"]))
    return affine(av * au, sd[prefix + "attention_weights.weight"], sd[prefix + "attention_weights.bias"])[:, 0]


def dtfd_forward(sd: SD, x: Tensor, training: bool, group: int = 5, distill: str = "AFS", act: str = "relu",
                 test_ids: Optional[Sequence[int]] = None) -> Tensor:
    """dtfd.py:168-272 (DTFD.train_forward / test_forward) with every dropout off.  x [N, D] -> slide prediction [1, C].
    training: contiguous pseudo-bags (np.array_split of range(N), :176-178); eval: pseudo-bags from the shuffled ids `test_ids`
    (:232-235; pass the permutation python's `random.shuffle` produced).  distill in ('AFS', 'MaxS', 'MaxMinS')."""
    import numpy as np
    n = x.shape[0]
    mid = apply_act(affine(x, sd["dimReduction.fc1.weight"]), act)                               # :81-83
    ids = list(range(n)) if training else list(test_ids)
    chunks = [torch.as_tensor(c, dtype=torch.long) for c in np.array_split(np.array(ids), group)]
    s_all = dtfd_gated_logits(sd, mid, "attention.")
    feats = []
    for idx in chunks:
        h = mid[idx]
        a = torch.softmax(s_all[idx], 0)                                                         # :191 / :239-240
        att = h * a[:, None]                                                                     # :193 tattFeats
        pooled = att.sum(0, keepdim=True)
        if distill == "AFS":
            feats.append(pooled)
            continue
        cam = att @ sd["classifier.fc.weight"].t()                                               # get_cam_1d: no bias (:29-32)
        order = torch.sort(torch.softmax(cam, 1)[:, -1], descending=True).indices                # :200-202
        sel = order[:1] if distill == "MaxS" else torch.cat([order[:1], order[-1:]])
        feats.append(h[sel])
    pseudo = torch.cat(feats, 0)
    a2 = torch.softmax(dtfd_gated_logits(sd, pseudo, "UClassifier.attention."), 0)
    return affine((a2[None] @ pseudo), sd["UClassifier.classifier.fc.weight"], sd["UClassifier.classifier.fc.bias"])


def smooth_top1_svm(x: Tensor, y: Tensor, tau: float = 1.0, alpha: float = 1.0, thresh: float = 1e3) -> Tensor:
    """SmoothTop1SVM(n_classes).forward (modules/topk/svm.py:84-108; functional.py:9-17 hard, :35-42 smooth; utils.py:8-20 delta, :36-42
    detect_large; polynomial/sp.py:100-106 log_sum_exp).  x [n, C] logits, y [n] int64 targets -> scalar."""
    import math
    n, C = x.shape
    top = x.topk(2, 1).values
    hard = ((top[:, 0] - top[:, 1]) >= 1 * tau * math.log(thresh)).detach()
    smooth = ~hard
    delta = alpha * (y[:, None] != torch.arange(C, device=x.device)[None, :]).to(x.dtype)
    loss = x.new_zeros(())
    if bool(smooth.any()):
        xs = x[smooth] + delta[smooth] - x[smooth].gather(1, y[smooth][:, None])
        xs = xs / tau
        mx = xs.max(1).values
        loss = loss + (tau * (mx + torch.log(torch.exp(xs - mx[:, None]).sum(1)))).sum() / n
    if bool(hard.any()):
        xh = x[hard]
        loss = loss + ((xh + delta[hard]).max(1).values - xh.gather(1, y[hard][:, None]).squeeze(1)).sum() / n
    return loss


def clam_attention_logits(sd: SD, h: Tensor, gate: bool, fc_drop: bool, drop_a: Optional[Tensor] = None, drop_b: Optional[Tensor] = None) -> Tensor:
    """Attn_Net / Attn_Net_Gated of clam.py:31-80 on h [N, 512] -> raw attention logits [N, K].  The attention net is the last entry of
    `attention_net` (index 2 without, 3 with the fc Dropout, clam.py:106-126); drop_a / drop_b: pre-scaled keep masks of its Dropout(0.25)s."""
    pre = f"attention_net.{3 if fc_drop else 2}."
    if gate:
        a = torch.tanh(affine(h, sd[pre + "attention_a.0.weight"], sd.get(pre + "attention_a.0.bias")))
        b = torch.sigmoid(affine(h, sd[pre + "attention_b.0.weight"], sd.get(pre + "attention_b.0.bias")))
        if drop_a is not None:
            a, b = a * drop_a, b * drop_b
        return affine(a * b, sd[pre + "attention_c.weight"], sd.get(pre + "attention_c.bias"))
    a = torch.tanh(affine(h, sd[pre + "module.0.weight"], sd[pre + "module.0.bias"]))
    if drop_a is not None:
        a = a * drop_a
    last = 3 if drop_a is not None or (pre + "module.3.weight") in sd else 2
    return affine(a, sd[pre + f"module.{last}.weight"], sd[pre + f"module.{last}.bias"])


def clam_forward(sd: SD, x: Tensor, multi_branch: bool, n_classes: int = 2, gate: bool = True, act: str = "relu", k_sample: int = 8,
                 subtyping: bool = False, label: Optional[int] = None, fc_drop: bool = False, drop_h: Optional[Tensor] = None,
                 drop_a: Optional[Tensor] = None, drop_b: Optional[Tensor] = None):
    """CLAM_SB.forward (clam.py:173-241) / CLAM_MB.forward (:278-331) for one bag x [N, D].  Returns (logits [1, C], instance_loss or None,
    A_raw [K, N]).  label: the bag's class index -> the instance-level branch (:186-209 / :293-311) runs, else it is skipped.
    fc_drop: the model was built with dropout != 0 (shifts the attention net's index); drop_*: pre-scaled keep masks (None = eval)."""
    h = apply_act(affine(x, sd["attention_net.0.weight"], sd.get("attention_net.0.bias")), act)
    if drop_h is not None:
        h = h * drop_h
    A_raw = clam_attention_logits(sd, h, gate, fc_drop, drop_a, drop_b).t()                         # [K, N]
    A = torch.softmax(A_raw, dim=1)
    inst_loss = None
    if label is not None:
        inst_loss = x.new_zeros(())
        for i in range(n_classes):
            Wi, bi = sd[f"instance_classifiers.{i}.weight"], sd[f"instance_classifiers.{i}.bias"]
            Ai = A[i] if multi_branch else A[0]
            if i == label:                                                                         # in-the-class (:137-154)
                top_p = torch.topk(Ai, k_sample).indices
                top_n = torch.topk(-Ai, k_sample).indices
                logits_i = affine(torch.cat([h[top_p], h[top_n]], 0), Wi, bi)
                tgt = torch.cat([torch.ones(k_sample, dtype=torch.long, device=h.device), torch.zeros(k_sample, dtype=torch.long, device=h.device)])
            elif subtyping:                                                                        # out-of-the-class (:157-167)
                top_p = torch.topk(Ai, k_sample).indices
                logits_i = affine(h[top_p], Wi, bi)
                tgt = torch.zeros(k_sample, dtype=torch.long, device=h.device)
            else:
                continue
            inst_loss = inst_loss + smooth_top1_svm(logits_i, tgt)
        if subtyping:
            inst_loss = inst_loss / n_classes
    M = A @ h                                                                                      # [K, 512]
    if multi_branch:
        logits = torch.stack([affine(M[c:c + 1], sd[f"classifiers.{c}.weight"], sd[f"classifiers.{c}.bias"])[0, 0] for c in range(n_classes)])[None]
    else:
        logits = affine(M, sd["classifiers.weight"], sd.get("classifiers.bias")).max(dim=0).values[None]   # [1, K = 1, C].max(dim=1)
    return logits, inst_loss, A_raw


# ----------------------------------------------------------------------------------------------
# Analytic ABMIL backward -- what the streaming backward kernel implements.
# ----------------------------------------------------------------------------------------------
def abmil_backward_analytic(x: Tensor, W1: Tensor, b1: Tensor, Wa: Tensor, ba: Optional[Tensor], wc: Tensor,
                            act: str, g_p: Tensor, Wb: Optional[Tensor] = None, bb: Optional[Tensor] = None):
    """Gradients of p = softmax_N(s) @ h wrt the weights, given g_p = dL/dp.  x [N,D]; wc [Da].

    Autograd of abmil.py:213-234 written out: g_s = a (h.g_p - p.g_p); g_h = a g_p + Wa^T g_u (+ Wb^T g_v);
    g_pre = g_h * act'(pre); dW1 = g_pre^T x, etc.
    """
    pre = affine(x, W1, b1)
    h = apply_act(pre, act)
    ua = affine(h, Wa, ba)
    t = torch.tanh(ua)
    if Wb is not None:
        vb = affine(h, Wb, bb)
        sg = torch.sigmoid(vb)
        gate = t * sg
    else:
        gate = t
    s = gate @ wc
    a = torch.softmax(s, 0)
    p = a @ h
    g_s = a * (h @ g_p - p @ g_p)
    g_gate = g_s[:, None] * wc[None]
    if Wb is not None:
        g_u = g_gate * sg * (1 - t * t)
        g_v = g_gate * t * sg * (1 - sg)
    else:
        g_u = g_gate * (1 - t * t)
        g_v = None
    g_h = a[:, None] * g_p[None] + g_u @ Wa
    if Wb is not None:
        g_h = g_h + g_v @ Wb
    if act == "relu":
        d = (pre > 0).to(pre.dtype)
    elif act == "gelu":
        d = 0.5 * (1 + torch.erf(pre / math.sqrt(2))) + pre * torch.exp(-0.5 * pre * pre) / math.sqrt(2 * math.pi)
    else:
        d = torch.ones_like(pre)
    g_pre = g_h * d
    out = {"W1": g_pre.t() @ x, "b1": g_pre.sum(0), "Wa": g_u.t() @ h, "ba": g_u.sum(0), "wc": g_s @ gate}
    if Wb is not None:
        out["Wb"] = g_v.t() @ h
        out["bb"] = g_v.sum(0)
    return out


# ----------------------------------------------------------------------------------------------
# f-1 : EMA teacher update
# ----------------------------------------------------------------------------------------------
def ema_update(params_q, params_k, mm: float):
    """engines/base_engine.py:155-167 (:478-489 for survival): `param_k.data.mul_(mm).add_(param_q.data, alpha=1. - mm)` over
    `zip(model.parameters(), model_ema.parameters())`.  Returns the new teacher tensors (inputs untouched).

    Arithmetic as torch executes it on fp32 tensors: the python doubles `mm` and `1 - mm` are each rounded to fp32; `k * mm` is
    rounded to fp32; `alpha * q + that` is one fused multiply-add (checked bit for bit against the literal loop in
    tests/test_oracle_golden.py)."""
    import numpy as np
    assert 0.0 <= mm <= 1.0, "Momentum needs to be between 0.0 and 1.0, got %.5f" % mm          # base_engine.py:164
    mm32, alpha32 = np.float32(mm), np.float32(1.0 - mm)
    out = []
    for q, k in zip(params_q, params_k):
        k1 = (k.detach().cpu().numpy() * mm32).astype(np.float32)
        fused = k1.astype(np.float64) + np.float64(alpha32) * q.detach().cpu().numpy().astype(np.float64)   # exact product, one rounding
        out.append(torch.from_numpy(fused.astype(np.float32)).reshape(k.shape))
    return out
