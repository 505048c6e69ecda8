"""MHIM with the reference's interface (modules/mhim.py:12-378): teacher scoring -> masked hard-instance selection ->
(merge) -> attention pooling -> logits + distillation loss.

Kernel mapping
  forward_teacher / forward_test / pure, baseline='attn', no autograd, dropout inactive:
      ONE fused tcgen05 pass over the bag (projection + attention logits + online softmax pool + h.W_pred), then
      mil_cam_score for the instance scores -- the N x 512 embedding is never written to HBM.
  forward (student, training): CUDA GEMM primitives with CUDA backward (ops.linear_act / ops.softmax_pool), top-k and
      mask_ids on the device (no host sync).
Known upstream defects that are NOT reproduced: `merge_enable=False` crashing forward() (Identity called with two
arguments, mhim.py:351) -- here the merge step is simply skipped.
"""
import torch
from torch import nn

from .. import ops
from . import _common as C
from .mhim_modules.baseline import DAttention, DSMIL, SAttention
from .mhim_modules.losses import SoftTargetCrossEntropy
from .mhim_modules.masking import mask_fn, select_mask_fn
from .mhim_modules.merge import Merge
from .mhim_modules.scoring import get_pseudo_score, get_pseudo_score_trans


class MHIM(C.MilModule):
    def __init__(self, input_dim=1024, mlp_dim=512, mask_ratio=0, n_classes=2, temp_t=1.0, dropout=0.25, act="relu", mask_ratio_h=0.0,
                 mrh_sche=None, mask_ratio_hr=0.0, mask_ratio_l=0.0, da_act="gelu", baseline="selfattn", head=8, attn2score=True,
                 merge_enable=True, merge_k=1, merge_mm=0.9998, merge_ratio=0.0, merge_test=False):
        super().__init__()
        self.mask_ratio, self.mask_ratio_h, self.mask_ratio_hr, self.mask_ratio_l = mask_ratio, mask_ratio_h, mask_ratio_hr, mask_ratio_l
        self.select_inv, self.msa_fusion, self.mrh_sche, self.attn_layer = False, "vote", mrh_sche, 0
        self.baseline, self.merge_test, self.attn2score, self.head = baseline, merge_test, attn2score, head
        self.act = act.lower() if act.lower() in ("relu", "gelu") else "none"
        feat = [nn.Linear(input_dim, mlp_dim)]
        if self.act != "none":
            feat += [C.act_module(self.act)]
        self.feature = nn.Sequential(*feat)
        self.dp = nn.Dropout(dropout) if dropout > 0.0 else nn.Identity()
        self.merge = Merge(mlp_dim, k=merge_k, g_q_mm=merge_mm, merge_ratio=merge_ratio, mask_type="random") if merge_enable else nn.Identity()
        if baseline == "selfattn":
            self.online_encoder = SAttention(mlp_dim=mlp_dim, head=head)
        elif baseline == "attn":
            self.online_encoder = DAttention(mlp_dim, da_act)
        elif baseline == "dsmil":
            self.online_encoder = DSMIL(n_classes=n_classes, mlp_dim=mlp_dim, mask_ratio=mask_ratio, cls_attn=self.attn2score)
        else:
            raise ValueError(baseline)
        self.predictor = nn.Linear(mlp_dim, n_classes)
        self.temp_t, self.temp_s = temp_t, 1.0
        self.cl_loss = SoftTargetCrossEntropy(self.temp_t, self.temp_s)
        self.predictor_cl = self.target_predictor = nn.Identity()
        self.precision = ops.DEFAULT_PRECISION
        C.init_linear_layers(self)

    # ------------------------------------------------------------------ helpers
    def _embed(self, x):
        """feature + dropout on every row (mhim.py:193-194 etc.); x [1,N,D] -> [1,N,512]"""
        C.require_cuda(x, "MHIM")
        if x.dim() != 3 or x.shape[0] != 1:
            raise RuntimeError("mhimk MHIM: input must be [1, N, D] (one bag)")
        return self.dp(C.lin(self.feature[0], x[0], self.act))[None]

    def _fusable(self, x):
        dp_active = self.training and isinstance(self.dp, nn.Dropout) and self.dp.p > 0
        enc = self.online_encoder
        return (self.baseline == "attn" and not enc.gated and not dp_active and not C.grad_needed(self, x) and x.is_cuda
                and x.dim() == 3 and x.shape[0] == 1 and x.shape[-1] % 32 == 0 and self.feature[0].out_features == 512
                and not (self.training and enc.attention.p_drop > 0))

    def _fused(self, x, want_scores=False, want_h=False, with_pred=False, with_logits=False):
        f0, att = self.feature[0], self.online_encoder.attention
        return ops.abmil_fused_forward(x[0], f0.weight, f0.bias, self.act, att.attention[0].weight, None, att.attention[-1].weight, None,
                                       att.act, Wp=self.predictor.weight if with_pred else None, want_scores=want_scores, want_h=want_h,
                                       precision=self.precision, Wcls=self.predictor.weight if with_logits else None,
                                       bcls=self.predictor.bias if with_logits else None, volatile=self.training)

    # ------------------------------------------------------------------ masking
    def get_mask(self, ps, i, attn, mrh=None):
        """(len_keep, mask_ids) exactly as mhim.py:109-179 stages it: random (v1), low (v1), then high-attention masking."""
        len_keep, mask_ids = ps, None
        if attn is not None and self.mask_ratio > 0.0:
            len_keep, mask_ids = select_mask_fn(ps, attn, False, self.mask_ratio, select_inv=self.select_inv, random_ratio=0.001,
                                                msa_fusion=self.msa_fusion)
        if attn is not None and self.mask_ratio_l > 0.0:
            if mask_ids is None:
                len_keep, mask_ids = select_mask_fn(ps, attn, False, self.mask_ratio_l, select_inv=self.select_inv, msa_fusion=self.msa_fusion)
            else:
                other = (mask_ids[:, :len_keep] if self.select_inv else mask_ids[:, len_keep:]).squeeze()
                len_keep, mask_ids = select_mask_fn(ps, attn, False, self.mask_ratio_l, select_inv=self.select_inv, mask_ids_other=mask_ids,
                                                    len_keep_other=ps, cls_attn_topk_idx_other=other, msa_fusion=self.msa_fusion)
        r_h = self.mask_ratio_h
        if self.mrh_sche is not None:
            r_h = self.mrh_sche[i]
        if mrh is not None:
            r_h = mrh
        if r_h > 0.0:
            if mask_ids is None:
                len_keep, mask_ids = select_mask_fn(ps, attn, largest=True, mask_ratio=r_h, len_keep_other=ps, random_ratio=self.mask_ratio_hr,
                                                    select_inv=self.select_inv, msa_fusion=self.msa_fusion)
            else:
                other = (mask_ids[:, :len_keep] if self.select_inv else mask_ids[:, len_keep:]).squeeze()
                len_keep, mask_ids = select_mask_fn(ps, attn, largest=True, mask_ratio=r_h, mask_ids_other=mask_ids, len_keep_other=ps,
                                                    cls_attn_topk_idx_other=other, random_ratio=self.mask_ratio_hr, select_inv=self.select_inv,
                                                    msa_fusion=self.msa_fusion)
        return len_keep, mask_ids

    # ------------------------------------------------------------------ entry points
    @torch.no_grad()
    def forward_teacher(self, x):
        """-> (cls_feat, score) (mhim.py:181-227)"""
        if self._fusable(x) and not self.merge_test and self.attn2score:
            out = self._fused(x, want_scores=True, with_pred=True)
            score = ops.cam_score(out["s"], out["t"], out["stats"], self.predictor.bias)     # bias[0] read on the device: no host sync
            return out["pooled"][None], score[None]
        h = self._embed(x)
        p = h.size(1)
        if self.merge_test:
            was, self.merge.training = self.merge.training, False
            h = self.merge(h)
            self.merge.training = was
        if self.baseline == "dsmil":
            _, feat, attn = self.online_encoder(h, return_attn=True)
            return feat, (attn[:, :p] if self.merge_test else attn)
        feat, attn, act = self.online_encoder(h, return_attn=True, return_act=True)
        if self.merge_test:
            attn = [a[:, :, :p] for a in attn] if isinstance(attn, (list, tuple)) else attn[:, :p]
        if self.attn2score:
            if self.baseline == "selfattn":
                attn = get_pseudo_score_trans(self.predictor, act, attn[0], self.online_encoder.layer1.attn.to_out)
            else:
                attn = get_pseudo_score(self.predictor, act, attn)
        elif isinstance(attn, (list, tuple)):
            attn = attn[self.attn_layer]
        return feat, attn

    @torch.no_grad()
    def forward_test(self, x, return_attn=False, no_norm=False, return_act=False, **kwargs):
        """inference (mhim.py:229-272)"""
        if self._fusable(x) and not self.merge_test and not return_act:
            out = self._fused(x, want_scores=return_attn, with_logits=True)
            logits = out["logits"]
            if not return_attn:
                return logits
            a = out["s"] if no_norm else torch.exp(out["s"] - out["stats"][0]) / out["stats"][1]
            return logits, a[None]
        h = self._embed(x)
        if self.merge_test:
            h = self.merge(h)
        a = None
        if return_attn:
            if return_act:
                y, a, act = self.online_encoder(h, return_attn=True, return_act=True, no_norm=no_norm, **kwargs)
                a = [a, act]
            elif self.baseline == "dsmil":
                y, _, a = self.online_encoder(h, return_attn=True, no_norm=no_norm, **kwargs)
            else:
                y, a = self.online_encoder(h, return_attn=True, no_norm=no_norm, **kwargs)
        else:
            y = self.online_encoder(h)
        if self.baseline != "dsmil":
            y = C.lin(self.predictor, y)
        return (y, a) if return_attn else y

    def pure(self, x):
        """no masking, no merging (mhim.py:274-298)"""
        ps = x.size(1)
        if self._fusable(x):
            y = self._fused(x, with_logits=True)["logits"]
        else:
            h = self._embed(x)
            if self.baseline == "dsmil":
                y, _ = self.online_encoder(h)
            else:
                y = C.lin(self.predictor, self.online_encoder(h))
        return (y, 0, ps, ps) if self.training else y

    def forward_loss(self, student_cls_feat, teacher_cls_feat):
        return self.cl_loss(student_cls_feat, teacher_cls_feat.detach()) if teacher_cls_feat is not None else 0.0

    def forward(self, x, attn=None, teacher_cls_feat=None, i=None, pos=None):
        """student pass (mhim.py:318-378) -> (logits, cls_loss, ps, len_keep)"""
        h = self._embed(x)
        ps = h.size(1)
        len_keep, mask_ids = self.get_mask(ps, i, attn)
        if mask_ids is None:
            raise RuntimeError("MHIM.forward needs teacher attention and a positive mask ratio (same precondition as the reference, "
                               "masking.py:104)")
        h = mask_fn(h, mask_ids, len_keep)
        ids_keep = mask_ids[:, :len_keep]
        attn = attn[:, :, ids_keep[0]] if attn.dim() > 2 else attn[:, ids_keep[0]]
        if isinstance(self.merge, Merge):
            h = self.merge(h, attn)
        len_keep = h.size(1)
        if self.baseline == "dsmil":
            logit, feat = self.online_encoder(h)
        else:
            feat = self.online_encoder(h)
            logit = C.lin(self.predictor, feat)
        return logit, self.forward_loss(feat, teacher_cls_feat), ps, len_keep
