"""Components of the MHIM model (mirror of the reference's modules/mhim_modules package)."""
from .baseline import DAttention, DSMIL, SAttention
from .losses import SoftTargetCrossEntropy
from .masking import mask_fn, select_mask_fn
from .merge import MCA, Merge
from .scoring import get_pseudo_score, get_pseudo_score_trans
from .utils import initialize_weights

__all__ = ["SAttention", "DAttention", "DSMIL", "select_mask_fn", "mask_fn", "get_pseudo_score", "get_pseudo_score_trans", "Merge", "MCA",
           "SoftTargetCrossEntropy", "initialize_weights"]
