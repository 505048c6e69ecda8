from .common_mil import CommonMIL  # noqa: F401
from .ema import ema_update  # noqa: F401
from .graphed import GraphedStep  # noqa: F401
from .loader import BagLoader  # noqa: F401
from .optim import FusedAdam  # noqa: F401
from .validate import binary_auroc, classification_metrics, collect_logits, validate  # noqa: F401
