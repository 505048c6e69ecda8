from .common_mil import CommonMIL  # noqa: F401
from .ema import ema_update  # noqa: F401
from .graphed import GraphedStep  # noqa: F401
