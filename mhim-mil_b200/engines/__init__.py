from .common_mil import CommonMIL  # noqa: F401
from .ema import ema_update  # noqa: F401
