"""Drop-in replacements for the reference's modules/ package (same class names, constructor kwargs, forward()
signatures and state_dict keys) whose arithmetic runs in libmhimk.so."""
from . import abmil, clam, dsmil, dtfd, mhim, nystrom_attention, transmil  # noqa: F401
from .abmil import AttentionGated, DAttention  # noqa: F401
from .clam import CLAM_MB, CLAM_SB  # noqa: F401
from .dsmil import MILNet  # noqa: F401
from .dtfd import DTFD  # noqa: F401
from .mhim import MHIM  # noqa: F401
from .nystrom_attention import NystromAttention  # noqa: F401
from .transmil import TransMIL  # noqa: F401
