"""Drop-in replacements for the reference's modules/ package (same class names, constructor kwargs, forward()
signatures and state_dict keys) whose arithmetic runs in libmhimk.so."""
from .abmil import AttentionGated, DAttention  # noqa: F401
