"""Init helper (reference: modules/mhim_modules/utils.py:8-22)."""
from .._common import init_linear_layers as initialize_weights  # noqa: F401
