"""mhimk -- B200 (sm_100a) kernels for MHIM-MIL's per-bag aggregation path behind the reference's module API.

The directory is named `mhim-mil_b200`; import it as `mhimk` (repo-root shim) or
`importlib.import_module("mhim-mil_b200")`.
"""
from . import _lib  # noqa: F401
from . import ops  # noqa: F401

__all__ = ["_lib", "ops"]
