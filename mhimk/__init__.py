"""Import alias: `import mhimk` -> the package in ./mhim-mil_b200 (a hyphen is not a valid identifier)."""
import importlib
import os
import sys

_root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if _root not in sys.path:
    sys.path.insert(0, _root)
_pkg = importlib.import_module("mhim-mil_b200")
sys.modules[__name__] = _pkg
