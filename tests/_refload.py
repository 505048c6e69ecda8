"""Import the live reference classes from /root/reference WITHOUT running modules/__init__.py
(which pulls in `future`/`timm`, absent here).  Test infrastructure only.  The GPU box has no /root/reference: there the
unmodified copies under oracle/_ref/ (made by oracle/make_ref.py, git-ignored) are used when present; guard with `have_reference()`.
"""
import importlib
import importlib.util
import os
import sys
import types

_SHIPPED = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_ref")     # oracle/make_ref.py
REF_ROOT = os.environ.get("MHIM_REFERENCE_ROOT") or ("/root/reference" if os.path.isfile("/root/reference/modules/mhim.py") else _SHIPPED)


def have_reference() -> bool:
    return os.path.isfile(os.path.join(REF_ROOT, "modules", "mhim.py"))


def load_reference():
    """Returns a namespace with the reference classes on the hot path."""
    if not have_reference():
        raise RuntimeError("reference tree not present")
    if "modules" not in sys.modules or getattr(sys.modules["modules"], "__mhim_stub__", False) is False:
        pkg = types.ModuleType("modules")
        pkg.__path__ = [os.path.join(REF_ROOT, "modules")]
        pkg.__mhim_stub__ = True
        sys.modules["modules"] = pkg
    ns = types.SimpleNamespace()
    ns.abmil = importlib.import_module("modules.abmil")
    ns.mhim = importlib.import_module("modules.mhim")
    ns.dsmil = importlib.import_module("modules.dsmil")
    ns.transmil = importlib.import_module("modules.transmil")
    ns.nystrom = importlib.import_module("modules.nystrom_attention")
    ns.masking = importlib.import_module("modules.mhim_modules.masking")
    ns.scoring = importlib.import_module("modules.mhim_modules.scoring")
    ns.merge = importlib.import_module("modules.mhim_modules.merge")
    ns.baseline = importlib.import_module("modules.mhim_modules.baseline")
    spec = importlib.util.spec_from_file_location("ref_common_mil", os.path.join(REF_ROOT, "engines", "common_mil.py"))
    cm = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(cm)
    ns.common_mil = cm
    return ns


def load_clam():
    """modules/clam.py of the reference (CLAM_SB / CLAM_MB).  Its `topk` package imports `future.builtins.range` (python-future, absent
    here): a two-line stand-in is registered.  `SmoothTop1SVM(2).cuda()` in the constructors (clam.py:129, :270) needs a CUDA device; on a
    CPU-only host the loss module's `.cuda()` is replaced by the no-op it is for a buffer-only module, so the classes can be built."""
    import torch
    if "future" not in sys.modules:
        fut, fb = types.ModuleType("future"), types.ModuleType("future.builtins")
        fb.range = range
        fut.builtins = fb
        sys.modules["future"], sys.modules["future.builtins"] = fut, fb
    load_reference()
    clam = importlib.import_module("modules.clam")
    if not torch.cuda.is_available():
        svm = importlib.import_module("modules.topk.svm")

        def _cpu_cuda(self, device=None):
            self.get_losses()
            return self
        svm._SVMLoss.cuda = _cpu_cuda
    return clam


def zero_dropout(module):
    """Neutralise every nn.Dropout inside a reference instance (MCA/Nystrom hard-code p=0.1)."""
    import torch.nn as nn
    for m in module.modules():
        if isinstance(m, nn.Dropout):
            m.p = 0.0
    return module
