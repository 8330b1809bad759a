"""oracle.reference_loader — import the UNMODIFIED reference from /root/reference (build
container only; the GPU box has no /root/reference).  TEST INFRASTRUCTURE.

Two shims: a stub `matplotlib.pyplot` (imported by `networks/casred.py:7`) and,
on a CPU-only box, `torch.Tensor.cuda = identity` because the networks call `.cuda()`
unconditionally (`networks/casred.py:34,176-189`, `modules/module.py:617-620`).
"""
from __future__ import annotations

import contextlib
import importlib
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("SATMVS_REFERENCE_ROOT", "/root/reference")


def available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "modules", "warping.py"))


def load():
    """Returns a namespace with the reference modules: warping, module, depth_range, casred, casmvs."""
    if not available():
        raise RuntimeError(f"reference not present at {REFERENCE_ROOT}")
    import torch
    if "matplotlib" not in sys.modules:
        try:
            import matplotlib.pyplot  # noqa: F401
        except Exception:
            mpl = types.ModuleType("matplotlib")
            mpl.pyplot = types.ModuleType("matplotlib.pyplot")
            sys.modules["matplotlib"] = mpl
            sys.modules["matplotlib.pyplot"] = mpl.pyplot
    if not torch.cuda.is_available():
        torch.Tensor.cuda = lambda self, *a, **k: self
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    ns = types.SimpleNamespace()
    with open(os.devnull, "w") as devnull, contextlib.redirect_stdout(devnull):
        ns.warping = importlib.import_module("modules.warping")
        ns.module = importlib.import_module("modules.module")
        ns.depth_range = importlib.import_module("modules.depth_range")
        ns.casred = importlib.import_module("networks.casred")
        ns.casmvs = importlib.import_module("networks.casmvs")
    return ns
