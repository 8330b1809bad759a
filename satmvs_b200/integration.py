"""Drop-in wiring for the reference's own scripts.

The reference star-imports its ops into the network modules (`networks/casred.py:4-6`,
`networks/casmvs.py:4-6`), so the names must be rebound in *those* namespaces, not only in
`modules.*`.  `patch_reference()` does that for an already imported reference tree; nothing here
imports the reference itself.
"""
from __future__ import annotations

import sys

from . import depth_range, module, warping

_OPS = {
    "rpc_warping": warping.rpc_warping,                      # modules/warping.py:310
    "rpc_warping_enisum": warping.rpc_warping_enisum,        # modules/warping.py:139
    "homo_warping": warping.homo_warping,                    # modules/warping.py:6
    "depth_regression": module.depth_regression,             # modules/module.py:433
    "get_depth_range_samples": depth_range.get_depth_range_samples,   # modules/depth_range.py:23
    "CostRegNet": module.CostRegNet,                         # modules/module.py:546
    "FeatureNet": module.FeatureNet,                         # modules/module.py:442
    "RED_Regularization": module.RED_Regularization,         # modules/module.py:595
    "slice_RED_Regularization": module.slice_RED_Regularization,      # modules/module.py:653
}
_TARGETS = ("modules.warping", "modules.module", "modules.depth_range", "networks.casred", "networks.casmvs", "networks.ucs")


def patch_reference(modules=None) -> dict[str, list[str]]:
    """Rebind the hot-path operators of an imported SatMVS tree to the B200 implementations.
    `modules`: iterable of module objects, default = those of `_TARGETS` present in sys.modules.
    Returns {module name: [rebound names]}.  Models must be (re)constructed after patching so the
    regulariser classes resolve to the new ones; reference checkpoints load unchanged."""
    if modules is None:
        modules = [sys.modules[n] for n in _TARGETS if n in sys.modules]
    done = {}
    for m in modules:
        hit = [name for name in _OPS if hasattr(m, name)]
        for name in hit:
            setattr(m, name, _OPS[name])
        done[m.__name__] = hit
    return done
