"""satmvs_b200 — B200-native (sm_100a) implementation of SatMVS's RPC plane-sweep hot path.

Host code mirrors the reference's operator interface (`modules/warping.py`, `modules/module.py`,
`modules/depth_range.py`); the arithmetic runs in hand-written CUDA kernels behind the C ABI of
`include/satmvs_b200.h` (libsatmvs_b200.so, built in-tree by `satmvs_b200.build`).
"""
from .warping import rpc_warping, rpc_warping_enisum, homo_warping, build_cost_volume  # noqa: F401
from .regress import softargmin, StreamingSoftArgmin  # noqa: F401
from .rpc_tensor import RPCModelParameter  # noqa: F401
from .module import RED_Regularization, slice_RED_Regularization, CostRegNet, FeatureNet, depth_regression, conv_block  # noqa: F401
from .stages import stage_train_red, stage_pred_red, stage_casmvs, cascade
from .depth_range import get_depth_range_samples, stage_depth_hypotheses  # noqa: F401
from . import training  # noqa: F401  (train() mode and loss.backward() of the patched operators, train.py:267-287)
from . import data_io  # noqa: F401  (PFM / RPC text formats)
from . import rpc_filter  # noqa: F401  (geometric-consistency filter, tools/rpc_filter.py)
