"""ctypes binding of libsatmvs_b200.so (include/satmvs_b200.h).

There is deliberately no fallback: if the library is missing or a call fails, a RuntimeError is
raised.  The product path never routes through torch ops or the CPU oracle.
"""
from __future__ import annotations

import ctypes as C
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("SATMVS_B200_LIB") or os.path.join(_HERE, "libsatmvs_b200.so")
_lib = None

_P, _I, _L = C.c_void_p, C.c_int, C.c_int64
_SIGNATURES = {
    "satmvs_abi_version": ([], _I),
    "satmvs_last_error": ([], C.c_char_p),
    "satmvs_async_error": ([], _I),
    "satmvs_profile_begin": ([], _I),
    "satmvs_profile_end": ([_P, _P], _I),
    "satmvs_profile_end_ex": ([_P, _P, _P], _I),
    "satmvs_cost_volume_rpc_fwd": ([_P, _P, _I, _P, _P, _P, _I, _I, _I, _I, _I, _P, _P], _I),
    "satmvs_cost_volume_homo_fwd": ([_P, _P, _I, _P, _P, _P, _I, _I, _I, _I, _I, _P, _P], _I),
    "satmvs_cost_volume_rpc_fwd_sharded": ([_P, _P, _I, _P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _P, _I, _P, C.c_size_t, _P], _I),
    "satmvs_cost_volume_homo_fwd_sharded": ([_P, _P, _I, _P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _P, _I, _P, C.c_size_t, _P], _I),
    "satmvs_rpc_warp_fwd": ([_P, _P, _P, _P, _I, _I, _I, _I, _I, _P, _P], _I),
    "satmvs_homo_warp_fwd": ([_P, _P, _P, _P, _I, _I, _I, _I, _I, _P, _P], _I),
    "satmvs_rpc_warp_bwd": ([_P, _P, _P, _P, _I, _I, _I, _I, _I, _P, _P], _I),
    "satmvs_homo_warp_bwd": ([_P, _P, _P, _P, _I, _I, _I, _I, _I, _P, _P], _I),
    "satmvs_cost_volume_rpc_bwd": ([_P, _P, _P, _I, _P, _P, _P, _I, _I, _I, _I, _I, _P, _P, _P], _I),
    "satmvs_cost_volume_homo_bwd": ([_P, _P, _P, _I, _P, _P, _P, _I, _I, _I, _I, _I, _P, _P, _P], _I),
    "satmvs_rpc_localise": ([_P, _P, _P, _P, _L, _P, _P, _P], _I),
    "satmvs_rpc_project": ([_P, _P, _P, _P, _L, _P, _P, _P], _I),
    "satmvs_remap_bilinear": ([_P, _I, _I, _P, _P, _L, C.c_float, _P, _P], _I),
    "satmvs_softargmin_fwd": ([_P, _P, _I, _I, _I, _I, _I, _P, _P, _P], _I),
    "satmvs_softargmin_stream_update": ([_P, _P, _I, _I, _I, _P, _P], _I),
    "satmvs_softargmin_stream_update_planes": ([_P, _P, _I, _I, _I, _I, _P, _P], _I),
    "satmvs_softargmin_stream_update_volume": ([_P, _P, _I, C.c_float, _I, _I, _I, _I, _P, _P], _I),
    "satmvs_softargmin_stream_finish": ([_P, _I, _I, _P, _P, _P], _I),
    "satmvs_resize_bilinear": ([_P, _I, _I, _I, _I, _I, _P, _P], _I),
    "satmvs_depth_hypotheses": ([_P, _I, _I, _P, _I, _I, C.c_float, _I, _I, _I, _I, _P, _P], _I),
    "satmvs_red_workspace_bytes": ([_I, _I, _I, _I], C.c_size_t),
    "satmvs_red_last_path": ([], _I),
    "satmvs_red_forward": ([_P, _P, _I, _I, _I, _I, _P, _P, _P, _P, C.c_size_t, _P], _I),
    "satmvs_red_pack_bytes": ([_I], C.c_size_t),
    "satmvs_red_forward_packed": ([_P, _P, _I, _I, _I, _I, _P, _P, _P, _P, C.c_size_t, _P, C.c_size_t, _P, _P], _I),
    "satmvs_conv_workspace_bytes": ([_I, _I, _I], C.c_size_t),
    "satmvs_conv_forward": ([_P, _I, _I, _I, _I, _P, _P, _P, _I, _I, _I, _I, C.c_float, _P, _I, _P, C.c_size_t, _P], _I),
    "satmvs_featurenet_workspace_bytes": ([_I, _I, _I, _I], C.c_size_t),
    "satmvs_featurenet_forward": ([_P, _P, _I, _I, _I, _I, _P, _P, _P, _P, C.c_size_t, _P], _I),
    "satmvs_costreg_workspace_bytes": ([_I, _I, _I, _I], C.c_size_t),
    "satmvs_costreg_forward": ([_P, _P, _I, _I, _I, _I, _I, _P, _P, C.c_size_t, _P], _I),
    "satmvs_conv3d_raw": ([_P, _I, _I, _I, _I, _P, C.c_longlong, C.c_longlong, _I, _I, _P, _I, _P], _I),
    "satmvs_conv3d_wgrad_workspace_bytes": ([_I, _I, _I, _I, _I, _I, _I], C.c_size_t),
    "satmvs_conv3d_wgrad": ([_P, _I, _I, _I, _I, _P, _I, _I, _I, _P, C.c_longlong, C.c_longlong, _I, _P, C.c_size_t, _P], _I),
    "satmvs_conv2d_raw": ([_P, _I, _I, _I, _I, _P, C.c_longlong, C.c_longlong, _I, _I, _P, _I, _P], _I),
    "satmvs_conv2d_wgrad_workspace_bytes": ([_I, _I, _I, _I, _I, _I, _I], C.c_size_t),
    "satmvs_conv2d_wgrad": ([_P, _I, _I, _I, _I, _P, _I, _I, _I, _P, C.c_longlong, C.c_longlong, _I, _P, C.c_size_t, _P], _I),
    "satmvs_bn_train_fwd": ([_P, _I, _I, C.c_longlong, _P, _P, C.c_float, _I, _P, _P, _P, _P, _P, _P], _I),
    "satmvs_bn_train_bwd": ([_P, _P, _P, _I, _I, C.c_longlong, _P, _P, _P, _P, C.c_float, _I, _P, _P, _P, _P, _P], _I),
    "satmvs_softargmin_bwd": ([_P, _P, _I, _I, _I, _I, _P, _P, _P], _I),
    "satmvs_red_workspace_layout": ([_I, _I, _I, _I, _P], _I),
    "satmvs_gn_act_fwd": ([_P, _P, _P, _P, _I, _I, _I, _I, C.c_float, _I, _P, _P, _P, _P], _I),
    "satmvs_elementwise": ([_P, C.c_longlong, _P, C.c_longlong, _P, C.c_longlong, _P, C.c_longlong, _P, C.c_longlong, C.c_float,
                            _P, C.c_longlong, _I, C.c_longlong, _P], _I),
    "satmvs_channel_sum": ([_P, _I, C.c_longlong, _P, _P, _P], _I),
    "satmvs_gn_param_grad": ([_P, _P, _P, _P, _I, _I, _I, _I, _P, _P, _P, _P], _I),
    "satmvs_red_recurrence_bwd": ([_P, _I, _P], _I),
}


def exported_symbols() -> list[str]:
    return sorted(_SIGNATURES)


def lib():
    """The loaded library; raises if it has not been built (python -m satmvs_b200.build)."""
    global _lib
    if _lib is None:
        if not os.path.isfile(LIB_PATH):
            raise RuntimeError(f"{LIB_PATH} is missing: build it with `python -m satmvs_b200.build` "
                               "(there is no non-CUDA fallback)")
        handle = C.CDLL(LIB_PATH)
        for name, (argtypes, restype) in _SIGNATURES.items():
            fn = getattr(handle, name)
            fn.argtypes, fn.restype = argtypes, restype
        if handle.satmvs_abi_version() != 1:
            raise RuntimeError("libsatmvs_b200.so ABI version mismatch; rebuild")
        _lib = handle
    return _lib


PROFILE_CLASSES = ("sweep", "conv_batched", "gru_gate_conv", "gru_output_conv", "gru_pointwise", "red_decoder",
                   "costreg", "heads", "featurenet", "train_conv", "train_wgrad", "train_norm")


class profile:
    """Context manager around satmvs_profile_begin/end_ex: per-kernel-class device time inside the library."""

    def __enter__(self):
        lib().satmvs_profile_begin()
        return self

    def __exit__(self, *a):
        ms = (C.c_float * len(PROFILE_CLASSES))()
        n = (C.c_int * len(PROFILE_CLASSES))()
        busy = (C.c_float * len(PROFILE_CLASSES))()
        lib().satmvs_profile_end_ex(ms, n, busy)
        self.ms = dict(zip(PROFILE_CLASSES, list(ms)))              # sum of the launch durations
        self.busy_ms = dict(zip(PROFILE_CLASSES, list(busy)))       # time with at least one launch of the class running
        self.launches = dict(zip(PROFILE_CLASSES, list(n)))


def check(rc: int, what: str) -> None:
    if rc != 0:
        raise RuntimeError(f"{what} failed ({rc}): {lib().satmvs_last_error().decode()}")


def stream_ptr(device) -> int:
    return torch.cuda.current_stream(device).cuda_stream


def require_cuda(t: torch.Tensor, name: str, dtype=torch.float32) -> torch.Tensor:
    if not t.is_cuda:
        raise RuntimeError(f"{name} must be a CUDA tensor: satmvs_b200 has no CPU path")
    if t.dtype != dtype:
        raise TypeError(f"{name} must be {dtype}, got {t.dtype}")
    return t.contiguous()


def ptr_array(ptrs):
    return (C.c_void_p * len(ptrs))(*ptrs)
