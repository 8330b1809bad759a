"""Host-side mirror of the reference's warping operators (`modules/warping.py`) on the
hand-written sm_100a kernels in libsatmvs_b200.so.

Same names, argument meaning and error behaviour as the reference:

    rpc_warping(src_fea, src_rpc, ref_rpc, depth_values, coef)      modules/warping.py:310
    rpc_warping_enisum(src_fea, src_rpc, ref_rpc, depth_values)     modules/warping.py:139
    homo_warping(src_fea, src_proj, ref_proj, depth_values)         modules/warping.py:6

plus the fused entry the cascade stages call instead of the per-view loop of
`networks/casred.py:26-53`:

    build_cost_volume(ref_fea, src_feas, ref_cam, src_cams, depth_values, geo_model)

Gradients flow to the feature maps only, as in the reference (grid built under no_grad).
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib

_SWEEP_WS: dict = {}


def sweep_workspace(nbytes: int, device) -> torch.Tensor:
    """Caller-owned scratch of the sweep (re-packed source features), cached per device."""
    key = (device.type, device.index)
    ws = _SWEEP_WS.get(key)
    if ws is None or ws.numel() < nbytes:
        ws = torch.empty(max(nbytes, 1), dtype=torch.uint8, device=device)
        _SWEEP_WS[key] = ws
    return ws


def host_f64(cam: torch.Tensor) -> np.ndarray:
    """Camera tensors travel to the kernels by value, so they are needed on the host.  CPU
    tensors are used directly; a CUDA tensor costs one small synchronising copy (<= 170 doubles
    per view) on EVERY call: nothing is cached, because the reference re-uploads `cam_para` each
    sample (`tools/utils.py:85`) and the caching allocator hands the same address to different
    cameras, so no key short of the content identifies them."""
    if not cam.is_cuda:
        return np.ascontiguousarray(cam.detach().numpy(), dtype=np.float64)
    return np.ascontiguousarray(cam.detach().cpu().numpy(), dtype=np.float64)


def _depth_arg(depth_values: torch.Tensor, B: int, H: int, W: int):
    """[B, D] -> per-plane mode, [B, D, H, W] -> per-pixel mode (`warping.py:329-332`)."""
    if depth_values.dim() == 2:
        return _lib.require_cuda(depth_values, "depth_values"), 0
    if depth_values.dim() == 4 and tuple(depth_values.shape[2:]) == (H, W):
        return _lib.require_cuda(depth_values, "depth_values"), 1
    raise ValueError(f"depth_values must be [B, D] or [B, D, {H}, {W}], got {tuple(depth_values.shape)}")


def _cam_len(geo_model: str) -> int:
    if geo_model == "rpc":
        return 170
    if geo_model == "pinhole":
        return 16
    raise ValueError(f"geo_model must be 'rpc' or 'pinhole', got {geo_model!r}")


def _check_cam(cam: np.ndarray, B: int, geo_model: str, name: str) -> np.ndarray:
    n = _cam_len(geo_model)
    cam = cam.reshape(cam.shape[0], -1)
    if cam.shape != (B, n):
        raise ValueError(f"{name} must be [B, {n if geo_model == 'rpc' else '4, 4'}] float64, got {cam.shape}")
    return cam


def _dptr(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)


class _WarpFn(torch.autograd.Function):
    """Single-view warp; backward scatters into the source feature map."""

    @staticmethod
    def forward(ctx, src_fea, depth_values, src_cam, ref_cam, geo_model):
        B, Cc, H, W = src_fea.shape
        src = _lib.require_cuda(src_fea, "src_fea")
        depth, per_pixel = _depth_arg(depth_values, B, H, W)
        D = depth.shape[1]
        out = torch.empty((B, Cc, D, H, W), dtype=torch.float32, device=src.device)
        fn = _lib.lib().satmvs_rpc_warp_fwd if geo_model == "rpc" else _lib.lib().satmvs_homo_warp_fwd
        with torch.cuda.device(src.device):
            st = _lib.stream_ptr(src.device)
            for b in range(B):
                _lib.check(fn(src[b].data_ptr(), _dptr(src_cam[b]), _dptr(ref_cam[b]), depth[b].data_ptr(), per_pixel,
                              Cc, D, H, W, out[b].data_ptr(), st), "warp_fwd")
        ctx.save_for_backward(depth)
        ctx.meta = (src_cam, ref_cam, geo_model, per_pixel, (B, Cc, D, H, W))
        return out

    @staticmethod
    def backward(ctx, grad_out):
        (depth,) = ctx.saved_tensors
        src_cam, ref_cam, geo_model, per_pixel, (B, Cc, D, H, W) = ctx.meta
        g = grad_out.contiguous()
        grad_src = torch.zeros((B, Cc, H, W), dtype=torch.float32, device=g.device)
        fn = _lib.lib().satmvs_rpc_warp_bwd if geo_model == "rpc" else _lib.lib().satmvs_homo_warp_bwd
        with torch.cuda.device(g.device):
            st = _lib.stream_ptr(g.device)
            for b in range(B):
                _lib.check(fn(g[b].data_ptr(), _dptr(src_cam[b]), _dptr(ref_cam[b]), depth[b].data_ptr(), per_pixel,
                              Cc, D, H, W, grad_src[b].data_ptr(), st), "warp_bwd")
        return grad_src, None, None, None, None


def _warp(src_fea, src_cam, ref_cam, depth_values, geo_model):
    if src_fea.dim() != 4:
        raise ValueError("src_fea must be [B, C, H, W]")
    B = src_fea.shape[0]
    s = _check_cam(host_f64(src_cam), B, geo_model, "src camera")
    r = _check_cam(host_f64(ref_cam), B, geo_model, "ref camera")
    return _WarpFn.apply(src_fea, depth_values, s, r, geo_model)


def rpc_warping(src_fea, src_rpc, ref_rpc, depth_values, coef=None):
    """`rpc_warping` (`modules/warping.py:310-365`).  src_fea [B,C,H,W] f32 (CUDA), src_rpc/ref_rpc
    [B,170] f64 (host or CUDA), depth_values [B,D] or [B,D,H,W] f32 -> [B,C,D,H,W] f32.
    `coef` is the reference's caller-owned fp64 scratch; it is accepted and ignored."""
    return _warp(src_fea, src_rpc, ref_rpc, depth_values, "rpc")


def homo_warping(src_fea, src_proj, ref_proj, depth_values):
    """`homo_warping` (`modules/warping.py:6-44`).  src_proj/ref_proj [B,4,4] f64 (K·E)."""
    return _warp(src_fea, src_proj, ref_proj, depth_values, "pinhole")


# index triple over x = (1, L, P, H) of each RPC00B monomial and its permutation count,
# the inverse of the 4x4x4 table `dataset/data_io.py:95-120`
_QC_TRIPLES = ((0, 0, 0, 1), (0, 0, 1, 3), (0, 0, 2, 3), (0, 0, 3, 3), (0, 1, 2, 6), (0, 1, 3, 6), (0, 2, 3, 6),
               (0, 1, 1, 3), (0, 2, 2, 3), (0, 3, 3, 3), (1, 2, 3, 6), (1, 1, 1, 1), (1, 2, 2, 3), (1, 3, 3, 3),
               (1, 1, 2, 3), (2, 2, 2, 1), (2, 3, 3, 3), (1, 1, 3, 3), (2, 2, 3, 3), (3, 3, 3, 1))


def qc_dict_to_rpc(rpc: dict) -> torch.Tensor:
    """The quaternary-cubic dict format of `rpc_warping_enisum` (`dataset/data_io.py:123-160`:
    `*_off`, `*_scale` [B] and `*_tensor` [B,4,4,4]) back to the [B,170] vector."""
    B = rpc["line_off"].shape[0]
    out = torch.zeros(B, 170, dtype=torch.float64)
    for i, k in enumerate(("line", "samp", "lat", "lon", "height")):
        out[:, i] = rpc[k + "_off"].detach().cpu().double()
        out[:, 5 + i] = rpc[k + "_scale"].detach().cpu().double()
    for k, at in (("line_num", 10), ("line_den", 30), ("samp_num", 50), ("samp_den", 70),
                  ("lat_num", 90), ("lat_den", 110), ("lon_num", 130), ("lon_den", 150)):
        T = rpc[k + "_tensor"].detach().cpu().double()
        for m, (i, j, l, n) in enumerate(_QC_TRIPLES):
            out[:, at + m] = T[:, i, j, l] * n
    return out


def rpc_warping_enisum(src_fea, src_rpc, ref_rpc, depth_values):
    """`rpc_warping_enisum` (`modules/warping.py:139-178`): same warp, cameras given in the 4x4x4
    tensor form.  The polynomial is identical, so it runs on the same kernel."""
    return rpc_warping(src_fea, qc_dict_to_rpc(src_rpc), qc_dict_to_rpc(ref_rpc), depth_values, None)


class _CostVolumeFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, depth_values, ref_cam, src_cams, geo_model, ref_fea, *src_feas):
        B, Cc, H, W = ref_fea.shape
        ref = _lib.require_cuda(ref_fea, "ref_fea")
        srcs = [_lib.require_cuda(s, "src_fea") for s in src_feas]
        for s in srcs:
            if s.shape != ref.shape or s.device != ref.device:
                raise ValueError("all feature maps must share shape and device")
        depth, per_pixel = _depth_arg(depth_values, B, H, W)
        D = depth.shape[1]
        n_src = len(srcs)
        out = torch.empty((B, Cc, D, H, W), dtype=torch.float32, device=ref.device)
        fn = (_lib.lib().satmvs_cost_volume_rpc_fwd_sharded if geo_model == "rpc"
              else _lib.lib().satmvs_cost_volume_homo_fwd_sharded)
        ws = sweep_workspace(n_src * Cc * H * W * 4, ref.device)
        with torch.cuda.device(ref.device):
            st = _lib.stream_ptr(ref.device)
            for b in range(B):
                ptrs = _lib.ptr_array([s[b].data_ptr() for s in srcs])
                cams = np.ascontiguousarray(src_cams[:, b])
                _lib.check(fn(ref[b].data_ptr(), ptrs, n_src, _dptr(ref_cam[b]), _dptr(cams), depth[b].data_ptr(),
                              per_pixel, Cc, D, H, W, 0, D, _lib.ptr_array([out[b].data_ptr()]), 1,
                              ws.data_ptr(), ws.numel(), st), "cost_volume_fwd")
        ctx.save_for_backward(depth, ref, *srcs)
        ctx.meta = (ref_cam, src_cams, geo_model, per_pixel, (B, Cc, D, H, W))
        return out

    @staticmethod
    def backward(ctx, grad_var):
        depth, ref, *srcs = ctx.saved_tensors
        ref_cam, src_cams, geo_model, per_pixel, (B, Cc, D, H, W) = ctx.meta
        g = grad_var.contiguous()
        n_src = len(srcs)
        grad_ref = torch.zeros_like(ref)
        grad_srcs = [torch.zeros_like(s) for s in srcs]
        fn = _lib.lib().satmvs_cost_volume_rpc_bwd if geo_model == "rpc" else _lib.lib().satmvs_cost_volume_homo_bwd
        with torch.cuda.device(g.device):
            st = _lib.stream_ptr(g.device)
            for b in range(B):
                sp = _lib.ptr_array([s[b].data_ptr() for s in srcs])
                gp = _lib.ptr_array([s[b].data_ptr() for s in grad_srcs])
                cams = np.ascontiguousarray(src_cams[:, b])
                _lib.check(fn(g[b].data_ptr(), ref[b].data_ptr(), sp, n_src, _dptr(ref_cam[b]), _dptr(cams),
                              depth[b].data_ptr(), per_pixel, Cc, D, H, W, grad_ref[b].data_ptr(), gp, st),
                           "cost_volume_bwd")
        return (None, None, None, None, grad_ref, *grad_srcs)


def build_cost_volume(ref_fea, src_feas, ref_cam, src_cams, depth_values, geo_model="rpc"):
    """Variance cost volume in ONE kernel: replaces the loop of `networks/casred.py:26-53`
    (`ref.repeat`, per-view warp, `+`, `**2`, `div_`, `sub_`, `pow_`).

    ref_fea [B,C,H,W]; src_feas list of V-1 [B,C,H,W]; ref_cam [B,170] / [B,4,4]; src_cams list of
    V-1 like ref_cam (or one [B,V-1,...] tensor); depth_values [B,D] or [B,D,H,W].
    Returns var [B,C,D,H,W] = Q/V - (S/V)^2."""
    src_feas = list(src_feas)
    if isinstance(src_cams, torch.Tensor):
        src_cams = list(torch.unbind(src_cams, 1))
    if len(src_cams) != len(src_feas) or not src_feas:
        raise AssertionError("Different number of images and projection matrices")   # casred.py:15
    B = ref_fea.shape[0]
    r = _check_cam(host_f64(ref_cam), B, geo_model, "ref camera")
    s = np.stack([_check_cam(host_f64(c), B, geo_model, "src camera") for c in src_cams])   # [V-1, B, n]
    return _CostVolumeFn.apply(depth_values, r, s, geo_model, ref_fea, *src_feas)
