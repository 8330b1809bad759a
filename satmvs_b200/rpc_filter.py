"""Mirror of the reference's RPC geometric-consistency filter (`tools/rpc_filter.py:9-112`): same three
functions, arguments and return values (host numpy in, host numpy out).  The four RPC evaluations per
view pair run on the point-list kernels behind `rpc_tensor.RPCModelParameter` (fp64, the same device
functions as the plane sweep) and the source-depth lookup on `satmvs_remap_bilinear` (a restatement of
the `cv2.remap` call at :29-30); the per-pixel thresholds and the averaging are a few numpy lines like
the reference's."""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib
from .rpc_tensor import RPCModelParameter


def remap_bilinear(src: torch.Tensor, mapx: torch.Tensor, mapy: torch.Tensor, border: float) -> torch.Tensor:
    """`cv2.remap(src, mapx, mapy, INTER_LINEAR, BORDER_CONSTANT, borderValue=border)` for float32 device tensors."""
    if not (src.is_cuda and mapx.is_cuda and mapy.is_cuda):
        raise RuntimeError("remap_bilinear runs on the GPU only (there is no CPU fallback)")
    if src.dim() != 2 or mapx.shape != mapy.shape:
        raise AssertionError("src is [H,W]; the two maps share one shape")
    src, mapx, mapy = (t.contiguous().float() for t in (src, mapx, mapy))
    out = torch.empty_like(mapx)
    with torch.cuda.device(src.device):
        _lib.check(_lib.lib().satmvs_remap_bilinear(src.data_ptr(), src.shape[0], src.shape[1], mapx.data_ptr(), mapy.data_ptr(),
                                                    mapx.numel(), C.c_float(border), out.data_ptr(), _lib.stream_ptr(src.device)),
                   "remap_bilinear")
    return out


def reproject_with_depth(depth_ref, rpc_ref, depth_src, rpc_src):
    """rpc_filter.py:9-45: reference pixels -> object space at the reference heights -> source view; source heights
    gathered there; back to object space at those heights -> reference view."""
    dev = torch.device("cuda", torch.cuda.current_device())
    m_ref, m_src = RPCModelParameter(rpc_ref), RPCModelParameter(rpc_src)
    height, width = depth_ref.shape
    y, x = torch.meshgrid(torch.arange(height, device=dev, dtype=torch.float64),
                          torch.arange(width, device=dev, dtype=torch.float64), indexing="ij")
    x, y = x.reshape(-1), y.reshape(-1)
    d_ref = torch.as_tensor(np.ascontiguousarray(depth_ref)).to(dev, torch.float64).reshape(-1)
    lat, lon = m_ref.RPC_PHOTO2OBJ_device(x, y, d_ref)
    x_src, y_src = m_src.RPC_OBJ2PHOTO_device(lat, lon, d_ref)
    d_src = torch.as_tensor(np.ascontiguousarray(depth_src)).to(dev, torch.float32)
    sampled = remap_bilinear(d_src, x_src.float(), y_src.float(), -999.0)          # rpc_filter.py:29-30
    lat, lon = m_src.RPC_PHOTO2OBJ_device(x_src, y_src, sampled.double())
    x_rep, y_rep = m_ref.RPC_OBJ2PHOTO_device(lat, lon, sampled.double())
    shape = (height, width)
    return (sampled.cpu().numpy().reshape(shape), x_rep.cpu().numpy().reshape(shape), y_rep.cpu().numpy().reshape(shape),
            x_src.cpu().numpy().reshape(shape), y_src.cpu().numpy().reshape(shape))


def check_geometric_consistency(depth_ref, rpc_ref, depth_src, rpc_src, p_ratio, d_ratio):
    """rpc_filter.py:48-65: a pixel is consistent when it reprojects within p_ratio pixels and d_ratio metres."""
    height, width = depth_ref.shape
    x_ref, y_ref = np.meshgrid(np.arange(0, width), np.arange(0, height))
    depth_reprojected, x_rep, y_rep, x_src, y_src = reproject_with_depth(depth_ref, rpc_ref, depth_src, rpc_src)
    dist = np.sqrt((x_rep - x_ref) ** 2 + (y_rep - y_ref) ** 2)
    depth_diff = np.abs(depth_reprojected - depth_ref)
    mask = np.logical_and(dist < p_ratio, depth_diff < d_ratio)
    depth_reprojected[~mask] = 0
    return mask, depth_reprojected, x_src, y_src


def filter_depth(depths, rpcs, p_ratio, d_ratio, geo_consist_num, prob=None, confidence_ratio=0.0):
    """rpc_filter.py:68-112: view 0 is the reference; returns (final mask, height averaged over the consistent views)."""
    ref_depth, ref_rpc = depths[0], rpcs[0]
    photo_mask = prob > confidence_ratio if prob is not None else np.ones_like(ref_depth, bool)
    geo_mask_sum = 0
    ests = []
    for v in range(1, depths.shape[0]):
        geo_mask, depth_reprojected, _, _ = check_geometric_consistency(ref_depth, ref_rpc, depths[v], rpcs[v], p_ratio, d_ratio)
        geo_mask_sum = geo_mask_sum + geo_mask.astype(np.int32)
        ests.append(depth_reprojected)
    depth_est_averaged = (sum(ests) + ref_depth) / (geo_mask_sum + 1)
    final_mask = np.logical_and(photo_mask, geo_mask_sum >= geo_consist_num)
    return final_mask, depth_est_averaged
