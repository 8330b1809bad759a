"""oracle.volume — variance cost volume over views.  TEST INFRASTRUCTURE (see oracle/__init__.py)."""
from __future__ import annotations

import torch

from . import geometry


def warp_view(src_fea, src_cam, ref_cam, depth_values, geo_model: str, sampler: str = "aten"):
    """Dispatch used by `networks/casred.py:38-45`: RPC or pin-hole warp of one source view."""
    if geo_model == "rpc":
        return geometry.rpc_warp(src_fea, src_cam, ref_cam, depth_values, sampler)
    if geo_model == "pinhole":
        return geometry.homo_warp(src_fea, src_cam, ref_cam, depth_values, sampler)
    raise ValueError(f"geo_model must be 'rpc' or 'pinhole', got {geo_model!r}")


def variance_cost_volume(features: list[torch.Tensor], cams: torch.Tensor, depth_values: torch.Tensor,
                         geo_model: str = "rpc", sampler: str = "aten") -> torch.Tensor:
    """`compute_depth_when_train` step 2 (`networks/casred.py:26-53`; same in `casmvs.py:30-59`).

    features: V tensors [B, C, H, W] (index 0 = reference view); cams [B, V, 170] or [B, V, 4, 4]
    fp64; depth_values [B, D] or [B, D, H, W] fp32.  Returns var [B, C, D, H, W] with
    S = ref + sum warp_v, Q = ref^2 + sum warp_v^2, var = Q/V - (S/V)^2, evaluated in that order.
    """
    V = len(features)
    cam = torch.unbind(cams.to(features[0].device), 1)   # device-agnostic: the same code is the eager-GPU baseline
    D = depth_values.shape[1]
    ref = features[0].unsqueeze(2).repeat(1, 1, D, 1, 1)
    s, q = ref, ref ** 2
    for v in range(1, V):
        w = warp_view(features[v], cam[v], cam[0], depth_values, geo_model, sampler)
        s = s + w
        q = q + w ** 2
    return q.div_(V).sub_(s.div_(V).pow_(2))
