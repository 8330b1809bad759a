"""oracle.stages — one cascade stage end to end (features in, depth out) and the cascade loop.
TEST INFRASTRUCTURE (see oracle/__init__.py)."""
from __future__ import annotations

import torch

from . import geometry, hypotheses, regnets, regress, volume


def stage_train_red(features, cams, depth_values, sd, geo_model="rpc", sampler="aten"):
    """`compute_depth_when_train` (`networks/casred.py:10-64`): whole-volume build, RED
    regulariser over D, softmax, expectation, max-probability confidence."""
    var = volume.variance_cost_volume(features, cams, depth_values, geo_model, sampler)
    logits = regnets.red_regularization(var, sd)
    depth, conf = regress.softargmin_red(logits, depth_values)
    return {"depth": depth, "photometric_confidence": conf}


def stage_casmvs(features, cams, depth_values, sd, geo_model="rpc", sampler="aten"):
    """`DepthNet.forward` (`networks/casmvs.py:15-76`), eval mode: CostRegNet regulariser."""
    var = volume.variance_cost_volume(features, cams, depth_values, geo_model, sampler)
    logits = regnets.costregnet(var, sd, training=False).squeeze(1)
    depth, conf = regress.softargmin_casmvs(logits, depth_values)
    return {"depth": depth, "photometric_confidence": conf}


def stage_pred_red(features, cams, depth_values, sd, geo_model="rpc", sampler="aten"):
    """`compute_depth_when_pred` (`networks/casred.py:161-238`): plane-by-plane build ->
    slice regulariser (states carried) -> streaming fp64 soft-argmin."""
    V = len(features)
    cam = torch.unbind(cams, 1)
    ref = features[0]
    B, _, H, W = ref.shape
    states = regnets.red_initial_states(B, H, W)
    head = regress.StreamingSoftArgmin(B, H, W)
    for d in range(depth_values.shape[1]):
        plane = depth_values[:, d:d + 1]
        s = ref.unsqueeze(2)
        q = s ** 2
        for v in range(1, V):
            w = volume.warp_view(features[v], cam[v], cam[0], plane, geo_model, sampler)
            s = s + w
            q = q + w ** 2
        var = q.div_(V).sub_(s.div_(V).pow_(2)).squeeze(2)
        reg, *states = regnets.red_slice(var, *states, sd)
        head.update(reg, plane if plane.dim() == 4 else plane.view(B, 1, 1, 1))
    depth, conf = head.finish()
    return {"depth": depth, "photometric_confidence": conf}


def cascade(features_per_stage, cams_per_stage, depth_range, weights_per_stage, *, img_hw,
            ndepths=(48, 32, 8), ratios=(4, 2, 1), min_interval=2.5, scales=(4, 2, 1),
            geo_model="rpc", head="red_train", sampler="aten"):
    """The stage loop shared by `CascadeREDNet.forward` (`networks/casred.py:125-154`),
    `Infer_CascadeREDNet.forward` (`:296-331`) and `CascadeMVSNet.forward` (`casmvs.py:138-168`),
    starting from per-stage feature maps (FeatureNet is outside the path).

    features_per_stage[s] = list of V [B, C_s, H_s, W_s]; cams_per_stage[s] = [B, V, ...]."""
    fn = {"red_train": stage_train_red, "red_pred": stage_pred_red, "casmvs": stage_casmvs}[head]
    outputs, depth = {}, None
    for s, nd in enumerate(ndepths):
        dv = hypotheses.stage_hypotheses(depth, depth_range, nd, ratios[s] * min_interval, img_hw, scales[s])
        out = fn(features_per_stage[s], cams_per_stage[s], dv, weights_per_stage[s], geo_model, sampler)
        depth = out["depth"]
        outputs[f"stage{s + 1}"] = out
    outputs.update(out)
    return outputs
