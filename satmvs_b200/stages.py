"""One cascade stage on the GPU kernels: features in, depth + confidence out.

Mirrors `compute_depth_when_train` (`networks/casred.py:10-64`), `compute_depth_when_pred`
(`networks/casred.py:161-238`) and `DepthNet.forward` (`networks/casmvs.py:15-76`) from the feature
maps onward; FeatureNet is outside the path.
"""
from __future__ import annotations

import torch

from .depth_range import stage_depth_hypotheses
from .regress import StreamingSoftArgmin, softargmin
from .warping import build_cost_volume


def _split(features, cams):
    if len(features) != cams.shape[1]:
        raise AssertionError("Different number of images and projection matrices")      # casred.py:15
    return features[0], list(features[1:]), cams[:, 0], [cams[:, v] for v in range(1, cams.shape[1])]


def stage_train_red(features, cams, depth_values, regulariser, geo_model="rpc"):
    """Whole-volume stage: fused sweep -> RED regulariser over D -> softmax/expectation/max-prob.
    features: V x [B,C,H,W] CUDA; cams [B,V,170] or [B,V,4,4] f64 (host preferred); depth_values
    [B,D] or [B,D,H,W]; regulariser: `satmvs_b200.module.RED_Regularization`."""
    ref, srcs, ref_cam, src_cams = _split(features, cams)
    var = build_cost_volume(ref, srcs, ref_cam, src_cams, depth_values, geo_model)
    logits = regulariser(var)
    if logits.requires_grad:      # loss.backward() (train.py:284): head with a gradient to the logits
        from .training import softargmin_train
        depth, conf = softargmin_train(logits, depth_values, "red")
    else:
        depth, conf = softargmin(logits, depth_values, "red")
    return {"depth": depth, "photometric_confidence": conf}


def stage_casmvs(features, cams, depth_values, regulariser, geo_model="rpc"):
    """CasMVSNet stage: fused sweep -> CostRegNet -> softmax/expectation/4-neighbour confidence."""
    ref, srcs, ref_cam, src_cams = _split(features, cams)
    var = build_cost_volume(ref, srcs, ref_cam, src_cams, depth_values, geo_model)
    logits = regulariser(var).squeeze(1)
    if logits.requires_grad:      # train(): the head carries a gradient to the logits (confidence has none, casmvs.py:69)
        from .training import softargmin_casmvs_train
        depth, conf = softargmin_casmvs_train(logits, depth_values)
    else:
        depth, conf = softargmin(logits, depth_values, "casmvs")
    return {"depth": depth, "photometric_confidence": conf}


def stage_pred_red(features, cams, depth_values, regulariser, geo_model="rpc", chunk=None):
    """Plane-streaming stage of the inference net (`compute_depth_when_pred`, `networks/casred.py:161-238`): sweep, one
    recurrent regulariser step per plane (states carried), streaming fp64 soft-argmin -- in chunks of `chunk` planes: ONE
    fused sweep builds the chunk's variance planes, ONE library call runs the recurrence over them (tensor-core cluster
    kernel, states in / out), ONE launch folds them into the fp64 running sums.  Memory stays O(chunk * C * H * W) whatever D
    is; chunk = 1 is the reference's literal per-plane loop.  regulariser: `satmvs_b200.module.slice_RED_Regularization`."""
    ref, srcs, ref_cam, src_cams = _split(features, cams)
    B, _, H, W = ref.shape
    dev = ref.device
    states = [torch.zeros((B, c, H >> l, W >> l), dtype=torch.float32, device=dev) for l, c in enumerate((8, 16, 32, 64))]
    head = StreamingSoftArgmin(B, H, W, dev)
    D = depth_values.shape[1]
    if chunk is None:      # planes per library call: SATMVS_PRED_CHUNK, default 32 (>= 32 planes per call
        # let the library overlap the batched convs with the recurrence inside the call: 2.36 -> 2.07 ms on the 256x128 cascade)
        import os
        chunk = int(os.environ.get("SATMVS_PRED_CHUNK", "32"))
    chunk = max(1, int(chunk))
    for d0 in range(0, D, chunk):
        planes = depth_values[:, d0:d0 + chunk].contiguous()
        var = build_cost_volume(ref, srcs, ref_cam, src_cams, planes, geo_model)
        if planes.shape[1] == 1:
            reg, *states = regulariser(var.squeeze(2), *states)
            head.update(reg, planes)
        else:
            reg, *states = regulariser.forward_planes(var, *states)
            head.update_planes(reg, planes)
    depth, conf = head.finish()
    return {"depth": depth, "photometric_confidence": conf}


_STAGE_FNS = {"red_train": stage_train_red, "red_pred": stage_pred_red, "casmvs": stage_casmvs}


def cascade(features_per_stage, cams_per_stage, depth_values, regularisers, *, img_hw, ndepths=(48, 32, 8),
            depth_interals_ratio=(4, 2, 1), min_interval=2.5, scales=(4, 2, 1), geo_model="rpc", head="red_train"):
    """The stage loop of `CascadeREDNet.forward` (`networks/casred.py:125-154`), `Infer_CascadeREDNet.forward`
    (`:296-331`) and `CascadeMVSNet.forward` (`networks/casmvs.py:138-168`) from per-stage feature maps on.
    Returns {"stage1": {...}, ..., "depth", "photometric_confidence"} like the reference."""
    fn = _STAGE_FNS[head]
    outputs, depth, out = {}, None, None
    for s, nd in enumerate(ndepths):
        dv = stage_depth_hypotheses(depth, depth_values, nd, depth_interals_ratio[s] * min_interval, img_hw, scales[s])
        out = fn(features_per_stage[s], cams_per_stage[s], dv, regularisers[s], geo_model)
        depth = out["depth"]
        outputs[f"stage{s + 1}"] = out
    outputs.update(out)
    return outputs
