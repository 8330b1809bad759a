"""Host-side mirror of `modules/depth_range.py` and of the cascade's resampling glue
(`networks/casred.py:132-145`) on the sm_100a kernel in csrc/hypotheses.cu.

    get_depth_range_samples(cur_depth, ndepth, depth_inteval_pixel, device, dtype, shape)   depth_range.py:23
    stage_depth_hypotheses(prev_depth, depth_values, ndepth, interval, img_hw, scale)       fused cascade glue
"""
from __future__ import annotations

import torch

from . import _lib


def _run(prev, rng, D, interval, Himg, Wimg, h, w, B, device):
    out = torch.empty((B, D, h, w), dtype=torch.float32, device=device)
    with torch.cuda.device(device):
        st = _lib.stream_ptr(device)
        for b in range(B):
            if prev is not None:
                rc = _lib.lib().satmvs_depth_hypotheses(prev[b].data_ptr(), prev.shape[1], prev.shape[2], None, 0, D,
                                                        float(interval), Himg, Wimg, h, w, out[b].data_ptr(), st)
            else:
                rc = _lib.lib().satmvs_depth_hypotheses(None, 0, 0, rng[b].data_ptr(), rng.shape[1], D, float(interval),
                                                        Himg, Wimg, h, w, out[b].data_ptr(), st)
            _lib.check(rc, "depth_hypotheses")
    return out


def get_depth_range_samples(cur_depth, ndepth, depth_inteval_pixel, device=None, dtype=None, shape=None):
    """`get_depth_range_samples` (`modules/depth_range.py:23-42`): cur_depth [B, 2+] (first stage) or
    [B, H, W]; returns [B, D, H, W] at the resolution given by `shape` = (B, H, W)."""
    cur = _lib.require_cuda(cur_depth, "cur_depth")
    B, H, W = shape
    if cur.dim() == 2:
        return _run(None, cur, ndepth, depth_inteval_pixel, H, W, H, W, B, cur.device)
    assert tuple(cur.shape) == tuple(shape), "cur_depth:{}, input shape:{}".format(cur.shape, shape)   # depth_range.py:13
    return _run(cur, None, ndepth, depth_inteval_pixel, H, W, H, W, B, cur.device)


def stage_depth_hypotheses(prev_depth, depth_values, ndepth, interval, img_hw, scale):
    """Hypotheses of one cascade stage directly on its grid: replaces
    `F.interpolate(prev)` + `get_depth_range_samples` + `F.interpolate(..., trilinear)`
    (`networks/casred.py:132-145`) without the full-resolution [B, D, Himg, Wimg] temporary.
    prev_depth [B, hp, wp] or None (first stage, uses depth_values [B, 2+]); returns [B, D, Himg/scale, Wimg/scale]."""
    Himg, Wimg = img_hw
    h, w = Himg // int(scale), Wimg // int(scale)
    if prev_depth is None:
        rng = _lib.require_cuda(depth_values, "depth_values")
        return _run(None, rng, ndepth, interval, Himg, Wimg, h, w, rng.shape[0], rng.device)
    prev = _lib.require_cuda(prev_depth, "prev_depth")
    return _run(prev, None, ndepth, interval, Himg, Wimg, h, w, prev.shape[0], prev.device)
