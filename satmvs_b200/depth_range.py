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


class _SamplesFn(torch.autograd.Function):
    """Hypotheses around a depth map: sample_d = cur + (d * nd / (nd - 1) - nd / 2) * interval, so d sample_d / d cur = 1 and the
    gradient to cur_depth is the sum over planes.  `CascadeREDNet.forward` (`networks/casred.py:132-154`) does NOT detach the
    previous stage's depth: the later stages' losses reach the earlier regularisers through the hypotheses."""

    @staticmethod
    def forward(ctx, cur, ndepth, interval, shape):
        B, H, W = shape
        ctx.dims = (B, ndepth, H, W)
        return _run(cur.detach(), None, ndepth, interval, H, W, H, W, B, cur.device)

    @staticmethod
    def backward(ctx, g):
        B, D, H, W = ctx.dims
        g = g.contiguous().float()
        out = torch.empty((B, H, W), dtype=torch.float32, device=g.device)
        ones = torch.ones(D, dtype=torch.float32, device=g.device)
        with torch.cuda.device(g.device):
            st = _lib.stream_ptr(g.device)
            for b in range(B):     # sum over planes = depth_regression of the gradient planes with unit "depths" (heads.cu, mode 2)
                _lib.check(_lib.lib().satmvs_softargmin_fwd(g[b].data_ptr(), ones.data_ptr(), 0, 2, D, H, W, out[b].data_ptr(),
                                                           None, st), "depth_range_samples_bwd")
        return out, None, None, None


def get_depth_range_samples(cur_depth, ndepth, depth_inteval_pixel, device=None, dtype=None, shape=None):
    """`get_depth_range_samples` (`modules/depth_range.py:23-42`): cur_depth [B, 2+] (first stage) or
    [B, H, W]; returns [B, D, H, W] at the resolution given by `shape` = (B, H, W).  Carries a gradient to a [B, H, W] cur_depth."""
    cur = _lib.require_cuda(cur_depth, "cur_depth")
    B, H, W = shape
    if cur.dim() == 2:
        return _run(None, cur, ndepth, depth_inteval_pixel, H, W, H, W, B, cur.device)
    assert tuple(cur.shape) == tuple(shape), "cur_depth:{}, input shape:{}".format(cur.shape, shape)   # depth_range.py:13
    if torch.is_grad_enabled() and cur.requires_grad:
        return _SamplesFn.apply(cur, ndepth, depth_inteval_pixel, tuple(shape))
    return _run(cur, None, ndepth, depth_inteval_pixel, H, W, H, W, B, cur.device)


def stage_depth_hypotheses(prev_depth, depth_values, ndepth, interval, img_hw, scale):
    """Hypotheses of one cascade stage directly on its grid: replaces
    `F.interpolate(prev)` + `get_depth_range_samples` + `F.interpolate(..., trilinear)`
    (`networks/casred.py:132-145`) without the full-resolution [B, D, Himg, Wimg] temporary.
    prev_depth [B, hp, wp] or None (first stage, uses depth_values [B, 2+]); returns [B, D, Himg/scale, Wimg/scale].
    No gradient flows to prev_depth (`grad_method="detach"`, the default of `CascadeMVSNet`, `networks/casmvs.py:145-146`); the
    non-detached cascade of `CascadeREDNet` in training goes through `get_depth_range_samples` + torch's `F.interpolate`."""
    Himg, Wimg = img_hw
    h, w = Himg // int(scale), Wimg // int(scale)
    if prev_depth is None:
        rng = _lib.require_cuda(depth_values, "depth_values")
        return _run(None, rng, ndepth, interval, Himg, Wimg, h, w, rng.shape[0], rng.device)
    prev = _lib.require_cuda(prev_depth, "prev_depth")
    return _run(prev, None, ndepth, interval, Himg, Wimg, h, w, prev.shape[0], prev.device)
