"""oracle.hypotheses — depth-hypothesis generation and the cascade resampling glue.
TEST INFRASTRUCTURE (see oracle/__init__.py)."""
from __future__ import annotations

import torch
import torch.nn.functional as F


def depth_range_samples(cur_depth: torch.Tensor, ndepth: int, interval: float,
                        shape: tuple[int, int, int]) -> torch.Tensor:
    """`get_depth_range_samples` (`modules/depth_range.py:23-42`, `:4-20`) -> [B, D, Himg, Wimg].

    cur_depth [B, 2+] (first stage: uniform planes between column 0 and the last column,
    inclusive) or [B, Himg, Wimg] (later stages: D planes centred on the previous depth,
    min = cur - D/2*I, max = cur + D/2*I, step (max-min)/(D-1))."""
    k = torch.arange(0, ndepth, dtype=cur_depth.dtype)
    if cur_depth.dim() == 2:
        lo, hi = cur_depth[:, 0], cur_depth[:, -1]
        step = (hi - lo) / (ndepth - 1)
        planes = lo.unsqueeze(1) + k.reshape(1, -1) * step.unsqueeze(1)
        return planes.unsqueeze(-1).unsqueeze(-1).repeat(1, 1, shape[1], shape[2])
    assert tuple(cur_depth.shape) == tuple(shape)
    lo = cur_depth - ndepth / 2 * interval
    hi = cur_depth + ndepth / 2 * interval
    step = (hi - lo) / (ndepth - 1)
    return lo.unsqueeze(1) + k.reshape(1, -1, 1, 1) * step.unsqueeze(1)


def stage_hypotheses(prev_depth, depth_range: torch.Tensor, ndepth: int, interval: float,
                     img_hw: tuple[int, int], scale: int) -> torch.Tensor:
    """Cascade glue of `networks/casred.py:132-145`: previous depth bilinearly up-sampled to image
    resolution, hypotheses generated there, then trilinearly resized (align_corners=False) to the
    stage's [D, Himg/scale, Wimg/scale]."""
    Himg, Wimg = img_hw
    B = depth_range.shape[0]
    if prev_depth is None:
        cur = depth_range
    else:
        cur = F.interpolate(prev_depth.unsqueeze(1), [Himg, Wimg], mode="bilinear", align_corners=False).squeeze(1)
    samples = depth_range_samples(cur, ndepth, interval, (B, Himg, Wimg))
    dv = F.interpolate(samples.unsqueeze(1), [ndepth, Himg // scale, Wimg // scale],
                       mode="trilinear", align_corners=False)
    return dv.squeeze(1)
