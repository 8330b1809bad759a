"""Soft-argmin heads on the sm_100a kernels (csrc/heads.cu).

    depth_regression(p, depth_values)                       modules/module.py:433   (kept for API parity)
    softargmin(logits, depth_values, head)                  casred.py:58-62 / casmvs.py:66-74, fused
    StreamingSoftArgmin                                     casred.py:182-184, :218-236 (fp64, plane by plane)
"""
from __future__ import annotations

import torch

from . import _lib


def resize_bilinear(x: torch.Tensor, H: int, W: int) -> torch.Tensor:
    """`F.interpolate(x, [H, W], mode="bilinear", align_corners=False)` for x [B,N,h,w] on the library kernel."""
    x = _lib.require_cuda(x, "x")
    B, N, h, w = x.shape
    out = torch.empty((B, N, H, W), dtype=torch.float32, device=x.device)
    with torch.cuda.device(x.device):
        _lib.check(_lib.lib().satmvs_resize_bilinear(x.data_ptr(), B * N, h, w, H, W, out.data_ptr(), _lib.stream_ptr(x.device)),
                   "resize_bilinear")
    return out


def _depth_regression_fwd(p, dv, per_pixel):
    B, D, H, W = p.shape
    depth = torch.empty((B, H, W), dtype=torch.float32, device=p.device)
    with torch.cuda.device(p.device):
        st = _lib.stream_ptr(p.device)
        for b in range(B):
            _lib.check(_lib.lib().satmvs_softargmin_fwd(p[b].data_ptr(), dv[b].data_ptr(), per_pixel, 2, D, H, W,
                                                       depth[b].data_ptr(), None, st), "depth_regression")
    return depth


class _DepthRegressionFn(torch.autograd.Function):
    """sum_d p*d with its gradient to the probabilities, d depth / d p_d = depth_values_d (`casmvs.py:66-68` trains through it)."""

    @staticmethod
    def forward(ctx, p, dv, per_pixel):
        ctx.save_for_backward(dv)
        ctx.cfg = (per_pixel, tuple(p.shape))
        return _depth_regression_fwd(p.detach(), dv, per_pixel)

    @staticmethod
    def backward(ctx, g):
        (dv,) = ctx.saved_tensors
        per_pixel, (B, D, H, W) = ctx.cfg
        g = g.contiguous().float()
        out = torch.empty((B, D, H, W), dtype=torch.float32, device=g.device)
        dvf = dv if per_pixel else dv.reshape(B, D, 1, 1).expand(B, D, H, W).contiguous()
        with torch.cuda.device(g.device):
            st = _lib.stream_ptr(g.device)
            for b in range(B):     # out[d] = depth_values[d] * g  (g broadcast over the planes: channel stride 0)
                _lib.check(_lib.lib().satmvs_elementwise(dvf[b].data_ptr(), H * W, None, 0, g[b].data_ptr(), 0, None, 0, None, 0, 1.0,
                                                         out[b].data_ptr(), H * W, D, H * W, st), "depth_regression_bwd")
        return out, None, None


def depth_regression(p: torch.Tensor, depth_values: torch.Tensor) -> torch.Tensor:
    """`depth_regression` (`modules/module.py:433-439`): sum_d p*d on given probabilities p [B,D,H,W];
    depth_values [B,D] or [B,D,h,w] (resized bilinearly to p's grid like the reference).  Carries a gradient to p."""
    p = _lib.require_cuda(p, "p")
    B, D, H, W = p.shape
    dv = _lib.require_cuda(depth_values.detach(), "depth_values")
    per_pixel = 0
    if dv.dim() > 2:
        per_pixel = 1
        if tuple(dv.shape[2:]) != (H, W):
            dv = resize_bilinear(dv, H, W)
    if torch.is_grad_enabled() and p.requires_grad:
        return _DepthRegressionFn.apply(p, dv, per_pixel)
    return _depth_regression_fwd(p, dv, per_pixel)


def softargmin(logits: torch.Tensor, depth_values: torch.Tensor, head: str = "red"):
    """softmax over D + expectation + confidence in one kernel.
    logits [B,D,H,W]; depth_values [B,D] or [B,D,H,W]; head 'red' (conf = max p) or 'casmvs'
    (conf = sum of 4 neighbouring probabilities).  Returns (depth [B,H,W], conf [B,H,W])."""
    mode = {"red": 0, "casmvs": 1}[head]
    lg = _lib.require_cuda(logits, "logits")
    B, D, H, W = lg.shape
    dv = _lib.require_cuda(depth_values, "depth_values")
    if dv.dim() == 2:
        per_pixel = 0
    elif tuple(dv.shape) == (B, D, H, W):
        per_pixel = 1
    else:
        dv = resize_bilinear(dv, H, W)      # depth_regression resizes hypotheses given at another resolution (module.py:437)
        per_pixel = 1
    depth = torch.empty((B, H, W), dtype=torch.float32, device=lg.device)
    conf = torch.empty_like(depth)
    with torch.cuda.device(lg.device):
        st = _lib.stream_ptr(lg.device)
        for b in range(B):
            _lib.check(_lib.lib().satmvs_softargmin_fwd(lg[b].data_ptr(), dv[b].data_ptr(), per_pixel, mode, D, H, W,
                                                       depth[b].data_ptr(), conf[b].data_ptr(), st), "softargmin_fwd")
    return depth, conf


class StreamingSoftArgmin:
    """Running fp64 (sum e, sum d*e, max e) with e = exp(reg), no max subtraction — the
    plane-by-plane head of `compute_depth_when_pred` (`networks/casred.py:218-236`)."""

    def __init__(self, B: int, H: int, W: int, device):
        self.B, self.H, self.W = B, H, W
        self.state = torch.zeros((B, 3, H, W), dtype=torch.float64, device=device)

    def update(self, reg: torch.Tensor, depth_plane: torch.Tensor) -> None:
        reg = _lib.require_cuda(reg, "reg").view(self.B, self.H, self.W)
        dp = _lib.require_cuda(depth_plane, "depth_plane")
        per_pixel = 1 if dp.numel() == self.B * self.H * self.W else 0
        dp = dp.reshape(self.B, -1)
        with torch.cuda.device(reg.device):
            st = _lib.stream_ptr(reg.device)
            for b in range(self.B):
                _lib.check(_lib.lib().satmvs_softargmin_stream_update(
                    reg[b].data_ptr(), dp[b].data_ptr(), per_pixel, self.H, self.W, self.state[b].data_ptr(), st),
                    "softargmin_stream_update")

    def update_planes(self, reg: torch.Tensor, depth_planes: torch.Tensor) -> None:
        """K planes at once: reg [B,K,H,W]; depth_planes [B,K,H,W] or [B,K].  Same arithmetic, plane by plane in order."""
        reg = _lib.require_cuda(reg, "reg")
        K = reg.shape[1]
        dp = _lib.require_cuda(depth_planes, "depth_planes")
        per_pixel = 1 if dp.dim() == 4 else 0
        with torch.cuda.device(reg.device):
            st = _lib.stream_ptr(reg.device)
            for b in range(self.B):
                _lib.check(_lib.lib().satmvs_softargmin_stream_update_planes(
                    reg[b].data_ptr(), dp[b].data_ptr(), per_pixel, K, self.H, self.W, self.state[b].data_ptr(), st),
                    "softargmin_stream_update_planes")

    def update_volume(self, var: torch.Tensor, depth_planes: torch.Tensor, scale: float = -1.0) -> None:
        """Plane sweep without a regulariser: reg_k = scale * mean_c var[:, c, k] folded plane by plane (var [B,C,K,H,W])."""
        var = _lib.require_cuda(var, "var")
        Cc, K = var.shape[1], var.shape[2]
        dp = _lib.require_cuda(depth_planes, "depth_planes")
        per_pixel = 1 if dp.dim() == 4 else 0
        with torch.cuda.device(var.device):
            st = _lib.stream_ptr(var.device)
            for b in range(self.B):
                _lib.check(_lib.lib().satmvs_softargmin_stream_update_volume(
                    var[b].data_ptr(), dp[b].data_ptr(), per_pixel, float(scale), Cc, K, self.H, self.W,
                    self.state[b].data_ptr(), st), "softargmin_stream_update_volume")

    def finish(self):
        depth = torch.empty((self.B, self.H, self.W), dtype=torch.float32, device=self.state.device)
        conf = torch.empty_like(depth)
        with torch.cuda.device(depth.device):
            st = _lib.stream_ptr(depth.device)
            for b in range(self.B):
                _lib.check(_lib.lib().satmvs_softargmin_stream_finish(
                    self.state[b].data_ptr(), self.H, self.W, depth[b].data_ptr(), conf[b].data_ptr(), st),
                    "softargmin_stream_finish")
        return depth, conf
