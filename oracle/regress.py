"""oracle.regress — soft-argmin depth regression and confidences.
TEST INFRASTRUCTURE (see oracle/__init__.py)."""
from __future__ import annotations

import torch
import torch.nn.functional as F


def depth_regression(p: torch.Tensor, depth_values: torch.Tensor) -> torch.Tensor:
    """`depth_regression` (`modules/module.py:433-439`): sum_d p*d; p [B, D, H, W];
    depth_values [B, D] or [B, D, H', W'] (bilinearly resized to p's H, W)."""
    if depth_values.dim() <= 2:
        dv = depth_values.view(*depth_values.shape, 1, 1)
    else:
        dv = F.interpolate(depth_values, [p.shape[2], p.shape[3]], mode="bilinear", align_corners=False)
    return torch.sum(p * dv, 1)


def softargmin_red(logits: torch.Tensor, depth_values: torch.Tensor):
    """RED train-style head (`networks/casred.py:58-62`): softmax over D, expectation, max-prob."""
    p = F.softmax(logits, dim=1)
    return depth_regression(p, depth_values), p.max(1)[0]


def softargmin_casmvs(logits: torch.Tensor, depth_values: torch.Tensor):
    """CasMVSNet head (`networks/casmvs.py:66-74`): confidence = sum of the 4 probabilities
    around the regressed plane index (pad 1 before, 2 after)."""
    D = logits.shape[1]
    p = F.softmax(logits, dim=1)
    depth = depth_regression(p, depth_values)
    sum4 = 4 * F.avg_pool3d(F.pad(p.unsqueeze(1), pad=(0, 0, 0, 0, 1, 2)), (4, 1, 1), stride=1, padding=0).squeeze(1)
    idx = depth_regression(p, torch.arange(D, dtype=torch.float, device=p.device)).long()   # casmvs.py:70 (device= p.device there)
    idx = idx.clamp(min=0, max=D - 1)
    return depth, torch.gather(sum4, 1, idx.unsqueeze(1)).squeeze(1)


class StreamingSoftArgmin:
    """Plane-by-plane head of the inference net (`networks/casred.py:182-184`, `:218-236`):
    fp64 running sum of e = exp(reg) (no max subtraction), sum d*e and max e."""

    def __init__(self, B: int, H: int, W: int):
        self.exp_sum = torch.zeros(B, 1, H, W, dtype=torch.float64)
        self.depth_acc = torch.zeros(B, 1, H, W, dtype=torch.float64)
        self.max_e = torch.zeros(B, 1, H, W, dtype=torch.float64)

    def update(self, reg: torch.Tensor, depth_plane: torch.Tensor) -> None:
        e = reg.double().exp()
        self.max_e = torch.where(self.max_e < e, e, self.max_e)
        self.depth_acc = depth_plane.double() * e + self.depth_acc
        self.exp_sum = self.exp_sum + e

    def finish(self):
        tot = self.exp_sum + 1e-10
        return (self.depth_acc / tot).squeeze(1).float(), (self.max_e / tot).squeeze(1).float()


class StreamingSoftArgminState(StreamingSoftArgmin):
    """The streaming head with the product's packed state [B, 3, H, W] fp64 = (sum e, sum d*e, max e) and the regulariser-free
    matching cost reg_k = scale * mean_c var[:, c, k] of a variance slab (what a depth-sharded sweep exchanges).
    Stand-in for `satmvs_b200.regress.StreamingSoftArgmin` in the CPU (gloo) tests of `sharded.sweep_depth_sharded`."""

    def __init__(self, B: int, H: int, W: int):
        self.state = torch.zeros(B, 3, H, W, dtype=torch.float64)

    def update_volume(self, var: torch.Tensor, depth_planes: torch.Tensor, scale: float = -1.0) -> None:
        B, C, K, H, W = var.shape
        for k in range(K):
            s = torch.zeros(B, H, W, dtype=torch.float32)
            for c in range(C):                       # channels in ascending order, fp32
                s = s + var[:, c, k]
            reg = (s * (1.0 / C)) * scale
            e = reg.double().exp()
            dp = depth_planes[:, k].double() if depth_planes.dim() == 4 else depth_planes[:, k].double().view(B, 1, 1)
            self.state[:, 0] = self.state[:, 0] + e
            self.state[:, 1] = dp * e + self.state[:, 1]
            self.state[:, 2] = torch.where(self.state[:, 2] < e, e, self.state[:, 2])

    def finish(self):
        tot = self.state[:, 0] + 1e-10
        return (self.state[:, 1] / tot).float(), (self.state[:, 2] / tot).float()
