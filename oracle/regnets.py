"""oracle.regnets — functional CPU restatement of the two cost-volume regularisers.
TEST INFRASTRUCTURE (see oracle/__init__.py).

Weights are passed as a flat dict with the reference's `state_dict()` names, so the same
dict loads into the reference modules (`make_golden.py`) and into `satmvs_b200.module`.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

BN_EPS = 1e-5   # nn.BatchNorm3d default, `modules/module.py:348`
GN_EPS = 1e-5   # nn.GroupNorm(1, C, 1e-5, True), `modules/module.py:15-20`


# ---------------------------------------------------------------------------------------
# CostRegNet: 3-D conv UNet, `modules/module.py:546-577`
# ---------------------------------------------------------------------------------------
def _bn_relu(x, sd, name, training):
    x = F.batch_norm(x, None if training else sd[f"{name}.bn.running_mean"],
                     None if training else sd[f"{name}.bn.running_var"],
                     sd[f"{name}.bn.weight"], sd[f"{name}.bn.bias"], training=training, eps=BN_EPS)
    return F.relu(x)


def _conv3d_block(x, sd, name, stride=1, training=False):
    # Conv3d(bias=False) + BatchNorm3d + ReLU, `modules/module.py:354-360`
    return _bn_relu(F.conv3d(x, sd[f"{name}.conv.weight"], stride=stride, padding=1), sd, name, training)


def _deconv3d_block(x, sd, name, training=False):
    # ConvTranspose3d(stride 2, pad 1, output_padding 1, bias=False) + BN + ReLU, `module.py:398-404`
    y = F.conv_transpose3d(x, sd[f"{name}.conv.weight"], stride=2, padding=1, output_padding=1)
    return _bn_relu(y, sd, name, training)


def costregnet(x: torch.Tensor, sd: dict, training: bool = False) -> torch.Tensor:
    """`CostRegNet.forward` (`modules/module.py:568-577`): [B, Cin, D, H, W] -> [B, 1, D, H, W]."""
    c0 = _conv3d_block(x, sd, "conv0", 1, training)
    c2 = _conv3d_block(_conv3d_block(c0, sd, "conv1", 2, training), sd, "conv2", 1, training)
    c4 = _conv3d_block(_conv3d_block(c2, sd, "conv3", 2, training), sd, "conv4", 1, training)
    y = _conv3d_block(_conv3d_block(c4, sd, "conv5", 2, training), sd, "conv6", 1, training)
    y = c4 + _deconv3d_block(y, sd, "conv7", training)
    y = c2 + _deconv3d_block(y, sd, "conv9", training)
    y = c0 + _deconv3d_block(y, sd, "conv11", training)
    return F.conv3d(y, sd["prob.weight"], padding=1)


# ---------------------------------------------------------------------------------------
# RED regulariser: 2-D conv-GRU UNet recurring over depth, `modules/module.py:595-693`
# ---------------------------------------------------------------------------------------
def conv_gru(x, h, sd, name):
    """`ConvGRUCell2.forward` (`modules/module.py:24-58`): returns the new hidden state."""
    C = h.shape[1]
    f = F.conv2d(torch.cat((x, h), 1), sd[f"{name}.gate_conv.weight"], sd[f"{name}.gate_conv.bias"], padding=1)
    r = torch.sigmoid(F.group_norm(f[:, :C], 1, sd[f"{name}.reset_gate_norm.weight"],
                                   sd[f"{name}.reset_gate_norm.bias"], GN_EPS))
    u = torch.sigmoid(F.group_norm(f[:, C:], 1, sd[f"{name}.update_gate_norm.weight"],
                                   sd[f"{name}.update_gate_norm.bias"], GN_EPS))
    o = F.conv2d(torch.cat((x, r * h), 1), sd[f"{name}.output_conv.weight"],
                 sd[f"{name}.output_conv.bias"], padding=1)
    y = torch.tanh(F.group_norm(o, 1, sd[f"{name}.output_norm.weight"], sd[f"{name}.output_norm.bias"], GN_EPS))
    return u * h + (1 - u) * y


def _down(x, sd, name):   # ConvReLU stride 2, `modules/module.py:178-184`
    return F.relu(F.conv2d(x, sd[f"{name}.conv.weight"], stride=2, padding=1))


def _up(x, sd, name):     # ConvTransReLU stride 2, `modules/module.py:208-215`
    return F.relu(F.conv_transpose2d(x, sd[f"{name}.conv.weight"], stride=2, padding=1, output_padding=1))


def red_slice(cost, s1, s2, s3, s4, sd):
    """`slice_RED_Regularization.forward` (`modules/module.py:672-693`): one depth slice.
    cost [B, C, H, W]; states [B, 8|16|32|64, H/1|2|4|8, W/...].  Returns (reg [B,1,H,W], s1..s4)."""
    x = -cost
    e1 = _down(x, sd, "conv1")
    e2 = _down(e1, sd, "conv2")
    e3 = _down(e2, sd, "conv3")
    s4 = conv_gru(e3, s4, sd, "conv_gru4")
    s3 = conv_gru(e2, s3, sd, "conv_gru3")
    u3 = _up(s4, sd, "upconv3") + s3
    s2 = conv_gru(e1, s2, sd, "conv_gru2")
    u2 = _up(u3, sd, "upconv2") + s2
    s1 = conv_gru(x, s1, sd, "conv_gru1")
    u1 = _up(u2, sd, "upconv1") + s1
    reg = F.conv_transpose2d(u1, sd["upconv2d.weight"], sd["upconv2d.bias"], stride=1, padding=1)
    return reg, s1, s2, s3, s4


def red_initial_states(B: int, H: int, W: int):
    # hard-coded 8/16/32/64 channels, `modules/module.py:617-620`, `networks/casred.py:176-179`
    return (torch.zeros(B, 8, H, W), torch.zeros(B, 16, H // 2, W // 2),
            torch.zeros(B, 32, H // 4, W // 4), torch.zeros(B, 64, H // 8, W // 8))


def red_regularization(volume: torch.Tensor, sd: dict) -> torch.Tensor:
    """`RED_Regularization.forward` (`modules/module.py:614-649`): [B, C, D, H, W] -> [B, D, H, W]."""
    B, _, D, H, W = volume.shape
    states = red_initial_states(B, H, W)
    out = []
    for d in range(D):
        reg, *states = red_slice(volume[:, :, d], *states, sd)
        out.append(reg)
    return torch.stack(out, dim=1).squeeze(2)


def featurenet(x: torch.Tensor, sd: dict, training: bool = False) -> dict:
    """`FeatureNet.forward` (`modules/module.py:506-543`), arch_mode "unet", three stages; BatchNorm on running statistics, or on
    batch statistics when `training` (`Conv2d` / `Deconv2d` blocks `:78-159`, `DeConv2dFuse` `:303-321`).
    x [B,3,H,W] -> {"stage1", "stage2", "stage3"}."""
    import torch.nn.functional as F

    def bn(y, name):
        if training:
            return F.batch_norm(y, None, None, sd[name + ".bn.weight"], sd[name + ".bn.bias"], True, 0.1, 1e-5)
        return F.batch_norm(y, sd[name + ".bn.running_mean"], sd[name + ".bn.running_var"], sd[name + ".bn.weight"],
                            sd[name + ".bn.bias"], False, 0.1, 1e-5)

    def conv(y, name, stride=1, pad=1):
        return F.relu(bn(F.conv2d(y, sd[name + ".conv.weight"], None, stride, pad), name))

    def fuse(x_pre, y, name):
        h, w = y.shape[2:]
        d = F.conv_transpose2d(y, sd[name + ".deconv.conv.weight"], None, 2, 1, 1)[:, :, :2 * h, :2 * w]
        d = F.relu(bn(d, name + ".deconv"))
        return conv(torch.cat((d, x_pre), 1), name + ".conv")

    conv0 = conv(conv(x, "conv0.0"), "conv0.1")
    conv1 = conv(conv(conv(conv0, "conv1.0", 2, 2), "conv1.1"), "conv1.2")
    conv2 = conv(conv(conv(conv1, "conv2.0", 2, 2), "conv2.1"), "conv2.2")
    out = {"stage1": F.conv2d(conv2, sd["out1.weight"])}
    f1 = fuse(conv1, conv2, "deconv1")
    out["stage2"] = F.conv2d(f1, sd["out2.weight"])
    f2 = fuse(conv0, f1, "deconv2")
    out["stage3"] = F.conv2d(f2, sd["out3.weight"])
    return out
