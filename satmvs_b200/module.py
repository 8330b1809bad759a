"""Host-side mirror of the reference's regulariser modules (`modules/module.py`) on the sm_100a
kernels of libsatmvs_b200.so.

The classes keep the reference's constructor signatures, sub-module names and parameter shapes, so
`load_state_dict` accepts a reference checkpoint unchanged (`train.py:216-219`); the sub-modules are
parameter containers only — `forward` hands raw pointers to the C ABI and never calls a torch conv.

    RED_Regularization(in_channels, base_channels=8).forward(volume)            modules/module.py:595-649
    slice_RED_Regularization(...).forward(cost, s1, s2, s3, s4)                 modules/module.py:653-693
    CostRegNet(in_channels, base_channels).forward(x)                           modules/module.py:546-577
    depth_regression(p, depth_values)                                           modules/module.py:433-439
"""
from __future__ import annotations

import ctypes as C

import torch
import torch.nn as nn

from . import _lib

_F = C.c_void_p


class _RedWeights(C.Structure):
    _fields_ = [(n, _F * 4) for n in ("gate_w", "gate_b", "out_w", "out_b", "rn_w", "rn_b", "un_w", "un_b", "on_w", "on_b")] + \
               [("conv_w", _F * 3), ("upconv_w", _F * 3), ("upconv2d_w", _F), ("upconv2d_b", _F)]


_WORKSPACES: dict = {}


def _workspace(nbytes: int, device) -> torch.Tensor:
    """Caller-owned scratch for the C ABI, cached per device and grown on demand."""
    key = (device.type, device.index)
    ws = _WORKSPACES.get(key)
    if ws is None or ws.numel() < nbytes:
        ws = torch.empty(nbytes, dtype=torch.uint8, device=device)
        _WORKSPACES[key] = ws
    return ws


def _param_key(module: nn.Module):
    """Identity of a module's parameters and buffers: (storage pointer, in-place version) of each.  Derived tensors (folded
    BatchNorm, packed tensor-core weights) are cached on this key and recomputed when a checkpoint is loaded, the module
    is moved, or an optimiser step modifies a parameter in place."""
    return tuple((t.data_ptr(), t._version) for t in list(module.parameters()) + list(module.buffers()))


def _fold_bn(blk):
    scale = blk.bn.weight.detach() * torch.rsqrt(blk.bn.running_var + blk.bn.eps)
    shift = blk.bn.bias.detach() - blk.bn.running_mean * scale
    return scale.contiguous(), shift.contiguous()


_ENGINES = {"auto": 0, "tcgen05": 1, "ffma": 2, "tcgen05_tn": 3}


def conv_block(x, weight, scale=None, shift=None, *, stride=1, relu=False, acc_scale=1.0, engine="auto"):
    """One convolution block of the regularisers on the library's kernels: `ConvReLU` (`modules/module.py:178-186`,
    weight [Cout,Cin,3,3], applied to every depth plane of x [B,Cin,D,H,W] or to x [B,Cin,H,W]) or `Conv3d`
    (`modules/module.py:324-366`, weight [Cout,Cin,3,3,3], x [B,Cin,D,H,W]); padding 1, per-channel scale / shift (folded
    BatchNorm or bias) and optional ReLU.  engine: "auto", "tcgen05" (tensor cores, 3xTF32 split; raises when the shape
    does not fit) or "ffma" (fp32 FFMA kernels).  Inference only (no autograd)."""
    nz = 3 if weight.dim() == 5 else 1
    squeeze = x.dim() == 4
    if squeeze:
        if nz == 3:
            raise ValueError("a 3x3x3 weight needs a 5-D input")
        x = x.unsqueeze(2)
    x = _lib.require_cuda(x, "x")
    w = _lib.require_cuda(weight, "weight")
    B, Cin, D, H, W = x.shape
    Cout = w.shape[0]
    if w.shape[1] != Cin or tuple(w.shape[2:]) != (3,) * (3 if nz == 3 else 2):
        want = "3, 3, 3" if nz == 3 else "3, 3"
        raise ValueError(f"weight must be [Cout, {Cin}, {want}], got {tuple(w.shape)}")
    sc = None if scale is None else _lib.require_cuda(scale, "scale")
    sh = None if shift is None else _lib.require_cuda(shift, "shift")
    Do = D // stride if nz == 3 else D
    out = torch.empty((B, Cout, Do, H // stride, W // stride), dtype=torch.float32, device=x.device)
    nbytes = _lib.lib().satmvs_conv_workspace_bytes(Cin, Cout, nz)
    ws = torch.empty(max(int(nbytes), 1), dtype=torch.uint8, device=x.device)   # not the shared cache: the packed weights live here
    with torch.cuda.device(x.device):
        st = _lib.stream_ptr(x.device)
        for b in range(B):
            _lib.check(_lib.lib().satmvs_conv_forward(
                x[b].data_ptr(), Cin, D, H, W, w.data_ptr(), sc.data_ptr() if sc is not None else None,
                sh.data_ptr() if sh is not None else None, Cout, nz, stride, int(bool(relu)), float(acc_scale), out[b].data_ptr(),
                _ENGINES[engine], ws.data_ptr(), ws.numel(), st), "conv_forward")
    return out.squeeze(2) if squeeze else out


class ConvGRUCell2(nn.Module):
    """Parameter container with the reference's names (`modules/module.py:6-22`)."""

    def __init__(self, input_channel, output_channel, kernel_size):
        super().__init__()
        k = input_channel + output_channel
        self.output_channel = output_channel
        self.gate_conv = nn.Conv2d(k, output_channel * 2, kernel_size, padding=1)
        self.reset_gate_norm = nn.GroupNorm(1, output_channel, 1e-5, True)
        self.update_gate_norm = nn.GroupNorm(1, output_channel, 1e-5, True)
        self.output_conv = nn.Conv2d(k, output_channel, kernel_size, padding=1)
        self.output_norm = nn.GroupNorm(1, output_channel, 1e-5, True)


class ConvReLU(nn.Module):
    def __init__(self, in_channels, out_channels, kernel_size=3, stride=1, pad=1):
        super().__init__()
        self.conv = nn.Conv2d(in_channels, out_channels, kernel_size, stride=stride, padding=pad, bias=False)


class ConvTransReLU(nn.Module):
    def __init__(self, in_channels, out_channels, kernel_size=3, stride=1, pad=1, output_pad=1):
        super().__init__()
        self.conv = nn.ConvTranspose2d(in_channels, out_channels, kernel_size=kernel_size, stride=stride, padding=pad,
                                       output_padding=output_pad, bias=False)


class _RedBase(nn.Module):
    def __init__(self, in_channels, base_channels=8):
        super().__init__()
        if base_channels != 8:
            # the reference hard-codes hidden states of 8/16/32/64 channels (module.py:617-620)
            raise ValueError("RED regulariser: base_channels must be 8")
        b = base_channels
        self.in_channels = in_channels
        self.base_channels = b
        self.conv_gru1 = ConvGRUCell2(in_channels, b, 3)
        self.conv_gru2 = ConvGRUCell2(b * 2, b * 2, 3)
        self.conv_gru3 = ConvGRUCell2(b * 4, b * 4, 3)
        self.conv_gru4 = ConvGRUCell2(b * 8, b * 8, 3)
        self.conv1 = ConvReLU(in_channels, b * 2, 3, 2, 1)
        self.conv2 = ConvReLU(b * 2, b * 4, 3, 2, 1)
        self.conv3 = ConvReLU(b * 4, b * 8, 3, 2, 1)
        self.upconv3 = ConvTransReLU(b * 8, b * 4, 3, 2, 1, 1)
        self.upconv2 = ConvTransReLU(b * 4, b * 2, 3, 2, 1, 1)
        self.upconv1 = ConvTransReLU(b * 2, b, 3, 2, 1, 1)
        self.upconv2d = nn.ConvTranspose2d(b, 1, kernel_size=3, stride=1, padding=1, output_padding=0)

    def _weights(self) -> _RedWeights:
        w = _RedWeights()
        keep = []

        def ptr(t):
            t = _lib.require_cuda(t.detach(), "parameter")
            keep.append(t)
            return t.data_ptr()

        for i, g in enumerate((self.conv_gru1, self.conv_gru2, self.conv_gru3, self.conv_gru4)):
            w.gate_w[i], w.gate_b[i] = ptr(g.gate_conv.weight), ptr(g.gate_conv.bias)
            w.out_w[i], w.out_b[i] = ptr(g.output_conv.weight), ptr(g.output_conv.bias)
            w.rn_w[i], w.rn_b[i] = ptr(g.reset_gate_norm.weight), ptr(g.reset_gate_norm.bias)
            w.un_w[i], w.un_b[i] = ptr(g.update_gate_norm.weight), ptr(g.update_gate_norm.bias)
            w.on_w[i], w.on_b[i] = ptr(g.output_norm.weight), ptr(g.output_norm.bias)
        for i, m in enumerate((self.conv1, self.conv2, self.conv3)):
            w.conv_w[i] = ptr(m.conv.weight)
        for i, m in enumerate((self.upconv1, self.upconv2, self.upconv3)):
            w.upconv_w[i] = ptr(m.conv.weight)
        w.upconv2d_w, w.upconv2d_b = ptr(self.upconv2d.weight), ptr(self.upconv2d.bias)
        w._keep = keep
        return w

    def _run(self, volume: torch.Tensor, states_in, want_states: bool):
        """volume [B,C,D,H,W] -> logits [B,D,H,W] (+ final states)."""
        if torch.is_grad_enabled() and (volume.requires_grad or any(p.requires_grad for p in self.parameters())):
            if states_in is None and not want_states:
                # train.py:284 loss.backward(): the whole-volume form carries a backward through the recurrence (training.py)
                from .training import red_train_forward
                return red_train_forward(self, volume), None
            # the per-slice form (predict path) has no backward.  Failing here beats returning logits without a grad_fn,
            # which would let a loss skip the regulariser silently.
            raise RuntimeError("satmvs_b200 slice_RED_Regularization is inference-only: call it under torch.no_grad() "
                               "(or with parameters and input that do not require grad)")
        vol = _lib.require_cuda(volume, "volume")
        B, Cc, D, H, W = vol.shape
        if Cc != self.in_channels:
            raise ValueError(f"expected {self.in_channels} channels, got {Cc}")
        nbytes = _lib.lib().satmvs_red_workspace_bytes(Cc, D, H, W)
        if nbytes == 0:
            raise ValueError("RED regulariser needs H and W to be multiples of 8")
        ws = _workspace(nbytes, vol.device)
        w = self._weights()
        # packed tensor-core weights live in a buffer of the module and are reused while the parameters are unchanged
        key = (_param_key(self), vol.device)
        pk = getattr(self, "_pack", None)
        if pk is None or pk[0] != key:
            buf = torch.empty(int(_lib.lib().satmvs_red_pack_bytes(Cc)) + 256, dtype=torch.uint8, device=vol.device)
            pk = (key, buf, C.c_ulonglong(0))
            object.__setattr__(self, "_pack", pk)
        pbuf, ptag = pk[1], pk[2]
        pptr = (pbuf.data_ptr() + 255) // 256 * 256
        logits = torch.empty((B, D, H, W), dtype=torch.float32, device=vol.device)
        out_states = None
        if want_states:
            out_states = [torch.empty((B, c, H >> l, W >> l), dtype=torch.float32, device=vol.device)
                          for l, c in enumerate((8, 16, 32, 64))]
        if states_in is not None:
            states_in = [_lib.require_cuda(s, "state") for s in states_in]
        with torch.cuda.device(vol.device):
            st = _lib.stream_ptr(vol.device)
            for b in range(B):
                sin = _lib.ptr_array([s[b].data_ptr() for s in states_in]) if states_in is not None else None
                sout = _lib.ptr_array([s[b].data_ptr() for s in out_states]) if out_states is not None else None
                _lib.check(_lib.lib().satmvs_red_forward_packed(C.byref(w), vol[b].data_ptr(), Cc, D, H, W, sin, sout,
                                                               logits[b].data_ptr(), ws.data_ptr(), ws.numel(), pptr,
                                                               pbuf.numel() - (pptr - pbuf.data_ptr()), C.byref(ptag), st),
                           "red_forward")
        return logits, out_states


class RED_Regularization(_RedBase):
    """`RED_Regularization` (`modules/module.py:595-649`): volume [B,C,D,H,W] -> [B,D,H,W]."""

    def forward(self, volume_variance):
        return self._run(volume_variance, None, False)[0]


class slice_RED_Regularization(_RedBase):
    """`slice_RED_Regularization` (`modules/module.py:653-693`): one depth slice with explicit states.
    cost [B,C,H,W]; returns (reg [B,1,H,W], state1..state4)."""

    def forward(self, cost, state1, state2, state3, state4):
        logits, st = self._run(cost.unsqueeze(2), [state1, state2, state3, state4], True)
        return (logits, *st)

    def forward_planes(self, volume, state1, state2, state3, state4):
        """K consecutive slices in one call: volume [B,C,K,H,W] -> (reg [B,K,H,W], state1..state4 after the last slice).
        Equal to K calls of `forward` with the states carried (the recurrence runs inside the library)."""
        logits, st = self._run(volume, [state1, state2, state3, state4], True)
        return (logits, *st)


class _CostRegWeights(C.Structure):
    _fields_ = [("conv_w", _F * 10), ("bn_scale", _F * 10), ("bn_shift", _F * 10), ("prob_w", _F)]


class Conv3d(nn.Module):
    """Parameter container for the reference's `Conv3d` block (`modules/module.py:324-366`)."""

    def __init__(self, in_channels, out_channels, kernel_size=3, stride=1, relu=True, bn=True, bn_momentum=0.1,
                 init_method="xavier", **kwargs):
        super().__init__()
        assert stride in [1, 2] and bn and relu
        self.out_channels, self.kernel_size, self.stride, self.relu = out_channels, kernel_size, stride, relu
        self.conv = nn.Conv3d(in_channels, out_channels, kernel_size, stride=stride, bias=False, **kwargs)
        self.bn = nn.BatchNorm3d(out_channels, momentum=bn_momentum)


class Deconv3d(nn.Module):
    """Parameter container for the reference's `Deconv3d` block (`modules/module.py:369-410`)."""

    def __init__(self, in_channels, out_channels, kernel_size=3, stride=1, relu=True, bn=True, bn_momentum=0.1,
                 init_method="xavier", **kwargs):
        super().__init__()
        assert stride in [1, 2] and bn and relu
        self.out_channels, self.stride, self.relu = out_channels, stride, relu
        self.conv = nn.ConvTranspose3d(in_channels, out_channels, kernel_size, stride=stride, bias=False, **kwargs)
        self.bn = nn.BatchNorm3d(out_channels, momentum=bn_momentum)


class CostRegNet(nn.Module):
    """`CostRegNet` (`modules/module.py:546-577`): x [B,Cin,D,H,W] -> [B,1,D,H,W].
    `.eval()`: BatchNorm on running statistics, folded into the convolutions (`satmvs_costreg_forward`).
    `.train()`: BatchNorm on batch statistics (running statistics updated) and a backward to x and every parameter
    (`satmvs_b200.training`, `csrc/train.cu`)."""

    _BLOCKS = ("conv0", "conv1", "conv2", "conv3", "conv4", "conv5", "conv6", "conv7", "conv9", "conv11")

    def __init__(self, in_channels, base_channels):
        super().__init__()
        b = base_channels
        self.in_channels, self.base_channels = in_channels, b
        self.conv0 = Conv3d(in_channels, b, padding=1)
        self.conv1 = Conv3d(b, b * 2, stride=2, padding=1)
        self.conv2 = Conv3d(b * 2, b * 2, padding=1)
        self.conv3 = Conv3d(b * 2, b * 4, stride=2, padding=1)
        self.conv4 = Conv3d(b * 4, b * 4, padding=1)
        self.conv5 = Conv3d(b * 4, b * 8, stride=2, padding=1)
        self.conv6 = Conv3d(b * 8, b * 8, padding=1)
        self.conv7 = Deconv3d(b * 8, b * 4, stride=2, padding=1, output_padding=1)
        self.conv9 = Deconv3d(b * 4, b * 2, stride=2, padding=1, output_padding=1)
        self.conv11 = Deconv3d(b * 2, b * 1, stride=2, padding=1, output_padding=1)
        self.prob = nn.Conv3d(b, 1, 3, stride=1, padding=1, bias=False)

    def _weights(self) -> _CostRegWeights:
        key = _param_key(self)
        hit = getattr(self, "_wcache", None)
        if hit is not None and hit[0] == key:
            return hit[1]
        w = _CostRegWeights()
        keep = []

        def ptr(t):
            t = _lib.require_cuda(t.detach(), "parameter")
            keep.append(t)
            return t.data_ptr()

        for i, name in enumerate(self._BLOCKS):
            blk = getattr(self, name)
            scale, shift = _fold_bn(blk)
            w.conv_w[i], w.bn_scale[i], w.bn_shift[i] = ptr(blk.conv.weight), ptr(scale), ptr(shift)
        w.prob_w = ptr(self.prob.weight)
        w._keep = keep
        object.__setattr__(self, "_wcache", (key, w))
        return w

    def forward(self, x):
        if self.training:
            # train.py:268 runs the network in train() mode: BatchNorm3d on batch statistics, backward to x and the parameters
            from .training import costreg_train_forward
            return costreg_train_forward(self, x)
        x = _lib.require_cuda(x, "x")
        B, Cc, D, H, W = x.shape
        if Cc != self.in_channels:
            raise ValueError(f"expected {self.in_channels} channels, got {Cc}")
        nbytes = _lib.lib().satmvs_costreg_workspace_bytes(self.base_channels, D, H, W)
        if nbytes == 0:
            raise ValueError("CostRegNet needs D, H and W to be multiples of 8")
        ws = _workspace(nbytes, x.device)
        w = self._weights()
        out = torch.empty((B, 1, D, H, W), dtype=torch.float32, device=x.device)
        with torch.cuda.device(x.device):
            st = _lib.stream_ptr(x.device)
            for b in range(B):
                _lib.check(_lib.lib().satmvs_costreg_forward(C.byref(w), x[b].data_ptr(), Cc, self.base_channels, D, H, W,
                                                            out[b].data_ptr(), ws.data_ptr(), ws.numel(), st), "costreg_forward")
        return out


class _FeatWeights(C.Structure):
    _fields_ = [("block", _F * 36), ("out_w", _F * 3)]        # 12 x (w, scale, shift)


class Conv2d(nn.Module):
    """Parameter container for the reference's `Conv2d` block (`modules/module.py:78-118`): conv (bias iff no bn) + bn + relu."""

    def __init__(self, in_channels, out_channels, kernel_size, stride=1, relu=True, bn=True, bn_momentum=0.1,
                 init_method="xavier", **kwargs):
        super().__init__()
        assert bn and relu
        self.conv = nn.Conv2d(in_channels, out_channels, kernel_size, stride=stride, bias=False, **kwargs)
        self.kernel_size, self.stride, self.relu = kernel_size, stride, relu
        self.bn = nn.BatchNorm2d(out_channels, momentum=bn_momentum)


class Deconv2d(nn.Module):
    """Parameter container for the reference's `Deconv2d` block (`modules/module.py:121-159`)."""

    def __init__(self, in_channels, out_channels, kernel_size, stride=1, relu=True, bn=True, bn_momentum=0.1,
                 init_method="xavier", **kwargs):
        super().__init__()
        assert bn and relu and stride == 2
        self.conv = nn.ConvTranspose2d(in_channels, out_channels, kernel_size, stride=stride, bias=False, **kwargs)
        self.stride, self.relu = stride, relu
        self.bn = nn.BatchNorm2d(out_channels, momentum=bn_momentum)


class DeConv2dFuse(nn.Module):
    """`DeConv2dFuse` (`modules/module.py:303-321`)."""

    def __init__(self, in_channels, out_channels, kernel_size, relu=True, bn=True, bn_momentum=0.1):
        super().__init__()
        self.deconv = Deconv2d(in_channels, out_channels, kernel_size, stride=2, padding=1, output_padding=1, bn=True, relu=relu,
                               bn_momentum=bn_momentum)
        self.conv = Conv2d(2 * out_channels, out_channels, kernel_size, stride=1, padding=1, bn=bn, relu=relu, bn_momentum=bn_momentum)


class FeatureNet(nn.Module):
    """`FeatureNet` (`modules/module.py:442-543`), arch_mode "unet", three stages: the reference's constructor, sub-module names
    and parameter shapes (reference checkpoints load unchanged), forward on the library's kernels (inference-mode BatchNorm).

    forward(x [B,3,H,W]) -> {"stage1": [B,4b,H/4,W/4], "stage2": [B,2b,H/2,W/2], "stage3": [B,b,H,W]} like the reference;
    forward_views([img_0 .. img_{V-1}]) runs every view of a stack through each layer in one launch and returns the V dicts
    (`networks/casred.py:116-119` calls the net once per view)."""

    def __init__(self, base_channels, num_stage=3, stride=4, arch_mode="unet"):
        super().__init__()
        if arch_mode != "unet" or num_stage != 3:
            raise NotImplementedError("satmvs_b200.FeatureNet implements the configuration the cascades use: unet, 3 stages")
        b = base_channels
        self.arch_mode, self.stride, self.base_channels, self.num_stage = arch_mode, stride, b, num_stage
        self.conv0 = nn.Sequential(Conv2d(3, b, 3, 1, padding=1), Conv2d(b, b, 3, 1, padding=1))
        self.conv1 = nn.Sequential(Conv2d(b, b * 2, 5, stride=2, padding=2), Conv2d(b * 2, b * 2, 3, 1, padding=1),
                                   Conv2d(b * 2, b * 2, 3, 1, padding=1))
        self.conv2 = nn.Sequential(Conv2d(b * 2, b * 4, 5, stride=2, padding=2), Conv2d(b * 4, b * 4, 3, 1, padding=1),
                                   Conv2d(b * 4, b * 4, 3, 1, padding=1))
        self.out1 = nn.Conv2d(b * 4, b * 4, 1, bias=False)
        self.deconv1 = DeConv2dFuse(b * 4, b * 2, 3)
        self.deconv2 = DeConv2dFuse(b * 2, b, 3)
        self.out2 = nn.Conv2d(b * 2, b * 2, 1, bias=False)
        self.out3 = nn.Conv2d(b, b, 1, bias=False)
        self.out_channels = [4 * b, 2 * b, b]

    def _weights(self) -> _FeatWeights:
        key = _param_key(self)
        hit = getattr(self, "_wcache", None)
        if hit is not None and hit[0] == key:
            return hit[1]
        w = _FeatWeights()
        keep = []

        def ptr(t):
            t = _lib.require_cuda(t.detach(), "parameter")
            keep.append(t)
            return t.data_ptr()

        blocks = [self.conv0[0], self.conv0[1], self.conv1[0], self.conv1[1], self.conv1[2], self.conv2[0], self.conv2[1],
                  self.conv2[2], self.deconv1.deconv, self.deconv1.conv, self.deconv2.deconv, self.deconv2.conv]
        for i, blk in enumerate(blocks):
            scale, shift = _fold_bn(blk)
            w.block[3 * i], w.block[3 * i + 1], w.block[3 * i + 2] = ptr(blk.conv.weight), ptr(scale), ptr(shift)
        for i, m in enumerate((self.out1, self.out2, self.out3)):
            w.out_w[i] = ptr(m.weight)
        w._keep = keep
        object.__setattr__(self, "_wcache", (key, w))
        return w

    def forward_views(self, images):
        if self.training:
            # train.py:268 model.train(): batch-statistics BatchNorm, one autograd node per block (training.py); the reference
            # calls the net once per view (casred.py:116-119), so the statistics are per view here too
            from .training import featurenet_train_forward
            return [featurenet_train_forward(self, _lib.require_cuda(i, "image")) for i in images]
        if torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters()):
            raise RuntimeError("satmvs_b200.FeatureNet in eval() mode has no backward: call it under torch.no_grad() "
                               "(a frozen extractor) or in train() mode")
        imgs = [_lib.require_cuda(i, "image") for i in images]
        B, c3, H, W = imgs[0].shape
        if c3 != 3 or any(i.shape != imgs[0].shape for i in imgs):
            raise ValueError("images must be V tensors [B, 3, H, W] of one shape")
        V, b = len(imgs), self.base_channels
        nbytes = _lib.lib().satmvs_featurenet_workspace_bytes(b, V, H, W)
        if nbytes == 0:
            raise ValueError("FeatureNet needs H and W to be multiples of 4")
        ws = _workspace(nbytes, imgs[0].device)
        x = torch.stack(imgs, dim=2).contiguous()                        # [B, 3, V, H, W]: the views ride the plane axis
        dev = x.device
        o1 = torch.empty((B, V, 4 * b, H // 4, W // 4), dtype=torch.float32, device=dev)      # view-major per sample
        o2 = torch.empty((B, V, 2 * b, H // 2, W // 2), dtype=torch.float32, device=dev)
        o3 = torch.empty((B, V, b, H, W), dtype=torch.float32, device=dev)
        w = self._weights()
        with torch.cuda.device(dev):
            st = _lib.stream_ptr(dev)
            for bi in range(B):
                _lib.check(_lib.lib().satmvs_featurenet_forward(C.byref(w), x[bi].data_ptr(), b, V, H, W, o1[bi].data_ptr(),
                                                               o2[bi].data_ptr(), o3[bi].data_ptr(), ws.data_ptr(), ws.numel(), st),
                           "featurenet_forward")
        # batch 1 (the reference's recipes): every view's [1,C,h,w] map is a contiguous slice, no copy
        return [{"stage1": o1[:, v].contiguous(), "stage2": o2[:, v].contiguous(), "stage3": o3[:, v].contiguous()} for v in range(V)]

    def forward(self, x):
        return self.forward_views([x])[0]


from .regress import depth_regression  # noqa: E402,F401  (`modules/module.py:433-439`, on the heads kernel)
