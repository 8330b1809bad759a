"""Training form of `CostRegNet` (`modules/module.py:546-577`; blocks `Conv3d:324-366`, `Deconv3d:369-410`) and of the CasMVS
soft-argmin head, for `train.py:267-287` (`model.train()`, `loss.backward()`): BatchNorm3d on batch statistics with running-stat
updates, and a hand-written backward to the input and to every parameter, all on the library's kernels
(`satmvs_conv3d_raw`, `satmvs_conv3d_wgrad`, `satmvs_bn_train_fwd/_bwd`, `satmvs_softargmin_bwd`; `csrc/train.cu`).
torch supplies the autograd tape hook (`torch.autograd.Function`), device memory and the stream; no torch operator touches an
activation."""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib

_EPS = 1e-5      # nn.BatchNorm3d default


def _ws(nbytes: int, device) -> torch.Tensor:
    return torch.empty(max(int(nbytes), 256), dtype=torch.uint8, device=device)


def conv3d_raw(x, w, mode: int, cout: int, w_co: int, w_ci: int, nz: int = 3):
    """`satmvs_conv3d_raw` on a batch: x [B,Cin,D,H,W] -> [B,cout,D',H',W'] (mode 0/2: same size, 1: halved, 3: doubled)."""
    B, cin, D, H, W = x.shape
    if mode == 1:
        od, oh, ow = (D // 2 if nz == 3 else D), H // 2, W // 2
    elif mode == 3:
        od, oh, ow = (2 * D if nz == 3 else D), 2 * H, 2 * W
    else:
        od, oh, ow = D, H, W
    out = torch.empty((B, cout, od, oh, ow), dtype=torch.float32, device=x.device)
    st = _lib.stream_ptr(x.device)
    for b in range(B):
        _lib.check(_lib.lib().satmvs_conv3d_raw(x[b].data_ptr(), cin, D, H, W, w.data_ptr(), w_co, w_ci, nz, mode,
                                                out[b].data_ptr(), cout, st), "conv3d_raw")
    return out


def conv3d_wgrad(x, dy, stride: int, dw, dw_co: int, dw_ci: int, nz: int = 3):
    """dw[co,ci,tap] = sum over the batch and all positions of dy[co,o] * x[ci, stride*o + k - 1] (`satmvs_conv3d_wgrad`)."""
    B, cin, D, H, W = x.shape
    cout = dy.shape[1]
    L = _lib.lib()
    nbytes = L.satmvs_conv3d_wgrad_workspace_bytes(cin, cout, D, H, W, nz, stride)
    if nbytes == 0:
        raise ValueError("conv3d_wgrad: shape not supported")
    ws = _ws(nbytes, x.device)
    st = _lib.stream_ptr(x.device)
    for b in range(B):
        _lib.check(L.satmvs_conv3d_wgrad(x[b].data_ptr(), cin, D, H, W, dy[b].data_ptr(), cout, nz, stride, dw.data_ptr(),
                                         dw_co, dw_ci, int(b > 0), ws.data_ptr(), ws.numel(), st), "conv3d_wgrad")
    return dw


def bn_train_fwd(y, gamma, beta, relu: bool, post_add=None):
    """z = relu(BN(y; batch statistics)) (+ post_add); returns z, mean, biased var."""
    B, Cc = y.shape[:2]
    n = y[0, 0].numel()
    z = torch.empty_like(y)
    mean = torch.empty(Cc, dtype=torch.float32, device=y.device)
    var = torch.empty(Cc, dtype=torch.float32, device=y.device)
    acc = torch.empty(2 * Cc, dtype=torch.float64, device=y.device)
    _lib.check(_lib.lib().satmvs_bn_train_fwd(y.data_ptr(), B, Cc, n, gamma.data_ptr(), beta.data_ptr(), _EPS, int(relu),
                                              post_add.data_ptr() if post_add is not None else None, z.data_ptr(),
                                              mean.data_ptr(), var.data_ptr(), acc.data_ptr(), _lib.stream_ptr(y.device)),
               "bn_train_fwd")
    return z, mean, var


def bn_train_bwd(dz, dz2, y, gamma, beta, mean, var, relu: bool):
    """Gradient at the conv output and d gamma, d beta from the gradient(s) at the block output."""
    B, Cc = y.shape[:2]
    n = y[0, 0].numel()
    dy = torch.empty_like(y)
    dg = torch.empty(Cc, dtype=torch.float32, device=y.device)
    db = torch.empty(Cc, dtype=torch.float32, device=y.device)
    acc = torch.empty(2 * Cc, dtype=torch.float64, device=y.device)
    _lib.check(_lib.lib().satmvs_bn_train_bwd(dz.data_ptr(), dz2.data_ptr() if dz2 is not None else None, y.data_ptr(), B, Cc, n,
                                              gamma.data_ptr(), beta.data_ptr(), mean.data_ptr(), var.data_ptr(), _EPS,
                                              int(relu), dy.data_ptr(), dg.data_ptr(), db.data_ptr(), acc.data_ptr(),
                                              _lib.stream_ptr(y.device)), "bn_train_bwd")
    return dy, dg, db


def _update_running_stats(bn, mean, var, count: int) -> None:
    """`nn.BatchNorm3d` bookkeeping in train mode: running stats with momentum, unbiased variance, batch counter."""
    if not bn.track_running_stats or bn.running_mean is None:
        return
    with torch.no_grad():
        bn.num_batches_tracked += 1
        m = bn.momentum if bn.momentum is not None else 1.0 / float(bn.num_batches_tracked)
        bn.running_mean.mul_(1.0 - m).add_(mean, alpha=m)
        bn.running_var.mul_(1.0 - m).add_(var, alpha=m * count / max(count - 1, 1))


class _CostRegTrainFn(torch.autograd.Function):
    """The whole CostRegNet as one tape entry: forward keeps every block's input, conv output and batch statistics."""

    @staticmethod
    def forward(ctx, net, x, *params):
        x = _lib.require_cuda(x.detach(), "x")
        b = net.base_channels
        blocks = [getattr(net, nme) for nme in net._BLOCKS]
        p = [t.detach() for t in params]
        W = [p[3 * i] for i in range(10)]
        G = [p[3 * i + 1] for i in range(10)]
        Bt = [p[3 * i + 2] for i in range(10)]
        wp = p[30]
        cin = [net.in_channels, b, 2 * b, 2 * b, 4 * b, 4 * b, 8 * b, 8 * b, 4 * b, 2 * b]
        cout = [b, 2 * b, 2 * b, 4 * b, 4 * b, 8 * b, 8 * b, 4 * b, 2 * b, b]
        mode = [0, 1, 0, 1, 0, 1, 0, 3, 3, 3]
        xs, ys, means, vars_ = [], [], [], []
        with torch.cuda.device(x.device):
            cur = x
            outs = []
            for i in range(10):
                if mode[i] == 3:      # ConvTranspose3d weight [Cin, Cout, 27]
                    y = conv3d_raw(cur, W[i], 3, cout[i], 27, cout[i] * 27)
                else:                 # Conv3d weight [Cout, Cin, 27]
                    y = conv3d_raw(cur, W[i], mode[i], cout[i], cin[i] * 27, 27)
                skip = {7: outs[4], 8: outs[2], 9: outs[0]}.get(i) if i >= 7 else None
                z, mean, var = bn_train_fwd(y, G[i], Bt[i], True, skip)
                _update_running_stats(blocks[i].bn, mean, var, y.shape[0] * y[0, 0].numel())
                xs.append(cur); ys.append(y); means.append(mean); vars_.append(var)
                outs.append(z)
                cur = z
            out = conv3d_raw(cur, wp, 0, 1, b * 27, 27)
        ctx.net_dims = (cin, cout, mode, b)
        ctx.save_for_backward(cur, wp, *xs, *ys, *means, *vars_, *W, *G, *Bt)
        return out

    @staticmethod
    def backward(ctx, g):
        cin, cout, mode, b = ctx.net_dims
        sv = ctx.saved_tensors
        x11, wp = sv[0], sv[1]
        xs, ys, means, vars_ = sv[2:12], sv[12:22], sv[22:32], sv[32:42]
        W, G, Bt = sv[42:52], sv[52:62], sv[62:72]
        g = g.contiguous().float()
        grads = [None] * 31
        with torch.cuda.device(g.device):
            dwp = torch.empty_like(wp)
            conv3d_wgrad(x11, g, 1, dwp, b * 27, 27)
            grads[30] = dwp
            dz = conv3d_raw(g, wp, 2, b, 27, b * 27)             # gradient at x11 = c0 + z11
            skip_grad = {}                                        # block index -> gradient arriving over the skip connection
            for i in range(9, -1, -1):
                if i >= 7:
                    skip_grad[{9: 0, 8: 2, 7: 4}[i]] = dz         # the skip tensor receives the same gradient
                dy, dg, db = bn_train_bwd(dz, skip_grad.pop(i, None), ys[i], G[i], Bt[i], means[i], vars_[i], True)
                dw = torch.empty_like(W[i])
                if mode[i] == 3:
                    conv3d_wgrad(dy, xs[i], 2, dw, cout[i] * 27, 27)                       # roles swapped (csrc/train.cu)
                    dx = conv3d_raw(dy, W[i], 1, cin[i], cout[i] * 27, 27)
                else:
                    conv3d_wgrad(xs[i], dy, 2 if mode[i] == 1 else 1, dw, cin[i] * 27, 27)
                    if i == 0 and not ctx.needs_input_grad[1]:
                        dx = None
                    elif mode[i] == 1:
                        dx = conv3d_raw(dy, W[i], 3, cin[i], 27, cin[i] * 27)
                    else:
                        dx = conv3d_raw(dy, W[i], 2, cin[i], 27, cin[i] * 27)
                grads[3 * i], grads[3 * i + 1], grads[3 * i + 2] = dw, dg, db
                dz = dx
        return (None, dz, *grads)


def costreg_train_forward(net, x):
    """`CostRegNet.forward` in train mode (`modules/module.py:568-577`) with a backward."""
    params = []
    for nme in net._BLOCKS:
        blk = getattr(net, nme)
        params += [blk.conv.weight, blk.bn.weight, blk.bn.bias]
    params.append(net.prob.weight)
    B, Cc, D, H, W = x.shape
    if Cc != net.in_channels:
        raise ValueError(f"expected {net.in_channels} channels, got {Cc}")
    if D % 8 or H % 8 or W % 8:
        raise ValueError("CostRegNet needs D, H and W to be multiples of 8")
    if (D * H * W // 512) % 4:
        raise ValueError("training form: the coarsest level must hold a multiple of 4 voxels per channel (D*H*W % 2048 == 0)")
    return _CostRegTrainFn.apply(net, x.contiguous().float(), *params)


class _SoftArgminDepthFn(torch.autograd.Function):
    """depth = sum_d softmax(logits)_d * depth_values_d with its gradient to the logits (`casmvs.py:66-68`)."""

    @staticmethod
    def forward(ctx, logits, depth_values):
        from .regress import softargmin
        depth, conf = softargmin(logits.detach(), depth_values, head="casmvs")
        ctx.save_for_backward(logits.detach(), depth_values)
        ctx.mark_non_differentiable(conf)
        return depth, conf

    @staticmethod
    def backward(ctx, gdepth, _gconf):
        logits, dv = ctx.saved_tensors
        B, D, H, W = logits.shape
        per_pixel = int(dv.dim() == 4)
        gdepth = gdepth.contiguous().float()
        out = torch.empty_like(logits)
        with torch.cuda.device(logits.device):
            st = _lib.stream_ptr(logits.device)
            for b in range(B):
                _lib.check(_lib.lib().satmvs_softargmin_bwd(logits[b].data_ptr(), dv[b].data_ptr(), per_pixel, D, H, W,
                                                            gdepth[b].data_ptr(), out[b].data_ptr(), st), "softargmin_bwd")
        return out, None


def softargmin_casmvs_train(logits, depth_values):
    """CasMVS head with a gradient to the logits: returns (depth [B,H,W], confidence [B,H,W] (no gradient, as in the reference's
    `torch.no_grad()` block `casmvs.py:69`))."""
    return _SoftArgminDepthFn.apply(logits.contiguous().float(), depth_values.contiguous().float())
