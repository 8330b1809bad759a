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
    return _CostRegTrainFn.apply(net, x.contiguous().float(), *params)


class _SoftArgminDepthFn(torch.autograd.Function):
    """depth = sum_d softmax(logits)_d * depth_values_d with its gradient to the logits (`casmvs.py:66-68`, `casred.py:58-61`);
    the confidence map carries no gradient (`casmvs.py:69` computes it under `torch.no_grad()`)."""

    @staticmethod
    def forward(ctx, logits, depth_values, head):
        from .regress import softargmin
        depth, conf = softargmin(logits.detach(), depth_values, head=head)
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
        return out, None, None


def softargmin_train(logits, depth_values, head="casmvs"):
    """Soft-argmin head with a gradient to the logits: returns (depth [B,H,W], confidence [B,H,W] without gradient);
    head "casmvs" (4-neighbour confidence) or "red" (max probability)."""
    return _SoftArgminDepthFn.apply(logits.contiguous().float(), depth_values.contiguous().float(), head)


def softargmin_casmvs_train(logits, depth_values):
    return softargmin_train(logits, depth_values, "casmvs")


# ---------------------------------------------------------------------------------------------------------------------------
# RED regulariser: backward through the recurrence (RED_Regularization.forward, modules/module.py:614-649)
# ---------------------------------------------------------------------------------------------------------------------------
_CH = (8, 16, 32, 64)
_GN_EPS = 1e-5


class _GruBwdLevel(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("S", "ru", "y", "opre", "gpre", "ob", "gb", "on_w", "rn_w", "un_w", "ostat", "gstat", "dec",
                                          "wo_h", "wg_h")] + [("w_ci", C.c_longlong)] + \
               [(n, C.c_void_p) for n in ("dyn", "dgn", "dO", "dG", "scratch")] + \
               [("ch", C.c_int), ("h", C.c_int), ("w", C.c_int), ("D", C.c_int)]


def _ew(a, out, b=None, mul=None, m1=None, m2=None, scale=1.0):
    """out = scale * (a [+ b]) [* mul], zero where (m1 [- m2]) <= 0; operands [C, ...] views whose channels are contiguous."""
    Cc, n = a.shape[0], a[0].numel()
    for t in (a, out, b, mul, m1, m2):
        assert t is None or (t.shape[0] == Cc and t[0].numel() == n and t[0].is_contiguous())
    p = lambda t: (t.data_ptr(), t.stride(0)) if t is not None else (None, 0)
    _lib.check(_lib.lib().satmvs_elementwise(*p(a), *p(b), *p(mul), *p(m1), *p(m2), float(scale), *p(out), Cc, n,
                                             _lib.stream_ptr(a.device)), "elementwise")
    return out


def _conv2d(x, w_ptr_tensor, mode, cout, w_co, w_ci):
    """Per-plane 3x3 conv of a [Cin, D, h, w] tensor (satmvs_conv3d_raw, NZ = 1)."""
    return conv3d_raw(x.unsqueeze(0), w_ptr_tensor, mode, cout, w_co, w_ci, nz=1)[0]


def _gn_act(pre, bias, gamma, beta, ch, groups, act):
    Cc, D, h, w = pre.shape
    out = torch.empty_like(pre)
    stats = torch.empty((D, groups, 2), dtype=torch.float32, device=pre.device)
    acc = torch.empty(D * groups * 2, dtype=torch.float64, device=pre.device)
    _lib.check(_lib.lib().satmvs_gn_act_fwd(pre.data_ptr(), bias.data_ptr(), gamma.data_ptr(), beta.data_ptr(), ch, groups, D, h * w,
                                            _GN_EPS, act, out.data_ptr(), stats.data_ptr(), acc.data_ptr(),
                                            _lib.stream_ptr(pre.device)), "gn_act_fwd")
    return out, stats


def _channel_sum(t):
    Cc = t.shape[0]
    out = torch.empty(Cc, dtype=torch.float32, device=t.device)
    acc = torch.empty(2 * Cc, dtype=torch.float64, device=t.device)
    _lib.check(_lib.lib().satmvs_channel_sum(t.data_ptr(), Cc, t[0].numel(), out.data_ptr(), acc.data_ptr(),
                                             _lib.stream_ptr(t.device)), "channel_sum")
    return out


def _gn_param_grad(dout, pre, bias, stats, ch, groups):
    Cc, D, h, w = pre.shape
    dg = torch.empty(Cc, dtype=torch.float32, device=pre.device)
    db = torch.empty(Cc, dtype=torch.float32, device=pre.device)
    acc = torch.empty(2 * Cc, dtype=torch.float64, device=pre.device)
    _lib.check(_lib.lib().satmvs_gn_param_grad(dout.data_ptr(), pre.data_ptr(), bias.data_ptr(), stats.data_ptr(), ch, groups, D,
                                               h * w, dg.data_ptr(), db.data_ptr(), acc.data_ptr(), _lib.stream_ptr(pre.device)),
               "gn_param_grad")
    return dg, db


def _wgrad2d(x, dy, stride, shape, co_stride):
    dw = torch.empty(shape, dtype=torch.float32, device=x.device)
    conv3d_wgrad(x.unsqueeze(0), dy.unsqueeze(0), stride, dw, co_stride, 9, nz=1)
    return dw


_RED_PARAM_NAMES = tuple(f"conv_gru{i + 1}.{n}" for i in range(4) for n in (
    "gate_conv.weight", "gate_conv.bias", "output_conv.weight", "output_conv.bias", "reset_gate_norm.weight", "reset_gate_norm.bias",
    "update_gate_norm.weight", "update_gate_norm.bias", "output_norm.weight", "output_norm.bias")) + (
    "conv1.conv.weight", "conv2.conv.weight", "conv3.conv.weight", "upconv1.conv.weight", "upconv2.conv.weight",
    "upconv3.conv.weight", "upconv2d.weight", "upconv2d.bias")


def _red_backward_sample(P: dict, vol, ws, dlog):
    """Gradients of one sample.  P: parameter name -> tensor; vol [C,D,H,W]; ws: the workspace the forward of THIS sample ran in;
    dlog [D,H,W].  Returns (dvol [C,D,H,W], {name: grad})."""
    dev = vol.device
    Cin, D, H, W = vol.shape
    offs = (C.c_size_t * 10)()
    _lib.check(_lib.lib().satmvs_red_workspace_layout(Cin, D, H, W, offs), "red_workspace_layout")

    def view(off, shape):
        n = 1
        for s in shape:
            n *= s
        return ws[off:off + 4 * n].view(torch.float32).view(shape)

    S = [view(offs[l], (_CH[l], D + 1, H >> l, W >> l)) for l in range(4)]
    E = [None] + [view(offs[3 + l], (_CH[l], D, H >> l, W >> l)) for l in (1, 2, 3)]
    U = [view(offs[7 + l], (_CH[l], D + 1, H >> l, W >> l)) for l in range(3)]
    g = {}
    X0 = _ew(vol, torch.empty_like(vol), scale=-1.0)                       # the network sees -cost (module.py:627, :640)
    X = [X0, E[1], E[2], E[3]]

    # ---- decoder, batched over planes (module.py:633-643) ----
    dlog4 = dlog.reshape(1, D, H, W).contiguous()
    w2d = P["upconv2d.weight"]
    dU = _conv2d(dlog4, w2d, 0, 8, 9, 9)
    U0c = _ew(U[0][:, 1:], torch.empty((8, D, H, W), dtype=torch.float32, device=dev))
    g["upconv2d.weight"] = _wgrad2d(U0c, dlog4, 1, (1, 8, 3, 3), 72).flip(2, 3).permute(1, 0, 2, 3).contiguous()
    g["upconv2d.bias"] = _channel_sum(dlog4)
    dec = [None] * 4
    ups = ("upconv1.conv.weight", "upconv2.conv.weight", "upconv3.conv.weight")
    for l in range(3):
        dec[l] = dU
        src = U[l + 1][:, 1:] if l + 1 < 3 else S[3][:, 1:]                                     # the transposed conv's input
        srcc = _ew(src, torch.empty((_CH[l + 1], D, H >> (l + 1), W >> (l + 1)), dtype=torch.float32, device=dev))
        wt = P[ups[l]]                                                                          # [ch_{l+1}, ch_l, 3, 3]
        # ReLU mask from the recomputed transposed conv: U_l - S_l would lose activations smaller than an ulp of S_l, and a mask
        # error is a 100 % error of that voxel's gradient
        pre = _conv2d(srcc, wt, 3, _CH[l], 9, _CH[l] * 9)
        dpre = _ew(dU, torch.empty_like(dU), m1=pre)
        del pre
        g[ups[l]] = _wgrad2d(dpre, srcc, 2, tuple(wt.shape), _CH[l] * 9)
        dU = _conv2d(dpre, wt, 1, _CH[l + 1], _CH[l] * 9, 9)
    dec[3] = dU

    # ---- per level: recompute the cell's pre-activations batched over planes ----
    L = _lib.lib()
    st = _lib.stream_ptr(dev)
    lv = (_GruBwdLevel * 4)()
    keep = []
    for l in range(4):
        ch, cx, h, w = _CH[l], (Cin if l == 0 else _CH[l]), H >> l, W >> l
        nm = f"conv_gru{l + 1}."
        Wg, bg, Wo, bo = P[nm + "gate_conv.weight"], P[nm + "gate_conv.bias"], P[nm + "output_conv.weight"], P[nm + "output_conv.bias"]
        rn_w, un_w, on_w = P[nm + "reset_gate_norm.weight"], P[nm + "update_gate_norm.weight"], P[nm + "output_norm.weight"]
        gam = torch.cat([rn_w, un_w]).contiguous()
        bet = torch.cat([P[nm + "reset_gate_norm.bias"], P[nm + "update_gate_norm.bias"]]).contiguous()
        Hp = S[l][:, :D]
        XH = torch.empty((cx + ch, D, h, w), dtype=torch.float32, device=dev)
        _ew(X[l], XH[:cx]); _ew(Hp, XH[cx:])
        Gpre = _conv2d(XH, Wg, 0, 2 * ch, (cx + ch) * 9, 9)
        RU, gstat = _gn_act(Gpre, bg, gam, bet, ch, 2, 0)
        XRH = torch.empty_like(XH)
        _ew(X[l], XRH[:cx]); _ew(RU[:ch], XRH[cx:], mul=Hp)
        Opre = _conv2d(XRH, Wo, 0, ch, (cx + ch) * 9, 9)
        Y, ostat = _gn_act(Opre, bo, on_w, P[nm + "output_norm.bias"], ch, 1, 1)
        dGO = torch.empty((3 * ch, D, h, w), dtype=torch.float32, device=dev)     # [dG (2ch) | dO (ch)]
        dGn = torch.empty((2 * ch, D, h, w), dtype=torch.float32, device=dev)
        dYn = torch.empty((ch, D, h, w), dtype=torch.float32, device=dev)
        scratch = torch.empty(12 * D + 4 + 14 * ch * h * w, dtype=torch.float32, device=dev)
        Wo_h, Wg_h = Wo.reshape(-1)[cx * 9:], Wg.reshape(-1)[cx * 9:]
        a = lv[l]
        a.S, a.ru, a.y, a.opre, a.gpre = S[l].data_ptr(), RU.data_ptr(), Y.data_ptr(), Opre.data_ptr(), Gpre.data_ptr()
        a.ob, a.gb, a.on_w, a.rn_w, a.un_w = bo.data_ptr(), bg.data_ptr(), on_w.data_ptr(), rn_w.data_ptr(), un_w.data_ptr()
        a.ostat, a.gstat, a.dec = ostat.data_ptr(), gstat.data_ptr(), dec[l].data_ptr()
        a.wo_h, a.wg_h, a.w_ci = Wo_h.data_ptr(), Wg_h.data_ptr(), (cx + ch) * 9
        a.dyn, a.dgn, a.dG, a.dO, a.scratch = dYn.data_ptr(), dGn.data_ptr(), dGO.data_ptr(), dGO[2 * ch:].data_ptr(), scratch.data_ptr()
        a.ch, a.h, a.w, a.D = ch, h, w, D
        keep.append((XH, XRH, Gpre, Opre, RU, Y, gstat, ostat, dGO, dGn, dYn, scratch, Wg, bg, Wo, bo, cx))
    # ---- sequential in depth: planes D-1 .. 0, the four levels on streams of their own (csrc/red_train.cu) ----
    _lib.check(L.satmvs_red_recurrence_bwd(lv, 4, st), "red_recurrence_bwd")
    # ---- batched: gradient to the cells' x inputs, filters, biases, GroupNorm affine parameters ----
    dX = [None] * 4
    for l in range(4):
        XH, XRH, Gpre, Opre, RU, Y, gstat, ostat, dGO, dGn, dYn, scratch, Wg, bg, Wo, bo, cx = keep[l]
        ch = _CH[l]
        nm = f"conv_gru{l + 1}."
        dG_all, dO_all = dGO[:2 * ch], dGO[2 * ch:]
        Wcat = torch.cat([Wg, Wo], 0).contiguous()
        dX[l] = _conv2d(dGO, Wcat, 2, cx, 9, (cx + ch) * 9)
        g[nm + "gate_conv.weight"] = _wgrad2d(XH, dG_all, 1, tuple(Wg.shape), (cx + ch) * 9)
        g[nm + "output_conv.weight"] = _wgrad2d(XRH, dO_all, 1, tuple(Wo.shape), (cx + ch) * 9)
        g[nm + "gate_conv.bias"] = _channel_sum(dG_all)
        g[nm + "output_conv.bias"] = _channel_sum(dO_all)
        dgam, dbet = _gn_param_grad(dGn, Gpre, bg, gstat, ch, 2)
        g[nm + "reset_gate_norm.weight"], g[nm + "update_gate_norm.weight"] = dgam[:ch].clone(), dgam[ch:].clone()
        g[nm + "reset_gate_norm.bias"], g[nm + "update_gate_norm.bias"] = dbet[:ch].clone(), dbet[ch:].clone()
        g[nm + "output_norm.weight"], g[nm + "output_norm.bias"] = _gn_param_grad(dYn, Opre, bo, ostat, ch, 1)

    # ---- encoders conv3, conv2, conv1 (ConvReLU stride 2, module.py:627-629), batched ----
    dE = dX[3]
    for l in (3, 2, 1):
        nm = f"conv{l}.conv.weight"
        wc = P[nm]                                                     # [ch_l, cin, 3, 3]
        cin = wc.shape[1]
        dpre = _ew(dE, torch.empty_like(dE), m1=E[l])
        g[nm] = _wgrad2d(X[l - 1], dpre, 2, tuple(wc.shape), cin * 9)
        back = _conv2d(dpre, wc, 3, cin, 9, cin * 9)
        dE = _ew(dX[l - 1], torch.empty_like(back), b=back, scale=(-1.0 if l == 1 else 1.0))
    return dE, g


class _RedTrainFn(torch.autograd.Function):
    """RED_Regularization.forward with a backward: the forward is the library's fast path (tensor-core recurrence) run in a
    workspace of its own per sample, which the backward reads the state history from."""

    @staticmethod
    def forward(ctx, net, volume, *params):
        vol = _lib.require_cuda(volume.detach(), "volume").contiguous().float()
        B, Cc, D, H, W = vol.shape
        L = _lib.lib()
        nbytes = L.satmvs_red_workspace_bytes(Cc, D, H, W)
        if nbytes == 0:
            raise ValueError("RED regulariser needs H and W to be multiples of 8")
        w = net._weights()
        logits = torch.empty((B, D, H, W), dtype=torch.float32, device=vol.device)
        wss = []
        with torch.cuda.device(vol.device):
            st = _lib.stream_ptr(vol.device)
            for b in range(B):
                ws = torch.empty(nbytes, dtype=torch.uint8, device=vol.device)
                _lib.check(L.satmvs_red_forward(C.byref(w), vol[b].data_ptr(), Cc, D, H, W, None, None, logits[b].data_ptr(),
                                                ws.data_ptr(), ws.numel(), st), "red_forward")
                wss.append(ws)
        ctx.save_for_backward(vol, *wss, *[p.detach() for p in params])
        ctx.nb = B
        return logits

    @staticmethod
    def backward(ctx, gl):
        sv = ctx.saved_tensors
        B = ctx.nb
        vol, wss, params = sv[0], sv[1:1 + B], sv[1 + B:]
        P = dict(zip(_RED_PARAM_NAMES, params))
        gl = gl.contiguous().float()
        dvol = torch.empty_like(vol)
        total = None
        with torch.cuda.device(vol.device):
            for b in range(B):
                dv, g = _red_backward_sample(P, vol[b], wss[b], gl[b])
                dvol[b] = dv
                if total is None:
                    total = g
                else:
                    for k in total:
                        total[k] = _ew(total[k].reshape(1, -1), torch.empty_like(total[k]).reshape(1, -1),
                                       b=g[k].reshape(1, -1)).reshape(total[k].shape)
        return (None, dvol, *[total[n] for n in _RED_PARAM_NAMES])


def red_train_forward(net, volume):
    """`RED_Regularization.forward` (`modules/module.py:614-649`) with gradients to the volume and to every parameter."""
    sd = dict(net.named_parameters())
    return _RedTrainFn.apply(net, volume, *[sd[n] for n in _RED_PARAM_NAMES])


# ---------------------------------------------------------------------------------------------------------------------------
# FeatureNet in train mode (modules/module.py:442-543): block-level autograd nodes on the 2-D primitives
# ---------------------------------------------------------------------------------------------------------------------------
def _conv2d_raw_b(x, w, K, mode, cout, w_co, w_ci):
    """satmvs_conv2d_raw on a batch [B, Cin, H, W] -> [B, cout, H', W']."""
    B, cin, H, W = x.shape
    oh, ow = (H // 2, W // 2) if mode == 1 else ((2 * H, 2 * W) if mode == 3 else (H, W))
    out = torch.empty((B, cout, oh, ow), dtype=torch.float32, device=x.device)
    st = _lib.stream_ptr(x.device)
    for b in range(B):
        _lib.check(_lib.lib().satmvs_conv2d_raw(x[b].data_ptr(), cin, 1, H, W, w.data_ptr(), w_co, w_ci, K, mode,
                                                out[b].data_ptr(), cout, st), "conv2d_raw")
    return out


def _conv2d_wgrad_b(x, dy, K, stride, shape, dw_co):
    """dw (given shape) = sum over the batch of the weight gradient with x [B,Cin,H,W] as the conv input, dy [B,Cout,H',W']."""
    B, cin, H, W = x.shape
    cout = dy.shape[1]
    L = _lib.lib()
    nbytes = L.satmvs_conv2d_wgrad_workspace_bytes(cin, cout, 1, H, W, K, stride)
    if nbytes == 0:
        raise ValueError("conv2d_wgrad: shape not supported")
    ws = _ws(nbytes, x.device)
    dw = torch.empty(shape, dtype=torch.float32, device=x.device)
    st = _lib.stream_ptr(x.device)
    for b in range(B):
        _lib.check(L.satmvs_conv2d_wgrad(x[b].data_ptr(), cin, 1, H, W, dy[b].data_ptr(), cout, K, stride, dw.data_ptr(), dw_co,
                                         K * K, int(b > 0), ws.data_ptr(), ws.numel(), st), "conv2d_wgrad")
    return dw


class _Conv2dBlockFn(torch.autograd.Function):
    """`Conv2d` / `Deconv2d` block of the reference (`modules/module.py:78-159`): conv (no bias) + BatchNorm2d on batch statistics +
    ReLU.  kind 0: conv stride 1, 1: conv stride 2, 3: transposed conv stride 2 (output_padding 1)."""

    @staticmethod
    def forward(ctx, x, w, gamma, beta, bn, K, kind):
        x = _lib.require_cuda(x.detach(), "x").contiguous().float()
        w_, g_, b_ = w.detach(), gamma.detach(), beta.detach()
        taps = K * K
        with torch.cuda.device(x.device):
            if kind == 3:      # ConvTranspose2d weight [Cin, Cout, K, K]
                cout = w_.shape[1]
                y = _conv2d_raw_b(x, w_, K, 3, cout, taps, cout * taps)
            else:              # Conv2d weight [Cout, Cin, K, K]
                cout = w_.shape[0]
                y = _conv2d_raw_b(x, w_, K, kind, cout, x.shape[1] * taps, taps)
            z, mean, var = bn_train_fwd(y, g_, b_, True)
        _update_running_stats(bn, mean, var, y.shape[0] * y[0, 0].numel())
        ctx.save_for_backward(x, y, w_, g_, b_, mean, var)
        ctx.cfg = (K, kind)
        return z

    @staticmethod
    def backward(ctx, dz):
        x, y, w, g, b, mean, var = ctx.saved_tensors
        K, kind = ctx.cfg
        taps = K * K
        cin = x.shape[1]
        dz = dz.contiguous().float()
        with torch.cuda.device(x.device):
            dy, dg, db = bn_train_bwd(dz, None, y, g, b, mean, var, True)
            if kind == 3:
                cout = w.shape[1]
                dw = _conv2d_wgrad_b(dy, x, K, 2, tuple(w.shape), cout * taps)            # roles swapped (csrc/train.cu)
                dx = _conv2d_raw_b(dy, w, K, 1, cin, cout * taps, taps) if ctx.needs_input_grad[0] else None
            else:
                dw = _conv2d_wgrad_b(x, dy, K, 2 if kind == 1 else 1, tuple(w.shape), cin * taps)
                if not ctx.needs_input_grad[0]:
                    dx = None
                elif kind == 1:
                    dx = _conv2d_raw_b(dy, w, K, 3, cin, taps, cin * taps)
                else:
                    dx = _conv2d_raw_b(dy, w, K, 2, cin, taps, cin * taps)
        return dx, dw, dg, db, None, None, None


class _Conv2dPlainFn(torch.autograd.Function):
    """Bare `nn.Conv2d(bias=False)`, stride 1 (the 1x1 output heads `out1..3`, `modules/module.py:472-483`)."""

    @staticmethod
    def forward(ctx, x, w):
        x = _lib.require_cuda(x.detach(), "x").contiguous().float()
        w_ = w.detach()
        K = w_.shape[2]
        with torch.cuda.device(x.device):
            y = _conv2d_raw_b(x, w_, K, 0, w_.shape[0], x.shape[1] * K * K, K * K)
        ctx.save_for_backward(x, w_)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, w = ctx.saved_tensors
        K, cin = w.shape[2], x.shape[1]
        dy = dy.contiguous().float()
        with torch.cuda.device(x.device):
            dw = _conv2d_wgrad_b(x, dy, K, 1, tuple(w.shape), cin * K * K)
            dx = _conv2d_raw_b(dy, w, K, 2, cin, K * K, cin * K * K) if ctx.needs_input_grad[0] else None
        return dx, dw


def _block2d(blk, x, kind):
    K = blk.conv.weight.shape[2]
    return _Conv2dBlockFn.apply(x, blk.conv.weight, blk.bn.weight, blk.bn.bias, blk.bn, K, kind)


def featurenet_train_forward(net, x):
    """`FeatureNet.forward` (`modules/module.py:506-543`, unet, three stages) in train mode: every block is an autograd node of its
    own on the library's kernels; the concatenations of `DeConv2dFuse` (`:303-321`) are torch memory copies."""
    if x.shape[2] % 4 or x.shape[3] % 4:
        raise ValueError("FeatureNet needs H and W to be multiples of 4")
    c0 = _block2d(net.conv0[1], _block2d(net.conv0[0], x, 0), 0)
    c1 = _block2d(net.conv1[2], _block2d(net.conv1[1], _block2d(net.conv1[0], c0, 1), 0), 0)
    c2 = _block2d(net.conv2[2], _block2d(net.conv2[1], _block2d(net.conv2[0], c1, 1), 0), 0)
    out = {"stage1": _Conv2dPlainFn.apply(c2, net.out1.weight)}
    f1 = _block2d(net.deconv1.conv, torch.cat((_block2d(net.deconv1.deconv, c2, 3), c1), dim=1), 0)
    out["stage2"] = _Conv2dPlainFn.apply(f1, net.out2.weight)
    f2 = _block2d(net.deconv2.conv, torch.cat((_block2d(net.deconv2.deconv, f1, 3), c0), dim=1), 0)
    out["stage3"] = _Conv2dPlainFn.apply(f2, net.out3.weight)
    return out
