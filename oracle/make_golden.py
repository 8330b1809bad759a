"""Generate tests/golden/*.npz by running the UNMODIFIED reference (imported from
/root/reference, see oracle/reference_loader.py) on seeded synthetic inputs.

Run in the build container:   python -m oracle.make_golden
The GPU box has no /root/reference; it only reads the committed .npz files.
TEST INFRASTRUCTURE (see oracle/__init__.py).
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import reference_loader  # noqa: E402
from satmvs_b200 import synth  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def _np(t):
    return t.detach().cpu().numpy()


def save(name, **arrays):
    os.makedirs(OUT, exist_ok=True)
    path = os.path.join(OUT, name + ".npz")
    np.savez_compressed(path, **{k: (_np(v) if torch.is_tensor(v) else np.asarray(v)) for k, v in arrays.items()})
    print(f"{name}: {os.path.getsize(path) / 1024:.0f} KiB")


def golden_rpc_geometry(ref):
    """fp64 localisation -> projection on a flat point list (`RPC_Photo2Obj`, `RPC_Obj2Photo`)."""
    B, N = 2, 257
    rng = np.random.default_rng(21)
    rpcs = synth.make_rpc_stack(B, 3, 96, 192, shift_geo=True)
    samp = torch.from_numpy(rng.uniform(-5, 197, (B, N)))
    line = torch.from_numpy(rng.uniform(-5, 101, (B, N)))
    hei = torch.from_numpy(rng.uniform(-50, 1050, (B, N)))
    coef = torch.ones(B, N, 20, dtype=torch.double)
    lat, lon = ref.warping.RPC_Photo2Obj(samp, line, hei, rpcs[:, 0], coef)
    s1, l1 = ref.warping.RPC_Obj2Photo(lat, lon, hei, rpcs[:, 1], coef)
    s2, l2 = ref.warping.RPC_Obj2Photo(lat, lon, hei, rpcs[:, 2], coef)
    save("rpc_geometry", rpcs=rpcs, samp=samp, line=line, hei=hei, lat=lat, lon=lon,
         samp1=s1, line1=l1, samp2=s2, line2=l2)


def golden_warps(ref):
    B, V, C, D, H, W = 2, 3, 4, 6, 24, 40
    fe = synth.make_features(B, V, C, H, W, seed=3)
    coef = torch.ones(B, D * H * W, 20, dtype=torch.double)
    for tag, shift in (("plain", False), ("shifted", True)):
        rp = synth.make_rpc_stack(B, V, H, W, shift_geo=shift)
        rp[1] = torch.from_numpy(np.stack([synth.make_rpc(v, H, W, seed=50 + v, shift_geo=shift) for v in range(V)]))
        dv4 = synth.make_depth_planes(B, D, H, W, seed=5)
        dv2 = synth.make_depth_planes(B, D, H, W, per_pixel=False)
        out = {"rpcs": rp, "depth4": dv4, "depth2": dv2}
        for v in (1, 2):
            out[f"fea{v}"] = fe[v]
            out[f"warp4_v{v}"] = ref.warping.rpc_warping(fe[v], rp[:, v], rp[:, 0], dv4, coef)
            out[f"warp2_v{v}"] = ref.warping.rpc_warping(fe[v], rp[:, v], rp[:, 0], dv2, coef)
        save(f"rpc_warp_{tag}", **out)
    pp = synth.make_pinhole_stack(B, V, H, W)
    dv4 = synth.make_depth_planes(B, D, H, W, lo=90, hi=110, jitter=0.2, seed=6)
    dv2 = synth.make_depth_planes(B, D, H, W, lo=90, hi=110, per_pixel=False)
    out = {"projs": pp, "depth4": dv4, "depth2": dv2}
    for v in (1, 2):
        out[f"fea{v}"] = fe[v]
        out[f"warp4_v{v}"] = ref.warping.homo_warping(fe[v], pp[:, v], pp[:, 0], dv4)
        out[f"warp2_v{v}"] = ref.warping.homo_warping(fe[v], pp[:, v], pp[:, 0], dv2)
    save("homo_warp", **out)


def golden_qc(ref):
    """20-term vs quaternary-cubic einsum form (`rpc_warping_enisum`, `warping.py:139-178`)."""
    B, V, C, D, H, W = 1, 2, 2, 3, 12, 20
    fe = synth.make_features(B, V, C, H, W, seed=4)
    rp = synth.make_rpc_stack(B, V, H, W)
    dv = synth.make_depth_planes(B, D, H, W, seed=8)

    def as_dict(r):   # the dict format `rpc_warping_enisum` takes, built like `data_io.py:123-160`
        from oracle.geometry import qc_tensor
        d = {"line_off": r[:, 0], "samp_off": r[:, 1], "lat_off": r[:, 2], "lon_off": r[:, 3], "height_off": r[:, 4],
             "line_scale": r[:, 5], "samp_scale": r[:, 6], "lat_scale": r[:, 7], "lon_scale": r[:, 8],
             "height_scale": r[:, 9]}
        for k, o in (("line_num", 10), ("line_den", 30), ("samp_num", 50), ("samp_den", 70),
                     ("lat_num", 90), ("lat_den", 110), ("lon_num", 130), ("lon_den", 150)):
            d[k + "_tensor"] = torch.stack([qc_tensor(r[b, o:o + 20]) for b in range(r.shape[0])])
        return d

    out = ref.warping.rpc_warping_enisum(fe[1], as_dict(rp[:, 1]), as_dict(rp[:, 0]), dv)
    save("rpc_warp_qc", rpcs=rp, depth4=dv, fea1=fe[1], warp=out)


def golden_stage_train(ref):
    """`compute_depth_when_train` on given features: captures the variance volume via the
    regulariser hook, then the real RED regulariser + heads."""
    B, V, C, D, H, W = 1, 3, 8, 8, 16, 24
    fe = synth.make_features(B, V, C, H, W, seed=9)
    sd = synth.make_red_weights(C)
    reg = ref.module.RED_Regularization(C, 8)
    reg.load_state_dict(sd)
    reg.eval()
    grabbed = {}

    def hook(vol):
        grabbed["var"] = vol.clone()
        grabbed["logits"] = reg(vol)
        return grabbed["logits"]

    for geo, cams, dv in (("rpc", synth.make_rpc_stack(B, V, H, W), synth.make_depth_planes(B, D, H, W, seed=10)),
                          ("pinhole", synth.make_pinhole_stack(B, V, H, W),
                           synth.make_depth_planes(B, D, H, W, lo=90, hi=110, jitter=0.2, seed=10))):
        with torch.no_grad():
            out = ref.casred.compute_depth_when_train(fe, cams, dv, D, hook, geo, False)
        save(f"stage_train_{geo}", cams=cams, depth_values=dv, **{f"fea{v}": fe[v] for v in range(V)},
             var=grabbed["var"], logits=grabbed["logits"], depth=out["depth"], conf=out["photometric_confidence"])


def golden_regularisers(ref):
    C, D, H, W = 8, 8, 16, 24
    rng = np.random.default_rng(31)
    x = torch.from_numpy(rng.standard_normal((1, C, D, H, W), dtype=np.float32)).abs_()
    sd = synth.make_costregnet_weights(C)
    m = ref.module.CostRegNet(C, 8)
    m.load_state_dict(sd)
    m.eval()
    with torch.no_grad():
        save("costregnet", x=x, y=m(x))
    sdr = synth.make_red_weights(C)
    s = ref.module.slice_RED_Regularization(C, 8)
    s.load_state_dict(sdr)
    s.eval()
    st = [torch.from_numpy(rng.standard_normal((1, c, H // k, W // k), dtype=np.float32) * 0.5)
          for c, k in ((8, 1), (16, 2), (32, 4), (64, 8))]
    with torch.no_grad():
        reg, n1, n2, n3, n4 = s(x[:, :, 0], *st)
    save("red_slice", cost=x[:, :, 0], s1=st[0], s2=st[1], s3=st[2], s4=st[3], reg=reg, n1=n1, n2=n2, n3=n3, n4=n4)


def golden_featurenet(ref):
    """`FeatureNet(8, num_stage=3, arch_mode="unet")` in eval mode on three small views (`modules/module.py:442-543`)."""
    import contextlib
    sd = synth.make_featurenet_weights(8)
    with open(os.devnull, "w") as devnull, contextlib.redirect_stdout(devnull):
        m = ref.module.FeatureNet(base_channels=8, stride=4, num_stage=3, arch_mode="unet")
    m.load_state_dict(sd)
    m.eval()
    g = torch.Generator().manual_seed(23)
    imgs = [torch.rand(1, 3, 40, 56, generator=g) for _ in range(3)]
    arrays = {}
    with torch.no_grad():
        for v, im in enumerate(imgs):
            out = m(im)
            arrays[f"img{v}"] = im
            for k in ("stage1", "stage2", "stage3"):
                arrays[f"{k}_v{v}"] = out[k]
    save("featurenet", **arrays)


def golden_heads(ref):
    rng = np.random.default_rng(41)
    B, D, H, W = 2, 8, 12, 20
    logits = torch.from_numpy(rng.standard_normal((B, D, H, W), dtype=np.float32) * 3)
    dv = synth.make_depth_planes(B, D, H, W, seed=12)
    import torch.nn.functional as F
    p = F.softmax(logits, dim=1)
    depth = ref.module.depth_regression(p, depth_values=dv)
    conf_red = p.max(1)[0]
    # CasMVS confidence: run the reference DepthNet with a regulariser returning fixed logits
    net = ref.casmvs.DepthNet()
    net.eval()
    fe = synth.make_features(B, 2, 2, H, W, seed=13)
    cams = synth.make_rpc_stack(B, 2, H, W)
    with torch.no_grad():
        out = net(fe, cams, dv, D, lambda vol: logits.unsqueeze(1), "rpc")
    save("heads", logits=logits, depth_values=dv, depth=depth, conf_red=conf_red,
         depth_casmvs=out["depth"], conf_casmvs=out["photometric_confidence"])


def golden_hypotheses(ref):
    B, Himg, Wimg = 2, 32, 64
    rng = np.random.default_rng(51)
    rng_depth = torch.tensor([[0.0, 1000.0], [100.0, 900.0]])
    first = ref.depth_range.get_depth_range_samples(rng_depth, 8, 10.0, "cpu", torch.float32, [B, Himg, Wimg])
    cur = torch.from_numpy(rng.uniform(200, 800, (B, Himg, Wimg)).astype(np.float32))
    later = ref.depth_range.get_depth_range_samples(cur, 6, 5.0, "cpu", torch.float32, [B, Himg, Wimg])
    save("hypotheses", depth_range=rng_depth, first=first, cur=cur, later=later)


def golden_cascade(ref):
    """Whole networks end to end (FeatureNet outputs captured as the path's inputs)."""
    B, V, Himg, Wimg = 1, 3, 64, 96
    rng = np.random.default_rng(61)
    imgs = torch.from_numpy(rng.standard_normal((B, V, 3, Himg, Wimg), dtype=np.float32))
    base = synth.make_rpc_stack(B, V, Himg, Wimg)
    cams = {"stage1": torch.from_numpy(synth.rescale_rpc(_np(base), 4)),
            "stage2": torch.from_numpy(synth.rescale_rpc(_np(base), 2)), "stage3": base}
    depth_range = torch.tensor([[0.0, 1000.0]])
    chans = (32, 16, 8)
    for tag, ctor, wfn, nd in (("red_train", ref.casred.CascadeREDNet, synth.make_red_weights, [8, 4, 4]),
                               ("red_pred", ref.casred.Infer_CascadeREDNet, synth.make_red_weights, [8, 4, 4]),
                               ("casmvs", ref.casmvs.CascadeMVSNet, synth.make_costregnet_weights, [16, 8, 8])):
        torch.manual_seed(0)
        kw = dict(arch_mode="unet") if tag == "casmvs" else {}
        import contextlib
        with open(os.devnull, "w") as dn, contextlib.redirect_stdout(dn):
            net = ctor("rpc", ndepths=nd, **kw)
        for s, c in enumerate(chans):
            net.cost_regularization[s].load_state_dict(wfn(c, seed=100 + s))
        net.eval()
        feats = []
        hook = net.feature.register_forward_hook(lambda m, i, o: feats.append({k: v.clone() for k, v in o.items()}))
        with torch.no_grad():
            out = net(imgs, cams, depth_range)
        hook.remove()
        arrays = {"depth_range": depth_range, "ndepths": np.array(nd), "img_hw": np.array([Himg, Wimg])}
        for s in range(3):
            arrays[f"cams{s + 1}"] = cams[f"stage{s + 1}"]
            for v in range(V):
                arrays[f"fea{s + 1}_{v}"] = feats[v][f"stage{s + 1}"]
            arrays[f"depth{s + 1}"] = out[f"stage{s + 1}"]["depth"]
            arrays[f"conf{s + 1}"] = out[f"stage{s + 1}"]["photometric_confidence"]
        save(f"cascade_{tag}", **arrays)



def golden_training(ref):
    """train() mode + loss.backward() of the UNMODIFIED reference modules (`train.py:267-287`): outputs, gradients to the input
    and to a few parameters, running statistics after the step -- pins the oracle's (and the library's) training form."""
    import contextlib
    rng = np.random.default_rng(41)
    # CostRegNet: batch-statistics BatchNorm3d
    C, D, H, W = 8, 8, 16, 32
    x = torch.from_numpy(rng.standard_normal((2, C, D, H, W), dtype=np.float32)).abs_().requires_grad_(True)
    m = ref.module.CostRegNet(C, 8)
    m.load_state_dict(synth.make_costregnet_weights(C, seed=3))
    m.train()
    y = m(x)
    gy = torch.from_numpy(rng.standard_normal(tuple(y.shape), dtype=np.float32))
    y.backward(gy)
    save("train_costregnet", x=x, gy=gy, y=y, dx=x.grad,
         d_conv0_w=m.conv0.conv.weight.grad, d_conv6_w=m.conv6.conv.weight.grad, d_conv7_w=m.conv7.conv.weight.grad,
         d_conv7_bn_w=m.conv7.bn.weight.grad, d_conv7_bn_b=m.conv7.bn.bias.grad, d_prob_w=m.prob.weight.grad,
         rm_conv0=m.conv0.bn.running_mean, rv_conv0=m.conv0.bn.running_var, rm_conv11=m.conv11.bn.running_mean,
         rv_conv11=m.conv11.bn.running_var)
    # RED regulariser: backward through the recurrence
    C, D, H, W = 8, 4, 16, 24
    v = torch.from_numpy(rng.standard_normal((1, C, D, H, W), dtype=np.float32)).abs_().requires_grad_(True)
    r = ref.module.RED_Regularization(C, 8)
    r.load_state_dict(synth.make_red_weights(C, seed=5))
    r.train()
    lg = r(v)
    gl = torch.from_numpy(rng.standard_normal(tuple(lg.shape), dtype=np.float32))
    lg.backward(gl)
    save("train_red", volume=v, gl=gl, logits=lg, dvolume=v.grad,
         d_gru1_gate_w=r.conv_gru1.gate_conv.weight.grad, d_gru1_gate_b=r.conv_gru1.gate_conv.bias.grad,
         d_gru4_out_w=r.conv_gru4.output_conv.weight.grad, d_gru2_rn_w=r.conv_gru2.reset_gate_norm.weight.grad,
         d_gru3_on_b=r.conv_gru3.output_norm.bias.grad, d_conv2_w=r.conv2.conv.weight.grad, d_upconv2_w=r.upconv2.conv.weight.grad,
         d_upconv2d_w=r.upconv2d.weight.grad)
    # FeatureNet: batch-statistics BatchNorm2d, three output heads
    with open(os.devnull, "w") as devnull, contextlib.redirect_stdout(devnull):
        f = ref.module.FeatureNet(base_channels=8, stride=4, num_stage=3, arch_mode="unet")
    f.load_state_dict(synth.make_featurenet_weights(8))
    f.train()
    img = torch.from_numpy(rng.random((2, 3, 32, 48), dtype=np.float32))
    out = f(img)
    gs = {k: torch.from_numpy(rng.standard_normal(tuple(o.shape), dtype=np.float32)) for k, o in out.items()}
    sum((out[k] * gs[k]).sum() for k in out).backward()
    save("train_featurenet", img=img, **{f"g_{k}": g for k, g in gs.items()}, **{k: o for k, o in out.items()},
         d_conv0_0_w=f.conv0[0].conv.weight.grad, d_conv1_0_w=f.conv1[0].conv.weight.grad, d_deconv1_deconv_w=f.deconv1.deconv.conv.weight.grad,
         d_deconv2_conv_bn_w=f.deconv2.conv.bn.weight.grad, d_out1_w=f.out1.weight.grad, d_out3_w=f.out3.weight.grad,
         rm_conv0_0=f.conv0[0].bn.running_mean, rv_conv0_0=f.conv0[0].bn.running_var)


def main_featurenet_only():
    golden_featurenet(reference_loader.load())


def main():
    torch.set_num_threads(8)
    ref = reference_loader.load()
    golden_rpc_geometry(ref)
    golden_warps(ref)
    golden_qc(ref)
    golden_stage_train(ref)
    golden_regularisers(ref)
    golden_heads(ref)
    golden_hypotheses(ref)
    golden_cascade(ref)
    golden_featurenet(ref)
    golden_training(ref)


if __name__ == "__main__":
    if "--featurenet" in sys.argv:
        main_featurenet_only()
    elif "--training" in sys.argv:
        golden_training(reference_loader.load())
    else:
        main()
