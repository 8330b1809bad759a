"""Training form of CostRegNet (train.py:267-287: model.train(), loss.backward()) on the library's kernels against autograd of
the CPU oracle: batch-statistics BatchNorm, running-stat updates, gradients to the input and to every parameter.
Tolerance: 1e-3 of each tensor's largest magnitude (fp32 on both sides, different summation orders)."""
import pytest
import torch

import satmvs_b200
from oracle import regnets, regress
from satmvs_b200 import synth, training

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def rel(a, b):
    a, b = a.detach().cpu().double(), b.detach().cpu().double()
    return (a - b).abs().max().item() / max(b.abs().max().item(), 1e-30)


@pytest.mark.parametrize("mode,cin,cout,D,H,W", [(0, 8, 16, 4, 8, 16), (1, 8, 16, 4, 8, 16), (3, 16, 8, 2, 4, 8), (0, 5, 3, 3, 5, 7),
                                                  (1, 3, 5, 4, 6, 10), (3, 4, 6, 3, 5, 7)])
def test_conv_primitives_match_autograd(mode, cin, cout, D, H, W):
    """satmvs_conv3d_raw in its forward arrangement, the arrangement that is its data gradient, and satmvs_conv3d_wgrad."""
    import torch.nn.functional as F
    g = torch.Generator().manual_seed(mode * 7 + cin)
    x = torch.randn(1, cin, D, H, W, generator=g, requires_grad=True)
    if mode == 3:
        w = torch.randn(cin, cout, 3, 3, 3, generator=g, requires_grad=True)
        y = F.conv_transpose3d(x, w, stride=2, padding=1, output_padding=1)
    else:
        w = torch.randn(cout, cin, 3, 3, 3, generator=g, requires_grad=True)
        y = F.conv3d(x, w, stride=2 if mode == 1 else 1, padding=1)
    gy = torch.randn(y.shape, generator=g)
    y.backward(gy)
    xd, wd, gyd = x.detach().to(DEV), w.detach().to(DEV), gy.to(DEV)
    if mode == 3:
        got = training.conv3d_raw(xd, wd, 3, cout, 27, cout * 27)
        dx = training.conv3d_raw(gyd, wd, 1, cin, cout * 27, 27)
        dw = training.conv3d_wgrad(gyd, xd, 2, torch.empty_like(wd), cout * 27, 27)
    else:
        got = training.conv3d_raw(xd, wd, mode, cout, cin * 27, 27)
        dx = training.conv3d_raw(gyd, wd, 3 if mode == 1 else 2, cin, 27, cin * 27)
        dw = training.conv3d_wgrad(xd, gyd, 2 if mode == 1 else 1, torch.empty_like(wd), cin * 27, 27)
    assert rel(got, y) < 1e-5
    assert rel(dx, x.grad) < 1e-5
    assert rel(dw, w.grad) < 1e-4


def test_batchnorm_train_matches_autograd():
    import torch.nn.functional as F
    g = torch.Generator().manual_seed(5)
    y = (torch.randn(2, 6, 4, 6, 8, generator=g) * 3 + 1).requires_grad_(True)
    gamma = torch.randn(6, generator=g).requires_grad_(True)
    beta = torch.randn(6, generator=g).requires_grad_(True)
    skip = torch.randn(2, 6, 4, 6, 8, generator=g)
    z = F.relu(F.batch_norm(y, None, None, gamma, beta, training=True, eps=1e-5)) + skip
    gz, gz2 = torch.randn(z.shape, generator=g), torch.randn(z.shape, generator=g)
    z.backward(gz + gz2)
    zd, mean, var = training.bn_train_fwd(y.detach().to(DEV), gamma.detach().to(DEV), beta.detach().to(DEV), True, skip.to(DEV))
    assert rel(zd, z) < 1e-5
    assert rel(mean, y.detach().mean((0, 2, 3, 4))) < 1e-5
    assert rel(var, y.detach().var((0, 2, 3, 4), unbiased=False)) < 1e-5
    dy, dg, db = training.bn_train_bwd(gz.to(DEV), gz2.to(DEV), y.detach().to(DEV), gamma.detach().to(DEV), beta.detach().to(DEV),
                                       mean, var, True)
    assert rel(dy, y.grad) < 1e-4
    assert rel(dg, gamma.grad) < 1e-4
    assert rel(db, beta.grad) < 1e-4


@pytest.mark.parametrize("C,B,D,H,W", [(8, 1, 8, 16, 32), (32, 1, 16, 32, 64), (16, 2, 8, 16, 32)])
def test_costregnet_train_step_matches_oracle_autograd(C, B, D, H, W):
    sd = synth.make_costregnet_weights(C, seed=3)
    net = satmvs_b200.CostRegNet(C, 8)
    net.load_state_dict(sd)
    net = net.to(DEV).train()
    x = synth.make_features(1, 1, B * C * D, H, W, seed=4)[0].view(B, C, D, H, W).abs()
    gen = torch.Generator().manual_seed(9)
    gout = torch.randn(B, 1, D, H, W, generator=gen)

    ref = {k: v.clone().requires_grad_(v.is_floating_point() and "running" not in k) for k, v in sd.items()}
    xr = x.clone().requires_grad_(True)
    want = regnets.costregnet(xr, ref, training=True)
    want.backward(gout)

    xg = x.to(DEV).requires_grad_(True)
    got = net(xg)
    assert got.shape == want.shape and got.requires_grad
    assert rel(got, want) < 1e-3
    got.backward(gout.to(DEV))
    assert rel(xg.grad, xr.grad) < 1e-3
    worst = 0.0
    for name, p in net.named_parameters():
        assert p.grad is not None, name
        r = rel(p.grad, ref[name].grad)
        worst = max(worst, r)
        assert r < 2e-3, (name, r)
    # running statistics after one step: (1 - 0.1) * old + 0.1 * batch (unbiased variance), module.py:348 momentum
    c0 = torch.nn.functional.conv3d(x, sd["conv0.conv.weight"], padding=1)
    m = c0.mean((0, 2, 3, 4))
    v = c0.var((0, 2, 3, 4), unbiased=True)
    assert rel(net.conv0.bn.running_mean, 0.9 * sd["conv0.bn.running_mean"] + 0.1 * m) < 1e-4
    assert rel(net.conv0.bn.running_var, 0.9 * sd["conv0.bn.running_var"] + 0.1 * v) < 1e-4
    assert int(net.conv0.bn.num_batches_tracked) == int(sd.get("conv0.bn.num_batches_tracked", 0)) + 1


def test_eval_after_train_uses_the_updated_running_statistics():
    sd = synth.make_costregnet_weights(8, seed=3)
    net = satmvs_b200.CostRegNet(8, 8)
    net.load_state_dict(sd)
    net = net.to(DEV).train()
    x = synth.make_features(1, 1, 8 * 8, 16, 32, seed=4)[0].view(1, 8, 8, 16, 32).abs()
    net(x.to(DEV))
    net.eval()
    with torch.no_grad():
        got = net(x.to(DEV))
    sd2 = {k: v.detach().cpu() for k, v in net.state_dict().items()}
    want = regnets.costregnet(x, sd2)
    assert rel(got, want) < 1e-3


@pytest.mark.parametrize("head", ["casmvs", "red"])
def test_softargmin_head_gradient(head):
    gen = torch.Generator().manual_seed(2)
    logits = torch.randn(2, 12, 9, 13, generator=gen).requires_grad_(True)
    dv = (torch.linspace(400, 600, 12).view(1, 12, 1, 1) + torch.randn(2, 12, 9, 13, generator=gen)).contiguous()
    want, _ = (regress.softargmin_casmvs if head == "casmvs" else regress.softargmin_red)(logits, dv)
    gd = torch.randn(2, 9, 13, generator=gen)
    want.backward(gd)
    lg = logits.detach().to(DEV).requires_grad_(True)
    depth, conf = training.softargmin_train(lg, dv.to(DEV), head)
    assert rel(depth, want) < 1e-5 and not conf.requires_grad
    depth.backward(gd.to(DEV))
    assert rel(lg.grad, logits.grad) < 1e-4


@pytest.mark.parametrize("C,B,D,H,W", [(8, 1, 3, 16, 24), (32, 1, 4, 32, 32), (16, 2, 5, 16, 16)])
def test_red_regulariser_backward_matches_oracle_autograd(C, B, D, H, W):
    """loss.backward() through RED_Regularization (train.py:284): gradient to the volume and to all 48 parameters against autograd
    of the oracle's restatement of modules/module.py:614-649."""
    sd = synth.make_red_weights(C, seed=5)
    net = satmvs_b200.RED_Regularization(C, 8)
    net.load_state_dict(sd)
    net = net.to(DEV).train()
    x = synth.make_features(1, 1, B * C * D, H, W, seed=6)[0].view(B, C, D, H, W).abs()
    gen = torch.Generator().manual_seed(10)
    gout = torch.randn(B, D, H, W, generator=gen)

    ref = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    xr = x.clone().requires_grad_(True)
    want = regnets.red_regularization(xr, ref)
    want.backward(gout)

    xg = x.to(DEV).requires_grad_(True)
    got = net(xg)
    assert got.requires_grad and rel(got, want) < 1e-4
    got.backward(gout.to(DEV))
    assert rel(xg.grad, xr.grad) < 1e-3
    for name, p in net.named_parameters():
        assert p.grad is not None, name
        r = rel(p.grad, ref[name].grad)
        assert r < 2e-3, (name, r)


def test_red_training_stage_reaches_the_feature_maps():
    """Fused sweep -> RED regulariser -> head, loss.backward(): gradients arrive at the feature maps and the parameters and match
    autograd of the oracle stage."""
    from oracle import volume
    B, V, C, D, H, W = 1, 3, 8, 4, 16, 24
    fe = synth.make_features(B, V, C, H, W, seed=2)
    rp = synth.make_rpc_stack(B, V, H, W)
    dv = synth.make_depth_planes(B, D, H, W)
    sd = synth.make_red_weights(C)
    gt = torch.full((B, H, W), float(dv.mean()))
    ref = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    fr = [f.clone().requires_grad_(True) for f in fe]
    logits = regnets.red_regularization(volume.variance_cost_volume(fr, rp, dv, "rpc"), ref)
    want, _ = regress.softargmin_red(logits, dv)
    torch.nn.functional.smooth_l1_loss(want, gt).backward()

    net = satmvs_b200.RED_Regularization(C, 8)
    net.load_state_dict(sd)
    net = net.to(DEV).train()
    fg = [f.to(DEV).requires_grad_(True) for f in fe]
    out = satmvs_b200.stage_train_red(fg, rp, dv.to(DEV), net, "rpc")
    assert rel(out["depth"], want) < 1e-5
    torch.nn.functional.smooth_l1_loss(out["depth"], gt.to(DEV)).backward()
    for a, b in zip(fg, fr):
        assert rel(a.grad, b.grad) < 2e-3
    assert rel(net.conv_gru1.gate_conv.weight.grad, ref["conv_gru1.gate_conv.weight"].grad) < 2e-3
    assert rel(net.upconv1.conv.weight.grad, ref["upconv1.conv.weight"].grad) < 2e-3
    # (upconv2d.bias shifts every logit of a pixel alike: the soft-argmin is invariant, its exact gradient is zero)
    assert net.upconv2d.bias.grad.abs().max().item() < 1e-4 * net.upconv2d.weight.grad.abs().max().item()


@pytest.mark.parametrize("K,mode,cin,cout,H,W", [(5, 1, 8, 16, 16, 24), (5, 1, 3, 5, 12, 20), (1, 0, 32, 32, 8, 12), (1, 0, 5, 7, 9, 11),
                                                  (3, 0, 8, 8, 16, 16), (3, 3, 16, 8, 8, 12)])
def test_conv2d_primitives_match_autograd(K, mode, cin, cout, H, W):
    """satmvs_conv2d_raw / satmvs_conv2d_wgrad for FeatureNet's layer shapes: 5x5 stride 2, 1x1, 3x3, transposed 3x3."""
    import torch.nn.functional as F
    g = torch.Generator().manual_seed(K * 10 + mode)
    x = torch.randn(2, cin, H, W, generator=g, requires_grad=True)
    taps = K * K
    if mode == 3:
        w = torch.randn(cin, cout, K, K, generator=g, requires_grad=True)
        y = F.conv_transpose2d(x, w, stride=2, padding=K // 2, output_padding=1)
    else:
        w = torch.randn(cout, cin, K, K, generator=g, requires_grad=True)
        y = F.conv2d(x, w, stride=2 if mode == 1 else 1, padding=K // 2)
    gy = torch.randn(y.shape, generator=g)
    y.backward(gy)
    xd, wd, gyd = x.detach().to(DEV), w.detach().to(DEV), gy.to(DEV)
    if mode == 3:
        got = training._conv2d_raw_b(xd, wd, K, 3, cout, taps, cout * taps)
        dx = training._conv2d_raw_b(gyd, wd, K, 1, cin, cout * taps, taps)
        dw = training._conv2d_wgrad_b(gyd, xd, K, 2, tuple(wd.shape), cout * taps)
    else:
        got = training._conv2d_raw_b(xd, wd, K, mode, cout, cin * taps, taps)
        dx = training._conv2d_raw_b(gyd, wd, K, 3 if mode == 1 else 2, cin, taps, cin * taps)
        dw = training._conv2d_wgrad_b(xd, gyd, K, 2 if mode == 1 else 1, tuple(wd.shape), cin * taps)
    assert rel(got, y) < 1e-5
    assert rel(dx, x.grad) < 1e-5
    assert rel(dw, w.grad) < 1e-4


@pytest.mark.parametrize("B,H,W", [(1, 32, 48), (2, 16, 16)])
def test_featurenet_train_step_matches_oracle_autograd(B, H, W):
    sd = synth.make_featurenet_weights(8)
    net = satmvs_b200.FeatureNet(8)
    net.load_state_dict(sd)
    net = net.to(DEV).train()
    gen = torch.Generator().manual_seed(21)
    x = torch.rand(B, 3, H, W, generator=gen)
    ref = {k: v.clone().requires_grad_(v.is_floating_point() and "running" not in k) for k, v in sd.items()}
    want = regnets.featurenet(x, ref, training=True)
    gouts = {k: torch.randn(v.shape, generator=gen) for k, v in want.items()}
    sum((want[k] * gouts[k]).sum() for k in want).backward()
    got = net(x.to(DEV))
    for k in want:
        assert rel(got[k], want[k]) < 1e-4, k
    sum((got[k] * gouts[k].to(DEV)).sum() for k in got).backward()
    for name, p in net.named_parameters():
        assert p.grad is not None, name
        assert rel(p.grad, ref[name].grad) < 2e-3, (name, rel(p.grad, ref[name].grad))
    assert int(net.conv0[0].bn.num_batches_tracked) == int(sd.get("conv0.0.bn.num_batches_tracked", 0)) + 1


@pytest.mark.parametrize("per_pixel", [False, True])
def test_depth_regression_gradient(per_pixel):
    """The patched reference trains through `depth_regression(F.softmax(...), depth_values)` (casmvs.py:66-68)."""
    gen = torch.Generator().manual_seed(4)
    logits = torch.randn(2, 6, 10, 12, generator=gen, requires_grad=True)
    dv = torch.rand(2, 6, 10, 12, generator=gen) * 100 if per_pixel else torch.rand(2, 6, generator=gen) * 100
    want = regress.depth_regression(torch.softmax(logits, 1), dv)
    gd = torch.randn(2, 10, 12, generator=gen)
    want.backward(gd)
    lg = logits.detach().to(DEV).requires_grad_(True)
    got = satmvs_b200.depth_regression(torch.softmax(lg, 1), dv.to(DEV))
    assert rel(got, want) < 1e-5
    got.backward(gd.to(DEV))
    assert rel(lg.grad, logits.grad) < 1e-5


def test_depth_range_samples_gradient():
    """casred.py:132-154 keeps the previous depth attached: d samples / d cur_depth = 1 per plane."""
    from oracle import hypotheses
    gen = torch.Generator().manual_seed(8)
    cur = (torch.rand(2, 12, 16, generator=gen) * 50 + 400).requires_grad_(True)
    nd, interval = 8, 2.5
    lo = cur - nd / 2 * interval
    want = lo.unsqueeze(1) + torch.arange(nd).view(1, -1, 1, 1) * (((cur + nd / 2 * interval) - lo) / (nd - 1)).unsqueeze(1)
    gd = torch.randn(2, nd, 12, 16, generator=gen)
    want.backward(gd)
    c = cur.detach().to(DEV).requires_grad_(True)
    got = satmvs_b200.get_depth_range_samples(c, nd, interval, shape=(2, 12, 16))
    assert rel(got, want) < 1e-6
    got.backward(gd.to(DEV))
    assert rel(c.grad, cur.grad) < 1e-5


def test_whole_casmvs_network_training_step():
    """Images -> FeatureNet (train) per view -> three CasMVS stages (fused sweep, train-mode CostRegNet, head; depth detached between
    stages as `casmvs.py:145-146`) -> sum of stage losses -> backward: gradients at the first FeatureNet filter and at every stage's
    regulariser against autograd of the oracle network."""
    from oracle import hypotheses, volume
    B, V, Himg, Wimg = 1, 3, 128, 128      # large enough that every BatchNorm sees >= 32 voxels per channel
    nds, ratios, scales = (16, 16, 8), (4, 2, 1), (4, 2, 1)
    gen = torch.Generator().manual_seed(31)
    imgs = [torch.rand(B, 3, Himg, Wimg, generator=gen) for _ in range(V)]
    fsd = synth.make_featurenet_weights(8)
    rsd = [synth.make_costregnet_weights(c, seed=40 + i) for i, c in enumerate((32, 16, 8))]
    cams = [synth.make_rpc_stack(B, V, Himg // s, Wimg // s) for s in scales]
    rng = torch.tensor([[400.0, 600.0]])
    gt = torch.full((B, Himg, Wimg), 500.0)

    # oracle network with autograd
    fref = {k: v.clone().requires_grad_(v.is_floating_point() and "running" not in k) for k, v in fsd.items()}
    rref = [{k: v.clone().requires_grad_(v.is_floating_point() and "running" not in k) for k, v in sd.items()} for sd in rsd]
    feats = [regnets.featurenet(i, fref, training=True) for i in imgs]
    depth, loss_ref = None, 0.0
    for s in range(3):
        h, w = Himg // scales[s], Wimg // scales[s]
        dv = hypotheses.stage_hypotheses(None if depth is None else depth.detach(), rng, nds[s], ratios[s] * 2.5, (Himg, Wimg), scales[s])
        var = volume.variance_cost_volume([f[f"stage{s + 1}"] for f in feats], cams[s], dv, "rpc")
        logits = regnets.costregnet(var, rref[s], training=True).squeeze(1)
        depth, _ = regress.softargmin_casmvs(logits, dv)
        loss_ref = loss_ref + torch.nn.functional.smooth_l1_loss(depth, gt[:, :h, :w])
    loss_ref.backward()

    fnet = satmvs_b200.FeatureNet(8)
    fnet.load_state_dict(fsd)
    fnet = fnet.to(DEV).train()
    regs = []
    for sd, c in zip(rsd, (32, 16, 8)):
        r = satmvs_b200.CostRegNet(c, 8)
        r.load_state_dict(sd)
        regs.append(r.to(DEV).train())
    fg = fnet.forward_views([i.to(DEV) for i in imgs])
    out = satmvs_b200.cascade([[f[f"stage{s + 1}"] for f in fg] for s in range(3)], cams, rng.to(DEV), regs, img_hw=(Himg, Wimg),
                              ndepths=nds, depth_interals_ratio=ratios, min_interval=2.5, scales=scales, geo_model="rpc", head="casmvs")
    loss = 0.0
    for s in range(3):
        h, w = Himg // scales[s], Wimg // scales[s]
        loss = loss + torch.nn.functional.smooth_l1_loss(out[f"stage{s + 1}"]["depth"], gt[:, :h, :w].to(DEV))
    assert abs(loss.item() - loss_ref.item()) < 1e-4 * abs(loss_ref.item())
    loss.backward()
    # a deep fp32 network with batch-statistics BatchNorm amplifies rounding differences: direction and size of each gradient
    errs = {}
    def close(a, b, what):
        a, b = a.detach().cpu().double().flatten(), b.double().flatten()
        cos = torch.dot(a, b).item() / (a.norm().item() * b.norm().item() + 1e-300)
        errs[what] = (round(1.0 - cos, 6), round(a.norm().item() / b.norm().item(), 4))
    close(fnet.conv0[0].conv.weight.grad, fref["conv0.0.conv.weight"].grad, "featurenet conv0.0")
    close(fnet.conv2[2].conv.weight.grad, fref["conv2.2.conv.weight"].grad, "featurenet conv2.2")
    close(fnet.out3.weight.grad, fref["out3.weight"].grad, "featurenet out3")
    for s in range(3):
        close(regs[s].conv0.conv.weight.grad, rref[s]["conv0.conv.weight"].grad, f"stage {s + 1} conv0")
        close(regs[s].prob.weight.grad, rref[s]["prob.weight"].grad, f"stage {s + 1} prob")
    print("whole-network training step: (1 - cosine, norm ratio) per gradient:", errs)
    for what, (c, r) in errs.items():
        assert c < 1e-3 and abs(r - 1.0) < 2e-2, (what, errs)


def test_training_goldens_from_the_unmodified_reference(golden):
    """CUDA train() + backward against outputs / gradients / running statistics the UNMODIFIED reference modules produced in
    train() mode (oracle/make_golden.py --training): CostRegNet, RED_Regularization, FeatureNet."""
    def todev(t):
        return t.to(DEV)

    g = golden("train_costregnet")
    net = satmvs_b200.CostRegNet(8, 8)
    net.load_state_dict(synth.make_costregnet_weights(8, seed=3))
    net = net.to(DEV).train()
    x = todev(g["x"]).requires_grad_(True)
    y = net(x)
    y.backward(todev(g["gy"]))
    assert rel(y, g["y"]) < 1e-3 and rel(x.grad, g["dx"]) < 1e-3
    for key, p in (("d_conv0_w", net.conv0.conv.weight), ("d_conv6_w", net.conv6.conv.weight), ("d_conv7_w", net.conv7.conv.weight),
                   ("d_conv7_bn_w", net.conv7.bn.weight), ("d_conv7_bn_b", net.conv7.bn.bias), ("d_prob_w", net.prob.weight)):
        assert rel(p.grad, g[key]) < 2e-3, key
    for key, buf in (("rm_conv0", net.conv0.bn.running_mean), ("rv_conv0", net.conv0.bn.running_var),
                     ("rm_conv11", net.conv11.bn.running_mean), ("rv_conv11", net.conv11.bn.running_var)):
        assert rel(buf, g[key]) < 1e-4, key

    g = golden("train_red")
    red = satmvs_b200.RED_Regularization(8, 8)
    red.load_state_dict(synth.make_red_weights(8, seed=5))
    red = red.to(DEV).train()
    v = todev(g["volume"]).requires_grad_(True)
    lg = red(v)
    lg.backward(todev(g["gl"]))
    assert rel(lg, g["logits"]) < 1e-4 and rel(v.grad, g["dvolume"]) < 1e-3
    for key, p in (("d_gru1_gate_w", red.conv_gru1.gate_conv.weight), ("d_gru1_gate_b", red.conv_gru1.gate_conv.bias),
                   ("d_gru4_out_w", red.conv_gru4.output_conv.weight), ("d_gru2_rn_w", red.conv_gru2.reset_gate_norm.weight),
                   ("d_gru3_on_b", red.conv_gru3.output_norm.bias), ("d_conv2_w", red.conv2.conv.weight),
                   ("d_upconv2_w", red.upconv2.conv.weight), ("d_upconv2d_w", red.upconv2d.weight)):
        assert rel(p.grad, g[key]) < 2e-3, key

    g = golden("train_featurenet")
    fnet = satmvs_b200.FeatureNet(8)
    fnet.load_state_dict(synth.make_featurenet_weights(8))
    fnet = fnet.to(DEV).train()
    out = fnet(todev(g["img"]))
    sum((out[k] * todev(g[f"g_{k}"])).sum() for k in out).backward()
    for k in out:
        assert rel(out[k], g[k]) < 1e-4, k
    for key, p in (("d_conv0_0_w", fnet.conv0[0].conv.weight), ("d_conv1_0_w", fnet.conv1[0].conv.weight),
                   ("d_deconv1_deconv_w", fnet.deconv1.deconv.conv.weight), ("d_deconv2_conv_bn_w", fnet.deconv2.conv.bn.weight),
                   ("d_out1_w", fnet.out1.weight), ("d_out3_w", fnet.out3.weight)):
        assert rel(p.grad, g[key]) < 2e-3, key
    assert rel(fnet.conv0[0].bn.running_mean, g["rm_conv0_0"]) < 1e-4 and rel(fnet.conv0[0].bn.running_var, g["rv_conv0_0"]) < 1e-4
