"""The oracle restatement against the committed reference outputs (tests/golden, produced by
oracle/make_golden.py from the unmodified reference).  CPU only."""
import pytest
import torch

from oracle import geometry, hypotheses, regnets, regress, stages, volume
from satmvs_b200 import synth


def maxdiff(a, b):
    return (a.double() - b.double()).abs().max().item()


def test_rpc_localise_project_bit_exact(golden):
    g = golden("rpc_geometry")
    lat, lon = geometry.rpc_localise(g["samp"], g["line"], g["hei"], g["rpcs"][:, 0])
    assert maxdiff(lat, g["lat"]) == 0.0 and maxdiff(lon, g["lon"]) == 0.0
    for v in (1, 2):
        s, l = geometry.rpc_project(lat, lon, g["hei"], g["rpcs"][:, v])
        assert maxdiff(s, g[f"samp{v}"]) == 0.0 and maxdiff(l, g[f"line{v}"]) == 0.0


def test_rpc_qc_form_matches_20_term(golden):
    """tools/rpc_tensor.py localisation->projection (4x4x4 tensor form) == 20-term form."""
    g = golden("rpc_geometry")
    for b in range(2):
        lat, lon = geometry.rpc_localise_qc(g["samp"][b], g["line"][b], g["hei"][b], g["rpcs"][b, 0])
        assert maxdiff(lat, g["lat"][b]) < 1e-12 and maxdiff(lon, g["lon"][b]) < 1e-12
        s, l = geometry.rpc_project_qc(lat, lon, g["hei"][b], g["rpcs"][b, 1])
        assert maxdiff(s, g["samp1"][b]) < 1e-9 and maxdiff(l, g["line1"][b]) < 1e-9


@pytest.mark.parametrize("tag", ["plain", "shifted"])
def test_rpc_warp(golden, tag):
    g = golden(f"rpc_warp_{tag}")
    for v in (1, 2):
        for dk in ("depth4", "depth2"):
            want = g[f"warp{dk[-1]}_v{v}"]
            got = geometry.rpc_warp(g[f"fea{v}"], g["rpcs"][:, v], g["rpcs"][:, 0], g[dk])
            assert maxdiff(got, want) == 0.0
            exp = geometry.rpc_warp(g[f"fea{v}"], g["rpcs"][:, v], g["rpcs"][:, 0], g[dk], sampler="explicit")
            assert maxdiff(exp, want) == 0.0
    assert (g["warp4_v1"] != 0).float().mean() > 0.5


def test_homo_warp(golden):
    g = golden("homo_warp")
    for v in (1, 2):
        for dk in ("depth4", "depth2"):
            want = g[f"warp{dk[-1]}_v{v}"]
            got = geometry.homo_warp(g[f"fea{v}"], g["projs"][:, v], g["projs"][:, 0], g[dk])
            assert maxdiff(got, want) == 0.0
            exp = geometry.homo_warp(g[f"fea{v}"], g["projs"][:, v], g["projs"][:, 0], g[dk], sampler="explicit")
            assert maxdiff(exp, want) == 0.0


def test_qc_warp_equals_20_term(golden):
    g = golden("rpc_warp_qc")
    got = geometry.rpc_warp(g["fea1"], g["rpcs"][:, 1], g["rpcs"][:, 0], g["depth4"])
    assert maxdiff(got, g["warp"]) < 1e-5


@pytest.mark.parametrize("geo", ["rpc", "pinhole"])
def test_stage_train(golden, geo):
    g = golden(f"stage_train_{geo}")
    fe = [g[f"fea{v}"] for v in range(3)]
    var = volume.variance_cost_volume(fe, g["cams"], g["depth_values"], geo)
    assert maxdiff(var, g["var"]) == 0.0
    sd = synth.make_red_weights(8)
    logits = regnets.red_regularization(var, sd)
    assert maxdiff(logits, g["logits"]) < 1e-5
    out = stages.stage_train_red(fe, g["cams"], g["depth_values"], sd, geo)
    assert maxdiff(out["depth"], g["depth"]) < 1e-3 * g["depth"].abs().max().item() * 1e-2
    assert maxdiff(out["photometric_confidence"], g["conf"]) < 1e-5


def test_costregnet(golden):
    g = golden("costregnet")
    assert maxdiff(regnets.costregnet(g["x"], synth.make_costregnet_weights(8)), g["y"]) < 1e-5


def test_red_slice(golden):
    g = golden("red_slice")
    out = regnets.red_slice(g["cost"], g["s1"], g["s2"], g["s3"], g["s4"], synth.make_red_weights(8))
    for got, key in zip(out, ("reg", "n1", "n2", "n3", "n4")):
        assert maxdiff(got, g[key]) < 1e-5


def test_heads(golden):
    g = golden("heads")
    d, c = regress.softargmin_red(g["logits"], g["depth_values"])
    assert maxdiff(d, g["depth"]) == 0.0 and maxdiff(c, g["conf_red"]) == 0.0
    d, c = regress.softargmin_casmvs(g["logits"], g["depth_values"])
    assert maxdiff(d, g["depth_casmvs"]) == 0.0 and maxdiff(c, g["conf_casmvs"]) == 0.0


def test_streaming_head_equals_softmax(golden):
    g = golden("heads")
    B, D, H, W = g["logits"].shape
    head = regress.StreamingSoftArgmin(B, H, W)
    for d in range(D):
        head.update(g["logits"][:, d:d + 1], g["depth_values"][:, d:d + 1])
    depth, conf = head.finish()
    assert maxdiff(depth, g["depth"]) < 1e-3 and maxdiff(conf, g["conf_red"]) < 1e-6


def test_hypotheses(golden):
    g = golden("hypotheses")
    B, D, Himg, Wimg = g["first"].shape
    assert maxdiff(hypotheses.depth_range_samples(g["depth_range"], D, 10.0, (B, Himg, Wimg)), g["first"]) == 0.0
    assert maxdiff(hypotheses.depth_range_samples(g["cur"], 6, 5.0, (B, Himg, Wimg)), g["later"]) == 0.0


@pytest.mark.parametrize("tag,head,wfn", [("red_train", "red_train", synth.make_red_weights),
                                          ("red_pred", "red_pred", synth.make_red_weights),
                                          ("casmvs", "casmvs", synth.make_costregnet_weights)])
def test_cascade(golden, tag, head, wfn):
    g = golden(f"cascade_{tag}")
    feats = [[g[f"fea{s}_{v}"] for v in range(3)] for s in (1, 2, 3)]
    cams = [g[f"cams{s}"] for s in (1, 2, 3)]
    weights = [wfn(c, seed=100 + s) for s, c in enumerate((32, 16, 8))]
    out = stages.cascade(feats, cams, g["depth_range"], weights, img_hw=tuple(g["img_hw"].tolist()),
                         ndepths=tuple(g["ndepths"].tolist()), head=head)
    for s in (1, 2, 3):
        want = g[f"depth{s}"]
        rel = maxdiff(out[f"stage{s}"]["depth"], want) / want.abs().max().item()
        assert rel < 1e-5, (s, rel)
        assert maxdiff(out[f"stage{s}"]["photometric_confidence"], g[f"conf{s}"]) < 1e-4


def test_remap_restatement_equals_cv2_bit_for_bit():
    """oracle.remap against outputs of cv2.remap itself (random coordinates, exact 1/32 ties, border, non-finite) and
    against the gather inside the reference filter (tests/golden/rpc_filter.npz: its own coordinates)."""
    import os
    import numpy as np
    from oracle import remap
    G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    g = np.load(os.path.join(G, "remap_cv2.npz"))
    assert np.array_equal(remap.remap_bilinear(g["src"], g["mapx"], g["mapy"], -999.0), g["out"])
    f = np.load(os.path.join(G, "rpc_filter.npz"))
    got = remap.remap_bilinear(f["depths"][1], f["x_src"].astype(np.float32), f["y_src"].astype(np.float32), -999.0)
    assert np.array_equal(got, f["sampled"])


def test_featurenet_restatement_equals_reference(golden):
    """oracle.regnets.featurenet against outputs of the unmodified reference FeatureNet (eval mode, three views)."""
    g = golden("featurenet")
    sd = synth.make_featurenet_weights(8)
    with torch.no_grad():
        for v in range(3):
            out = regnets.featurenet(g[f"img{v}"], sd)
            for k in ("stage1", "stage2", "stage3"):
                assert maxdiff(out[k], g[f"{k}_v{v}"]) == 0.0


def _rel(a, b):
    a, b = a.detach().double(), b.detach().double()
    return (a - b).abs().max().item() / max(b.abs().max().item(), 1e-30)


def test_training_form_of_the_oracle_equals_reference(golden):
    """The oracle's train-mode restatements (batch-statistics BatchNorm, autograd through the recurrence) against outputs and
    gradients of the UNMODIFIED reference modules in train() mode (`train.py:267-287`; oracle/make_golden.py --training)."""
    g = golden("train_costregnet")
    sd = {k: v.clone().requires_grad_(v.is_floating_point() and "running" not in k)
          for k, v in synth.make_costregnet_weights(8, seed=3).items()}
    x = g["x"].clone().requires_grad_(True)
    y = regnets.costregnet(x, sd, training=True)
    y.backward(g["gy"])
    assert _rel(y, g["y"]) < 1e-5 and _rel(x.grad, g["dx"]) < 1e-4
    for key, name in (("d_conv0_w", "conv0.conv.weight"), ("d_conv6_w", "conv6.conv.weight"), ("d_conv7_w", "conv7.conv.weight"),
                      ("d_conv7_bn_w", "conv7.bn.weight"), ("d_conv7_bn_b", "conv7.bn.bias"), ("d_prob_w", "prob.weight")):
        assert _rel(sd[name].grad, g[key]) < 1e-4, name

    g = golden("train_red")
    sd = {k: v.clone().requires_grad_(True) for k, v in synth.make_red_weights(8, seed=5).items()}
    v = g["volume"].clone().requires_grad_(True)
    lg = regnets.red_regularization(v, sd)
    lg.backward(g["gl"])
    assert _rel(lg, g["logits"]) < 1e-5 and _rel(v.grad, g["dvolume"]) < 1e-4
    for key, name in (("d_gru1_gate_w", "conv_gru1.gate_conv.weight"), ("d_gru1_gate_b", "conv_gru1.gate_conv.bias"),
                      ("d_gru4_out_w", "conv_gru4.output_conv.weight"), ("d_gru2_rn_w", "conv_gru2.reset_gate_norm.weight"),
                      ("d_gru3_on_b", "conv_gru3.output_norm.bias"), ("d_conv2_w", "conv2.conv.weight"),
                      ("d_upconv2_w", "upconv2.conv.weight"), ("d_upconv2d_w", "upconv2d.weight")):
        assert _rel(sd[name].grad, g[key]) < 1e-4, name

    g = golden("train_featurenet")
    sd = {k: v.clone().requires_grad_(v.is_floating_point() and "running" not in k) for k, v in synth.make_featurenet_weights(8).items()}
    out = regnets.featurenet(g["img"], sd, training=True)
    sum((out[k] * g[f"g_{k}"]).sum() for k in out).backward()
    for k in out:
        assert _rel(out[k], g[k]) < 1e-5, k
    for key, name in (("d_conv0_0_w", "conv0.0.conv.weight"), ("d_conv1_0_w", "conv1.0.conv.weight"),
                      ("d_deconv1_deconv_w", "deconv1.deconv.conv.weight"), ("d_deconv2_conv_bn_w", "deconv2.conv.bn.weight"),
                      ("d_out1_w", "out1.weight"), ("d_out3_w", "out3.weight")):
        assert _rel(sd[name].grad, g[key]) < 1e-4, name
