"""Parity AT the BASELINE.json configurations themselves (not reduced shapes): the CUDA path through the public
operator API against the CPU oracle on the same seeded synthetic inputs.  north_star bound: final depth within 1e-3
relative L-inf (fp32); variance volumes are additionally checked voxel by voxel.

    cfg-1  3-view 256x128, cascade 32/16/8 planes, model=red PREDICT path (plane streaming, fp64 online head)
    cfg-2  3-view 768x384, 64 planes, casred stage 1  (1,3,32,64,96,192): the cluster recurrence carries 64 planes
    cfg-3  3-view 768x384, cascade 48/32/8, full casred (stages 2/3 at 192x384 and 384x768)
    cfg-4  5-view 1536x768, 192x384 grid, 5 views (the > 2-source-view sweep kernel), 8 of the 192 planes
    cfg-5  pin-hole homography sweep 3-view 768x384, 64 planes
"""
import pytest
import torch

import satmvs_b200
from oracle import stages, volume
from satmvs_b200 import synth

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
VOL_TOL = 2e-4        # N(0,1) features; same bound as tests/test_gpu_sweep.py
DEPTH_REL = 1e-3      # north_star


def cu(t):
    return t.to(DEV)


def maxdiff(a, b):
    return (a.detach().cpu().double() - b.detach().cpu().double()).abs().max().item()


def _red(cls, C, seed):
    m = getattr(satmvs_b200, cls)(C, 8)
    m.load_state_dict(synth.make_red_weights(C, seed=seed))
    return m.to(DEV).eval()


def test_cfg2_stage1_volume_and_depth_vs_oracle():
    B, V, C, D, H, W = 1, 3, 32, 64, 96, 192
    fe = synth.make_features(B, V, C, H, W, seed=0)
    rp = synth.make_rpc_stack(B, V, H, W)
    dv = synth.make_depth_planes(B, D, H, W)
    sd = synth.make_red_weights(C)
    with torch.no_grad():
        want_var = volume.variance_cost_volume(fe, rp, dv, "rpc")
        want = stages.stage_train_red(fe, rp, dv, sd, "rpc")
        got_var = satmvs_b200.build_cost_volume(cu(fe[0]), [cu(f) for f in fe[1:]], rp[:, 0], rp[:, 1:], cu(dv), "rpc")
        reg = _red("RED_Regularization", C, 13)
        got = satmvs_b200.stage_train_red([cu(f) for f in fe], rp, cu(dv), reg, "rpc")
    assert maxdiff(got_var, want_var) < VOL_TOL
    assert (got_var.cpu() == want_var).float().mean() > 0.99
    rel = maxdiff(got["depth"], want["depth"]) / want["depth"].abs().max().item()
    assert rel < DEPTH_REL, rel
    assert maxdiff(got["photometric_confidence"], want["photometric_confidence"]) < 1e-3


def _cascade_inputs(img_hw, chans=(32, 16, 8), scales=(4, 2, 1), V=3, seed=0):
    feats = [synth.make_features(1, V, c, img_hw[0] // sc, img_hw[1] // sc, seed=seed * 10 + i)
             for i, (c, sc) in enumerate(zip(chans, scales))]
    cams = [synth.make_rpc_stack(1, V, img_hw[0] // sc, img_hw[1] // sc) for sc in scales]
    return feats, cams, torch.tensor([[0.0, 1000.0]])


def test_cfg3_full_cascade_vs_oracle():
    img_hw, ndepths = (384, 768), (48, 32, 8)
    feats, cams, drange = _cascade_inputs(img_hw)
    sds = [synth.make_red_weights(c, seed=100 + i) for i, c in enumerate((32, 16, 8))]
    with torch.no_grad():
        want = stages.cascade(feats, cams, drange, sds, img_hw=img_hw, ndepths=ndepths, head="red_train")
        regs = [_red("RED_Regularization", c, 100 + i) for i, c in enumerate((32, 16, 8))]
        got = satmvs_b200.cascade([[cu(f) for f in fs] for fs in feats], cams, cu(drange), regs, img_hw=img_hw,
                                  ndepths=ndepths, head="red_train")
    for s in (1, 2, 3):
        w = want[f"stage{s}"]["depth"]
        rel = maxdiff(got[f"stage{s}"]["depth"], w) / w.abs().max().item()
        assert rel < DEPTH_REL, (s, rel)


def test_cfg1_predict_path_vs_oracle():
    """configs[0]: 256x128 image, 32/16/8 planes, the plane-streaming inference net (networks/casred.py:161-238)."""
    img_hw, ndepths = (128, 256), (32, 16, 8)
    feats, cams, drange = _cascade_inputs(img_hw, seed=1)
    sds = [synth.make_red_weights(c, seed=200 + i) for i, c in enumerate((32, 16, 8))]
    with torch.no_grad():
        want = stages.cascade(feats, cams, drange, sds, img_hw=img_hw, ndepths=ndepths, head="red_pred")
        regs = [_red("slice_RED_Regularization", c, 200 + i) for i, c in enumerate((32, 16, 8))]
        got = satmvs_b200.cascade([[cu(f) for f in fs] for fs in feats], cams, cu(drange), regs, img_hw=img_hw,
                                  ndepths=ndepths, head="red_pred")
    for s in (1, 2, 3):
        w = want[f"stage{s}"]["depth"]
        rel = maxdiff(got[f"stage{s}"]["depth"], w) / w.abs().max().item()
        assert rel < DEPTH_REL, (s, rel)
    assert maxdiff(got["photometric_confidence"], want["photometric_confidence"]) < 1e-3


def test_cfg4_five_view_volume_vs_oracle():
    """configs[3] grid (192x384, C 32, 5 views); 8 planes taken from across the 192-plane range."""
    B, V, C, D, H, W = 1, 5, 32, 192, 192, 384
    fe = synth.make_features(B, V, C, H, W, seed=4)
    rp = synth.make_rpc_stack(B, V, H, W)
    dv = synth.make_depth_planes(B, D, H, W)[:, ::24].contiguous()          # planes 0, 24, ..., 168
    with torch.no_grad():
        want = volume.variance_cost_volume(fe, rp, dv, "rpc")
        got = satmvs_b200.build_cost_volume(cu(fe[0]), [cu(f) for f in fe[1:]], rp[:, 0], rp[:, 1:], cu(dv), "rpc")
    assert got.shape == (B, C, 8, H, W)
    assert maxdiff(got, want) < VOL_TOL
    assert (got.cpu() == want).float().mean() > 0.99


def test_cfg5_pinhole_volume_vs_oracle():
    B, V, C, D, H, W = 1, 3, 32, 64, 96, 192
    fe = synth.make_features(B, V, C, H, W, seed=5)
    pp = synth.make_pinhole_stack(B, V, H, W)
    dv = synth.make_depth_planes(B, D, H, W, lo=90, hi=110, jitter=0.2)
    with torch.no_grad():
        want = volume.variance_cost_volume(fe, pp, dv, "pinhole")
        got = satmvs_b200.build_cost_volume(cu(fe[0]), [cu(f) for f in fe[1:]], pp[:, 0], pp[:, 1:], cu(dv), "pinhole")
    assert maxdiff(got, want) < VOL_TOL


def test_fresh_cuda_cameras_are_never_cached():
    """The reference uploads `cam_para` afresh every sample (tools/utils.py:85): a new CUDA tensor of the same shape can get
    the address of the previous one from the caching allocator.  Two calls with different cameras at (likely) the same
    address must give the two different volumes."""
    B, V, C, D, H, W = 1, 3, 4, 3, 16, 24
    fe = [cu(f) for f in synth.make_features(B, V, C, H, W, seed=3)]
    dv = cu(synth.make_depth_planes(B, D, H, W))
    rp_a = synth.make_rpc_stack(B, V, H, W)
    rp_b = rp_a.clone()
    rp_b[:, 1:, synth.SAMP_OFF] += 1.5
    outs, ptrs = [], []
    for rp in (rp_a, rp_b, rp_a):
        cams = rp.to(DEV)                       # fresh allocation, _version == 0
        ptrs.append(cams.data_ptr())
        w = satmvs_b200.rpc_warping(fe[1], cams[:, 1], cams[:, 0], dv, None)
        v = satmvs_b200.build_cost_volume(fe[0], fe[1:], cams[:, 0], [cams[:, 1], cams[:, 2]], dv, "rpc")
        outs.append((w.clone(), v.clone()))
        del cams, w, v
    want_a = satmvs_b200.build_cost_volume(fe[0], fe[1:], rp_a[:, 0], [rp_a[:, 1], rp_a[:, 2]], dv, "rpc")
    want_b = satmvs_b200.build_cost_volume(fe[0], fe[1:], rp_b[:, 0], [rp_b[:, 1], rp_b[:, 2]], dv, "rpc")
    assert not torch.equal(want_a, want_b)
    assert torch.equal(outs[0][1], want_a) and torch.equal(outs[1][1], want_b) and torch.equal(outs[2][1], want_a)
    assert not torch.equal(outs[0][0], outs[1][0]) and torch.equal(outs[0][0], outs[2][0])


def test_depth_regression_and_resize_match_torch():
    """`depth_regression` (modules/module.py:433-439) incl. its bilinear resize of the hypotheses."""
    g = torch.Generator().manual_seed(0)
    p = torch.softmax(torch.randn(2, 6, 12, 20, generator=g), 1)
    dv = torch.rand(2, 6, 6, 10, generator=g) * 100
    want = (p * torch.nn.functional.interpolate(dv, [12, 20], mode="bilinear", align_corners=False)).sum(1)
    got = satmvs_b200.depth_regression(cu(p), cu(dv))
    assert maxdiff(got, want) < 1e-4
    got2 = satmvs_b200.depth_regression(cu(p), cu(dv[:, :, 0, 0].contiguous()))
    assert maxdiff(got2, (p * dv[:, :, 0, 0].view(2, 6, 1, 1)).sum(1)) < 1e-4
