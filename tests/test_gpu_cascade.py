"""GPU parity of hypothesis generation and of the whole cascade (three stages, features in ->
depth out) against the reference networks' committed outputs.  north_star bound: depth within
1e-3 relative L-inf of the reference."""
import pytest
import torch

import satmvs_b200
from oracle import hypotheses
from satmvs_b200 import synth

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def maxdiff(a, b):
    return (a.detach().cpu().double() - b.detach().cpu().double()).abs().max().item()


def test_hypotheses_golden(golden):
    g = golden("hypotheses")
    B, D, Himg, Wimg = g["first"].shape
    first = satmvs_b200.get_depth_range_samples(g["depth_range"].to(DEV), D, 10.0, DEV, torch.float32, [B, Himg, Wimg])
    assert maxdiff(first, g["first"]) == 0.0
    later = satmvs_b200.get_depth_range_samples(g["cur"].to(DEV), 6, 5.0, DEV, torch.float32, [B, Himg, Wimg])
    assert maxdiff(later, g["later"]) < 1e-4


@pytest.mark.parametrize("scale", [1, 2, 4])
def test_stage_hypotheses_vs_oracle(golden, scale):
    g = golden("hypotheses")
    B, Himg, Wimg = g["cur"].shape
    prev = torch.nn.functional.avg_pool2d(g["cur"].unsqueeze(1), 2).squeeze(1)       # a half-resolution previous stage
    want = hypotheses.stage_hypotheses(prev, g["depth_range"], 8, 5.0, (Himg, Wimg), scale)
    got = satmvs_b200.stage_depth_hypotheses(prev.to(DEV), None, 8, 5.0, (Himg, Wimg), scale)
    assert got.shape == want.shape
    assert maxdiff(got, want) < 2e-4
    want0 = hypotheses.stage_hypotheses(None, g["depth_range"], 8, 5.0, (Himg, Wimg), scale)
    got0 = satmvs_b200.stage_depth_hypotheses(None, g["depth_range"].to(DEV), 8, 5.0, (Himg, Wimg), scale)
    assert maxdiff(got0, want0) < 2e-4


@pytest.mark.parametrize("tag,head,cls,wfn", [
    ("red_train", "red_train", "RED_Regularization", synth.make_red_weights),
    ("red_pred", "red_pred", "slice_RED_Regularization", synth.make_red_weights),
    ("casmvs", "casmvs", "CostRegNet", synth.make_costregnet_weights)])
def test_cascade_golden(golden, tag, head, cls, wfn):
    g = golden(f"cascade_{tag}")
    feats = [[g[f"fea{s}_{v}"].to(DEV) for v in range(3)] for s in (1, 2, 3)]
    cams = [g[f"cams{s}"] for s in (1, 2, 3)]
    regs = []
    for s, c in enumerate((32, 16, 8)):
        m = getattr(satmvs_b200, cls)(c, 8)
        m.load_state_dict(wfn(c, seed=100 + s))
        regs.append(m.to(DEV).eval())
    with torch.no_grad():
        out = satmvs_b200.cascade(feats, cams, g["depth_range"].to(DEV), regs, img_hw=tuple(g["img_hw"].tolist()),
                                  ndepths=tuple(g["ndepths"].tolist()), head=head)
    for s in (1, 2, 3):
        want = g[f"depth{s}"]
        rel = maxdiff(out[f"stage{s}"]["depth"], want) / want.abs().max().item()
        assert rel < 1e-3, (s, rel)
    assert maxdiff(out["depth"], g["depth3"]) / g["depth3"].abs().max().item() < 1e-3
