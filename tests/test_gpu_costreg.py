"""GPU parity of CostRegNet (3-D conv engine through the C ABI) against the committed reference
vectors and the CPU oracle; fp32 throughout, tolerance 2e-4 of the logit range."""
import pytest
import torch

import satmvs_b200
from oracle import regnets, stages
from satmvs_b200 import synth

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def maxdiff(a, b):
    return (a.detach().cpu().double() - b.detach().cpu().double()).abs().max().item()


def make(C, seed=11):
    m = satmvs_b200.CostRegNet(C, 8)
    m.load_state_dict(synth.make_costregnet_weights(C, seed=seed))
    return m.to(DEV).eval()


def test_golden(golden):
    g = golden("costregnet")
    y = make(8)(g["x"].to(DEV))
    assert y.shape == g["y"].shape
    assert maxdiff(y, g["y"]) < 2e-4 * max(1.0, g["y"].abs().max().item())


@pytest.mark.parametrize("C,D,H,W", [(32, 16, 32, 64), (16, 8, 24, 40), (8, 8, 8, 8)])
def test_vs_oracle(C, D, H, W):
    sd = synth.make_costregnet_weights(C, seed=3)
    m = satmvs_b200.CostRegNet(C, 8)
    m.load_state_dict(sd)
    m = m.to(DEV).eval()
    x = synth.make_features(1, 1, C * D, H, W, seed=4)[0].view(1, C, D, H, W).abs()
    want = regnets.costregnet(x, sd)
    got = m(x.to(DEV))
    assert maxdiff(got, want) < 2e-4 * max(1.0, want.abs().max().item())


def test_training_mode_runs_on_batch_statistics():
    """train() is the mode train.py:268 uses: batch-statistics BatchNorm and a backward (tests/test_gpu_train.py has the parity)."""
    m = make(8).train()
    x = torch.rand(2, 8, 8, 16, 24, device=DEV)          # coarsest level: 1 x 2 x 3 voxels per sample (a ragged BatchNorm row)
    y = m(x)
    assert y.shape == (2, 1, 8, 16, 24) and y.requires_grad
    want = regnets.costregnet(x.cpu(), {k: v.detach().cpu() for k, v in m.state_dict().items()}, training=True)
    assert maxdiff(y, want) < 1e-3 * max(1.0, want.abs().max().item())


def test_casmvs_stage_vs_oracle():
    B, V, C, D, H, W = 1, 3, 8, 8, 16, 24
    fe = synth.make_features(B, V, C, H, W, seed=6)
    rp = synth.make_rpc_stack(B, V, H, W)
    dv = synth.make_depth_planes(B, D, H, W)
    sd = synth.make_costregnet_weights(C)
    want = stages.stage_casmvs(fe, rp, dv, sd, "rpc")
    got = satmvs_b200.stage_casmvs([f.to(DEV) for f in fe], rp, dv.to(DEV), make(C), "rpc")
    rel = maxdiff(got["depth"], want["depth"]) / want["depth"].abs().max().item()
    assert rel < 1e-4, rel
    assert ((got["photometric_confidence"].cpu() - want["photometric_confidence"]).abs() < 1e-4).float().mean() > 0.99


def test_reference_checkpoint_keys():
    have = satmvs_b200.CostRegNet(32, 8).state_dict()
    want = synth.make_costregnet_weights(32)
    assert set(have) == set(want) and all(have[k].shape == want[k].shape for k in want)
