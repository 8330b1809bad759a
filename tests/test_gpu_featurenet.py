"""GPU parity of FeatureNet (csrc/featurenet.cu through the C ABI) against vectors from the unmodified reference
(tests/golden/featurenet.npz) and against the oracle at the BASELINE image size (3 views, 768 x 384).  fp32 FFMA
convolutions with another summation order than oneDNN: features are checked to 2e-5 of their range."""
import pytest
import torch

import satmvs_b200
from oracle import regnets
from satmvs_b200 import synth

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def maxdiff(a, b):
    return (a.detach().cpu().double() - b.detach().cpu().double()).abs().max().item()


def make_net(seed=17):
    m = satmvs_b200.FeatureNet(8, num_stage=3, stride=4, arch_mode="unet")
    m.load_state_dict(synth.make_featurenet_weights(8, seed=seed))
    return m.to(DEV).eval()


def test_featurenet_golden(golden):
    g = golden("featurenet")
    m = make_net()
    with torch.no_grad():
        outs = m.forward_views([g[f"img{v}"].to(DEV) for v in range(3)])
        single = m(g["img1"].to(DEV))
    for v in range(3):
        for k in ("stage1", "stage2", "stage3"):
            want = g[f"{k}_v{v}"]
            assert outs[v][k].shape == want.shape
            assert maxdiff(outs[v][k], want) < 2e-5 * max(1.0, want.abs().max().item()), (v, k)
    for k in ("stage1", "stage2", "stage3"):
        assert torch.equal(single[k], outs[1][k])          # one view alone == the same view inside a stack


def test_featurenet_baseline_size_vs_oracle():
    """3 views of 768 x 384 (BASELINE configs[1..2]): all views through every layer in one launch."""
    g = torch.Generator().manual_seed(3)
    imgs = [torch.rand(1, 3, 384, 768, generator=g) for _ in range(3)]
    sd = synth.make_featurenet_weights(8, seed=5)
    m = make_net(seed=5)
    with torch.no_grad():
        got = m.forward_views([i.to(DEV) for i in imgs])
        for v in (0, 2):
            want = regnets.featurenet(imgs[v], sd)
            for k in ("stage1", "stage2", "stage3"):
                assert maxdiff(got[v][k], want[k]) < 2e-5 * max(1.0, want[k].abs().max().item()), (v, k)


def test_featurenet_checkpoint_keys_and_modes():
    m = satmvs_b200.FeatureNet(8)
    want = synth.make_featurenet_weights(8)
    have = m.state_dict()
    assert set(have) == set(want) and all(have[k].shape == want[k].shape for k in want)
    assert m.out_channels == [32, 16, 8]
    m = m.to(DEV)
    out = m(torch.rand(1, 3, 8, 8, device=DEV))            # training mode: batch statistics, autograd nodes (test_gpu_train.py)
    assert out["stage3"].requires_grad and out["stage1"].shape == (1, 32, 2, 2)
    with pytest.raises(RuntimeError):
        m.eval()(torch.rand(1, 3, 8, 8, device=DEV))       # eval mode with gradients enabled: no backward, says so
    with pytest.raises(NotImplementedError):
        satmvs_b200.FeatureNet(8, arch_mode="fpn")
