"""GPU parity of the RED regulariser (conv engine + GRU kernels through the C ABI) against the
committed reference vectors and the CPU oracle.  fp32 convolutions with a different summation
order than oneDNN's: logits are checked to 2e-4 of their range, depth to north_star's 1e-3
relative L-inf (measured values are ~1e-6, asserted at 1e-4)."""
import pytest
import torch

import satmvs_b200
from oracle import regnets, stages
from satmvs_b200 import synth

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.fixture(autouse=True)
def _inference_mode(request):
    """Inference tests run under no_grad; the slice form says so when gradients are on (test_slice_form_requires_no_grad)."""
    if "requires_no_grad" in request.node.name:
        yield
        return
    with torch.no_grad():
        yield


def test_slice_form_requires_no_grad():
    """The per-slice (predict) form has no backward and says so; the whole-volume form trains (tests/test_gpu_train.py)."""
    m = satmvs_b200.slice_RED_Regularization(8, 8).to(DEV)
    x = torch.rand(1, 8, 8, 8, device=DEV)
    st = [torch.zeros(1, c, 8 >> l, 8 >> l, device=DEV) for l, c in enumerate((8, 16, 32, 64))]
    with pytest.raises(RuntimeError, match="inference-only"):
        m(x, *st)
    with torch.no_grad():
        assert m(x, *st)[0].shape == (1, 1, 8, 8)


def maxdiff(a, b):
    return (a.detach().cpu().double() - b.detach().cpu().double()).abs().max().item()


def make_reg(cls, C, seed=13):
    m = cls(C, 8)
    m.load_state_dict(synth.make_red_weights(C, seed=seed))
    return m.to(DEV).eval()


def test_slice_golden(golden):
    g = golden("red_slice")
    m = make_reg(satmvs_b200.slice_RED_Regularization, 8)
    out = m(g["cost"].to(DEV), *[g[k].to(DEV) for k in ("s1", "s2", "s3", "s4")])
    errs = {key: maxdiff(got, g[key]) / max(1.0, g[key].abs().max().item())
            for got, key in zip(out, ("reg", "n1", "n2", "n3", "n4"))}
    assert all(got.shape == g[key].shape for got, key in zip(out, ("reg", "n1", "n2", "n3", "n4")))
    assert max(errs.values()) < 2e-4, errs


@pytest.mark.parametrize("geo", ["rpc", "pinhole"])
def test_volume_golden(golden, geo):
    g = golden(f"stage_train_{geo}")
    m = make_reg(satmvs_b200.RED_Regularization, 8)
    logits = m(g["var"].to(DEV))
    assert logits.shape == g["logits"].shape
    assert maxdiff(logits, g["logits"]) < 2e-4 * max(1.0, g["logits"].abs().max().item())
    fe = [g[f"fea{v}"].to(DEV) for v in range(3)]
    out = satmvs_b200.stage_train_red(fe, g["cams"], g["depth_values"].to(DEV), m, geo)
    rel = maxdiff(out["depth"], g["depth"]) / g["depth"].abs().max().item()
    assert rel < 1e-4, rel
    assert maxdiff(out["photometric_confidence"], g["conf"]) < 1e-4


@pytest.mark.parametrize("C,D,H,W", [(32, 12, 32, 64), (16, 3, 16, 24), (8, 1, 8, 8),
                                     (8, 2, 8, 512),      # wide planes: 8 window positions per thread in the tcgen05 conv
                                     (16, 5, 24, 40)])    # ragged tiles, odd plane count
def test_volume_vs_oracle(C, D, H, W):
    sd = synth.make_red_weights(C, seed=5)
    m = satmvs_b200.RED_Regularization(C, 8)
    m.load_state_dict(sd)
    m = m.to(DEV)
    x = synth.make_features(1, 1, C * D, H, W, seed=9)[0].view(1, C, D, H, W).abs()
    want = regnets.red_regularization(x, sd)
    got = m(x.to(DEV))
    assert maxdiff(got, want) < 2e-4 * max(1.0, want.abs().max().item())


@pytest.mark.parametrize("C,D,H,W", [(32, 12, 32, 64), (32, 4, 96, 192), (8, 3, 8, 32), (16, 2, 40, 96), (8, 5, 64, 96)])
def test_recurrence_paths_agree(C, D, H, W, monkeypatch):
    """Three implementations of the depth recurrence compute the same arithmetic graph: the tensor-core cluster kernel
    (red_tc.cuh: 3xTF32 split, K-split partial sums), the FFMA cluster kernel (red_cluster.cuh) and the per-plane kernel
    chain; they differ in summation order only."""
    from satmvs_b200 import _lib
    sd = synth.make_red_weights(C, seed=7)
    m = satmvs_b200.RED_Regularization(C, 8)
    m.load_state_dict(sd)
    m = m.to(DEV)
    x = synth.make_features(1, 1, C * D, H, W, seed=11)[0].view(1, C, D, H, W).abs().to(DEV)
    outs, paths = {}, {}
    for name, env in (("tc", {}), ("tc_seq", {"SATMVS_RED_NO_OVERLAP": "1"}), ("cluster", {"SATMVS_RED_NO_TC": "1"}),
                      ("chain", {"SATMVS_RED_NO_CLUSTER": "1"})):
        for k in ("SATMVS_RED_NO_TC", "SATMVS_RED_NO_CLUSTER", "SATMVS_RED_NO_OVERLAP"):
            monkeypatch.delenv(k, raising=False)
        for k, v in env.items():
            monkeypatch.setenv(k, v)
        outs[name] = m(x).clone()
        torch.cuda.synchronize()
        paths[name] = _lib.lib().satmvs_red_last_path()
    for k in ("SATMVS_RED_NO_TC", "SATMVS_RED_NO_CLUSTER", "SATMVS_RED_NO_OVERLAP"):
        monkeypatch.delenv(k, raising=False)
    assert paths["chain"] == 0
    assert torch.equal(outs["tc"], outs["tc_seq"])          # same kernels, same bits: only the launch order differs
    if W % 32 == 0:                       # every level keeps rows of a multiple of 4 pixels: the tensor-core kernel takes it
        assert paths["tc"] in (2, 3), paths   # 3 = overlapped with the batched convs on a side stream
    want = regnets.red_regularization(x.cpu(), sd)
    scale = max(1.0, want.abs().max().item())
    errs = {k: maxdiff(v, want) / scale for k, v in outs.items()}
    assert maxdiff(outs["tc"], outs["chain"]) < 5e-5 * scale, errs
    assert maxdiff(outs["cluster"], outs["chain"]) < 5e-5 * scale, errs
    assert max(errs.values()) < 2e-4, errs


def test_tensor_core_recurrence_carries_states():
    """slice form (D = 1, explicit states in and out) through the tensor-core kernel == oracle."""
    from satmvs_b200 import _lib
    C, H, W = 16, 32, 64
    sd = synth.make_red_weights(C, seed=3)
    m = make_reg(satmvs_b200.slice_RED_Regularization, C, seed=3)
    g = torch.Generator().manual_seed(5)
    cost = torch.rand(1, C, H, W, generator=g)
    states = [torch.randn(1, c, H >> l, W >> l, generator=g) * 0.5 for l, c in enumerate((8, 16, 32, 64))]
    want = regnets.red_slice(cost, *states, sd)
    got = m(cost.to(DEV), *[s.to(DEV) for s in states])
    torch.cuda.synchronize()
    assert _lib.lib().satmvs_red_last_path() in (2, 3)
    for a, b in zip(got, want):
        assert maxdiff(a, b) < 2e-4 * max(1.0, b.abs().max().item())


def test_pred_stage_equals_train_stage():
    """Plane-streaming inference form == whole-volume form (the reference's two nets agree to
    1.7e-5 relative)."""
    B, V, C, D, H, W = 1, 3, 8, 6, 16, 24
    fe = [f.to(DEV) for f in synth.make_features(B, V, C, H, W, seed=2)]
    rp = synth.make_rpc_stack(B, V, H, W)
    dv = synth.make_depth_planes(B, D, H, W).to(DEV)
    sd = synth.make_red_weights(C)
    a = make_reg(satmvs_b200.RED_Regularization, C)
    b = make_reg(satmvs_b200.slice_RED_Regularization, C)
    ta = satmvs_b200.stage_train_red(fe, rp, dv, a, "rpc")
    tb = satmvs_b200.stage_pred_red(fe, rp, dv, b, "rpc")
    scale = ta["depth"].abs().max().item()
    assert maxdiff(ta["depth"], tb["depth"]) < 1e-4 * scale
    want = stages.stage_pred_red([f.cpu() for f in fe], rp, dv.cpu(), sd, "rpc")
    assert maxdiff(tb["depth"], want["depth"]) < 1e-4 * scale
    assert maxdiff(tb["photometric_confidence"], want["photometric_confidence"]) < 1e-4


@pytest.mark.parametrize("chunk", [1, 4, 16])
def test_pred_stage_chunked_equals_per_plane_loop(chunk):
    """The chunked plane-streaming stage (one sweep + one recurrence call + one head update per chunk) == the reference's
    literal per-plane loop (chunk = 1) == the oracle."""
    B, V, C, D, H, W = 1, 3, 16, 10, 32, 64
    fe = [f.to(DEV) for f in synth.make_features(B, V, C, H, W, seed=4)]
    rp = synth.make_rpc_stack(B, V, H, W)
    dv = synth.make_depth_planes(B, D, H, W).to(DEV)
    sd = synth.make_red_weights(C)
    m = make_reg(satmvs_b200.slice_RED_Regularization, C)
    got = satmvs_b200.stage_pred_red(fe, rp, dv, m, "rpc", chunk=chunk)
    want = stages.stage_pred_red([f.cpu() for f in fe], rp, dv.cpu(), sd, "rpc")
    scale = want["depth"].abs().max().item()
    assert maxdiff(got["depth"], want["depth"]) < 1e-4 * scale
    assert maxdiff(got["photometric_confidence"], want["photometric_confidence"]) < 1e-4


def test_packed_weight_cache_follows_the_parameters():
    """The packed tensor-core weights are cached in the module; an in-place parameter update (optimiser step, checkpoint
    load) must invalidate them."""
    C, D, H, W = 16, 3, 32, 64
    sd = synth.make_red_weights(C, seed=9)
    m = make_reg(satmvs_b200.RED_Regularization, C, seed=9)
    x = synth.make_features(1, 1, C * D, H, W, seed=2)[0].view(1, C, D, H, W).abs()
    a = m(x.to(DEV)).clone()
    assert torch.equal(a, m(x.to(DEV)))                       # second call: cached packs, same bits
    with torch.no_grad():
        for name in ("conv_gru1.gate_conv.weight", "conv_gru3.output_conv.weight", "conv2.conv.weight"):
            dict(m.named_parameters())[name].mul_(0.5)
            sd[name] = sd[name] * 0.5
    b = m(x.to(DEV))
    want = regnets.red_regularization(x, sd)
    assert maxdiff(b, want) < 2e-4 * max(1.0, want.abs().max().item())
    assert maxdiff(a, b) > 1e-3


def test_reference_checkpoint_keys():
    """state_dict keys/shapes are the reference's (`train.py:216-219` checkpoints must load)."""
    m = satmvs_b200.RED_Regularization(32, 8)
    want = synth.make_red_weights(32)
    have = m.state_dict()
    assert set(have) == set(want)
    assert all(have[k].shape == want[k].shape for k in want)
