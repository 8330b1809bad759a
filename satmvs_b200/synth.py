"""Seeded synthetic inputs for the plane-sweep path.

Everything here is host-side numpy/torch: RPC vectors in the reference's 170-double
layout (`dataset/data_io.py:78-92`), pin-hole `K·E` matrices (`dataset/virdataset.py:67-70`),
feature maps, depth hypotheses and regulariser weights.  The generators use numpy's
PCG64 streams only, so the same seed gives the same bytes in this container, on the GPU
box and in `oracle/make_golden.py`.
"""
from __future__ import annotations

import numpy as np
import torch

# 170-vector layout (reference `tools/RPCCore.py:8-28`)
LINE_OFF, SAMP_OFF, LAT_OFF, LON_OFF, HEI_OFF = 0, 1, 2, 3, 4
LINE_SCALE, SAMP_SCALE, LAT_SCALE, LON_SCALE, HEI_SCALE = 5, 6, 7, 8, 9
LINE_NUM, LINE_DEN, SAMP_NUM, SAMP_DEN = 10, 30, 50, 70
LAT_NUM, LAT_DEN, LON_NUM, LON_DEN = 90, 110, 130, 150
RPC_LEN = 170


def make_rpc(view: int, height: int, width: int, *, seed: int | None = None,
             num_noise: float = 1e-4, den_noise: float = 1e-5,
             shift_geo: bool = False) -> np.ndarray:
    """One near-affine push-broom RPC (float64[170]) for a `height`×`width` raster.

    `view` = k scales the height parallax (0 for the reference view).  `shift_geo`
    perturbs LAT/LON offsets and LON_SCALE of the view so the ref→src affine
    composition is non-trivial (SURVEY.md §7 "hard parts" probe).
    """
    rng = np.random.default_rng(view if seed is None else seed)
    r = np.zeros(RPC_LEN, dtype=np.float64)
    r[LINE_OFF], r[SAMP_OFF] = height / 2.0, width / 2.0
    r[LAT_OFF], r[LON_OFF], r[HEI_OFF] = 30.0, -135.0, 500.0
    r[LINE_SCALE], r[SAMP_SCALE] = height / 2.0, width / 2.0
    r[LAT_SCALE], r[LON_SCALE], r[HEI_SCALE] = 0.01, 0.01, 500.0
    for num in (LINE_NUM, SAMP_NUM, LAT_NUM, LON_NUM):
        r[num:num + 20] = rng.normal(0.0, num_noise, 20)
        r[num] = 0.0
    for den in (LINE_DEN, SAMP_DEN, LAT_DEN, LON_DEN):
        r[den:den + 20] = rng.normal(0.0, den_noise, 20)
        r[den] = 1.0
    k = float(view)
    # monomial order [1, L, P, H, ...]: forward has L=lon, P=lat; inverse has L=line, P=samp
    r[LINE_NUM + 2], r[LINE_NUM + 3] = -1.0, 0.01 * k
    r[SAMP_NUM + 1], r[SAMP_NUM + 3] = 1.0, 0.05 * k
    r[LAT_NUM + 1], r[LAT_NUM + 3] = -1.0, 0.01 * k
    r[LON_NUM + 2], r[LON_NUM + 3] = 1.0, -0.05 * k
    if shift_geo and view > 0:
        r[LAT_OFF] += 0.003 * k
        r[LON_OFF] += 0.002 * k
        r[LON_SCALE] *= 1.1
    return r


def rescale_rpc(rpc: np.ndarray, factor: float) -> np.ndarray:
    """Per-stage RPC rescale: divide LINE/SAMP offset and scale (entries 0,1,5,6)
    exactly like `dataset/satmvsdataset.py:83-93`."""
    out = np.array(rpc, dtype=np.float64, copy=True)
    out[..., [LINE_OFF, SAMP_OFF, LINE_SCALE, SAMP_SCALE]] /= factor
    return out


def make_rpc_stack(batch: int, views: int, height: int, width: int, **kw) -> torch.Tensor:
    """[B, V, 170] float64, the `cam_para[stage]` format (`networks/casred.py:13`)."""
    one = np.stack([make_rpc(v, height, width, **kw) for v in range(views)])
    return torch.from_numpy(np.broadcast_to(one, (batch,) + one.shape).copy())


def make_pinhole_stack(batch: int, views: int, height: int, width: int,
                       depth_mid: float = 100.0) -> torch.Tensor:
    """[B, V, 4, 4] float64 `K·E` projections (`dataset/virdataset.py:67-70`)."""
    f = 2.0 * width
    K = np.array([[f, 0, width / 2.0], [0, f, height / 2.0], [0, 0, 1.0]])
    mats = []
    for k in range(views):
        rng = np.random.default_rng(1000 + k)
        ang = np.deg2rad(rng.uniform(-2.0, 2.0, 3)) * (1.0 if k else 0.0)
        cx, cy, cz = np.cos(ang)
        sx, sy, sz = np.sin(ang)
        Rx = np.array([[1, 0, 0], [0, cx, -sx], [0, sx, cx]])
        Ry = np.array([[cy, 0, sy], [0, 1, 0], [-sy, 0, cy]])
        Rz = np.array([[cz, -sz, 0], [sz, cz, 0], [0, 0, 1]])
        E = np.eye(4)
        E[:3, :3] = Rz @ Ry @ Rx
        E[:3, 3] = [k * 0.1 * depth_mid, 0.0, 0.0]
        P = np.eye(4)
        P[:3, :4] = K @ E[:3, :4]
        mats.append(P)
    one = np.stack(mats)
    return torch.from_numpy(np.broadcast_to(one, (batch,) + one.shape).copy())


def make_features(batch: int, views: int, channels: int, height: int, width: int,
                  seed: int = 0) -> list[torch.Tensor]:
    """V feature maps [B, C, H, W] float32, standard normal."""
    rng = np.random.default_rng(seed)
    return [torch.from_numpy(rng.standard_normal((batch, channels, height, width), dtype=np.float32))
            for _ in range(views)]


def make_depth_planes(batch: int, planes: int, height: int, width: int, *,
                      lo: float = 0.0, hi: float = 1000.0, per_pixel: bool = True,
                      jitter: float = 1.0, seed: int = 7) -> torch.Tensor:
    """Depth hypotheses: `linspace(lo, hi, D)` either as [B, D] or broadcast to
    [B, D, H, W] with N(0, jitter) noise (exercises the per-pixel path)."""
    base = torch.linspace(lo, hi, planes, dtype=torch.float32)
    if not per_pixel:
        return base.view(1, planes).repeat(batch, 1).contiguous()
    rng = np.random.default_rng(seed)
    noise = torch.from_numpy(rng.standard_normal((batch, planes, height, width), dtype=np.float32))
    return (base.view(1, planes, 1, 1) + jitter * noise).contiguous()


def _uniform(rng: np.random.Generator, shape, bound: float) -> torch.Tensor:
    return torch.from_numpy(rng.uniform(-bound, bound, size=shape).astype(np.float32))


def make_costregnet_weights(in_channels: int, base: int = 8, seed: int = 11) -> dict[str, torch.Tensor]:
    """State-dict for `CostRegNet(in_channels, base)` (`modules/module.py:546-566`) with the
    reference's parameter names.  BN running stats are non-trivial so eval-mode folding is tested."""
    rng = np.random.default_rng(seed)
    sd: dict[str, torch.Tensor] = {}

    def block(name, cin, cout, transposed=False):
        shape = (cin, cout, 3, 3, 3) if transposed else (cout, cin, 3, 3, 3)
        sd[f"{name}.conv.weight"] = _uniform(rng, shape, (1.0 / (cin * 27)) ** 0.5 * 1.7)
        sd[f"{name}.bn.weight"] = torch.from_numpy(rng.uniform(0.5, 1.5, cout).astype(np.float32))
        sd[f"{name}.bn.bias"] = torch.from_numpy(rng.normal(0, 0.1, cout).astype(np.float32))
        sd[f"{name}.bn.running_mean"] = torch.from_numpy(rng.normal(0, 0.1, cout).astype(np.float32))
        sd[f"{name}.bn.running_var"] = torch.from_numpy(rng.uniform(0.5, 1.5, cout).astype(np.float32))
        sd[f"{name}.bn.num_batches_tracked"] = torch.tensor(1, dtype=torch.long)

    b = base
    block("conv0", in_channels, b)
    block("conv1", b, 2 * b)
    block("conv2", 2 * b, 2 * b)
    block("conv3", 2 * b, 4 * b)
    block("conv4", 4 * b, 4 * b)
    block("conv5", 4 * b, 8 * b)
    block("conv6", 8 * b, 8 * b)
    block("conv7", 8 * b, 4 * b, transposed=True)
    block("conv9", 4 * b, 2 * b, transposed=True)
    block("conv11", 2 * b, b, transposed=True)
    sd["prob.weight"] = _uniform(rng, (1, b, 3, 3, 3), (1.0 / (b * 27)) ** 0.5 * 1.7)
    return sd


def make_featurenet_weights(base: int = 8, seed: int = 17) -> dict[str, torch.Tensor]:
    """State-dict for `FeatureNet(base, num_stage=3, arch_mode="unet")` (`modules/module.py:442-480`) with the reference's
    parameter names; BN running statistics are non-trivial so that the eval-mode folding is exercised."""
    rng = np.random.default_rng(seed)
    sd: dict[str, torch.Tensor] = {}

    def block(name, cin, cout, k, transposed=False):
        shape = (cin, cout, k, k) if transposed else (cout, cin, k, k)
        sd[f"{name}.conv.weight"] = _uniform(rng, shape, (1.0 / (cin * k * k)) ** 0.5 * 1.7)
        sd[f"{name}.bn.weight"] = torch.from_numpy(rng.uniform(0.5, 1.5, cout).astype(np.float32))
        sd[f"{name}.bn.bias"] = torch.from_numpy(rng.normal(0, 0.1, cout).astype(np.float32))
        sd[f"{name}.bn.running_mean"] = torch.from_numpy(rng.normal(0, 0.1, cout).astype(np.float32))
        sd[f"{name}.bn.running_var"] = torch.from_numpy(rng.uniform(0.5, 1.5, cout).astype(np.float32))
        sd[f"{name}.bn.num_batches_tracked"] = torch.tensor(1, dtype=torch.long)

    b = base
    block("conv0.0", 3, b, 3); block("conv0.1", b, b, 3)
    block("conv1.0", b, 2 * b, 5); block("conv1.1", 2 * b, 2 * b, 3); block("conv1.2", 2 * b, 2 * b, 3)
    block("conv2.0", 2 * b, 4 * b, 5); block("conv2.1", 4 * b, 4 * b, 3); block("conv2.2", 4 * b, 4 * b, 3)
    sd["out1.weight"] = _uniform(rng, (4 * b, 4 * b, 1, 1), (1.0 / (4 * b)) ** 0.5 * 1.7)
    block("deconv1.deconv", 4 * b, 2 * b, 3, transposed=True); block("deconv1.conv", 4 * b, 2 * b, 3)
    block("deconv2.deconv", 2 * b, b, 3, transposed=True); block("deconv2.conv", 2 * b, b, 3)
    sd["out2.weight"] = _uniform(rng, (2 * b, 2 * b, 1, 1), (1.0 / (2 * b)) ** 0.5 * 1.7)
    sd["out3.weight"] = _uniform(rng, (b, b, 1, 1), (1.0 / b) ** 0.5 * 1.7)
    return sd


def make_red_weights(in_channels: int, base: int = 8, seed: int = 13) -> dict[str, torch.Tensor]:
    """State-dict for `RED_Regularization` / `slice_RED_Regularization(in_channels, base)`
    (`modules/module.py:595-610`, `:653-668`; identical keys)."""
    rng = np.random.default_rng(seed)
    sd: dict[str, torch.Tensor] = {}

    def gru(name, cin, cout):
        k = cin + cout
        bound = (1.0 / (k * 9)) ** 0.5
        sd[f"{name}.gate_conv.weight"] = _uniform(rng, (2 * cout, k, 3, 3), bound * 1.7)
        sd[f"{name}.gate_conv.bias"] = _uniform(rng, (2 * cout,), bound)
        sd[f"{name}.output_conv.weight"] = _uniform(rng, (cout, k, 3, 3), bound * 1.7)
        sd[f"{name}.output_conv.bias"] = _uniform(rng, (cout,), bound)
        for g in ("reset_gate_norm", "update_gate_norm", "output_norm"):
            sd[f"{name}.{g}.weight"] = torch.from_numpy(rng.uniform(0.5, 1.5, cout).astype(np.float32))
            sd[f"{name}.{g}.bias"] = torch.from_numpy(rng.normal(0, 0.1, cout).astype(np.float32))

    b = base
    gru("conv_gru1", in_channels, b)
    gru("conv_gru2", 2 * b, 2 * b)
    gru("conv_gru3", 4 * b, 4 * b)
    gru("conv_gru4", 8 * b, 8 * b)
    for name, cin, cout in (("conv1", in_channels, 2 * b), ("conv2", 2 * b, 4 * b), ("conv3", 4 * b, 8 * b)):
        sd[f"{name}.conv.weight"] = _uniform(rng, (cout, cin, 3, 3), (1.0 / (cin * 9)) ** 0.5 * 1.7)
    for name, cin, cout in (("upconv3", 8 * b, 4 * b), ("upconv2", 4 * b, 2 * b), ("upconv1", 2 * b, b)):
        sd[f"{name}.conv.weight"] = _uniform(rng, (cin, cout, 3, 3), (1.0 / (cin * 9)) ** 0.5 * 1.7)
    sd["upconv2d.weight"] = _uniform(rng, (b, 1, 3, 3), (1.0 / (b * 9)) ** 0.5 * 1.7)
    sd["upconv2d.bias"] = _uniform(rng, (1,), 0.1)
    return sd
