"""RPC geometric-consistency filter (tools/rpc_filter.py) on the GPU against vectors produced by the unmodified reference
filter with OpenCV's remap and the reference's numpy RPC model (`oracle/make_golden_filter.py`)."""
import os

import numpy as np
import pytest
import torch

from satmvs_b200 import rpc_filter

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "rpc_filter.npz")


@pytest.fixture(scope="module")
def g():
    return np.load(G)


def test_remap_kernel_equals_cv2_bit_for_bit(g):
    """satmvs_remap_bilinear on the reference's own coordinates and on a cv2.remap stress fixture (random positions,
    exact ties of cvRound(x*32), border crossings, non-finite coordinates): integer / index work, so equality."""
    dev = "cuda:0"
    c = np.load(os.path.join(os.path.dirname(G), "remap_cv2.npz"))
    out = rpc_filter.remap_bilinear(torch.from_numpy(c["src"]).to(dev), torch.from_numpy(c["mapx"]).to(dev),
                                    torch.from_numpy(c["mapy"]).to(dev), -999.0).cpu().numpy()
    assert np.array_equal(out, c["out"])
    got = rpc_filter.remap_bilinear(torch.from_numpy(g["depths"][1]).to(dev), torch.from_numpy(g["x_src"]).float().to(dev),
                                    torch.from_numpy(g["y_src"]).float().to(dev), -999.0).cpu().numpy()
    assert np.array_equal(got, g["sampled"])


def test_reproject_with_depth_golden(g):
    sampled, xr, yr, xs, ys = rpc_filter.reproject_with_depth(g["depths"][0], g["rpcs"][0], g["depths"][1], g["rpcs"][1])
    assert np.abs(xs - g["x_src"]).max() < 1e-6 and np.abs(ys - g["y_src"]).max() < 1e-6          # fp64 projections (numpy twin of the model: other op order)
    # the gather follows OpenCV's 1/32-pixel weights: equal up to fp32 rounding of the four products, except where a
    # coordinate sits on a rounding boundary of the 1/32 grid (none allowed to differ by more than one weight step)
    # the gather is bit-exact given equal coordinates (test above); with coordinates that differ by ~1e-9 px the fp32 cast
    # can differ by one ulp and flip a 1/32-pixel cell: away from those ties the heights are equal bit for bit
    d = np.abs(sampled - g["sampled"])
    same_map = (xs.astype(np.float32) == g["x_src"].astype(np.float32)) & (ys.astype(np.float32) == g["y_src"].astype(np.float32))
    assert np.array_equal(sampled[same_map], g["sampled"][same_map])
    assert same_map.mean() > 0.95 and np.mean(d > 0) < 2e-3 and d.max() < 50.0, (same_map.mean(), np.mean(d > 0), d.max())
    ok = d < 1e-3
    assert np.abs(xr - g["x_reproj"])[ok].max() < 1e-3 and np.abs(yr - g["y_reproj"])[ok].max() < 1e-3


def test_filter_depth_golden(g):
    for args, mk, ak in (((1.0, 2.5, 1, g["prob"], 0.3), "mask", "avg"), ((0.5, 1.0, 2), "mask2", "avg2")):
        mask, avg = rpc_filter.filter_depth(g["depths"], g["rpcs"], *args)
        assert mask.shape == g[mk].shape and mask.dtype == bool
        assert np.mean(mask != g[mk]) < 2e-3
        if mk == "mask":      # single source pair decides the first view: masks equal wherever no coordinate sits on a tie
            m1, *_ = rpc_filter.check_geometric_consistency(g["depths"][0], g["rpcs"][0], g["depths"][1], g["rpcs"][1], 1.0, 2.5)
            sampled, xr, yr, xs, ys = rpc_filter.reproject_with_depth(g["depths"][0], g["rpcs"][0], g["depths"][1], g["rpcs"][1])
            same_map = (xs.astype(np.float32) == g["x_src"].astype(np.float32)) & (ys.astype(np.float32) == g["y_src"].astype(np.float32))
            H, W = g["depths"][0].shape
            xx, yy = np.meshgrid(np.arange(W), np.arange(H))
            ref_m = (np.sqrt((g["x_reproj"] - xx) ** 2 + (g["y_reproj"] - yy) ** 2) < 1.0) & (np.abs(g["sampled"] - g["depths"][0]) < 2.5)
            margin = np.abs(np.sqrt((g["x_reproj"] - xx) ** 2 + (g["y_reproj"] - yy) ** 2) - 1.0) > 1e-6
            sel = same_map & margin
            assert np.array_equal(m1[sel], ref_m[sel])
        same = mask == g[mk]
        rel = np.abs(avg - g[ak])[same] / np.abs(g[ak])[same].clip(1.0)
        assert np.quantile(rel, 0.998) < 1e-5


def test_remap_matches_integer_positions_and_border():
    src = torch.arange(12, dtype=torch.float32, device="cuda").reshape(3, 4)
    mx = torch.tensor([0.0, 3.0, 1.5, -1.0, 3.5, -5.0], device="cuda")
    my = torch.tensor([0.0, 2.0, 0.5, 0.0, 2.0, 7.0], device="cuda")
    out = rpc_filter.remap_bilinear(src, mx, my, -999.0).cpu().numpy()
    want = [0.0, 11.0, 3.5, -999.0, -494.0, -999.0]          # cv2.remap(..., INTER_LINEAR, BORDER_CONSTANT, -999) on the same inputs
    assert np.allclose(out, want)
