"""Golden vectors for satmvs_b200/rpc_filter.py from the UNMODIFIED reference filter (tools/rpc_filter.py:9-112).
The reference imports its CuPy RPC model (tools/rpc_tensor.py; CuPy is absent here and from the reference's own
environment.yml), so the module source is executed with the reference's numpy twin of that class
(tools/RPCCore.py:RPCModelParameter, same maths: SURVEY.md §8c) and with the numpy-1 aliases it uses (np.float, np.bool).
Test infrastructure only; run in the build container:  python -m oracle.make_golden_filter"""
import os
import sys

import numpy as np

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def reference_filter():
    sys.path.insert(0, "/root/reference")
    if not hasattr(np, "float"):
        np.float = float
    if not hasattr(np, "bool"):
        np.bool = bool
    from tools.RPCCore import RPCModelParameter
    src = open("/root/reference/tools/rpc_filter.py").read().replace("from tools.rpc_tensor import RPCModelParameter", "")
    ns = {"RPCModelParameter": RPCModelParameter}
    exec(compile(src, "rpc_filter.py", "exec"), ns)
    return ns


def main():
    from satmvs_b200 import synth
    ns = reference_filter()
    V, H, W = 3, 48, 80
    rp = synth.make_rpc_stack(1, V, H, W)[0].numpy()                      # [V,170] float64, near-affine pushbroom set
    rng = np.random.default_rng(3)
    yy, xx = np.meshgrid(np.arange(H), np.arange(W), indexing="ij")
    base = 500.0 + 40.0 * np.sin(xx / 9.0) + 25.0 * np.cos(yy / 7.0)      # a smooth height field seen by every view
    depths = np.stack([base + rng.normal(0, s, base.shape) for s in (0.0, 0.4, 1.5)]).astype(np.float32)
    prob = rng.uniform(0, 1, (H, W)).astype(np.float32)
    sampled, xr, yr, xs, ys = ns["reproject_with_depth"](depths[0], rp[0], depths[1], rp[1])
    mask, avg = ns["filter_depth"](depths, rp, 1.0, 2.5, 1, prob, 0.3)
    mask2, avg2 = ns["filter_depth"](depths, rp, 0.5, 1.0, 2)
    np.savez_compressed(os.path.join(OUT, "rpc_filter.npz"), rpcs=rp, depths=depths, prob=prob, sampled=sampled, x_reproj=xr,
                        y_reproj=yr, x_src=xs, y_src=ys, mask=mask, avg=avg, mask2=mask2, avg2=avg2)
    # cv2.remap alone on random coordinates (inside, on the 1/32 rounding boundaries, across the border): pins the
    # numpy restatement oracle/remap.py and the CUDA kernel satmvs_remap_bilinear bit for bit
    import cv2
    src = depths[2]
    mx = rng.uniform(-3, W + 3, (64, 96)).astype(np.float32)
    my = rng.uniform(-3, H + 3, (64, 96)).astype(np.float32)
    mx[0, :32] = (np.arange(32) + 0.5) / 32 + 7            # exact ties of cvRound(x * 32): round-half-even
    my[0, :32] = (np.arange(32) * 2 + 1) / 64 + 5
    mx[1, :4] = [np.nan, np.inf, -np.inf, 1e30]
    out = cv2.remap(src, mx, my, interpolation=cv2.INTER_LINEAR, borderMode=cv2.BORDER_CONSTANT, borderValue=-999)
    np.savez_compressed(os.path.join(OUT, "remap_cv2.npz"), src=src, mapx=mx, mapy=my, out=out)
    print("mask fraction", mask.mean(), mask2.mean(), "sampled range", sampled.min(), sampled.max())


if __name__ == "__main__":
    main()
