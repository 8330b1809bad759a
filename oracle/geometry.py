"""oracle.geometry — CPU restatement of the RPC / pin-hole plane-sweep geometry and the
bilinear sampler.  TEST INFRASTRUCTURE (see oracle/__init__.py).

fp64 geometry, fp32 sampling, same operation order as the reference so that the fp64
intermediates are bit-identical to `modules/warping.py` run on the same machine.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

# 170-vector slices, reference `dataset/data_io.py:78-92` / `tools/RPCCore.py:8-28`
_OFF = dict(line=0, samp=1, lat=2, lon=3, hei=4)
_SCL = dict(line=5, samp=6, lat=7, lon=8, hei=9)
_POLY = dict(line_num=10, line_den=30, samp_num=50, samp_den=70,
             lat_num=90, lat_den=110, lon_num=130, lon_den=150)


def _normalise(v: torch.Tensor, rpc: torch.Tensor, key: str) -> torch.Tensor:
    # (v - OFF) / SCALE, two separate fp64 ops like `warping.py:229-236`, `:270-278`
    return (v - rpc[:, _OFF[key]].view(-1, 1)) / rpc[:, _SCL[key]].view(-1, 1)


def _denormalise(v: torch.Tensor, rpc: torch.Tensor, key: str) -> torch.Tensor:
    # v * SCALE + OFF, `warping.py:246-250`, `:293-297`
    return v * rpc[:, _SCL[key]].view(-1, 1) + rpc[:, _OFF[key]].view(-1, 1)


def plh_monomials(P: torch.Tensor, L: torch.Tensor, H: torch.Tensor) -> torch.Tensor:
    """The 20 cubic monomials in RPC00B order, [B, N, 20] fp64.

    Follows `RPC_PLH_COEF` (`modules/warping.py:183-207`): the same products in the same
    association, e.g. term 10 = P*(L*H), term 14 = L*(L*P), so every entry is bit-identical.
    Argument order is (P, L, H); column 1 is L and column 2 is P.
    """
    LP, LH, PH = L * P, L * H, P * H
    LL, PP, HH = L * L, P * P, H * H
    cols = [torch.ones_like(P), L, P, H, LP, LH, PH, LL, PP, HH,
            P * LH, L * LL, L * PP, L * HH, L * LP, P * PP, P * HH, L * LH, P * PH, H * HH]
    return torch.stack(cols, dim=-1)


def _ratio(mono: torch.Tensor, rpc: torch.Tensor, num: str, den: str) -> torch.Tensor:
    # sum(coef * NUM) / sum(coef * DEN) over the 20 terms, `warping.py:241-244`, `:285-288`
    n0, d0 = _POLY[num], _POLY[den]
    top = torch.sum(mono * rpc[:, n0:n0 + 20].view(-1, 1, 20), dim=-1)
    bot = torch.sum(mono * rpc[:, d0:d0 + 20].view(-1, 1, 20), dim=-1)
    return top / bot


def rpc_localise(samp: torch.Tensor, line: torch.Tensor, hei: torch.Tensor,
                 rpc: torch.Tensor) -> tuple[torch.Tensor, torch.Tensor]:
    """Image + height -> (lat, lon): `RPC_Photo2Obj`, `modules/warping.py:255-307`.
    samp/line/hei [B, N] fp64, rpc [B, 170] fp64."""
    s = _normalise(samp, rpc, "samp")
    l = _normalise(line, rpc, "line")
    h = _normalise(hei, rpc, "hei")
    mono = plh_monomials(s, l, h)          # P = samp, L = line (`warping.py:280`)
    lat = _denormalise(_ratio(mono, rpc, "lat_num", "lat_den"), rpc, "lat")
    lon = _denormalise(_ratio(mono, rpc, "lon_num", "lon_den"), rpc, "lon")
    return lat, lon


def rpc_project(lat: torch.Tensor, lon: torch.Tensor, hei: torch.Tensor,
                rpc: torch.Tensor) -> tuple[torch.Tensor, torch.Tensor]:
    """(lat, lon, height) -> (samp, line): `RPC_Obj2Photo`, `modules/warping.py:218-252`."""
    p = _normalise(lat, rpc, "lat")
    l = _normalise(lon, rpc, "lon")
    h = _normalise(hei, rpc, "hei")
    mono = plh_monomials(p, l, h)          # P = lat, L = lon (`warping.py:238`)
    samp = _denormalise(_ratio(mono, rpc, "samp_num", "samp_den"), rpc, "samp")
    line = _denormalise(_ratio(mono, rpc, "line_num", "line_den"), rpc, "line")
    return samp, line


# ---------------------------------------------------------------------------------------
# quaternary-cubic ("QC") tensor form: `dataset/data_io.py:95-120` == `tools/rpc_tensor.py:24-50`
# ---------------------------------------------------------------------------------------
# monomial index -> sorted index triple over x = (1, L, P, H)
_QC_TRIPLES = [(0, 0, 0), (0, 0, 1), (0, 0, 2), (0, 0, 3), (0, 1, 2), (0, 1, 3), (0, 2, 3),
               (0, 1, 1), (0, 2, 2), (0, 3, 3), (1, 2, 3), (1, 1, 1), (1, 2, 2), (1, 3, 3),
               (1, 1, 2), (2, 2, 2), (2, 3, 3), (1, 1, 3), (2, 2, 3), (3, 3, 3)]


def qc_tensor(c20: torch.Tensor) -> torch.Tensor:
    """Symmetric 4x4x4 tensor T with sum_ijk T_ijk x_i x_j x_k = the 20-term polynomial.
    A coefficient is split over the distinct permutations of its index triple (1, 3 or 6),
    which is what the literal table in `data_io.py:95-120` spells out."""
    import itertools
    T = torch.zeros(4, 4, 4, dtype=torch.float64)
    for m, tri in enumerate(_QC_TRIPLES):
        perms = set(itertools.permutations(tri))
        for (i, j, k) in perms:
            T[i, j, k] = c20[m] / float(len(perms))
    return T


def qc_eval(x: torch.Tensor, T: torch.Tensor) -> torch.Tensor:
    """`QC_cal_en` (`tools/rpc_tensor.py:72-77`) / `cal_qc` (`modules/warping.py:54-57`):
    x [..., 4], T [4, 4, 4] -> [...]"""
    return torch.einsum("...i,...j,...k,ijk->...", x, x, x, T)


def rpc_localise_qc(samp, line, hei, rpc1: torch.Tensor):
    """`RPCModelParameter.RPC_PHOTO2OBJ` (`tools/rpc_tensor.py:138-165`) on flat point lists,
    single camera rpc1 [170]; x = (1, line_n, samp_n, h_n)."""
    l = (line - rpc1[0]) / rpc1[5]
    s = (samp - rpc1[1]) / rpc1[6]
    h = (hei - rpc1[4]) / rpc1[9]
    x = torch.stack((torch.ones_like(l), l, s, h), dim=-1)
    lat = qc_eval(x, qc_tensor(rpc1[90:110])) / qc_eval(x, qc_tensor(rpc1[110:130]))
    lon = qc_eval(x, qc_tensor(rpc1[130:150])) / qc_eval(x, qc_tensor(rpc1[150:170]))
    return lat * rpc1[7] + rpc1[2], lon * rpc1[8] + rpc1[3]


def rpc_project_qc(lat, lon, hei, rpc1: torch.Tensor):
    """`RPCModelParameter.RPC_OBJ2PHOTO` (`tools/rpc_tensor.py:109-136`); x = (1, lon_n, lat_n, h_n)."""
    lo = (lon - rpc1[3]) / rpc1[8]
    la = (lat - rpc1[2]) / rpc1[7]
    h = (hei - rpc1[4]) / rpc1[9]
    x = torch.stack((torch.ones_like(lo), lo, la, h), dim=-1)
    samp = qc_eval(x, qc_tensor(rpc1[50:70])) / qc_eval(x, qc_tensor(rpc1[70:90]))
    line = qc_eval(x, qc_tensor(rpc1[10:30])) / qc_eval(x, qc_tensor(rpc1[30:50]))
    return samp * rpc1[6] + rpc1[1], line * rpc1[5] + rpc1[0]


# ---------------------------------------------------------------------------------------
# sweep coordinates
# ---------------------------------------------------------------------------------------
def _expand_depth(depth_values: torch.Tensor, H: int, W: int) -> torch.Tensor:
    # [B, D] -> [B, D, H, W] broadcast (`warping.py:329-332`)
    if depth_values.dim() == 2:
        B, D = depth_values.shape
        return depth_values.view(B, D, 1, 1).expand(B, D, H, W)
    return depth_values


def rpc_sweep_coords(ref_rpc: torch.Tensor, src_rpc: torch.Tensor, depth_values: torch.Tensor,
                     H: int, W: int) -> tuple[torch.Tensor, torch.Tensor]:
    """Source-image (samp, line) in fp64 for every (b, d, y, x) of the reference grid:
    localisation with the reference camera then projection with the source camera
    (`modules/warping.py:322-341`).  Returns two [B, D, H, W] fp64 tensors."""
    B, D = depth_values.shape[:2]
    h = _expand_depth(depth_values, H, W).double().reshape(B, -1)
    dev = depth_values.device
    ys, xs = torch.meshgrid(torch.arange(H, dtype=torch.float64, device=dev),
                            torch.arange(W, dtype=torch.float64, device=dev), indexing="ij")
    x = xs.reshape(1, 1, H, W).expand(B, D, H, W).reshape(B, -1)
    y = ys.reshape(1, 1, H, W).expand(B, D, H, W).reshape(B, -1)
    lat, lon = rpc_localise(x, y, h, ref_rpc)
    samp, line = rpc_project(lat, lon, h, src_rpc)
    return samp.view(B, D, H, W), line.view(B, D, H, W)


def homo_sweep_coords(ref_proj: torch.Tensor, src_proj: torch.Tensor, depth_values: torch.Tensor,
                      H: int, W: int) -> tuple[torch.Tensor, torch.Tensor]:
    """Pin-hole plane sweep (`modules/warping.py:18-34`): P = src·ref^-1 (fp64),
    p = R·(x, y, 1)·d + t, (u, v) = p_xy / p_z.  No z<=0 guard, as in the reference."""
    B, D = depth_values.shape[:2]
    proj = torch.matmul(src_proj, torch.inverse(ref_proj))
    R, t = proj[:, :3, :3], proj[:, :3, 3:4]
    dev = depth_values.device
    ys, xs = torch.meshgrid(torch.arange(H, dtype=torch.float32, device=dev),
                            torch.arange(W, dtype=torch.float32, device=dev), indexing="ij")
    pix = torch.stack((xs.reshape(-1), ys.reshape(-1), torch.ones(H * W, device=dev))).double()   # [3, HW]
    rot = torch.matmul(R, pix.unsqueeze(0).expand(B, 3, H * W))                       # [B, 3, HW]
    d = _expand_depth(depth_values, H, W).reshape(B, 1, D, H * W).double()
    p = rot.unsqueeze(2) * d + t.view(B, 3, 1, 1)                                     # [B, 3, D, HW]
    uv = p[:, :2] / p[:, 2:3]
    return uv[:, 0].reshape(B, D, H, W), uv[:, 1].reshape(B, D, H, W)


def _fma32(a: torch.Tensor, b: torch.Tensor, c) -> torch.Tensor:
    """fp32 fused multiply-add emulated through fp64 (the product of two fp32 is exact in fp64)."""
    return (a.double() * b.double() + (c.double() if torch.is_tensor(c) else c)).float()


def to_tap_coords(samp: torch.Tensor, line: torch.Tensor, H: int, W: int, *, rpc_style: bool):
    """fp64 source coords -> fp32 *pixel-space tap positions* as `F.grid_sample(align_corners=False)`
    sees them.  RPC path casts to fp32 first and normalises in fp32 with a true division
    (`warping.py:347-351`); the homography path normalises in fp64 and then casts
    (`warping.py:35-38`).  The un-normalise step is ATen-CPU's `(g + 1) * (size / 2) - 0.5`
    evaluated as one fp32 add followed by one fused multiply-add (probed bit-exact against
    `F.grid_sample` on 8e5 samples; the CUDA eager path's `((g + 1) * size - 1) / 2` differs
    from it by <= 1 ulp)."""
    if rpc_style:
        gx = samp.float() / ((W - 1) / 2) - 1
        gy = line.float() / ((H - 1) / 2) - 1
    else:
        gx = (samp / ((W - 1) / 2) - 1).float()
        gy = (line / ((H - 1) / 2) - 1).float()
    ix = _fma32(gx + 1, torch.tensor(W / 2, dtype=torch.float32, device=gx.device), -0.5)
    iy = _fma32(gy + 1, torch.tensor(H / 2, dtype=torch.float32, device=gx.device), -0.5)
    return gx, gy, ix, iy


def bilinear_sample_zeros(fea: torch.Tensor, ix: torch.Tensor, iy: torch.Tensor) -> torch.Tensor:
    """Explicit restatement of `grid_sampler_2d` (bilinear, zeros padding): fea [B, C, H, W],
    ix/iy [B, D, H, W] fp32 pixel positions -> [B, C, D, H, W].  Corner weights are the
    products of differences used by ATen (nw = (x1-ix)(y1-iy), ...); the four corners are
    accumulated nw, ne, sw, se with fused multiply-adds; out-of-range corners add 0.
    Bit-exact against ATen-CPU `F.grid_sample` (tests/test_oracle_golden.py)."""
    B, C, H, W = fea.shape
    shp = ix.shape
    ix, iy = ix.reshape(B, -1), iy.reshape(B, -1)
    x0, y0 = torch.floor(ix), torch.floor(iy)
    x1, y1 = x0 + 1, y0 + 1
    flat = fea.reshape(B, C, H * W)
    out = torch.zeros(B, C, ix.shape[1], dtype=fea.dtype, device=fea.device)
    for xc, yc, wgt in ((x0, y0, (x1 - ix) * (y1 - iy)), (x1, y0, (ix - x0) * (y1 - iy)),
                        (x0, y1, (x1 - ix) * (iy - y0)), (x1, y1, (ix - x0) * (iy - y0))):
        ok = (xc >= 0) & (xc <= W - 1) & (yc >= 0) & (yc <= H - 1)
        idx = (yc.clamp(0, H - 1) * W + xc.clamp(0, W - 1)).long()
        val = torch.gather(flat, 2, idx.unsqueeze(1).expand(B, C, -1))
        out = _fma32(val, torch.where(ok, wgt, torch.zeros_like(wgt)).unsqueeze(1), out)
    return out.view(B, C, *shp[1:])


def _sample(src_fea, gx, gy, ix, iy, sampler: str):
    B, C, H, W = src_fea.shape
    D = gx.shape[1]
    if sampler == "aten":
        grid = torch.stack((gx.reshape(B, D, H * W), gy.reshape(B, D, H * W)), dim=3)
        out = F.grid_sample(src_fea, grid.view(B, D * H, W, 2), mode="bilinear",
                            padding_mode="zeros", align_corners=False)
        return out.view(B, C, D, H, W)
    return bilinear_sample_zeros(src_fea, ix, iy)


def rpc_warp(src_fea: torch.Tensor, src_rpc: torch.Tensor, ref_rpc: torch.Tensor,
             depth_values: torch.Tensor, sampler: str = "aten") -> torch.Tensor:
    """`rpc_warping` (`modules/warping.py:310-365`) -> [B, C, D, H, W] fp32."""
    H, W = src_fea.shape[2:]
    samp, line = rpc_sweep_coords(ref_rpc, src_rpc, depth_values, H, W)
    return _sample(src_fea, *to_tap_coords(samp, line, H, W, rpc_style=True), sampler)


def homo_warp(src_fea: torch.Tensor, src_proj: torch.Tensor, ref_proj: torch.Tensor,
              depth_values: torch.Tensor, sampler: str = "aten") -> torch.Tensor:
    """`homo_warping` (`modules/warping.py:6-44`) -> [B, C, D, H, W] fp32."""
    H, W = src_fea.shape[2:]
    u, v = homo_sweep_coords(ref_proj, src_proj, depth_values, H, W)
    return _sample(src_fea, *to_tap_coords(u, v, H, W, rpc_style=False), sampler)
