// geometry.cuh — per-hypothesis camera geometry of the plane sweep, fp64 on the device.
//
// Replaces the ATen op chains of RPC_Photo2Obj / RPC_Obj2Photo / RPC_PLH_COEF
// (modules/warping.py:183-307) and of homo_warping's projection (modules/warping.py:18-38).
// The camera "packs" below are built on the host from the reference's raw 170-double RPC
// vectors (dataset/data_io.py:78-92) or 4x4 K·E matrices (dataset/virdataset.py:67-70) and are
// passed to the kernels BY VALUE (__grid_constant__), so every coefficient is a constant-bank
// operand of a DFMA: no loads, no shared memory, no global state.
#pragma once
#include <cmath>
#include "common.cuh"

namespace satmvs {

// ------------------------------------------------------------------------------------------
// 20-term cubic, RPC00B monomial order of RPC_PLH_COEF (warping.py:189-207):
// [1, L, P, H, LP, LH, PH, LL, PP, HH, PLH, LLL, LPP, LHH, LLP, PPP, PHH, LLH, PPH, HHH]
// ------------------------------------------------------------------------------------------
struct Poly20 { double c[20]; };

// The polynomial regrouped as a cubic in H for fixed (L, P): a0 + a1 H + a2 H^2 + a3 H^3.
struct CubicH { double a0, a1, a2, a3; };

__host__ __device__ __forceinline__ CubicH collapse_lp(const Poly20& q, double L, double P) {
  const double* c = q.c;
  CubicH r;
  // H^0 : c0 + c1 L + c2 P + c4 LP + c7 LL + c8 PP + c11 LLL + c12 LPP + c14 LLP + c15 PPP  (Horner in P, then L)
  double p0 = fma(L, fma(L, fma(L, c[11], c[7]), c[1]), c[0]);
  double p1 = fma(L, fma(L, c[14], c[4]), c[2]);
  double p2 = fma(L, c[12], c[8]);
  r.a0 = fma(P, fma(P, fma(P, c[15], p2), p1), p0);
  // H^1 : c3 + c5 L + c6 P + c10 LP + c17 LL + c18 PP
  double q0 = fma(L, fma(L, c[17], c[5]), c[3]);
  double q1 = fma(L, c[10], c[6]);
  r.a1 = fma(P, fma(P, c[18], q1), q0);
  // H^2 : c9 + c13 L + c16 P ;  H^3 : c19
  r.a2 = fma(P, c[16], fma(L, c[13], c[9]));
  r.a3 = c[19];
  return r;
}

__host__ __device__ __forceinline__ double eval_h(const CubicH& k, double H) {
  return fma(H, fma(H, fma(H, k.a3, k.a2), k.a1), k.a0);
}

__host__ __device__ __forceinline__ double poly20(const Poly20& q, double L, double P, double H) {
  return eval_h(collapse_lp(q, L, P), H);   // 19 FMAs, no monomial table
}

// The same 19-FMA nesting evaluated for N hypothesis planes in lock step: each coefficient is fetched
// from the constant bank once and used N times, and the N chains are independent (ILP).
template <int N>
__device__ __forceinline__ void poly20_many(const Poly20& q, const double (&L)[N], const double (&P)[N],
                                            const double (&H)[N], double (&out)[N]) {
  const double* c = q.c;
  double a0[N], a1[N], a2[N], t[N];
#pragma unroll
  for (int k = 0; k < N; ++k) a0[k] = fma(L[k], c[11], c[7]);
#pragma unroll
  for (int k = 0; k < N; ++k) a0[k] = fma(L[k], a0[k], c[1]);
#pragma unroll
  for (int k = 0; k < N; ++k) a0[k] = fma(L[k], a0[k], c[0]);               // p0
#pragma unroll
  for (int k = 0; k < N; ++k) t[k] = fma(L[k], c[14], c[4]);
#pragma unroll
  for (int k = 0; k < N; ++k) t[k] = fma(L[k], t[k], c[2]);                 // p1
#pragma unroll
  for (int k = 0; k < N; ++k) a1[k] = fma(L[k], c[12], c[8]);               // p2
#pragma unroll
  for (int k = 0; k < N; ++k) a1[k] = fma(P[k], c[15], a1[k]);
#pragma unroll
  for (int k = 0; k < N; ++k) a1[k] = fma(P[k], a1[k], t[k]);
#pragma unroll
  for (int k = 0; k < N; ++k) a0[k] = fma(P[k], a1[k], a0[k]);              // H^0 term
#pragma unroll
  for (int k = 0; k < N; ++k) a1[k] = fma(L[k], c[17], c[5]);
#pragma unroll
  for (int k = 0; k < N; ++k) a1[k] = fma(L[k], a1[k], c[3]);               // q0
#pragma unroll
  for (int k = 0; k < N; ++k) t[k] = fma(L[k], c[10], c[6]);                // q1
#pragma unroll
  for (int k = 0; k < N; ++k) t[k] = fma(P[k], c[18], t[k]);
#pragma unroll
  for (int k = 0; k < N; ++k) a1[k] = fma(P[k], t[k], a1[k]);               // H^1 term
#pragma unroll
  for (int k = 0; k < N; ++k) a2[k] = fma(L[k], c[13], c[9]);
#pragma unroll
  for (int k = 0; k < N; ++k) a2[k] = fma(P[k], c[16], a2[k]);              // H^2 term
#pragma unroll
  for (int k = 0; k < N; ++k) out[k] = fma(H[k], fma(H[k], fma(H[k], c[19], a2[k]), a1[k]), a0[k]);
}

// 1/x to ~1 ulp: hardware seed (MUFU.RCP64H) + two Newton steps.  The reference divides with
// IEEE fp64; the difference (<= ~2 ulp of fp64) is 9 orders below the fp32 tap-coordinate ulp.
__device__ __forceinline__ double fast_rcp(double x) {
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
  double e = fma(-x, r, 1.0);
  r = fma(r, e, r);
  e = fma(-x, r, 1.0);
  r = fma(r, e, r);
  return r;
}

// ------------------------------------------------------------------------------------------
// RPC packs
// ------------------------------------------------------------------------------------------
struct RpcRefPack {            // localisation (image + height -> normalised lat/lon), warping.py:255-307
  double samp_off, samp_iscale, line_off, line_iscale, hei_off, hei_iscale;
  Poly20 lat_num, lat_den, lon_num, lon_den;
  double lat_scale, lat_off, lon_scale, lon_off;   // only used by the point-list op
};

struct RpcSrcPack {            // projection (lat/lon/height -> image), warping.py:218-252
  // normalised source coordinates from the reference camera's NORMALISED lat/lon (affine composed
  // on the host in fp64, so degree-valued lat/lon are never formed on the device):
  double p_a, p_b;             // P = lat_n_ref * p_a + p_b
  double l_a, l_b;             // L = lon_n_ref * l_a + l_b
  double h_a, h_b;             // H = h * h_a + h_b   (h in metres)
  Poly20 samp_num, samp_den, line_num, line_den;
  double samp_scale, samp_off, line_scale, line_off;
  double lat_off, lat_iscale, lon_off, lon_iscale;  // only used by the point-list op
};

inline void load_poly(Poly20& q, const double* rpc, int at) { for (int i = 0; i < 20; ++i) q.c[i] = rpc[at + i]; }

inline RpcRefPack make_rpc_ref_pack(const double* r) {
  RpcRefPack p;
  p.line_off = r[0]; p.samp_off = r[1]; p.hei_off = r[4];
  p.line_iscale = 1.0 / r[5]; p.samp_iscale = 1.0 / r[6]; p.hei_iscale = 1.0 / r[9];
  load_poly(p.lat_num, r, 90); load_poly(p.lat_den, r, 110);
  load_poly(p.lon_num, r, 130); load_poly(p.lon_den, r, 150);
  p.lat_scale = r[7]; p.lat_off = r[2]; p.lon_scale = r[8]; p.lon_off = r[3];
  return p;
}

inline RpcSrcPack make_rpc_src_pack(const double* s, const double* ref) {
  RpcSrcPack p;
  // lat = lat_n_ref*LAT_SCALE_ref + LAT_OFF_ref ; P = (lat - LAT_OFF_src)/LAT_SCALE_src
  p.p_a = ref[7] / s[7]; p.p_b = (ref[2] - s[2]) / s[7];
  p.l_a = ref[8] / s[8]; p.l_b = (ref[3] - s[3]) / s[8];
  p.h_a = 1.0 / s[9];    p.h_b = -s[4] / s[9];
  load_poly(p.line_num, s, 10); load_poly(p.line_den, s, 30);
  load_poly(p.samp_num, s, 50); load_poly(p.samp_den, s, 70);
  p.samp_scale = s[6]; p.samp_off = s[1]; p.line_scale = s[5]; p.line_off = s[0];
  p.lat_off = s[2]; p.lat_iscale = 1.0 / s[7]; p.lon_off = s[3]; p.lon_iscale = 1.0 / s[8];
  return p;
}

// x / b for a launch constant b, correctly rounded: q = RN(x*y), r = x - b*q (exact, fma),
// q' = RN(q + r*y) with y = RN(1/b) from the host (Markstein).  3 instructions instead of the
// ~12 + slow-path call of an IEEE fp32 division; equals torch's true division bit for bit
// (checked exhaustively on 4e6 samples per divisor in DESIGN.md §numerics).
__device__ __forceinline__ float div_const(float x, float b, float inv_b) {
  float q = __fmul_rn(x, inv_b);
  float r = __fmaf_rn(-b, q, x);
  return __fmaf_rn(r, inv_b, q);
}

// num/den for two ratios with ONE reciprocal: a = an/ad, b = bn/bd
__device__ __forceinline__ void ratio2(double an, double ad, double bn, double bd, double& a, double& b) {
  double r = fast_rcp(ad * bd);
  a = an * bd * r;
  b = bn * ad * r;
}

template <int NSRC>
struct RpcSweep {
  static constexpr int kNumSrc = NSRC;
  RpcRefPack ref;
  RpcSrcPack src[NSRC];
  float half_wm1, half_hm1;      // (W-1)/2, (H-1)/2 as fp32 (warping.py:350-351)
  float inv_half_wm1, inv_half_hm1;

  struct Pixel { CubicH lat_num, lat_den, lon_num, lon_den; };
  struct Plane { double lat_n, lon_n, h; };

  // once per reference pixel: the four localisation polynomials collapse to cubics in H,
  // because (samp, line) of the pixel are fixed and only the hypothesis height varies with d
  __device__ __forceinline__ Pixel pixel(int x, int y) const {
    double s = ((double)x - ref.samp_off) * ref.samp_iscale;   // P = samp
    double l = ((double)y - ref.line_off) * ref.line_iscale;   // L = line   (warping.py:280)
    Pixel p;
    p.lat_num = collapse_lp(ref.lat_num, l, s);
    p.lat_den = collapse_lp(ref.lat_den, l, s);
    p.lon_num = collapse_lp(ref.lon_num, l, s);
    p.lon_den = collapse_lp(ref.lon_den, l, s);
    return p;
  }

  __device__ __forceinline__ Plane plane(const Pixel& p, float h32) const {
    Plane pl;
    pl.h = (double)h32;                                        // h.double(), warping.py:337
    double hn = (pl.h - ref.hei_off) * ref.hei_iscale;
    ratio2(eval_h(p.lat_num, hn), eval_h(p.lat_den, hn), eval_h(p.lon_num, hn), eval_h(p.lon_den, hn),
           pl.lat_n, pl.lon_n);
    return pl;
  }

  // normalised grid coordinates of source view v (fp32), following warping.py:347-351:
  // cast samp/line to fp32 first, then x / ((W-1)/2) - 1 with a true fp32 division.
  __device__ __forceinline__ void project(int v, const Pixel&, const Plane& pl, float& gx, float& gy) const {
    const RpcSrcPack& s = src[v];
    double P = fma(pl.lat_n, s.p_a, s.p_b);                    // P = lat, L = lon (warping.py:238)
    double L = fma(pl.lon_n, s.l_a, s.l_b);
    double H = fma(pl.h, s.h_a, s.h_b);
    double sn, ln;
    ratio2(poly20(s.samp_num, L, P, H), poly20(s.samp_den, L, P, H),
           poly20(s.line_num, L, P, H), poly20(s.line_den, L, P, H), sn, ln);
    float samp = (float)fma(sn, s.samp_scale, s.samp_off);
    float line = (float)fma(ln, s.line_scale, s.line_off);
    gx = __fsub_rn(div_const(samp, half_wm1, inv_half_wm1), 1.0f);
    gy = __fsub_rn(div_const(line, half_hm1, inv_half_hm1), 1.0f);
  }

  // grid coordinates of N planes x all source views for one reference pixel (same arithmetic as
  // pixel/plane/project, planes innermost so every constant-bank coefficient is fetched once per N)
  template <int N, class Emit>
  __device__ __forceinline__ void grid_coords(const Pixel& px, const float (&h32)[N], Emit&& emit) const {
    double latn[N], lonn[N], hh[N];
#pragma unroll
    for (int k = 0; k < N; ++k) {
      const Plane pl = plane(px, h32[k]);
      latn[k] = pl.lat_n; lonn[k] = pl.lon_n; hh[k] = pl.h;
    }
#pragma unroll 1          // one copy of the 4 x 19 x N FMA body in the instruction cache, views iterate over it
    for (int v = 0; v < NSRC; ++v) {
      const RpcSrcPack& s = src[v];
      double P[N], L[N], H[N], sn[N], sd[N], ln[N], ld[N];
#pragma unroll
      for (int k = 0; k < N; ++k) {
        P[k] = fma(latn[k], s.p_a, s.p_b);
        L[k] = fma(lonn[k], s.l_a, s.l_b);
        H[k] = fma(hh[k], s.h_a, s.h_b);
      }
      poly20_many<N>(s.samp_num, L, P, H, sn);
      poly20_many<N>(s.samp_den, L, P, H, sd);
      poly20_many<N>(s.line_num, L, P, H, ln);
      poly20_many<N>(s.line_den, L, P, H, ld);
#pragma unroll
      for (int k = 0; k < N; ++k) {
        double a, b;
        ratio2(sn[k], sd[k], ln[k], ld[k], a, b);
        const float samp = (float)fma(a, s.samp_scale, s.samp_off);
        const float line = (float)fma(b, s.line_scale, s.line_off);
        emit(k, v, __fsub_rn(div_const(samp, half_wm1, inv_half_wm1), 1.0f),
             __fsub_rn(div_const(line, half_hm1, inv_half_hm1), 1.0f));
      }
    }
  }
};

// ------------------------------------------------------------------------------------------
// pin-hole homography, modules/warping.py:18-38
// ------------------------------------------------------------------------------------------
struct HomoSrcPack { double r[9]; double t[3]; };   // proj = src_proj * inv(ref_proj): rot, trans

// 4x4 inverse by Gauss-Jordan with partial pivoting (torch.inverse, warping.py:19); returns false if singular
inline bool invert4(const double* m, double* inv) {
  double a[4][8];
  for (int i = 0; i < 4; ++i)
    for (int j = 0; j < 4; ++j) { a[i][j] = m[i * 4 + j]; a[i][4 + j] = (i == j) ? 1.0 : 0.0; }
  for (int col = 0; col < 4; ++col) {
    int piv = col;
    for (int i = col + 1; i < 4; ++i) if (std::fabs(a[i][col]) > std::fabs(a[piv][col])) piv = i;
    if (a[piv][col] == 0.0) return false;
    if (piv != col) for (int j = 0; j < 8; ++j) { double t = a[col][j]; a[col][j] = a[piv][j]; a[piv][j] = t; }
    double d = a[col][col];
    for (int j = 0; j < 8; ++j) a[col][j] /= d;
    for (int i = 0; i < 4; ++i) {
      if (i == col) continue;
      double f = a[i][col];
      if (f != 0.0) for (int j = 0; j < 8; ++j) a[i][j] -= f * a[col][j];
    }
  }
  for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) inv[i * 4 + j] = a[i][4 + j];
  return true;
}

inline bool make_homo_src_pack(HomoSrcPack& p, const double* src_proj, const double* ref_proj) {
  double inv[16];
  if (!invert4(ref_proj, inv)) return false;
  for (int i = 0; i < 3; ++i) {
    for (int j = 0; j < 4; ++j) {
      double acc = 0.0;
      for (int k = 0; k < 4; ++k) acc += src_proj[i * 4 + k] * inv[k * 4 + j];
      if (j < 3) p.r[i * 3 + j] = acc; else p.t[i] = acc;
    }
  }
  return true;
}

template <int NSRC>
struct HomoSweep {
  static constexpr int kNumSrc = NSRC;
  HomoSrcPack src[NSRC];
  double inv_half_wm1, inv_half_hm1;   // 1/((W-1)/2), 1/((H-1)/2) in fp64 (warping.py:35-36)

  struct Pixel { double rx[NSRC], ry[NSRC], rz[NSRC]; };   // R·(x, y, 1) per view (warping.py:30)
  struct Plane { double d; };

  __device__ __forceinline__ Pixel pixel(int x, int y) const {
    Pixel p;
#pragma unroll
    for (int v = 0; v < NSRC; ++v) {
      const double* r = src[v].r;
      p.rx[v] = fma(r[0], (double)x, fma(r[1], (double)y, r[2]));
      p.ry[v] = fma(r[3], (double)x, fma(r[4], (double)y, r[5]));
      p.rz[v] = fma(r[6], (double)x, fma(r[7], (double)y, r[8]));
    }
    return p;
  }
  __device__ __forceinline__ Plane plane(const Pixel&, float d32) const { return Plane{(double)d32}; }

  // homo_warping normalises in fp64 and casts to fp32 last (warping.py:34-38)
  __device__ __forceinline__ void project(int v, const Pixel& p, const Plane& pl, float& gx, float& gy) const {
    double X = fma(p.rx[v], pl.d, src[v].t[0]);
    double Y = fma(p.ry[v], pl.d, src[v].t[1]);
    double Z = fma(p.rz[v], pl.d, src[v].t[2]);
    double iz = fast_rcp(Z);             // Z == 0 -> NaN coordinates -> every tap out of range -> 0, like the reference's inf
    gx = (float)(X * iz * inv_half_wm1 - 1.0);
    gy = (float)(Y * iz * inv_half_hm1 - 1.0);
  }

  template <int N, class Emit>
  __device__ __forceinline__ void grid_coords(const Pixel& px, const float (&d32)[N], Emit&& emit) const {
#pragma unroll
    for (int k = 0; k < N; ++k) {
      const Plane pl = plane(px, d32[k]);
#pragma unroll
      for (int v = 0; v < NSRC; ++v) {
        float gx, gy;
        project(v, px, pl, gx, gy);
        emit(k, v, gx, gy);
      }
    }
  }
};

// ------------------------------------------------------------------------------------------
// bilinear tap record: grid_sampler_2d(bilinear, zeros, align_corners=False) semantics
// ------------------------------------------------------------------------------------------
struct Tap {
  int off;                     // y*W + x of the clamped 2x2 window origin (always in bounds)
  float w00, w01, w10, w11;    // corner weights in window order; out-of-range corners carry 0
};

// Same record with the clamped window origin kept as (x, y) and a flag telling whether any corner carries
// weight (used by the shared-memory-staged kernels to bound the source window of a CTA).
struct TapXY { int xc, yc; float w00, w01, w10, w11; bool live; };

// un-normalise exactly like ATen: (g + 1) rounded, then one fused multiply-add with size/2 and -0.5
__device__ __forceinline__ TapXY make_tap_xy(float gx, float gy, int H, int W, float half_w, float half_h) {
  float ix = __fmaf_rn(__fadd_rn(gx, 1.0f), half_w, -0.5f);
  float iy = __fmaf_rn(__fadd_rn(gy, 1.0f), half_h, -0.5f);
  float x0f = floorf(ix), y0f = floorf(iy);
  float wx0 = __fsub_rn(__fadd_rn(x0f, 1.0f), ix), wx1 = __fsub_rn(ix, x0f);   // (x1 - ix), (ix - x0)
  float wy0 = __fsub_rn(__fadd_rn(y0f, 1.0f), iy), wy1 = __fsub_rn(iy, y0f);
  // NaN/inf-safe integer corner: anything outside [-2, size] has no in-range tap
  int X0 = (int)fminf(fmaxf(x0f, -2.0f), (float)W);
  int Y0 = (int)fminf(fmaxf(y0f, -2.0f), (float)H);
  int xc = min(max(X0, 0), W - 2), yc = min(max(Y0, 0), H - 2);
  float wxa = (X0 == xc) ? wx0 : ((X0 == xc - 1) ? wx1 : 0.0f);
  float wxb = (X0 == xc) ? wx1 : ((X0 == xc + 1) ? wx0 : 0.0f);
  float wya = (Y0 == yc) ? wy0 : ((Y0 == yc - 1) ? wy1 : 0.0f);
  float wyb = (Y0 == yc) ? wy1 : ((Y0 == yc + 1) ? wy0 : 0.0f);
  TapXY t;
  t.xc = xc; t.yc = yc;
  t.w00 = __fmul_rn(wxa, wya); t.w01 = __fmul_rn(wxb, wya);
  t.w10 = __fmul_rn(wxa, wyb); t.w11 = __fmul_rn(wxb, wyb);
  t.live = (t.w00 != 0.0f) || (t.w01 != 0.0f) || (t.w10 != 0.0f) || (t.w11 != 0.0f);
  return t;
}

__device__ __forceinline__ Tap make_tap(float gx, float gy, int H, int W, float half_w, float half_h) {
  const TapXY r = make_tap_xy(gx, gy, H, W, half_w, half_h);
  Tap t;
  t.off = r.yc * W + r.xc;
  t.w00 = r.w00; t.w01 = r.w01; t.w10 = r.w10; t.w11 = r.w11;
  return t;
}

// nw, ne, sw, se accumulated with fused multiply-adds, the order ATen uses
// f0 = channel plane, f1 = f0 + W (both warp-uniform), so the per-thread part of each address is t.off only
__device__ __forceinline__ float tap_fetch(const float* __restrict__ f0, const float* __restrict__ f1, const Tap& t) {
  float v00 = __ldg(f0 + t.off), v01 = __ldg(f0 + t.off + 1), v10 = __ldg(f1 + t.off), v11 = __ldg(f1 + t.off + 1);
  return __fmaf_rn(v11, t.w11, __fmaf_rn(v10, t.w10, __fmaf_rn(v01, t.w01, __fmul_rn(v00, t.w00))));
}

}  // namespace satmvs
