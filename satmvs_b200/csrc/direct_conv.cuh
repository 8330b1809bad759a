// direct_conv.cuh — register-tiled direct 3x3 / 3x3x3 convolution (stride 1 or 2, padding 1) for the large,
// throughput-bound layers: RED's batched encoders and GRU x-halves (2-D taps applied to every depth
// plane of a [C,D,h,w] tensor) and CostRegNet's dense 3-D convs.
//
// The implicit-GEMM engine (conv_engine.cuh) spends ~40 % of its instructions gathering im2col
// operands (one predicated load + index arithmetic per (k, pixel) element).  Here one thread owns
// 8 output channels x 4 consecutive output pixels of a row and keeps the three input rows of a
// (channel, kz) slice in registers, so the 3 horizontal taps of all 4 pixels come out of ONE
// 128-bit load + 2 scalars per row: 9 loads + 18 broadcast LDS.128 of weights feed 288 FMAs (88 % FMA
// density), with a register ring prefetching the next slices.  Requires widths that are multiples of 4
// (16-byte aligned rows); the caller falls back to the engine otherwise.
#pragma once
#include "common.cuh"

namespace satmvs {

struct DirectConv {
  const float* in;        // [.][Di][Hi][Wi]; first used channel already applied to the pointer
  const float* w;         // w[co * w_co + ci * w_ci + tap], tap = (kz*3 + ky)*3 + kx  (kz only for NZ = 3)
  const float* scale;     // [Cout] or null
  const float* shift;     // [Cout] or null
  const float* post_add;  // indexed like out, added after the activation, or null
  float* out;             // [Cout][Do][Ho][Wo]
  int Cin, Cout, Di, Hi, Wi, Do, Ho, Wo;
  long long w_co, w_ci;
  long long in_cs;        // input channel stride in elements (0: Di*Hi*Wi)
  float acc_scale;
  int relu;
  int flip;               // 1: taps mirrored (tap -> TAPS-1-tap): a stride-1 transposed conv as a correlation
};

constexpr int kDcWarps = 4, kDcCo = 8, kDcPx = 4, kDcCiChunk = 16;

template <int NZ, int S>
__global__ void __launch_bounds__(kDcWarps * 32)
direct_conv_kernel(const __grid_constant__ DirectConv a) {
  constexpr int TAPS = NZ * 9;
  constexpr int RW = (S == 1) ? 6 : 9;                 // input values per row feeding 4 output pixels
  __shared__ __align__(16) float wsm[kDcCiChunk * TAPS * kDcCo];     // [ci][tap][co]

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int co0 = blockIdx.y * kDcCo;
  const int npx = a.Ho * a.Wo;
  // lanes enumerate groups of 4 output pixels over ALL output planes (no idle lanes on small planes);
  // Wo % 4 == 0 keeps a group inside one row
  const long long g0 = ((long long)blockIdx.x * kDcWarps + warp) * (32 * kDcPx) + lane * kDcPx;
  const bool ok = g0 < (long long)npx * a.Do;
  const int oz = ok ? (int)(g0 / npx) : 0;
  const int p0 = ok ? (int)(g0 - (long long)oz * npx) : 0;                     // first output pixel of this lane
  const int oy = p0 / a.Wo, ox = p0 - oy * a.Wo;
  const int iy0 = oy * S - 1, ix0 = ox * S;                                      // top row, first aligned column
  const long long in_plane = (long long)a.Hi * a.Wi, in_cs = a.in_cs ? a.in_cs : in_plane * a.Di;

  // validity of the NZ x 3 rows and of the two edge columns
  bool zok[NZ], yok[3];
#pragma unroll
  for (int kz = 0; kz < NZ; ++kz) {
    const int iz = (NZ == 1) ? oz : oz * S - 1 + kz;
    zok[kz] = ok && (unsigned)iz < (unsigned)a.Di;
  }
#pragma unroll
  for (int ky = 0; ky < 3; ++ky) yok[ky] = (unsigned)(iy0 + ky) < (unsigned)a.Hi;
  const bool lok = ix0 > 0, rok = (S == 1) && (ix0 + kDcPx < a.Wi);
  const float* in0 = a.in + (long long)((NZ == 1) ? oz : oz * S - 1) * in_plane + (long long)iy0 * a.Wi + ix0;

  // rows[ky][0 .. RW-1] = in[iz][iy0+ky][ix0-1 .. ix0-1+RW-1] of one (channel, kz) slice
  auto load_slice = [&](int ci, int kz, float (&r)[3][RW]) {
    const float* sp = in0 + ci * in_cs + kz * in_plane;
#pragma unroll
    for (int ky = 0; ky < 3; ++ky) {
      const float* rp = sp + ky * a.Wi;
      float4 m0 = make_float4(0.f, 0.f, 0.f, 0.f), m1 = m0;
      float lft = 0.f, rgt = 0.f;
      if (zok[kz] && yok[ky]) {
        m0 = __ldg(reinterpret_cast<const float4*>(rp));
        if (S == 2) m1 = __ldg(reinterpret_cast<const float4*>(rp + 4));
        if (lok) lft = __ldg(rp - 1);
        if (rok) rgt = __ldg(rp + kDcPx);
      }
      r[ky][0] = lft; r[ky][1] = m0.x; r[ky][2] = m0.y; r[ky][3] = m0.z; r[ky][4] = m0.w;
      if (S == 1) { r[ky][5] = rgt; }
      else { r[ky][5] = m1.x; r[ky][6] = m1.y; r[ky][7] = m1.z; r[ky][8] = m1.w; }
    }
  };

  float acc[kDcCo][kDcPx];
#pragma unroll
  for (int i = 0; i < kDcCo; ++i)
#pragma unroll
    for (int j = 0; j < kDcPx; ++j) acc[i][j] = 0.0f;

  auto fma_slice = [&](int cil, int kz, const float (&r)[3][RW]) {
#pragma unroll
    for (int t = 0; t < 9; ++t) {
      const float* wp = &wsm[((cil * NZ + kz) * 9 + t) * kDcCo];
      const float4 w0 = *reinterpret_cast<const float4*>(wp), w1 = *reinterpret_cast<const float4*>(wp + 4);
      const float wv[kDcCo] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
#pragma unroll
      for (int i = 0; i < kDcCo; ++i)
#pragma unroll
        for (int j = 0; j < kDcPx; ++j) acc[i][j] = fmaf(wv[i], r[t / 3][j * S + t % 3], acc[i][j]);
    }
  };

  for (int c0 = 0; c0 < a.Cin; c0 += kDcCiChunk) {
    const int nci = min(kDcCiChunk, a.Cin - c0);
    __syncthreads();                                   // previous chunk's weights no longer in use
    {  // stage this chunk's weights: [co][ci][tap] in global -> [ci][tap][co] in smem, loads batched
      constexpr int kPer = (kDcCiChunk * TAPS * kDcCo + kDcWarps * 32 - 1) / (kDcWarps * 32);
      float t[kPer];
#pragma unroll
      for (int k = 0; k < kPer; ++k) {
        const int e = tid + k * (kDcWarps * 32);
        const int co = e / (kDcCiChunk * TAPS), rr = e - co * (kDcCiChunk * TAPS);
        const int ci = rr / TAPS, tp = rr - ci * TAPS;
        t[k] = (co < kDcCo && ci < nci && co0 + co < a.Cout)
                   ? __ldg(a.w + (long long)(co0 + co) * a.w_co + (long long)(c0 + ci) * a.w_ci + (a.flip ? TAPS - 1 - tp : tp)) : 0.0f;
      }
#pragma unroll
      for (int k = 0; k < kPer; ++k) {
        const int e = tid + k * (kDcWarps * 32);
        const int co = e / (kDcCiChunk * TAPS), rr = e - co * (kDcCiChunk * TAPS);
        if (co < kDcCo) wsm[rr * kDcCo + co] = t[k];
      }
    }
    __syncthreads();
    // slices s = (ci, kz) of this chunk through a 3-slot register ring: 2 slices in flight under the FMAs
    const int nsl = nci * NZ;
    float ring[3][3][RW];
#pragma unroll
    for (int s = 0; s < 2; ++s)
      if (s < nsl) load_slice(c0 + s / NZ, s % NZ, ring[s]);
#pragma unroll 1
    for (int s = 0; s < nsl; s += 3) {
#pragma unroll
      for (int u = 0; u < 3; ++u) {
        const int cur = s + u, nxt = cur + 2;
        if (nxt < nsl) load_slice(c0 + nxt / NZ, nxt % NZ, ring[(u + 2) % 3]);
        if (cur < nsl) fma_slice(cur / NZ, cur % NZ, ring[u]);
      }
    }
  }

  if (!ok) return;
  const long long out_plane = (long long)a.Ho * a.Wo;
#pragma unroll
  for (int i = 0; i < kDcCo; ++i) {
    const int co = co0 + i;
    if (co >= a.Cout) break;
    const long long idx = ((long long)co * a.Do + oz) * out_plane + p0;
    const float sc = a.scale ? __ldg(a.scale + co) : 1.0f, sh = a.shift ? __ldg(a.shift + co) : 0.0f;
    float v[kDcPx];
#pragma unroll
    for (int j = 0; j < kDcPx; ++j) {
      v[j] = acc[i][j] * a.acc_scale * sc + sh;
      if (a.relu) v[j] = fmaxf(v[j], 0.0f);
    }
    if (a.post_add) {
      const float4 pa = __ldg(reinterpret_cast<const float4*>(a.post_add + idx));
      v[0] += pa.x; v[1] += pa.y; v[2] += pa.z; v[3] += pa.w;
    }
    *reinterpret_cast<float4*>(a.out + idx) = make_float4(v[0], v[1], v[2], v[3]);
  }
}

// true if the layer shape is one the direct kernel handles (16-byte aligned rows)
inline bool direct_conv_supported(const DirectConv& p, int NZ, int S) {
  if (p.Wi % 4 || p.Wo % 4) return false;
  if (S == 2 && (p.Wi != 2 * p.Wo || p.Hi != 2 * p.Ho)) return false;
  if (S == 1 && (p.Wi != p.Wo || p.Hi != p.Ho)) return false;
  if (NZ == 3 && ((S == 2 && p.Di != 2 * p.Do) || (S == 1 && p.Di != p.Do))) return false;
  if (NZ == 1 && p.Di != p.Do) return false;
  return (reinterpret_cast<uintptr_t>(p.in) % 16 == 0) && (reinterpret_cast<uintptr_t>(p.out) % 16 == 0) &&
         (p.post_add == nullptr || reinterpret_cast<uintptr_t>(p.post_add) % 16 == 0);
}

inline int direct_conv_launch(const DirectConv& p, int NZ, int S, cudaStream_t st, const char* what);
// 2-D stride-1 launch with an explicit input channel stride (input tensor with extra planes per channel)
inline int direct_conv_launch_cs(DirectConv p, long long in_cs, cudaStream_t st, const char* what) {
  p.in_cs = in_cs;
  return direct_conv_launch(p, 1, 1, st, what);
}

inline int direct_conv_launch(const DirectConv& p, int NZ, int S, cudaStream_t st, const char* what) {
  dim3 grid(ceil_div((long long)p.Ho * p.Wo * p.Do, kDcWarps * 32 * kDcPx), ceil_div(p.Cout, kDcCo), 1);
  if (NZ == 1 && S == 1) direct_conv_kernel<1, 1><<<grid, kDcWarps * 32, 0, st>>>(p);
  else if (NZ == 1 && S == 2) direct_conv_kernel<1, 2><<<grid, kDcWarps * 32, 0, st>>>(p);
  else if (NZ == 3 && S == 1) direct_conv_kernel<3, 1><<<grid, kDcWarps * 32, 0, st>>>(p);
  else direct_conv_kernel<3, 2><<<grid, kDcWarps * 32, 0, st>>>(p);
  return check_launch(what);
}

}  // namespace satmvs
