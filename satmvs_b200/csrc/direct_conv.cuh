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
#include "packed.cuh"

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
  int ksplit;             // > 1: grid.z CTAs share the input channels (small deep layers: too few output tiles to fill the GPU);
  float* partial;         //      raw sums go to partial[z][Cout][Do*Ho*Wo], direct_conv_finish_kernel adds them and applies the epilogue
};

constexpr int kDcWarps = 4, kDcCo = 8, kDcPx = 4, kDcCiChunk = 16;

#ifndef SATMVS_DC_MINB
#define SATMVS_DC_MINB 4   // 16 warps per SM: measured 1.28 -> 0.99 ms on the batched RED convs against (1, ring 3)
#endif
#ifndef SATMVS_DC_RING
#define SATMVS_DC_RING 2
#endif
template <int NZ, int S>
__global__ void __launch_bounds__(kDcWarps * 32, SATMVS_DC_MINB)
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

  // accumulators as (co, co+1) pairs: one FFMA2 = 2 output channels x 1 pixel; the weight pairs come
  // straight out of the 128-bit shared-memory loads, only the input value is duplicated (once per row value)
  u64 acc[kDcCo / 2][kDcPx];
#pragma unroll
  for (int i = 0; i < kDcCo / 2; ++i)
#pragma unroll
    for (int j = 0; j < kDcPx; ++j) acc[i][j] = 0ULL;

  auto fma_slice = [&](int cil, int kz, const float (&r)[3][RW]) {
    u64 rr[3][RW];
#pragma unroll
    for (int ky = 0; ky < 3; ++ky)
#pragma unroll
      for (int c = 0; c < RW; ++c) rr[ky][c] = pk(r[ky][c], r[ky][c]);
#pragma unroll
    for (int t = 0; t < 9; ++t) {
      const float* wp = &wsm[((cil * NZ + kz) * 9 + t) * kDcCo];
      const ulonglong2 w0 = *reinterpret_cast<const ulonglong2*>(wp), w1 = *reinterpret_cast<const ulonglong2*>(wp + 4);
      const u64 wv[kDcCo / 2] = {w0.x, w0.y, w1.x, w1.y};
#pragma unroll
      for (int i = 0; i < kDcCo / 2; ++i)
#pragma unroll
        for (int j = 0; j < kDcPx; ++j) acc[i][j] = ffma2(wv[i], rr[t / 3][j * S + t % 3], acc[i][j]);
    }
  };

  const int c_per = a.ksplit > 1 ? (a.Cin / a.ksplit + kDcCiChunk - 1) / kDcCiChunk * kDcCiChunk : a.Cin;
  const int c_begin = a.ksplit > 1 ? (int)blockIdx.z * c_per : 0, c_end = min(a.Cin, c_begin + c_per);
  for (int c0 = c_begin; c0 < c_end; c0 += kDcCiChunk) {
    const int nci = min(kDcCiChunk, c_end - c0);
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
    constexpr int RD = SATMVS_DC_RING;                 // ring slots: RD - 1 slices in flight under the FMAs
    float ring[RD][3][RW];
#pragma unroll
    for (int s = 0; s < RD - 1; ++s)
      if (s < nsl) load_slice(c0 + s / NZ, s % NZ, ring[s]);
#pragma unroll 1
    for (int s = 0; s < nsl; s += RD) {
#pragma unroll
      for (int u = 0; u < RD; ++u) {
        const int cur = s + u, nxt = cur + RD - 1;
        if (nxt < nsl) load_slice(c0 + nxt / NZ, nxt % NZ, ring[(u + RD - 1) % RD]);
        if (cur < nsl) fma_slice(cur / NZ, cur % NZ, ring[u]);
      }
    }
  }

  if (!ok) return;
  const long long out_plane = (long long)a.Ho * a.Wo;
  if (a.ksplit > 1) {        // raw partial sums; the epilogue runs in direct_conv_finish_kernel
    float* pp = a.partial + (long long)blockIdx.z * a.Cout * a.Do * out_plane;
#pragma unroll
    for (int i = 0; i < kDcCo; ++i) {
      const int co = co0 + i;
      if (co >= a.Cout) break;
      float v[kDcPx];
#pragma unroll
      for (int j = 0; j < kDcPx; ++j) { float alo, ahi; upk(acc[i >> 1][j], alo, ahi); v[j] = (i & 1) ? ahi : alo; }
      *reinterpret_cast<float4*>(pp + ((long long)co * a.Do + oz) * out_plane + p0) = make_float4(v[0], v[1], v[2], v[3]);
    }
    return;
  }
#pragma unroll
  for (int i = 0; i < kDcCo; ++i) {
    const int co = co0 + i;
    if (co >= a.Cout) break;
    const long long idx = ((long long)co * a.Do + oz) * out_plane + p0;
    const float sc = a.scale ? __ldg(a.scale + co) : 1.0f, sh = a.shift ? __ldg(a.shift + co) : 0.0f;
    float v[kDcPx];
#pragma unroll
    for (int j = 0; j < kDcPx; ++j) {
      float alo, ahi;
      upk(acc[i >> 1][j], alo, ahi);
      v[j] = ((i & 1) ? ahi : alo) * a.acc_scale * sc + sh;
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
// out = epilogue(sum over the ksplit partial tensors, in their order): scale / shift, ReLU, + post_add
template <int kUnused>     // template only so the header can be included from several translation units
__global__ void __launch_bounds__(256) direct_conv_finish_kernel(const DirectConv a) {
  const long long per_c = (long long)a.Do * a.Ho * a.Wo, total = per_c * a.Cout;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int co = (int)(i / per_c);
    float v = 0.0f;
    for (int z = 0; z < a.ksplit; ++z) v += a.partial[(long long)z * total + i];
    v = v * a.acc_scale * (a.scale ? __ldg(a.scale + co) : 1.0f) + (a.shift ? __ldg(a.shift + co) : 0.0f);
    if (a.relu) v = fmaxf(v, 0.0f);
    if (a.post_add) v += __ldg(a.post_add + i);
    a.out[i] = v;
  }
}

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
  dim3 grid(ceil_div((long long)p.Ho * p.Wo * p.Do, kDcWarps * 32 * kDcPx), ceil_div(p.Cout, kDcCo), p.ksplit > 1 ? p.ksplit : 1);
  if (NZ == 1 && S == 1) direct_conv_kernel<1, 1><<<grid, kDcWarps * 32, 0, st>>>(p);
  else if (NZ == 1 && S == 2) direct_conv_kernel<1, 2><<<grid, kDcWarps * 32, 0, st>>>(p);
  else if (NZ == 3 && S == 1) direct_conv_kernel<3, 1><<<grid, kDcWarps * 32, 0, st>>>(p);
  else direct_conv_kernel<3, 2><<<grid, kDcWarps * 32, 0, st>>>(p);
  if (p.ksplit > 1) {
    const long long total = (long long)p.Cout * p.Do * p.Ho * p.Wo;
    direct_conv_finish_kernel<0><<<(int)((total + 255) / 256 < 4 * kNumSMs ? (total + 255) / 256 : 4 * kNumSMs), 256, 0, st>>>(p);
  }
  return check_launch(what);
}

}  // namespace satmvs

// ---------------------------------------------------------------------------------------------
// Direct 2-D transposed convolution, k = 3, stride 2, padding 1, output_padding 1 (ConvTransReLU,
// modules/module.py:208-215), applied to every plane of a [C,D,h,w] tensor.
// One thread = 4 output channels x 4 input pixels of a row = a 2 x 8 output block; the 9 taps of every
// input pixel are used exactly once (o = 2 i - 1 + k):
//   O[2q  ][2x  ] = I[q][x] w11                    O[2q  ][2x+1] = I[q][x+1] w10 + I[q][x] w12
//   O[2q+1][2x  ] = I[q+1][x] w01 + I[q][x] w21    O[2q+1][2x+1] = I[q+1][x+1] w00 + I[q+1][x] w02 + I[q][x+1] w20 + I[q][x] w22
// Weights [Cin][Cout][3][3].  Epilogue: ReLU, then + skip tensor (indexed like the output).
// ---------------------------------------------------------------------------------------------
namespace satmvs {

struct DirectDeconv {
  const float* in;   long long in_cs;    // [Cin] channels of Dn planes of Hi x Wi; channel stride in elements
  const float* w;                        // [Cin][Cout][9]
  const float* scale; const float* shift;   // [Cout] folded BatchNorm applied before the ReLU, or null
  const float* post_add;                 // like out, or null
  float* out;        long long out_cs;   // [Cout] channels of Dn planes of 2Hi x 2Wi
  int Cin, Cout, Dn, Hi, Wi;
  int relu;
};

constexpr int kDdCo = 4, kDdPx = 4, kDdThreads = 128, kDdCiChunk = 32;

template <int kUnused>     // template only so the header can be included from several translation units
__global__ void __launch_bounds__(kDdThreads)
direct_deconv2x_kernel(const __grid_constant__ DirectDeconv a) {
  __shared__ __align__(16) float wsm[kDdCiChunk * 9 * kDdCo];       // [ci][tap][co]
  const int tid = threadIdx.x;
  const int co0 = blockIdx.y * kDdCo;
  const int npx = a.Hi * a.Wi;
  const long long g0 = ((long long)blockIdx.x * kDdThreads + tid) * kDdPx;
  const bool ok = g0 < (long long)npx * a.Dn;
  const int z = ok ? (int)(g0 / npx) : 0;
  const int p0 = ok ? (int)(g0 - (long long)z * npx) : 0;
  const int q = p0 / a.Wi, x0 = p0 - q * a.Wi;
  const bool row1 = ok && (q + 1 < a.Hi), rok = x0 + kDdPx < a.Wi;
  const float* in0 = a.in + (long long)z * npx + p0;

  float acc[kDdCo][2][2 * kDdPx];
#pragma unroll
  for (int i = 0; i < kDdCo; ++i)
#pragma unroll
    for (int r = 0; r < 2; ++r)
#pragma unroll
      for (int j = 0; j < 2 * kDdPx; ++j) acc[i][r][j] = 0.0f;

  for (int c0 = 0; c0 < a.Cin; c0 += kDdCiChunk) {
    const int nci = min(kDdCiChunk, a.Cin - c0);
    __syncthreads();
    {
      constexpr int kPer = (kDdCiChunk * 9 * kDdCo + kDdThreads - 1) / kDdThreads;   // 9
      float t[kPer];
#pragma unroll
      for (int k = 0; k < kPer; ++k) {
        const int e = tid + k * kDdThreads;                 // e = (ci * kDdCo + co) * 9 + tap : contiguous in global per ci
        const int ci = e / (kDdCo * 9), rr = e - ci * (kDdCo * 9);
        const int co = rr / 9;
        t[k] = (ci < nci && co0 + co < a.Cout) ? __ldg(a.w + ((long long)(c0 + ci) * a.Cout + co0) * 9 + rr) : 0.0f;
      }
#pragma unroll
      for (int k = 0; k < kPer; ++k) {
        const int e = tid + k * kDdThreads;
        const int ci = e / (kDdCo * 9), rr = e - ci * (kDdCo * 9);
        const int co = rr / 9, tp = rr - co * 9;
        if (ci < kDdCiChunk) wsm[(ci * 9 + tp) * kDdCo + co] = t[k];
      }
    }
    __syncthreads();
    // rows of input channel c (two rows x 5 values); the next channel's loads are issued before this channel's FMAs
    auto load_rows = [&](int c, float (&r0)[kDdPx + 1], float (&r1)[kDdPx + 1]) {
      const float* rp = in0 + (long long)(c0 + c) * a.in_cs;
      float4 m = make_float4(0.f, 0.f, 0.f, 0.f), n = m;
      float e0 = 0.f, e1 = 0.f;
      if (ok) {
        m = __ldg(reinterpret_cast<const float4*>(rp));
        if (rok) e0 = __ldg(rp + kDdPx);
        if (row1) {
          n = __ldg(reinterpret_cast<const float4*>(rp + a.Wi));
          if (rok) e1 = __ldg(rp + a.Wi + kDdPx);
        }
      }
      r0[0] = m.x; r0[1] = m.y; r0[2] = m.z; r0[3] = m.w; r0[4] = e0;
      r1[0] = n.x; r1[1] = n.y; r1[2] = n.z; r1[3] = n.w; r1[4] = e1;
    };
    float ra0[kDdPx + 1], ra1[kDdPx + 1], rb0[kDdPx + 1], rb1[kDdPx + 1];
    load_rows(0, ra0, ra1);
    auto fma_rows = [&](int c, const float (&r0)[kDdPx + 1], const float (&r1)[kDdPx + 1]) {
      float wv[9][kDdCo];
#pragma unroll
      for (int t = 0; t < 9; ++t) {
        const float4 w4 = *reinterpret_cast<const float4*>(&wsm[(c * 9 + t) * kDdCo]);
        wv[t][0] = w4.x; wv[t][1] = w4.y; wv[t][2] = w4.z; wv[t][3] = w4.w;
      }
#pragma unroll
      for (int i = 0; i < kDdCo; ++i)
#pragma unroll
        for (int j = 0; j < kDdPx; ++j) {
          acc[i][0][2 * j] = fmaf(r0[j], wv[4][i], acc[i][0][2 * j]);
          acc[i][0][2 * j + 1] = fmaf(r0[j + 1], wv[3][i], fmaf(r0[j], wv[5][i], acc[i][0][2 * j + 1]));
          acc[i][1][2 * j] = fmaf(r1[j], wv[1][i], fmaf(r0[j], wv[7][i], acc[i][1][2 * j]));
          acc[i][1][2 * j + 1] = fmaf(r1[j + 1], wv[0][i], fmaf(r1[j], wv[2][i],
                                 fmaf(r0[j + 1], wv[6][i], fmaf(r0[j], wv[8][i], acc[i][1][2 * j + 1]))));
        }
    };
#pragma unroll 1
    for (int c = 0; c < nci; c += 2) {
      if (c + 1 < nci) load_rows(c + 1, rb0, rb1);
      fma_rows(c, ra0, ra1);
      if (c + 2 < nci) load_rows(c + 2, ra0, ra1);
      if (c + 1 < nci) fma_rows(c + 1, rb0, rb1);
    }
  }
  if (!ok) return;
  const int Wo = 2 * a.Wi;
  const long long oplane = 4LL * npx;
#pragma unroll
  for (int i = 0; i < kDdCo; ++i) {
    if (co0 + i >= a.Cout) break;
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      const long long idx = (long long)(co0 + i) * a.out_cs + (long long)z * oplane + (long long)(2 * q + r) * Wo + 2 * x0;
      float v[8];
      const float sc = a.scale ? __ldg(a.scale + co0 + i) : 1.0f, sh = a.shift ? __ldg(a.shift + co0 + i) : 0.0f;
#pragma unroll
      for (int j = 0; j < 8; ++j) { const float t = fmaf(acc[i][r][j], sc, sh); v[j] = a.relu ? fmaxf(t, 0.0f) : t; }
      if (a.post_add) {
        const float4 p0v = __ldg(reinterpret_cast<const float4*>(a.post_add + idx));
        const float4 p1v = __ldg(reinterpret_cast<const float4*>(a.post_add + idx + 4));
        v[0] += p0v.x; v[1] += p0v.y; v[2] += p0v.z; v[3] += p0v.w; v[4] += p1v.x; v[5] += p1v.y; v[6] += p1v.z; v[7] += p1v.w;
      }
      *reinterpret_cast<float4*>(a.out + idx) = make_float4(v[0], v[1], v[2], v[3]);
      *reinterpret_cast<float4*>(a.out + idx + 4) = make_float4(v[4], v[5], v[6], v[7]);
    }
  }
}

inline bool direct_deconv_supported(const DirectDeconv& p) {
  return p.Wi % 4 == 0 && reinterpret_cast<uintptr_t>(p.in) % 16 == 0 && reinterpret_cast<uintptr_t>(p.out) % 16 == 0 &&
         (p.post_add == nullptr || reinterpret_cast<uintptr_t>(p.post_add) % 16 == 0) && p.in_cs % 4 == 0 && p.out_cs % 4 == 0;
}

inline int direct_deconv_launch(const DirectDeconv& p, cudaStream_t st, const char* what) {
  dim3 grid(ceil_div((long long)p.Hi * p.Wi * p.Dn, kDdThreads * kDdPx), ceil_div(p.Cout, kDdCo), 1);
  direct_deconv2x_kernel<0><<<grid, kDdThreads, 0, st>>>(p);
  return check_launch(what);
}

}  // namespace satmvs

// ---------------------------------------------------------------------------------------------
// 3-D transposed convolution, k 3, stride 2, padding 1, output_padding 1 (CostRegNet's Deconv3d blocks, modules/module.py:369-410,
// and the data gradient of its stride-2 convs): the 2-D scheme above with the plane axis added.  One thread = 2 output channels
// x 4 input voxels of a row = a 2 x 2 x 8 output block; per axis an even output uses tap 1 of the same input index, an odd
// output tap 0 of the next index and tap 2 of the same one, so no multiplication touches a structural zero.
// Weight element (ci, co, tap) at w[ci * w_ci + co * w_co + tap].  Epilogue: scale / shift (folded BatchNorm), ReLU, + skip.
// ---------------------------------------------------------------------------------------------
namespace satmvs {

struct DirectDeconv3d {
  const float* in;        // [Cin][Di][Hi][Wi]
  const float* w; long long w_ci, w_co;
  const float* scale; const float* shift; const float* post_add;
  float* out;             // [Cout][2Di][2Hi][2Wi]
  int Cin, Cout, Di, Hi, Wi;
  int relu;
};

constexpr int kD3Co = 2, kD3Px = 4, kD3Threads = 128, kD3CiChunk = 16;

template <int kUnused>
__global__ void __launch_bounds__(kD3Threads, 3)
direct_deconv3d_kernel(const __grid_constant__ DirectDeconv3d a) {
  __shared__ __align__(8) float wsm[kD3CiChunk * 27 * kD3Co];       // [ci][tap][co]
  const int tid = threadIdx.x;
  const int co0 = blockIdx.y * kD3Co;
  const int npx = a.Hi * a.Wi;
  const long long g0 = ((long long)blockIdx.x * kD3Threads + tid) * kD3Px;
  const bool ok = g0 < (long long)npx * a.Di;
  const int z = ok ? (int)(g0 / npx) : 0;
  const int p0 = ok ? (int)(g0 - (long long)z * npx) : 0;
  const int q = p0 / a.Wi, x0 = p0 - q * a.Wi;
  const bool zn = ok && (z + 1 < a.Di), yn = ok && (q + 1 < a.Hi), xn = x0 + kD3Px < a.Wi;
  const long long in_cs = (long long)a.Di * npx;
  const float* in0 = a.in + (long long)z * npx + p0;

  float acc[kD3Co][2][2][2 * kD3Px];
#pragma unroll
  for (int i = 0; i < kD3Co; ++i)
#pragma unroll
    for (int pz = 0; pz < 2; ++pz)
#pragma unroll
      for (int py = 0; py < 2; ++py)
#pragma unroll
        for (int j = 0; j < 2 * kD3Px; ++j) acc[i][pz][py][j] = 0.0f;

  for (int c0 = 0; c0 < a.Cin; c0 += kD3CiChunk) {
    const int nci = min(kD3CiChunk, a.Cin - c0);
    __syncthreads();
    for (int e = tid; e < kD3CiChunk * 27 * kD3Co; e += kD3Threads) {
      const int ci = e / (27 * kD3Co), rr = e - ci * (27 * kD3Co), tp = rr / kD3Co, co = rr - tp * kD3Co;
      wsm[e] = (ci < nci && co0 + co < a.Cout) ? __ldg(a.w + (long long)(c0 + ci) * a.w_ci + (long long)(co0 + co) * a.w_co + tp) : 0.0f;
    }
    __syncthreads();
#pragma unroll 1
    for (int c = 0; c < nci; ++c) {
      // I[dz][dy][0..4]: the voxel rows (z + dz, q + dy), columns x0 .. x0 + 4 (zero beyond the tensor)
      float I[2][2][kD3Px + 1];
      const float* rp = in0 + (long long)(c0 + c) * in_cs;
#pragma unroll
      for (int dz = 0; dz < 2; ++dz)
#pragma unroll
        for (int dy = 0; dy < 2; ++dy) {
          float4 m = make_float4(0.f, 0.f, 0.f, 0.f);
          float e = 0.f;
          if (ok && (dz == 0 || zn) && (dy == 0 || yn)) {
            const float* r = rp + (long long)dz * npx + dy * a.Wi;
            m = __ldg(reinterpret_cast<const float4*>(r));
            if (xn) e = __ldg(r + kD3Px);
          }
          I[dz][dy][0] = m.x; I[dz][dy][1] = m.y; I[dz][dy][2] = m.z; I[dz][dy][3] = m.w; I[dz][dy][4] = e;
        }
      const float* wc = wsm + c * 27 * kD3Co;
      // parity p of an output coordinate: terms (tap k, input offset d): p = 0 -> (1, 0); p = 1 -> (0, 1), (2, 0)
#pragma unroll
      for (int pz = 0; pz < 2; ++pz)
#pragma unroll
        for (int tz = 0; tz <= pz; ++tz) {
          const int kz = pz == 0 ? 1 : (tz == 0 ? 0 : 2), dz = (pz == 1 && tz == 0) ? 1 : 0;
#pragma unroll
          for (int py = 0; py < 2; ++py)
#pragma unroll
            for (int ty = 0; ty <= py; ++ty) {
              const int ky = py == 0 ? 1 : (ty == 0 ? 0 : 2), dy = (py == 1 && ty == 0) ? 1 : 0;
#pragma unroll
              for (int px = 0; px < 2; ++px)
#pragma unroll
                for (int tx = 0; tx <= px; ++tx) {
                  const int kx = px == 0 ? 1 : (tx == 0 ? 0 : 2), dx = (px == 1 && tx == 0) ? 1 : 0;
                  const int tap = (kz * 3 + ky) * 3 + kx;
                  const float2 w2 = *reinterpret_cast<const float2*>(wc + tap * kD3Co);
                  const float w0 = w2.x, w1 = w2.y;
#pragma unroll
                  for (int j = 0; j < kD3Px; ++j) {
                    const float v = I[dz][dy][j + dx];
                    acc[0][pz][py][2 * j + px] = fmaf(v, w0, acc[0][pz][py][2 * j + px]);
                    acc[1][pz][py][2 * j + px] = fmaf(v, w1, acc[1][pz][py][2 * j + px]);
                  }
                }
            }
        }
    }
  }
  if (!ok) return;
  const int Wo = 2 * a.Wi, Ho = 2 * a.Hi;
  const long long oplane = (long long)Ho * Wo, out_cs = 2LL * a.Di * oplane;
#pragma unroll
  for (int i = 0; i < kD3Co; ++i) {
    if (co0 + i >= a.Cout) break;
    const float sc = a.scale ? __ldg(a.scale + co0 + i) : 1.0f, sh = a.shift ? __ldg(a.shift + co0 + i) : 0.0f;
#pragma unroll
    for (int pz = 0; pz < 2; ++pz)
#pragma unroll
      for (int py = 0; py < 2; ++py) {
        const long long idx = (long long)(co0 + i) * out_cs + (long long)(2 * z + pz) * oplane + (long long)(2 * q + py) * Wo + 2 * x0;
        float v[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          v[j] = acc[i][pz][py][j] * sc + sh;
          if (a.relu) v[j] = fmaxf(v[j], 0.0f);
        }
        if (a.post_add) {
          const float4 p0v = __ldg(reinterpret_cast<const float4*>(a.post_add + idx));
          const float4 p1v = __ldg(reinterpret_cast<const float4*>(a.post_add + idx + 4));
          v[0] += p0v.x; v[1] += p0v.y; v[2] += p0v.z; v[3] += p0v.w; v[4] += p1v.x; v[5] += p1v.y; v[6] += p1v.z; v[7] += p1v.w;
        }
        *reinterpret_cast<float4*>(a.out + idx) = make_float4(v[0], v[1], v[2], v[3]);
        *reinterpret_cast<float4*>(a.out + idx + 4) = make_float4(v[4], v[5], v[6], v[7]);
      }
  }
}

inline bool direct_deconv3d_supported(const DirectDeconv3d& p) {
  return p.Wi % 4 == 0 && reinterpret_cast<uintptr_t>(p.in) % 16 == 0 && reinterpret_cast<uintptr_t>(p.out) % 16 == 0 &&
         (p.post_add == nullptr || reinterpret_cast<uintptr_t>(p.post_add) % 16 == 0);
}

inline int direct_deconv3d_launch(const DirectDeconv3d& p, cudaStream_t st, const char* what) {
  dim3 grid(ceil_div((long long)p.Di * p.Hi * p.Wi, kD3Threads * kD3Px), ceil_div(p.Cout, kD3Co), 1);
  direct_deconv3d_kernel<0><<<grid, kD3Threads, 0, st>>>(p);
  return check_launch(what);
}

}  // namespace satmvs

// ---------------------------------------------------------------------------------------------
// 3x3x3 (NZ 3) or per-plane 3x3 (NZ 1) stride-1 convolution to ONE output channel (CostRegNet's `prob` head, modules/module.py:566;
// RED's final upconv2d, :610): direct_conv_kernel computes
// 8 output channels per thread, 7 of them on zero filters here.  One thread = 8 consecutive output voxels of a row x 1 channel.
// ---------------------------------------------------------------------------------------------
namespace satmvs {

constexpr int kC1Px = 8, kC1Threads = 128, kC1MaxCin = 64;

template <int NZ>
__global__ void __launch_bounds__(kC1Threads)
direct_conv_c1_kernel(const __grid_constant__ DirectConv a) {
  constexpr int TAPS = NZ * 9;
  __shared__ float wsm[kC1MaxCin * TAPS];
  const int tid = threadIdx.x;
  for (int e = tid; e < a.Cin * TAPS; e += kC1Threads) {
    const int ci = e / TAPS, tp = e - ci * TAPS;
    wsm[e] = __ldg(a.w + (long long)ci * a.w_ci + (a.flip ? TAPS - 1 - tp : tp));
  }
  __syncthreads();
  const int npx = a.Hi * a.Wi;
  const long long g0 = ((long long)blockIdx.x * kC1Threads + tid) * kC1Px;
  if (g0 >= (long long)npx * a.Di) return;
  const int oz = (int)(g0 / npx), p0 = (int)(g0 - (long long)oz * npx), oy = p0 / a.Wi, ox = p0 - oy * a.Wi;
  const long long in_cs = a.in_cs ? a.in_cs : (long long)a.Di * npx;
  const bool lok = ox > 0, rok = ox + kC1Px < a.Wi;
  float acc[kC1Px];
#pragma unroll
  for (int j = 0; j < kC1Px; ++j) acc[j] = 0.0f;
#pragma unroll 1
  for (int ci = 0; ci < a.Cin; ++ci) {
    const float* wc = wsm + ci * TAPS;
#pragma unroll
    for (int kz = 0; kz < NZ; ++kz) {
      const int iz = NZ == 1 ? oz : oz - 1 + kz;
      if ((unsigned)iz >= (unsigned)a.Di) continue;
#pragma unroll
      for (int ky = 0; ky < 3; ++ky) {
        const int iy = oy - 1 + ky;
        if ((unsigned)iy >= (unsigned)a.Hi) continue;
        const float* rp = a.in + ci * in_cs + ((long long)iz * a.Hi + iy) * a.Wi + ox;
        const float4 m0 = __ldg(reinterpret_cast<const float4*>(rp)), m1 = __ldg(reinterpret_cast<const float4*>(rp + 4));
        const float r[kC1Px + 2] = {lok ? __ldg(rp - 1) : 0.0f, m0.x, m0.y, m0.z, m0.w, m1.x, m1.y, m1.z, m1.w, rok ? __ldg(rp + kC1Px) : 0.0f};
        const float w0 = wc[(kz * 3 + ky) * 3], w1 = wc[(kz * 3 + ky) * 3 + 1], w2 = wc[(kz * 3 + ky) * 3 + 2];
#pragma unroll
        for (int j = 0; j < kC1Px; ++j) acc[j] = fmaf(r[j], w0, fmaf(r[j + 1], w1, fmaf(r[j + 2], w2, acc[j])));
      }
    }
  }
  const float sc = (a.scale ? __ldg(a.scale) : 1.0f) * a.acc_scale, sh = a.shift ? __ldg(a.shift) : 0.0f;
  float v[kC1Px];
#pragma unroll
  for (int j = 0; j < kC1Px; ++j) {
    v[j] = acc[j] * sc + sh;
    if (a.relu) v[j] = fmaxf(v[j], 0.0f);
  }
  float* op = a.out + (long long)oz * npx + p0;
  if (a.post_add) {
    const float4 q0 = __ldg(reinterpret_cast<const float4*>(a.post_add + (long long)oz * npx + p0));
    const float4 q1 = __ldg(reinterpret_cast<const float4*>(a.post_add + (long long)oz * npx + p0 + 4));
    v[0] += q0.x; v[1] += q0.y; v[2] += q0.z; v[3] += q0.w; v[4] += q1.x; v[5] += q1.y; v[6] += q1.z; v[7] += q1.w;
  }
  *reinterpret_cast<float4*>(op) = make_float4(v[0], v[1], v[2], v[3]);
  *reinterpret_cast<float4*>(op + 4) = make_float4(v[4], v[5], v[6], v[7]);
}

inline bool direct_conv3d_c1_supported(const DirectConv& p) {
  return p.Cout == 1 && p.Cin <= kC1MaxCin && p.Wi % kC1Px == 0 && p.Wo == p.Wi && p.Ho == p.Hi && p.Do == p.Di && p.in_cs % 4 == 0 &&
         reinterpret_cast<uintptr_t>(p.in) % 16 == 0 && reinterpret_cast<uintptr_t>(p.out) % 16 == 0 &&
         (p.post_add == nullptr || reinterpret_cast<uintptr_t>(p.post_add) % 16 == 0);
}

inline int direct_conv3d_c1_launch(const DirectConv& p, cudaStream_t st, const char* what, int NZ = 3) {
  const int grid = ceil_div((long long)p.Di * p.Hi * p.Wi, kC1Threads * kC1Px);
  if (NZ == 3) direct_conv_c1_kernel<3><<<grid, kC1Threads, 0, st>>>(p);
  else direct_conv_c1_kernel<1><<<grid, kC1Threads, 0, st>>>(p);
  return check_launch(what);
}

}  // namespace satmvs

// ---------------------------------------------------------------------------------------------
// 1x1 convolution (FeatureNet's output heads, modules/module.py:469-483): out[co][p] = sum_ci w[co][ci] in[ci][p].  Memory-bound;
// one thread = 8 output channels x 4 consecutive positions, the filters in shared memory.
// ---------------------------------------------------------------------------------------------
namespace satmvs {

struct Conv1x1 {           // in [Cin][n], n % 4 == 0; out [Cout][n], or per view [n / view_n][Cout][view_n] when view_n > 0 (view_n % 4 == 0)
  const float* in; const float* w; float* out; long long n; int Cin, Cout; long long view_n;
};

constexpr int kP1Co = 8, kP1MaxCin = 64;

template <int kUnused>
__global__ void __launch_bounds__(256) conv1x1_kernel(const Conv1x1 a) {
  __shared__ float wsm[kP1MaxCin * kP1Co];          // [ci][co]
  const int co0 = blockIdx.y * kP1Co;
  for (int e = threadIdx.x; e < a.Cin * kP1Co; e += blockDim.x) {
    const int ci = e / kP1Co, co = e - ci * kP1Co;
    wsm[e] = co0 + co < a.Cout ? __ldg(a.w + (long long)(co0 + co) * a.Cin + ci) : 0.0f;
  }
  __syncthreads();
  const long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (i >= a.n) return;
  float4 acc[kP1Co];
#pragma unroll
  for (int c = 0; c < kP1Co; ++c) acc[c] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 4
  for (int ci = 0; ci < a.Cin; ++ci) {
    const float4 v = __ldg(reinterpret_cast<const float4*>(a.in + (long long)ci * a.n + i));
    const float4 w0 = *reinterpret_cast<const float4*>(&wsm[ci * kP1Co]), w1 = *reinterpret_cast<const float4*>(&wsm[ci * kP1Co + 4]);
    const float wv[kP1Co] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
#pragma unroll
    for (int c = 0; c < kP1Co; ++c) {
      acc[c].x = fmaf(v.x, wv[c], acc[c].x); acc[c].y = fmaf(v.y, wv[c], acc[c].y);
      acc[c].z = fmaf(v.z, wv[c], acc[c].z); acc[c].w = fmaf(v.w, wv[c], acc[c].w);
    }
  }
  long long obase = i, ocs = a.n;
  if (a.view_n > 0) { const long long v = i / a.view_n; obase = v * a.Cout * a.view_n + (i - v * a.view_n); ocs = a.view_n; }
#pragma unroll
  for (int c = 0; c < kP1Co; ++c)
    if (co0 + c < a.Cout) *reinterpret_cast<float4*>(a.out + (long long)(co0 + c) * ocs + obase) = acc[c];
}

inline bool conv1x1_supported(const Conv1x1& p) {
  return p.Cin <= kP1MaxCin && p.n % 4 == 0 && p.view_n % 4 == 0 && reinterpret_cast<uintptr_t>(p.in) % 16 == 0 &&
         reinterpret_cast<uintptr_t>(p.out) % 16 == 0;
}

inline int conv1x1_launch(const Conv1x1& p, cudaStream_t st, const char* what) {
  conv1x1_kernel<0><<<dim3(ceil_div(p.n / 4, 256), ceil_div(p.Cout, kP1Co)), 256, 0, st>>>(p);
  return check_launch(what);
}

}  // namespace satmvs

// ---------------------------------------------------------------------------------------------
// 5x5 stride-2 convolution, padding 2, per plane (FeatureNet conv1.0 / conv2.0, modules/module.py:456-466): one thread = 8 output
// channels x 4 consecutive output pixels; the five input rows of a channel (11 values each) sit in registers, weights broadcast
// from shared memory as (co, co+1) pairs, accumulation in FFMA2.  Epilogue: scale / shift (folded BatchNorm) + ReLU.
// ---------------------------------------------------------------------------------------------
namespace satmvs {

struct DirectConv5 {
  const float* in; long long in_cs;      // [Cin] channels of N planes of Hi x Wi
  const float* w;                        // [Cout][Cin][25]
  const float* scale; const float* shift;
  float* out; long long out_cs;          // [Cout] channels of N planes of Hi/2 x Wi/2
  int Cin, Cout, N, Hi, Wi, relu;
};

constexpr int kC5Co = 8, kC5Px = 4, kC5Threads = 128, kC5CiChunk = 8;

template <int kUnused>
__global__ void __launch_bounds__(kC5Threads, 3)
direct_conv5x5s2_kernel(const __grid_constant__ DirectConv5 a) {
  __shared__ __align__(16) float wsm[kC5CiChunk * 25 * kC5Co];      // [ci][tap][co]
  const int tid = threadIdx.x;
  const int co0 = blockIdx.y * kC5Co;
  const int Ho = a.Hi >> 1, Wo = a.Wi >> 1, npx = Ho * Wo;
  const long long g0 = ((long long)blockIdx.x * kC5Threads + tid) * kC5Px;
  const bool ok = g0 < (long long)npx * a.N;
  const int z = ok ? (int)(g0 / npx) : 0;
  const int p0 = ok ? (int)(g0 - (long long)z * npx) : 0;
  const int oy = p0 / Wo, ox = p0 - oy * Wo;
  const int iy0 = 2 * oy - 2, ix0 = 2 * ox;                        // first input row; first ALIGNED input column (tap 2 of pixel 0)
  const bool lok = ix0 >= 4, rok = ix0 + 8 < a.Wi;
  const float* in0 = a.in + (long long)z * a.Hi * a.Wi + (long long)iy0 * a.Wi + ix0;

  u64 acc[kC5Co / 2][kC5Px];
#pragma unroll
  for (int i = 0; i < kC5Co / 2; ++i)
#pragma unroll
    for (int j = 0; j < kC5Px; ++j) acc[i][j] = 0ULL;

  for (int c0 = 0; c0 < a.Cin; c0 += kC5CiChunk) {
    const int nci = min(kC5CiChunk, a.Cin - c0);
    __syncthreads();
    for (int e = tid; e < kC5CiChunk * 25 * kC5Co; e += kC5Threads) {
      const int ci = e / (25 * kC5Co), rr = e - ci * (25 * kC5Co), tp = rr / kC5Co, co = rr - tp * kC5Co;
      wsm[e] = (ci < nci && co0 + co < a.Cout) ? __ldg(a.w + ((long long)(co0 + co) * a.Cin + c0 + ci) * 25 + tp) : 0.0f;
    }
    __syncthreads();
#pragma unroll 1
    for (int c = 0; c < nci; ++c) {
      const float* cp = in0 + (long long)(c0 + c) * a.in_cs;
#pragma unroll
      for (int ky = 0; ky < 5; ++ky) {
        // r[0..10] = in[iy0 + ky][ix0 - 2 .. ix0 + 8]  (zero outside the plane)
        float r[11];
        const bool yok = ok && (unsigned)(iy0 + ky) < (unsigned)a.Hi;
        const float* rp = cp + (long long)ky * a.Wi;
        float4 m0 = make_float4(0.f, 0.f, 0.f, 0.f), m1 = m0;
        float2 l2 = make_float2(0.f, 0.f);
        float e8 = 0.f;
        if (yok) {
          m0 = __ldg(reinterpret_cast<const float4*>(rp));
          m1 = __ldg(reinterpret_cast<const float4*>(rp + 4));
          if (lok) l2 = __ldg(reinterpret_cast<const float2*>(rp - 2));
          if (rok) e8 = __ldg(rp + 8);
        }
        r[0] = l2.x; r[1] = l2.y; r[2] = m0.x; r[3] = m0.y; r[4] = m0.z; r[5] = m0.w; r[6] = m1.x; r[7] = m1.y; r[8] = m1.z; r[9] = m1.w; r[10] = e8;
        u64 rr2[11];
#pragma unroll
        for (int k = 0; k < 11; ++k) rr2[k] = pk(r[k], r[k]);
#pragma unroll
        for (int kx = 0; kx < 5; ++kx) {
          const float* wp = &wsm[(c * 25 + ky * 5 + kx) * kC5Co];
          const ulonglong2 w0 = *reinterpret_cast<const ulonglong2*>(wp), w1 = *reinterpret_cast<const ulonglong2*>(wp + 4);
          const u64 wv[kC5Co / 2] = {w0.x, w0.y, w1.x, w1.y};
#pragma unroll
          for (int i = 0; i < kC5Co / 2; ++i)
#pragma unroll
            for (int j = 0; j < kC5Px; ++j) acc[i][j] = ffma2(wv[i], rr2[2 * j + kx], acc[i][j]);
        }
      }
    }
  }
  if (!ok) return;
#pragma unroll
  for (int i = 0; i < kC5Co; ++i) {
    const int co = co0 + i;
    if (co >= a.Cout) break;
    const float sc = a.scale ? __ldg(a.scale + co) : 1.0f, sh = a.shift ? __ldg(a.shift + co) : 0.0f;
    float v[kC5Px];
#pragma unroll
    for (int j = 0; j < kC5Px; ++j) {
      float alo, ahi;
      upk(acc[i >> 1][j], alo, ahi);
      v[j] = fmaf((i & 1) ? ahi : alo, sc, sh);
      if (a.relu) v[j] = fmaxf(v[j], 0.0f);
    }
    *reinterpret_cast<float4*>(a.out + (long long)co * a.out_cs + (long long)z * npx + p0) = make_float4(v[0], v[1], v[2], v[3]);
  }
}

inline bool direct_conv5_supported(const DirectConv5& p) {
  return p.Wi % 8 == 0 && p.Hi % 2 == 0 && p.in_cs % 4 == 0 && p.out_cs % 4 == 0 &&
         reinterpret_cast<uintptr_t>(p.in) % 16 == 0 && reinterpret_cast<uintptr_t>(p.out) % 16 == 0;
}

inline int direct_conv5_launch(const DirectConv5& p, cudaStream_t st, const char* what) {
  dim3 grid(ceil_div((long long)p.N * (p.Hi / 2) * (p.Wi / 2), kC5Threads * kC5Px), ceil_div(p.Cout, kC5Co), 1);
  direct_conv5x5s2_kernel<0><<<grid, kC5Threads, 0, st>>>(p);
  return check_launch(what);
}

}  // namespace satmvs
