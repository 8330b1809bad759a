// red_cluster.cuh — the RED recurrence (phase B of red.cu) as ONE launch of four thread-block clusters.
//
// Reference: the plane loop of RED_Regularization.forward (modules/module.py:625-644) through
// ConvGRUCell2.forward (:27-58).  The four GRUs of the UNet never talk to each other inside the loop
// (GRU_l sees encoder l's output and its own previous state; the decoder runs afterwards, batched), so
// each level is an independent recurrence over the D planes.  One 16-CTA cluster owns one level for the
// whole sweep over depth:
//   * the hidden-state halves of the level's gate / output conv weights are staged into shared memory
//     ONCE (CTA (strip, cg) keeps the 8 hidden channels [8cg, 8cg+8): 16 gate + 8 output filters);
//   * a CTA owns a strip of rows x 8 hidden channels; per plane it stages the state tile (all input
//     channels, strip + halo rows) from L2 into shared memory and runs the two convolutions from there
//     (thread = 8 output channels x 4 pixels x 8 / 4 input channels in FFMA2, partial sums over the
//     input-channel chunks exchanged through shared memory in a fixed order);
//   * the GroupNorm sums of a level are exchanged through distributed shared memory, and the four
//     dependencies of a plane (gate sums, r*h halo, output sums, new-state halo) are hardware cluster
//     barriers (barrier.cluster, ~0.2 us) instead of kernel boundaries (~7 us each in the PDL chain).
// Pointwise steps (module.py:40-43, :54-57) run on the values the CTA already holds in shared memory.
// Peer-written tensors (state history, r*h) travel through global memory / L2 and are read with
// ld.global.cg after the cluster barrier (release / acquire at cluster scope).
#pragma once
#include <cooperative_groups.h>
#include <cstdlib>
#include "common.cuh"
#include "packed.cuh"

namespace satmvs {

constexpr int kClSize = 16;              // CTAs per cluster (non-portable size, opt-in)
constexpr int kClThreads = 576;          // 18 warps: 576 = 2 co-groups x 288 pixel quads at 96x192, every level
constexpr int kClWarps = kClThreads / 32;
constexpr int kClK = 8;                  // hidden channels per CTA

struct ClLevel {
  float* s; long long s_cs;              // state history [ch][D+1][px]: channel stride; slot stride = px
  const float* gx; long long g_cs;       // gate x-halves [2ch][D][px]
  const float* ox; long long o_cs;       // output x-halves [ch][D][px]
  float* rh;                             // [ch][px]
  const float* gate_w; const float* out_w; long long w_co;   // hidden-state halves of the conv weights
  const float *rn_w, *rn_b, *un_w, *un_b, *on_w, *on_b;
  double inv_n;
  int ch, h, w, px;
  int CG, R;                             // channel groups (ch / 8), rows per strip (ceil(h / (16 / CG)))
};
struct ClArgs { ClLevel l[4]; int D; };

struct ClSmemPlan { int wg, wo, tile, part, keep, total_floats; };
__host__ __device__ inline ClSmemPlan cl_smem_plan(int ch, int w, int R) {
  ClSmemPlan p;
  const int npx = R * w;
  p.wg = ch * 9 * 16;
  p.wo = ch * 9 * 8;
  p.tile = ch * (R + 2) * (w + 8);
  p.part = (ch / 4) * 8 * npx;           // gates: 2 co-groups x ch/8 chunks; output: ch/4 chunks
  p.keep = 8 * npx;
  p.total_floats = p.wg + p.wo + p.tile + p.part + 2 * p.keep;
  return p;
}
constexpr int kClFixedSmemBytes = (4 * kClWarps + 2 * 4 + 4) * 8 + 3 * 8 * 2 * 4;

// Pointwise transcendental functions on the SFU (ex2.approx + rcp.approx: ~2 ulp each).  Inside the recurrence every
// SM of a cluster owns 1/16 of a level's elements, so the ~40-instruction libm forms would cost as much as the convs.
__device__ __forceinline__ float cl_sigmoid(float x) { return __fdividef(1.0f, 1.0f + __expf(-x)); }
__device__ __forceinline__ float cl_tanh(float x) { return fmaf(2.0f, __fdividef(1.0f, 1.0f + __expf(-2.0f * x)), -1.0f); }

// n / d for n, d < 65536 as one multiply-high (exact: n * d < 2^32); d == 1 handled by a select
struct ClDiv { unsigned d, m; };
__device__ __forceinline__ ClDiv cl_mkdiv(int d) { ClDiv r; r.d = (unsigned)d; r.m = d > 1 ? 0xFFFFFFFFu / (unsigned)d + 1u : 0u; return r; }
__device__ __forceinline__ int cl_div(int n, const ClDiv& dv) { return dv.d <= 1u ? n : (int)__umulhi((unsigned)n, dv.m); }

// One unit: CI input channels x 9 taps x 8 output channels x 4 pixels, accumulated as (co, co+1) pairs.
template <int CI, int COT>
__device__ __forceinline__ void cl_conv_unit(const float* __restrict__ tile_px, int ci_stride, int pitch,
                                             const float* __restrict__ wsm, u64 (&acc2)[4][4]) {
#pragma unroll 2
  for (int c = 0; c < CI; ++c) {
    const float* tp = tile_px + c * ci_stride;
    u64 rr[3][6];
#pragma unroll
    for (int dy = 0; dy < 3; ++dy) {
      const float* rp = tp + dy * pitch;
      const float4 m = *reinterpret_cast<const float4*>(rp);
      const float lft = rp[-1], rgt = rp[4];
      rr[dy][0] = pk(lft, lft); rr[dy][1] = pk(m.x, m.x); rr[dy][2] = pk(m.y, m.y);
      rr[dy][3] = pk(m.z, m.z); rr[dy][4] = pk(m.w, m.w); rr[dy][5] = pk(rgt, rgt);
    }
    const float* wp = wsm + c * 9 * COT;
#pragma unroll
    for (int t = 0; t < 9; ++t) {
      const ulonglong2 w0 = *reinterpret_cast<const ulonglong2*>(wp + t * COT);
      const ulonglong2 w1 = *reinterpret_cast<const ulonglong2*>(wp + t * COT + 4);
      const u64 wv[4] = {w0.x, w0.y, w1.x, w1.y};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc2[i][j] = ffma2(wv[i], rr[t / 3][j + t % 3], acc2[i][j]);
    }
  }
}

__global__ void __launch_bounds__(kClThreads, 1)
red_cluster_kernel(const __grid_constant__ ClArgs a) {
  namespace cg = cooperative_groups;
  cg::cluster_group cluster = cg::this_cluster();
  extern __shared__ __align__(16) float cl_smem[];
  __shared__ double red[4][kClWarps];
  __shared__ double stat_out[2][4];         // this CTA's partial sums: [0] = gates (r sum, r sq, u sum, u sq), [1] = output
  __shared__ float coef[3][kClK][2];        // GroupNorm scale / shift of r, u, o for this CTA's 8 channels

  const ClLevel& L = a.l[blockIdx.x / kClSize];
  const int rank = (int)cluster.block_rank();
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int CG = L.CG, R = L.R, ch = L.ch, w = L.w;
  const int strip = rank / CG, cgp = rank - strip * CG;
  const int y0 = strip * R;
  int nrows = L.h - y0; nrows = nrows > R ? R : nrows; nrows = nrows < 0 ? 0 : nrows;
  const int pitch = w + 8, w4 = w >> 2;
  const int NPX = R * w;                    // channel stride of the partial / keep buffers
  const int npx4 = nrows * w4;              // pixel quads of this strip
  const int KS = ch >> 3, KS2 = ch >> 2;    // input-channel chunks: gates 8 channels, output 4 channels
  const int ci_stride = (R + 2) * pitch;
  const int c_own = cgp * kClK;             // first hidden channel of this CTA
  const ClDiv dv_npx4 = cl_mkdiv(npx4), dv_w4 = cl_mkdiv(w4), dv_perci = cl_mkdiv((nrows + 2) * w4);
  const int ksh = __ffs(KS) - 1;            // KS is a power of two (ch 8 / 16 / 32 / 64)

  const ClSmemPlan sp = cl_smem_plan(ch, w, R);
  float* wg = cl_smem;                      // [ci][tap][16]: r filters of the 8 channels, then u filters
  float* wo = wg + sp.wg;                   // [ci][tap][8]
  float* tile = wo + sp.wo;                 // [ci][R+2][pitch], pixel x at column 4 + x; borders stay zero
  float* part = tile + sp.tile;             // [chunk][8][NPX]
  float* keepA = part + sp.part;            // u * h
  float* keepB = keepA + sp.keep;           // 1 - u

  // ---- once: weights resident, tile borders zero ----
  for (int i = tid; i < ch * 9 * 16; i += kClThreads) {
    const int co = i & 15, r = i >> 4;      // r = ci * 9 + tap
    const int cglob = (co >> 3) * ch + c_own + (co & 7);
    wg[i] = __ldg(L.gate_w + (long long)cglob * L.w_co + r);
  }
  for (int i = tid; i < ch * 9 * 8; i += kClThreads) {
    const int co = i & 7, r = i >> 3;
    wo[i] = __ldg(L.out_w + (long long)(c_own + co) * L.w_co + r);
  }
  for (int i = tid; i < sp.tile; i += kClThreads) tile[i] = 0.0f;
  __syncthreads();

  // stage src[ci][gy][x] (channel stride cs, plane already selected) for rows y0-1 .. y0+nrows into the tile
  auto stage_tile = [&](const float* src, long long cs) {
    const int per_ci = (nrows + 2) * w4, items = ch * per_ci;
    for (int i0 = tid; i0 < items; i0 += 4 * kClThreads) {
      float4 v[4]; int dst[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int i = i0 + u * kClThreads;
        dst[u] = -1;
        if (i < items) {
          const int ci = cl_div(i, dv_perci), r = i - ci * per_ci;
          const int lr = cl_div(r, dv_w4), x4 = r - lr * w4;
          const int gy = y0 - 1 + lr;
          if (gy >= 0 && gy < L.h) {
            v[u] = __ldcg(reinterpret_cast<const float4*>(src + ci * cs + (long long)gy * w) + x4);
            dst[u] = ci * ci_stride + lr * pitch + 4 + 4 * x4;
          }
        }
      }
#pragma unroll
      for (int u = 0; u < 4; ++u)
        if (dst[u] >= 0) *reinterpret_cast<float4*>(tile + dst[u]) = v[u];
    }
  };

  // block-wide sums of up to 4 doubles -> stat_out[which]
  auto publish_stats = [&](int which, int n, double (&v)[4]) {
#pragma unroll
    for (int k = 0; k < 4; ++k)
      if (k < n) {
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) v[k] += __shfl_xor_sync(0xffffffffu, v[k], off);
        if (lane == 0) red[k][warp] = v[k];
      }
    __syncthreads();
    if (tid < n) {
      double s = 0.0;
      for (int i = 0; i < kClWarps; ++i) s += red[tid][i];
      stat_out[which][tid] = s;
    }
  };
  // after the cluster barrier: warp k < nnorm gathers the (sum, sum of squares) of GroupNorm `first + k` from the 16
  // ranks through distributed shared memory (fixed order: butterfly over the ranks) and lanes 0..7 turn them into the
  // scale / shift of this CTA's 8 channels
  auto gather_coef = [&](int which, int nnorm, int first) {
    if (warp < nnorm) {
      const int idx = 2 * warp + (lane >> 4), rk = lane & 15;
      const double* remote = cluster.map_shared_rank(&stat_out[which][idx], rk);
      double v = *remote;
#pragma unroll
      for (int off = 8; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
      const double S = __shfl_sync(0xffffffffu, v, 0), Q = __shfl_sync(0xffffffffu, v, 16);
      if (lane < 8) {
        const int norm = first + warp;
        const float* gw = norm == 0 ? L.rn_w : (norm == 1 ? L.un_w : L.on_w);
        const float* gb = norm == 0 ? L.rn_b : (norm == 1 ? L.un_b : L.on_b);
        const double mean = S * L.inv_n;
        const float var = (float)fmax(Q * L.inv_n - mean * mean, 0.0);
        const float rstd = rsqrtf(var + 1e-5f);
        const float ca = __ldg(gw + c_own + lane) * rstd;
        coef[norm][lane][0] = ca; coef[norm][lane][1] = __ldg(gb + c_own + lane) - (float)mean * ca;
      }
    }
    __syncthreads();
  };

  for (int d = 0; d < a.D; ++d) {
    const long long plane_g = (long long)d * L.px;
    // ================= P1: gates = GX[d] + conv(h[d]) =================
    {
      const int nunits = 2 * KS * npx4;
      // accumulators of the first chunk start from the x-half (read-only data of an earlier kernel)
      int u = tid;
      u64 acc2[4][4];
      int q = 0, gk = 0;
      bool have = u < nunits;
      if (have) { gk = cl_div(u, dv_npx4); q = u - gk * npx4; }
      const int g = gk >> ksh, kc = gk - g * KS;
      const int ly = cl_div(q, dv_w4), x = (q - ly * w4) << 2;
      float4 pv[8];
#pragma unroll
      for (int co = 0; co < 8; ++co) pv[co] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (have && kc == 0) {
        const float* pre = L.gx + (long long)(g * ch + c_own) * L.g_cs + plane_g + (long long)(y0 + ly) * w + x;
#pragma unroll
        for (int co = 0; co < 8; ++co) pv[co] = __ldg(reinterpret_cast<const float4*>(pre + co * L.g_cs));
        if (d + 1 < a.D) {
#pragma unroll
          for (int co = 0; co < 8; ++co) asm volatile("prefetch.global.L2 [%0];" :: "l"(pre + co * L.g_cs + L.px));
        }
      }
      stage_tile(L.s + plane_g, L.s_cs);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        acc2[i][0] = pk(pv[2 * i].x, pv[2 * i + 1].x); acc2[i][1] = pk(pv[2 * i].y, pv[2 * i + 1].y);
        acc2[i][2] = pk(pv[2 * i].z, pv[2 * i + 1].z); acc2[i][3] = pk(pv[2 * i].w, pv[2 * i + 1].w);
      }
      __syncthreads();
      for (; u < nunits; u += kClThreads) {
        int q_ = q, gk_ = gk;
        if (u != tid) {                      // further rounds (shapes with more units than threads): plain start
          gk_ = cl_div(u, dv_npx4); q_ = u - gk_ * npx4;
          const int g2 = gk_ >> ksh, kc2 = gk_ - g2 * KS;
          const int ly2 = cl_div(q_, dv_w4), x2 = (q_ - ly2 * w4) << 2;
#pragma unroll
          for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc2[i][j] = 0ULL;
          if (kc2 == 0) {
            const float* pre = L.gx + (long long)(g2 * ch + c_own) * L.g_cs + plane_g + (long long)(y0 + ly2) * w + x2;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const float4 p0 = __ldg(reinterpret_cast<const float4*>(pre + (2 * i) * L.g_cs));
              const float4 p1 = __ldg(reinterpret_cast<const float4*>(pre + (2 * i + 1) * L.g_cs));
              acc2[i][0] = pk(p0.x, p1.x); acc2[i][1] = pk(p0.y, p1.y); acc2[i][2] = pk(p0.z, p1.z); acc2[i][3] = pk(p0.w, p1.w);
            }
          }
        }
        const int g2 = gk_ >> ksh, kc2 = gk_ - g2 * KS;
        const int ly2 = cl_div(q_, dv_w4), x2 = (q_ - ly2 * w4) << 2;
        cl_conv_unit<8, 16>(tile + (kc2 * 8) * ci_stride + (ly2 + 0) * pitch + 4 + x2, ci_stride, pitch,
                            wg + (kc2 * 8) * 9 * 16 + g2 * 8, acc2);
        float* pp = part + (long long)(gk_ * 8) * NPX + 4 * q_;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          float lo[4], hi[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) upk(acc2[i][j], lo[j], hi[j]);
          *reinterpret_cast<float4*>(pp + (2 * i) * NPX) = make_float4(lo[0], lo[1], lo[2], lo[3]);
          *reinterpret_cast<float4*>(pp + (2 * i + 1) * NPX) = make_float4(hi[0], hi[1], hi[2], hi[3]);
        }
      }
      __syncthreads();
      // reduce the chunks (fixed order), keep the gates in chunk 0 of each co-group, GroupNorm sums of r and u
      double st_rs = 0.0, st_rq = 0.0, st_us = 0.0, st_uq = 0.0;
      const int items = 16 * npx4;
      for (int o = tid; o < items; o += kClThreads) {
        const int gc = cl_div(o, dv_npx4), p4 = o - gc * npx4;
        const int g2 = gc >> 3, co = gc & 7;
        float* p0 = part + (long long)((g2 * KS) * 8 + co) * NPX + 4 * p4;
        float4 v = *reinterpret_cast<const float4*>(p0);
        for (int k = 1; k < KS; ++k) {
          const float4 t = *reinterpret_cast<const float4*>(p0 + (long long)k * 8 * NPX);
          v.x += t.x; v.y += t.y; v.z += t.z; v.w += t.w;
        }
        *reinterpret_cast<float4*>(p0) = v;
        const float s4 = (v.x + v.y) + (v.z + v.w);
        const float q4 = fmaf(v.x, v.x, v.y * v.y) + fmaf(v.z, v.z, v.w * v.w);
        if (g2) { st_us += (double)s4; st_uq += (double)q4; } else { st_rs += (double)s4; st_rq += (double)q4; }
      }
      double st[4] = {st_rs, st_rq, st_us, st_uq};
      publish_stats(0, 4, st);
    }
    cluster.sync();                                          // #1: gate sums of every CTA of the level are published
    gather_coef(0, 2, 0);
    // ================= E1: rh = sigmoid(GN_r(G_r)) * h;  keep u*h and 1-u =================
    {
      const int items = 8 * npx4;
      for (int o = tid; o < items; o += kClThreads) {
        const int co = cl_div(o, dv_npx4), p4 = o - co * npx4;
        const int ly = cl_div(p4, dv_w4), x = (p4 - ly * w4) << 2;
        const float4 gr = *reinterpret_cast<const float4*>(part + (long long)co * NPX + 4 * p4);
        const float4 gu = *reinterpret_cast<const float4*>(part + (long long)(KS * 8 + co) * NPX + 4 * p4);
        const float4 hv = *reinterpret_cast<const float4*>(tile + (c_own + co) * ci_stride + (ly + 1) * pitch + 4 + x);
        const float ra = coef[0][co][0], rb = coef[0][co][1], ua = coef[1][co][0], ub = coef[1][co][1];
        float4 rh, ka, kb;
        rh.x = cl_sigmoid(fmaf(gr.x, ra, rb)) * hv.x; rh.y = cl_sigmoid(fmaf(gr.y, ra, rb)) * hv.y;
        rh.z = cl_sigmoid(fmaf(gr.z, ra, rb)) * hv.z; rh.w = cl_sigmoid(fmaf(gr.w, ra, rb)) * hv.w;
        const float u0 = cl_sigmoid(fmaf(gu.x, ua, ub)), u1 = cl_sigmoid(fmaf(gu.y, ua, ub));
        const float u2 = cl_sigmoid(fmaf(gu.z, ua, ub)), u3 = cl_sigmoid(fmaf(gu.w, ua, ub));
        ka = make_float4(u0 * hv.x, u1 * hv.y, u2 * hv.z, u3 * hv.w);
        kb = make_float4(1.0f - u0, 1.0f - u1, 1.0f - u2, 1.0f - u3);
        *reinterpret_cast<float4*>(L.rh + (long long)(c_own + co) * L.px + (long long)(y0 + ly) * w + x) = rh;
        *reinterpret_cast<float4*>(keepA + (long long)co * NPX + 4 * p4) = ka;
        *reinterpret_cast<float4*>(keepB + (long long)co * NPX + 4 * p4) = kb;
      }
    }
    cluster.sync();                                          // #2: r*h of the whole level is in L2
    // ================= P2: O = OX[d] + conv(rh) =================
    {
      const int nunits = KS2 * npx4;
      double st[4] = {0.0, 0.0, 0.0, 0.0};
      u64 acc2[4][4];
      int u = tid;
      const bool have = u < nunits;
      int kc = 0, q = 0;
      if (have) { kc = cl_div(u, dv_npx4); q = u - kc * npx4; }
      const int ly = cl_div(q, dv_w4), x = (q - ly * w4) << 2;
      float4 pv[8];
#pragma unroll
      for (int co = 0; co < 8; ++co) pv[co] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (have && kc == 0) {
        const float* pre = L.ox + (long long)c_own * L.o_cs + plane_g + (long long)(y0 + ly) * w + x;
#pragma unroll
        for (int co = 0; co < 8; ++co) pv[co] = __ldg(reinterpret_cast<const float4*>(pre + co * L.o_cs));
        if (d + 1 < a.D) {
#pragma unroll
          for (int co = 0; co < 8; ++co) asm volatile("prefetch.global.L2 [%0];" :: "l"(pre + co * L.o_cs + L.px));
        }
      }
      stage_tile(L.rh, L.px);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        acc2[i][0] = pk(pv[2 * i].x, pv[2 * i + 1].x); acc2[i][1] = pk(pv[2 * i].y, pv[2 * i + 1].y);
        acc2[i][2] = pk(pv[2 * i].z, pv[2 * i + 1].z); acc2[i][3] = pk(pv[2 * i].w, pv[2 * i + 1].w);
      }
      __syncthreads();
      for (; u < nunits; u += kClThreads) {
        int q_ = q, kc_ = kc;
        if (u != tid) {
          kc_ = cl_div(u, dv_npx4); q_ = u - kc_ * npx4;
          const int ly2 = cl_div(q_, dv_w4), x2 = (q_ - ly2 * w4) << 2;
#pragma unroll
          for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc2[i][j] = 0ULL;
          if (kc_ == 0) {
            const float* pre = L.ox + (long long)c_own * L.o_cs + plane_g + (long long)(y0 + ly2) * w + x2;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const float4 p0 = __ldg(reinterpret_cast<const float4*>(pre + (2 * i) * L.o_cs));
              const float4 p1 = __ldg(reinterpret_cast<const float4*>(pre + (2 * i + 1) * L.o_cs));
              acc2[i][0] = pk(p0.x, p1.x); acc2[i][1] = pk(p0.y, p1.y); acc2[i][2] = pk(p0.z, p1.z); acc2[i][3] = pk(p0.w, p1.w);
            }
          }
        }
        const int ly2 = cl_div(q_, dv_w4), x2 = (q_ - ly2 * w4) << 2;
        cl_conv_unit<4, 8>(tile + (kc_ * 4) * ci_stride + ly2 * pitch + 4 + x2, ci_stride, pitch, wo + (kc_ * 4) * 9 * 8, acc2);
        float* pp = part + (long long)(kc_ * 8) * NPX + 4 * q_;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          float lo[4], hi[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) upk(acc2[i][j], lo[j], hi[j]);
          *reinterpret_cast<float4*>(pp + (2 * i) * NPX) = make_float4(lo[0], lo[1], lo[2], lo[3]);
          *reinterpret_cast<float4*>(pp + (2 * i + 1) * NPX) = make_float4(hi[0], hi[1], hi[2], hi[3]);
        }
      }
      __syncthreads();
      const int items = 8 * npx4;
      for (int o = tid; o < items; o += kClThreads) {
        const int co = cl_div(o, dv_npx4), p4 = o - co * npx4;
        float* p0 = part + (long long)co * NPX + 4 * p4;
        float4 v = *reinterpret_cast<const float4*>(p0);
        for (int k = 1; k < KS2; ++k) {
          const float4 t = *reinterpret_cast<const float4*>(p0 + (long long)k * 8 * NPX);
          v.x += t.x; v.y += t.y; v.z += t.z; v.w += t.w;
        }
        *reinterpret_cast<float4*>(p0) = v;
        const float s4 = (v.x + v.y) + (v.z + v.w);
        const float q4 = fmaf(v.x, v.x, v.y * v.y) + fmaf(v.z, v.z, v.w * v.w);
        st[0] += (double)s4; st[1] += (double)q4;
      }
      publish_stats(1, 2, st);
    }
    cluster.sync();                                          // #3: output-conv sums are published
    gather_coef(1, 1, 2);
    // ================= E2: h[d+1] = u*h + (1-u)*tanh(GN_o(O)) =================
    {
      const int items = 8 * npx4;
      float* hn = L.s + plane_g + L.px;
      for (int o = tid; o < items; o += kClThreads) {
        const int co = cl_div(o, dv_npx4), p4 = o - co * npx4;
        const int ly = cl_div(p4, dv_w4), x = (p4 - ly * w4) << 2;
        const float4 y = *reinterpret_cast<const float4*>(part + (long long)co * NPX + 4 * p4);
        const float4 ka = *reinterpret_cast<const float4*>(keepA + (long long)co * NPX + 4 * p4);
        const float4 kb = *reinterpret_cast<const float4*>(keepB + (long long)co * NPX + 4 * p4);
        const float oa = coef[2][co][0], ob = coef[2][co][1];
        float4 hv;
        hv.x = ka.x + kb.x * cl_tanh(fmaf(y.x, oa, ob)); hv.y = ka.y + kb.y * cl_tanh(fmaf(y.y, oa, ob));     // module.py:57
        hv.z = ka.z + kb.z * cl_tanh(fmaf(y.z, oa, ob)); hv.w = ka.w + kb.w * cl_tanh(fmaf(y.w, oa, ob));
        *reinterpret_cast<float4*>(hn + (long long)(c_own + co) * L.s_cs + (long long)(y0 + ly) * w + x) = hv;
      }
    }
    cluster.sync();                                          // #4: the new state of the whole level is in L2
  }
}

// Shapes the cluster kernel takes: every level's width a multiple of 4 (vector rows) and the per-CTA
// working set within the shared-memory opt-in limit.  Returns the dynamic shared-memory bytes, 0 if unsupported.
inline size_t red_cluster_smem_bytes(const ClArgs& a, int smem_optin) {
  size_t need = 0;
  for (int l = 0; l < 4; ++l) {
    const ClLevel& L = a.l[l];
    if (L.w % 4 || L.ch % 8 || L.ch / 8 > kClSize || kClSize % (L.ch / 8)) return 0;
    const size_t b = (size_t)cl_smem_plan(L.ch, L.w, L.R).total_floats * sizeof(float);
    need = b > need ? b : need;
  }
  if (need + kClFixedSmemBytes + 1024 > (size_t)smem_optin) return 0;
  return need;
}

// Launches the recurrence as 4 clusters of 16 CTAs; *launched stays false when the shape or the device does not
// take it (the caller then runs the per-plane kernel chain).
inline int red_cluster_launch(ClArgs& a, cudaStream_t st, bool* launched) {
  *launched = false;
  int dev = 0, optin = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
  for (int l = 0; l < 4; ++l) {
    ClLevel& L = a.l[l];
    L.CG = L.ch / kClK;
    const int PS = kClSize / (L.CG > 0 ? L.CG : 1);
    L.R = (L.h + PS - 1) / PS;
  }
  const size_t smem = red_cluster_smem_bytes(a, optin);
  static const bool verbose = getenv("SATMVS_RED_DEBUG") != nullptr;
  auto declined = [&](const char* why, cudaError_t e) {
    if (verbose) fprintf(stderr, "red_cluster_launch: falling back to the kernel chain: %s (%s)\n", why, cudaGetErrorString(e));
    cudaGetLastError();
    return SATMVS_OK;
  };
  if (smem == 0) return declined("shape not supported", cudaSuccess);
  cudaError_t e;
  if ((e = cudaFuncSetAttribute(red_cluster_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1)) != cudaSuccess)
    return declined("non-portable cluster size", e);
  if ((e = cudaFuncSetAttribute(red_cluster_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)) != cudaSuccess)
    return declined("dynamic shared memory", e);
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(4 * kClSize); cfg.blockDim = dim3(kClThreads); cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = kClSize; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  int nclusters = 0;
  if ((e = cudaOccupancyMaxActiveClusters(&nclusters, red_cluster_kernel, &cfg)) != cudaSuccess || nclusters < 1)
    return declined("no co-resident cluster", e);
  if ((e = cudaLaunchKernelEx(&cfg, red_cluster_kernel, a)) != cudaSuccess) return declined("launch", e);
  *launched = true;
  return SATMVS_OK;
}

}  // namespace satmvs
