// red_cluster.cuh — the RED recurrence (phase B of red.cu) as ONE launch of six thread-block clusters.
//
// Reference: the plane loop of RED_Regularization.forward (modules/module.py:625-644) through
// ConvGRUCell2.forward (:27-58).  The four GRUs of the UNet never talk to each other inside the loop
// (GRU_l sees encoder l's output and its own previous state; the decoder runs afterwards, batched), so
// each level is an independent recurrence over the D planes.  A 16-CTA cluster "A" owns one level for the whole
// sweep over depth and a cluster "B" serves two levels (4 + 2 clusters: a B200 keeps seven 16-CTA clusters resident):
//   cluster A (critical path)  r = sigmoid(GN_r(GX_r + conv(h; Wr)));  rh = r*h;
//                              o = GN_o(OX + conv(rh; Wo));  h' = u*h + (1-u)*tanh(o)          (module.py:33-57)
//   cluster B (off the path)   u = sigmoid(GN_u(GX_u + conv(h; Wu)))                            (module.py:33-43)
// B needs h' of the previous plane and A needs u only at the very end of a plane, so the two exchange one
// release/acquire flag per plane through L2 and B's third of the convolution work leaves the critical path.
//   * the hidden-state halves of the filters are staged into shared memory ONCE (CTA (strip, cg) keeps the
//     8 hidden channels [8cg, 8cg+8));
//   * a CTA owns a strip of rows x 8 hidden channels; per plane it stages the input tile (all input
//     channels, strip + halo rows) from L2 into shared memory and convolves from there (thread = 8 output
//     channels x 4 pixels x 4 input channels in FFMA2; the partial sums over the input-channel chunks are
//     exchanged through shared memory and added in a fixed order);
//   * the GroupNorm sums of a level are exchanged through distributed shared memory, and the dependencies
//     inside a plane (sums, r*h halo, new-state halo) are hardware cluster barriers (barrier.cluster,
//     ~0.2 us) instead of kernel boundaries (~7 us each in the per-plane kernel chain).
// Peer-written tensors (state history, r*h, u) travel through global memory / L2 and are read with
// ld.global.cg after the barrier / flag (release / acquire).
#pragma once
#include <cooperative_groups.h>
#include <cstdlib>
#include "common.cuh"
#include "packed.cuh"

namespace satmvs {

#ifndef SATMVS_CL_UNROLL
#define SATMVS_CL_UNROLL 4
#endif

#ifndef SATMVS_CL_UNROLL
#define SATMVS_CL_UNROLL 4
#endif
constexpr int kClUnroll = SATMVS_CL_UNROLL;   // input channels per unrolled step of the conv loop (tuning knob)
constexpr int kClSize = 16;              // CTAs per cluster (non-portable size, opt-in)
constexpr int kClThreads = 576;          // 18 warps: 576 = ch/4 chunks x 288 / (ch/8) pixel quads at 96x192, every level
constexpr int kClWarps = kClThreads / 32;
constexpr int kClK = 8;                  // hidden channels per CTA
constexpr int kClFlagStride = 32;        // ints between the two flags of a level (separate 128-byte lines)

struct ClLevel {
  float* s; long long s_cs;              // state history [ch][D+1][px]: channel stride; slot stride = px
  const float* gx; long long g_cs;       // gate x-halves [2ch][D][px]
  const float* ox; long long o_cs;       // output x-halves [ch][D][px]
  float* rh;                             // [ch][px]
  float* ub;                             // [2][ch][px] update gate of the current / next plane (clusters B -> cluster A)
  int* flags;                            // [0] = planes of h' published by A, [kClFlagStride] = CTA-shares of u published by B (16 per plane)
  const float* gate_w; const float* out_w; long long w_co;   // hidden-state halves of the conv weights
  const float *rn_w, *rn_b, *un_w, *un_b, *on_w, *on_b;
  double inv_n;
  int ch, h, w, px;
  int CG, R;                             // channel groups (ch / 8), rows per strip (ceil(h / (16 / CG)))
};
struct ClArgs { ClLevel l[4]; int D; int nb; unsigned long long* dbg; int* err; };   // nb: clusters in role B (2 or 3);   // dbg: [6 clusters][16 slots] SM cycles per phase (SATMVS_RED_DEBUG)

struct ClSmemPlan { int wsm, tile, part, keep, total_floats; };
__host__ __device__ inline ClSmemPlan cl_smem_plan(int ch, int w, int R) {
  ClSmemPlan p;
  const int npx = R * w;
  p.wsm = ch * 9 * 8;                    // one filter bank (8 output channels); two banks per CTA
  p.tile = ch * (R + 2) * (w + 8);
  p.part = (ch / 4) * 8 * npx;           // ch/4 input-channel chunks x 8 output channels
  p.keep = 8 * npx;
  p.total_floats = 2 * p.wsm + p.tile + p.part + p.keep;
  return p;
}
constexpr int kClStaticSmemBytes = 1024;

// Pointwise transcendental functions on the SFU (ex2.approx + rcp.approx: ~2 ulp each).  Inside the recurrence every
// SM of a cluster owns 1/16 of a level's elements, so the ~40-instruction libm forms would cost as much as the convs.
__device__ __forceinline__ float cl_sigmoid(float x) { return __fdividef(1.0f, 1.0f + __expf(-x)); }
__device__ __forceinline__ float cl_tanh(float x) { return fmaf(2.0f, __fdividef(1.0f, 1.0f + __expf(-2.0f * x)), -1.0f); }

// n / d for n, d < 65536 as one multiply-high (exact: n * d < 2^32); d <= 1 handled by a select
struct ClDiv { unsigned d, m; };
__device__ __forceinline__ ClDiv cl_mkdiv(int d) { ClDiv r; r.d = (unsigned)d; r.m = d > 1 ? 0xFFFFFFFFu / (unsigned)d + 1u : 0u; return r; }
__device__ __forceinline__ int cl_div(int n, const ClDiv& dv) { return dv.d <= 1u ? n : (int)__umulhi((unsigned)n, dv.m); }

__device__ __forceinline__ int cl_ld_acquire(const int* p) { int v; asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v; }
__device__ __forceinline__ void cl_st_release(int* p, int v) { asm volatile("st.release.gpu.global.s32 [%0], %1;" :: "l"(p), "r"(v) : "memory"); }

// One unit: CI input channels x 9 taps x 8 output channels x 4 pixels, accumulated as (co, co+1) pairs.
template <int CI>
__device__ __forceinline__ void cl_conv_unit(const float* __restrict__ tile_px, int ci_stride, int pitch,
                                             const float* __restrict__ wsm, u64 (&acc2)[4][4]) {
#pragma unroll kClUnroll
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
    const float* wp = wsm + c * 9 * 8;
#pragma unroll
    for (int t = 0; t < 9; ++t) {
      const ulonglong2 w0 = *reinterpret_cast<const ulonglong2*>(wp + t * 8);
      const ulonglong2 w1 = *reinterpret_cast<const ulonglong2*>(wp + t * 8 + 4);
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
  __shared__ double red[2][kClWarps];
  __shared__ double stat_out[2][2];         // this CTA's partial (sum, sum of squares): [0] first conv of a plane, [1] second
  __shared__ float coef[kClK][2];           // GroupNorm scale / shift of the current norm for this CTA's 8 channels

  // clusters 0..3: role A of level cid; clusters 4, 5: role B of levels {0, 1} and {2, 3} (B carries a third of A's
  // convolution work per level; a B200 keeps seven such clusters resident, so a B cluster serves two levels and the
  // second level's tile is prefetched behind the first level's tail.  Measured with the clock64 marks below: a B cluster
  // needs ~42 k cycles per plane for its two levels against ~37 k for an A cluster, so A waits ~10 % of a plane for u)
  const int cid = blockIdx.x / kClSize;
  const bool roleB = cid >= 4;
  const int rank = (int)cluster.block_rank();
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  // ---- geometry of the level this CTA is working on (role B switches level twice per plane) ----
  const ClLevel* Lp = nullptr;
  int R, ch, w, y0, nrows, pitch, w4, NPX, npx4, KS2, ci_stride, c_own, lr_lo, per_ci, stage_items, nunits, pw_items;
  int kc0, q0, ly0, x0; bool have0; long long px_off0;
  ClDiv dv_npx4, dv_w4, dv_perci;
  auto setup = [&](int lv) {
    Lp = &a.l[lv];
    const int CG = Lp->CG;
    R = Lp->R; ch = Lp->ch; w = Lp->w;
    const int strip = rank / CG, cgp = rank - strip * CG;
    y0 = strip * R;
    nrows = Lp->h - y0; nrows = nrows > R ? R : nrows; nrows = nrows < 0 ? 0 : nrows;
    pitch = w + 8; w4 = w >> 2;
    NPX = R * w;                            // channel stride of the partial / keep buffers
    npx4 = nrows * w4;                      // pixel quads of this strip
    KS2 = ch >> 2;                          // input-channel chunks of 4
    ci_stride = (R + 2) * pitch;
    c_own = cgp * kClK;                     // first hidden channel of this CTA
    // tile rows that exist in the image: local rows lr_lo .. lr_lo + nlr - 1 (local row lr = image row y0 - 1 + lr)
    lr_lo = (y0 == 0) ? 1 : 0;
    int lr_hi = nrows + 1; if (y0 - 1 + lr_hi >= Lp->h) lr_hi = Lp->h - y0;
    const int nlr = (nrows > 0 && lr_hi >= lr_lo) ? lr_hi - lr_lo + 1 : 0;
    per_ci = nlr * w4; stage_items = ch * per_ci;
    dv_npx4 = cl_mkdiv(npx4); dv_w4 = cl_mkdiv(w4); dv_perci = cl_mkdiv(per_ci);
    // this thread's first conv unit (the same for every plane and conv of the level): chunk kc, pixel quad q
    nunits = KS2 * npx4; pw_items = 8 * npx4;
    have0 = tid < nunits;
    kc0 = have0 ? cl_div(tid, dv_npx4) : 0; q0 = have0 ? tid - kc0 * npx4 : 0;
    ly0 = cl_div(q0, dv_w4); x0 = (q0 - ly0 * w4) << 2;
    px_off0 = (long long)(y0 + ly0) * w + x0;
  };

  // shared-memory layout.  A: [Wr][Wo][tile][part][own h]; B: [Wu of levels 0..3][tile (largest level)][part (largest level)]
  float *w0s, *w1s, *tile, *part, *keepH;
  int tile_floats;
  if (!roleB) {
    const ClSmemPlan sp = cl_smem_plan(a.l[cid].ch, a.l[cid].w, a.l[cid].R);
    w0s = cl_smem;                          // [ci][tap][8] reset-gate filters
    w1s = w0s + sp.wsm;                     // [ci][tap][8] output-conv filters
    tile = w1s + sp.wsm;                    // [ci][R+2][pitch], pixel x at column 4 + x; borders stay zero
    part = tile + sp.tile;                  // [chunk][8][NPX]
    keepH = part + sp.part;                 // this CTA's own channels of h
    tile_floats = sp.tile;
  } else {
    // levels la = 2 (cid - 4) and lb = la + 1: both filter banks and both tiles stay resident
    const int la = 2 * (cid - 4);
    const ClSmemPlan spa = cl_smem_plan(a.l[la].ch, a.l[la].w, a.l[la].R), spb = cl_smem_plan(a.l[la + 1].ch, a.l[la + 1].w, a.l[la + 1].R);
    w0s = cl_smem;                          // Wu of level la
    w1s = w0s + spa.wsm;                    // Wu of level lb
    tile = w1s + spb.wsm;                   // tile of level la, then tile of level lb
    part = tile + spa.tile + spb.tile;
    keepH = part;
    tile_floats = spa.tile + spb.tile;
  }
  float* const tile0 = tile;
  float* const tile1 = roleB ? tile + cl_smem_plan(a.l[2 * (cid - 4)].ch, a.l[2 * (cid - 4)].w, a.l[2 * (cid - 4)].R).tile : tile;

  // ---- once: filters resident, tile zero ----
  if (!roleB) {
    setup(cid);
    for (int i = tid; i < ch * 9 * 8; i += kClThreads) {
      const int co = i & 7, r = i >> 3;     // r = ci * 9 + tap
      w0s[i] = __ldg(Lp->gate_w + (long long)(c_own + co) * Lp->w_co + r);
      w1s[i] = __ldg(Lp->out_w + (long long)(c_own + co) * Lp->w_co + r);
    }
  } else {
    for (int k = 0; k < 2; ++k) {
      setup(2 * (cid - 4) + k);
      float* dst = k ? w1s : w0s;
      for (int i = tid; i < ch * 9 * 8; i += kClThreads) {
        const int co = i & 7, r = i >> 3;
        dst[i] = __ldg(Lp->gate_w + (long long)(ch + c_own + co) * Lp->w_co + r);
      }
    }
  }
  for (int i = tid; i < tile_floats; i += kClThreads) tile[i] = 0.0f;
  __syncthreads();

  // debug (-DSATMVS_CL_TIMERS + SATMVS_RED_DEBUG=1): SM cycles rank 0 / thread 0 of every cluster spends per phase slot
#ifdef SATMVS_CL_TIMERS
  long long t_prev = 0, t_acc[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) t_acc[i] = 0;
  auto mark = [&](int slot) {
    if (a.dbg && tid == 0) { const long long t = clock64(); if (slot >= 0) t_acc[slot] += t - t_prev; t_prev = t; }
  };
#else
  auto mark = [](int) {};
#endif
  // stage src[ci][gy][x] (channel stride cs, plane already selected) for the existing rows around the strip
  // (16-byte cp.async.cg: L2-coherent, no registers, every row piece in flight at once; the caller waits with
  // cp.async.wait_all + __syncthreads before the tile is read)
  auto stage_tile = [&](const float* src, long long cs) {
    const float* sbase = src + (long long)(y0 - 1 + lr_lo) * w;
    const unsigned tbase = (unsigned)__cvta_generic_to_shared(tile + lr_lo * pitch + 4);
    for (int i = tid; i < stage_items; i += kClThreads) {
      const int ci = cl_div(i, dv_perci), r = i - ci * per_ci;
      const int lr = cl_div(r, dv_w4);
      const float* g = sbase + ci * cs + 4 * r;                                       // rows are contiguous: r = lr * w4 + x4
      const unsigned dst = tbase + 4u * (unsigned)(ci * ci_stride + lr * pitch + ((r - lr * w4) << 2));
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"(dst), "l"(g) : "memory");
    }
  };

  // One convolution of a plane: part[0][co][px] = pre[co][px] + sum_ci conv3x3(src[ci]; wsm[ci][.][co]) for this CTA's
  // strip and 8 output channels; returns this thread's share of (sum, sum of squares) over those outputs.
  // `before_stage` runs after the x-half loads are issued and before the tile is staged (the wait half of a split
  // cluster barrier goes there, so the loads fly across it).
  auto conv_phase = [&](const float* src, long long src_cs, const float* wsm, const float* pre, long long pre_cs,
                        bool prefetch_next, double& st_s, double& st_q, auto&& before_stage, bool staged_already = false) {
    u64 acc2[4][4];
    float4 pv[8];
#pragma unroll
    for (int co = 0; co < 8; ++co) pv[co] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (have0 && kc0 == 0) {               // the first chunk starts from the x-half (read-only data of an earlier kernel)
      const float* pp = pre + px_off0;
#pragma unroll
      for (int co = 0; co < 8; ++co) pv[co] = __ldg(reinterpret_cast<const float4*>(pp + co * pre_cs));
      if (prefetch_next) {
#pragma unroll
        for (int co = 0; co < 8; ++co) asm volatile("prefetch.global.L2 [%0];" :: "l"(pp + co * pre_cs + Lp->px));
      }
    }
    before_stage();
    if (!staged_already) stage_tile(src, src_cs);
    mark(13);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      acc2[i][0] = pk(pv[2 * i].x, pv[2 * i + 1].x); acc2[i][1] = pk(pv[2 * i].y, pv[2 * i + 1].y);
      acc2[i][2] = pk(pv[2 * i].z, pv[2 * i + 1].z); acc2[i][3] = pk(pv[2 * i].w, pv[2 * i + 1].w);
    }
    asm volatile("cp.async.wait_all;" ::: "memory");
    __syncthreads();
    mark(14);
    for (int u = tid; u < nunits; u += kClThreads) {
      int kc = kc0, q = q0, ly = ly0, x = x0;
      if (u != tid) {                        // further rounds (shapes with more units than threads)
        kc = cl_div(u, dv_npx4); q = u - kc * npx4;
        ly = cl_div(q, dv_w4); x = (q - ly * w4) << 2;
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) acc2[i][j] = 0ULL;
        if (kc == 0) {
          const float* pp = pre + (long long)(y0 + ly) * w + x;
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const float4 p0 = __ldg(reinterpret_cast<const float4*>(pp + (2 * i) * pre_cs));
            const float4 p1 = __ldg(reinterpret_cast<const float4*>(pp + (2 * i + 1) * pre_cs));
            acc2[i][0] = pk(p0.x, p1.x); acc2[i][1] = pk(p0.y, p1.y); acc2[i][2] = pk(p0.z, p1.z); acc2[i][3] = pk(p0.w, p1.w);
          }
        }
      }
      cl_conv_unit<4>(tile + (kc * 4) * ci_stride + ly * pitch + 4 + x, ci_stride, pitch, wsm + (kc * 4) * 9 * 8, acc2);
      float* pp = part + (long long)(kc * 8) * NPX + 4 * q;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        float lo[4], hi[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) upk(acc2[i][j], lo[j], hi[j]);
        *reinterpret_cast<float4*>(pp + (2 * i) * NPX) = make_float4(lo[0], lo[1], lo[2], lo[3]);
        *reinterpret_cast<float4*>(pp + (2 * i + 1) * NPX) = make_float4(hi[0], hi[1], hi[2], hi[3]);
      }
    }
    mark(15);
    __syncthreads();
    mark(3);
    // add the chunks in a fixed order, keep the result in chunk 0, GroupNorm sums
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
      st_s += (double)s4; st_q += (double)q4;
    }
  };

  // block-wide (sum, sum of squares) -> stat_out[which]
  auto publish_stats = [&](int which, double s, double q) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) { s += __shfl_xor_sync(0xffffffffu, s, off); q += __shfl_xor_sync(0xffffffffu, q, off); }
    if (lane == 0) { red[0][warp] = s; red[1][warp] = q; }
    __syncthreads();
    if (tid < 2) {
      double t = 0.0;
      for (int i = 0; i < kClWarps; ++i) t += red[tid][i];
      stat_out[which][tid] = t;
    }
  };
  // after the cluster barrier: warp 0 gathers the level's (sum, sum of squares) from the 16 ranks through distributed
  // shared memory (fixed order: butterfly over the ranks); lanes 0..7 turn them into the scale / shift of this CTA's channels
  auto gather_coef = [&](int which, const float* gw, const float* gb) {
    if (warp == 0) {
      const int idx = lane >> 4, rk = lane & 15;
      const double* remote = cluster.map_shared_rank(&stat_out[which][idx], rk);
      double v = *remote;
#pragma unroll
      for (int off = 8; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
      const double S = __shfl_sync(0xffffffffu, v, 0), Q = __shfl_sync(0xffffffffu, v, 16);
      if (lane < 8) {
        const double mean = S * Lp->inv_n;
        const float var = (float)fmax(Q * Lp->inv_n - mean * mean, 0.0);
        const float rstd = rsqrtf(var + 1e-5f);
        const float ca = __ldg(gw + c_own + lane) * rstd;
        coef[lane][0] = ca; coef[lane][1] = __ldg(gb + c_own + lane) - (float)mean * ca;
      }
    }
    __syncthreads();
  };
  auto wait_flag = [&](const int* flag, int target) {
    if (tid == 0) {
      // bounded spin: the partner cluster is co-resident (checked by the host), so this normally takes a few polls; if it
      // never arrives (~4 s) the kernel traps and the error surfaces at the caller's next synchronisation instead of a hang
      const long long t0 = clock64();
      while (cl_ld_acquire(flag) < target)
        if (clock64() - t0 > (1LL << 32)) { if (a.err) *reinterpret_cast<volatile int*>(a.err) = 3; break; }   // report (common.cuh), do not trap the context
    }
    __syncthreads();
  };


  if (roleB) {
    // ================= clusters B: u[d] = sigmoid(GN_u(GX_u[d] + conv(h[d]; Wu))) for levels la, lb in turn =================
    const int la = 2 * (cid - 4), lb = la + 1;
    // the tail of a task: sums -> scale / shift -> u -> L2, then this CTA's share is counted (16 per plane and level)
    auto finish_task = [&](int which, int d) {
      cluster.sync();
      gather_coef(which, Lp->un_w, Lp->un_b);
      for (int o = tid; o < pw_items; o += kClThreads) {
        const int co = cl_div(o, dv_npx4), p4 = o - co * npx4;
        const int ly = cl_div(p4, dv_w4), x = (p4 - ly * w4) << 2;
        const float4 g = *reinterpret_cast<const float4*>(part + (long long)co * NPX + 4 * p4);
        const float ca = coef[co][0], cb = coef[co][1];
        const float4 u = make_float4(cl_sigmoid(fmaf(g.x, ca, cb)), cl_sigmoid(fmaf(g.y, ca, cb)),
                                     cl_sigmoid(fmaf(g.z, ca, cb)), cl_sigmoid(fmaf(g.w, ca, cb)));
        *reinterpret_cast<float4*>(Lp->ub + ((long long)(d & 1) * ch + c_own + co) * Lp->px + (long long)(y0 + ly) * w + x) = u;
      }
      __syncthreads();
      if (tid == 0) { __threadfence(); atomicAdd(Lp->flags + kClFlagStride, 1); }
    };
    for (int d = 0; d < a.D; ++d) {
      double ss = 0.0, sq = 0.0;
      // ---- level la ----
      setup(la); tile = tile0;
      mark(-1);
      wait_flag(Lp->flags, d);                               // h[d] published by cluster A (slot 0: before the launch)
      mark(0);
      conv_phase(Lp->s + (long long)d * Lp->px, Lp->s_cs, w0s, Lp->gx + (long long)(ch + c_own) * Lp->g_cs + (long long)d * Lp->px,
                 Lp->g_cs, d + 1 < a.D, ss, sq, [] {});
      publish_stats(0, ss, sq);
      // level lb's tile flies (cp.async into its own buffer) while la's sums cross the cluster
      setup(lb); tile = tile1;
      mark(1);
      wait_flag(Lp->flags, d);
      mark(2);
      stage_tile(Lp->s + (long long)d * Lp->px, Lp->s_cs);
      setup(la); tile = tile0;
      finish_task(0, d);
      mark(4);
      // ---- level lb ----
      setup(lb); tile = tile1;
      ss = 0.0; sq = 0.0;
      conv_phase(Lp->s + (long long)d * Lp->px, Lp->s_cs, w1s, Lp->gx + (long long)(ch + c_own) * Lp->g_cs + (long long)d * Lp->px,
                 Lp->g_cs, d + 1 < a.D, ss, sq, [] {}, true);
      publish_stats(1, ss, sq);
      mark(5);
      finish_task(1, d);
      mark(6);
    }
  } else {
    // ================= cluster A =================
    for (int d = 0; d < a.D; ++d) {
      const long long plane_g = (long long)d * Lp->px;
      // ---- P1: G_r = GX_r[d] + conv(h[d]; Wr) ----
      double ss = 0.0, sq = 0.0;
      mark(-1);
      conv_phase(Lp->s + plane_g, Lp->s_cs, w0s, Lp->gx + (long long)c_own * Lp->g_cs + plane_g, Lp->g_cs, d + 1 < a.D, ss, sq, [&] {
        if (d > 0) {                                         // #4 of the previous plane: its new state is in L2 -> tell cluster B
          cluster.barrier_wait();
          if (rank == 0 && tid == 0) { __threadfence(); cl_st_release(Lp->flags, d); }
        }
      });
      mark(0);
      publish_stats(0, ss, sq);
      mark(1);
      cluster.sync();                                        // #1: reset-gate sums of every CTA of the level are published
      mark(2);
      gather_coef(0, Lp->rn_w, Lp->rn_b);
      mark(3);
      // ---- E1: rh = sigmoid(GN_r(G_r)) * h; keep this CTA's channels of h ----
      for (int o = tid; o < pw_items; o += kClThreads) {
        const int co = cl_div(o, dv_npx4), p4 = o - co * npx4;
        const int ly = cl_div(p4, dv_w4), x = (p4 - ly * w4) << 2;
        const float4 g = *reinterpret_cast<const float4*>(part + (long long)co * NPX + 4 * p4);
        const float4 hv = *reinterpret_cast<const float4*>(tile + (c_own + co) * ci_stride + (ly + 1) * pitch + 4 + x);
        const float ca = coef[co][0], cb = coef[co][1];
        const float4 rh = make_float4(cl_sigmoid(fmaf(g.x, ca, cb)) * hv.x, cl_sigmoid(fmaf(g.y, ca, cb)) * hv.y,
                                      cl_sigmoid(fmaf(g.z, ca, cb)) * hv.z, cl_sigmoid(fmaf(g.w, ca, cb)) * hv.w);
        *reinterpret_cast<float4*>(Lp->rh + (long long)(c_own + co) * Lp->px + (long long)(y0 + ly) * w + x) = rh;
        *reinterpret_cast<float4*>(keepH + (long long)co * NPX + 4 * p4) = hv;
      }
      mark(4);
      cluster.barrier_arrive();                              // #2 (wait inside conv_phase): r*h of the whole level is in L2
      mark(5);
      // ---- P2: O = OX[d] + conv(rh; Wo) ----
      ss = 0.0; sq = 0.0;
      conv_phase(Lp->rh, Lp->px, w1s, Lp->ox + (long long)c_own * Lp->o_cs + plane_g, Lp->o_cs, d + 1 < a.D, ss, sq,
                 [&] { cluster.barrier_wait(); });
      mark(6);
      publish_stats(1, ss, sq);
      cluster.barrier_arrive();                              // #3: output-conv sums are published
      mark(7);
      // the update gate of this plane (cluster B, normally long finished): loads fly across the barrier
      wait_flag(Lp->flags + kClFlagStride, kClSize * (d + 1));
      float4 uv[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int o = tid + k * kClThreads;
        uv[k] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (o < pw_items) {
          const int co = cl_div(o, dv_npx4), p4 = o - co * npx4;
          const int ly = cl_div(p4, dv_w4), x = (p4 - ly * w4) << 2;
          uv[k] = __ldcg(reinterpret_cast<const float4*>(Lp->ub + ((long long)(d & 1) * ch + c_own + co) * Lp->px + (long long)(y0 + ly) * w + x));
        }
      }
      mark(8);
      cluster.barrier_wait();
      mark(9);
      gather_coef(1, Lp->on_w, Lp->on_b);
      mark(10);
      // ---- E2: h[d+1] = u*h + (1-u)*tanh(GN_o(O)) ----
      float* hn = Lp->s + plane_g + Lp->px;
      auto update = [&](int o, const float4* upre) {
        const int co = cl_div(o, dv_npx4), p4 = o - co * npx4;
        const int ly = cl_div(p4, dv_w4), x = (p4 - ly * w4) << 2;
        const long long goff = (long long)(y0 + ly) * w + x;
        const float4 u = upre ? *upre : __ldcg(reinterpret_cast<const float4*>(Lp->ub + ((long long)(d & 1) * ch + c_own + co) * Lp->px + goff));
        const float4 y = *reinterpret_cast<const float4*>(part + (long long)co * NPX + 4 * p4);
        const float4 hv = *reinterpret_cast<const float4*>(keepH + (long long)co * NPX + 4 * p4);
        const float ca = coef[co][0], cb = coef[co][1];
        float4 hh;
        hh.x = u.x * hv.x + (1.0f - u.x) * cl_tanh(fmaf(y.x, ca, cb)); hh.y = u.y * hv.y + (1.0f - u.y) * cl_tanh(fmaf(y.y, ca, cb));   // module.py:57
        hh.z = u.z * hv.z + (1.0f - u.z) * cl_tanh(fmaf(y.z, ca, cb)); hh.w = u.w * hv.w + (1.0f - u.w) * cl_tanh(fmaf(y.w, ca, cb));
        *reinterpret_cast<float4*>(hn + (long long)(c_own + co) * Lp->s_cs + goff) = hh;
      };
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int o = tid + k * kClThreads;
        if (o < pw_items) update(o, &uv[k]);
      }
      for (int o = tid + 4 * kClThreads; o < pw_items; o += kClThreads) update(o, nullptr);
      mark(11);
      cluster.barrier_arrive();                              // #4 (wait inside the next plane's conv_phase)
      mark(12);
    }
    cluster.barrier_wait();
  }
#ifdef SATMVS_CL_TIMERS
  if (a.dbg && rank == 0 && tid == 0)
#pragma unroll
    for (int i = 0; i < 16; ++i) a.dbg[cid * 16 + i] = (unsigned long long)t_acc[i];
#endif
  cluster.sync();                                            // no CTA leaves while a peer may still read its shared memory
}

// Shapes the cluster kernel takes: every level's width a multiple of 4 (vector rows) and the per-CTA
// working set within the shared-memory opt-in limit.  Returns the dynamic shared-memory bytes, 0 if unsupported.
inline size_t red_cluster_smem_bytes(const ClArgs& a, int smem_optin) {
  size_t need = 0;
  ClSmemPlan sp[4];
  for (int l = 0; l < 4; ++l) {
    const ClLevel& L = a.l[l];
    if (L.w % 4 || L.ch % 8 || L.ch / 8 > kClSize || kClSize % (L.ch / 8)) return 0;
    sp[l] = cl_smem_plan(L.ch, L.w, L.R);
    if (L.ch * (L.R + 2) * (L.w / 4) >= 65536 || 8 * L.R * (L.w / 4) >= 65536) return 0;     // cl_div range
    const size_t b = (size_t)sp[l].total_floats * sizeof(float);                             // role A of level l
    need = b > need ? b : need;
  }
  for (int c = 0; c < 2; ++c) {                                                                // role B of levels 2c, 2c + 1
    const ClSmemPlan &p = sp[2 * c], &q = sp[2 * c + 1];
    const size_t b = (size_t)(p.wsm + q.wsm + p.tile + q.tile + (p.part > q.part ? p.part : q.part)) * sizeof(float);
    need = b > need ? b : need;
  }
  if (need + kClStaticSmemBytes > (size_t)smem_optin) return 0;
  return need;
}

// Launches the recurrence as 6 clusters of 16 CTAs (all co-resident: they exchange flags through L2);
// *launched stays false when the shape or the device does not take it (the caller then runs the per-plane chain).
inline int red_cluster_launch(ClArgs& a, int* flags_base, cudaStream_t st, bool* launched) {
  *launched = false;
  a.err = async_error_devptr();
  int dev = 0, optin = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
  for (int l = 0; l < 4; ++l) {
    ClLevel& L = a.l[l];
    L.CG = L.ch / kClK;
    const int PS = kClSize / (L.CG > 0 ? L.CG : 1);
    L.R = (L.h + PS - 1) / PS;
    L.flags = flags_base + l * 2 * kClFlagStride;
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
  cfg.gridDim = dim3(6 * kClSize); cfg.blockDim = dim3(kClThreads); cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = kClSize; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  int nclusters = 0;
  if ((e = cudaOccupancyMaxActiveClusters(&nclusters, red_cluster_kernel, &cfg)) != cudaSuccess || nclusters < 6)
    { if (verbose) fprintf(stderr, "red_cluster_launch: max active clusters %d\n", nclusters); return declined("fewer than 6 co-resident clusters", e); }
  a.nb = 2;
  cfg.gridDim = dim3(6 * kClSize);
  cudaMemsetAsync(flags_base, 0, 4 * 2 * kClFlagStride * sizeof(int), st);
  a.dbg = verbose ? reinterpret_cast<unsigned long long*>(flags_base + 4 * 2 * kClFlagStride) : nullptr;
  if ((e = cudaLaunchKernelEx(&cfg, red_cluster_kernel, a)) != cudaSuccess) return declined("launch", e);
#ifdef SATMVS_CL_TIMERS
  if (verbose) {
    unsigned long long h[7 * 16] = {};
    cudaStreamSynchronize(st);
    cudaMemcpy(h, a.dbg, sizeof(h), cudaMemcpyDeviceToHost);
    for (int c = 0; c < 4 + a.nb; ++c) {
      fprintf(stderr, "red_cluster_launch: cluster %d kcycles per phase slot:", c);
      for (int i = 0; i < 16; ++i) fprintf(stderr, " %.0f", h[c * 16 + i] * 1e-3);
      fprintf(stderr, "\n");
    }
  }
#endif
  *launched = true;
  return SATMVS_OK;
}

}  // namespace satmvs
