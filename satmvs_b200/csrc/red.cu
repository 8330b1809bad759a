// red.cu — the RED regulariser (2-D conv-GRU UNet recurring over depth planes) on the conv engine.
//
// Reference: RED_Regularization.forward (modules/module.py:614-649), slice_RED_Regularization.forward
// (:672-693), ConvGRUCell2 (:6-58).  The reference runs ~15 convolutions + 12 GroupNorms + ~25
// pointwise ops per depth plane, strictly one plane after another (~70 launches x D planes).
//
// Restructuring (DESIGN.md "RED"): a convolution over the concatenation [x, h] is the sum of a
// convolution over x and one over h, so everything that does not depend on the hidden state is
// computed for ALL planes at once on [C, D, h, w] tensors:
//   A. batched:    E1..E3 = the stride-2 encoders of -cost; GX_l, OX_l = the x-halves of every
//                  GRU's gate / output convolution (bias included)
//   B. recurrent:  per plane d and level l (the 4 levels run in the same grouped launches)
//        P1  G = GX[d] + conv(h; Wg_h)                 (+ GroupNorm sums for r and u)
//        E1  rh = sigmoid(GN_r(G_r)) * h
//        P2  O = OX[d] + conv(rh; Wo_h)                (+ GroupNorm sums)
//        E2  h' = u*h + (1-u)*tanh(GN_o(O)),  u = sigmoid(GN_u(G_u));  h' is kept for every plane
//   C. batched:    the decoder (3 stride-2 transposed convs with skip adds, final 8->1 transposed
//                  conv with bias) over the stored states of all planes.
// Only B is sequential in D: 4 launches per plane instead of ~70.
#include <cooperative_groups.h>
#include <type_traits>
#include "conv_engine.cuh"
#include "direct_conv.cuh"
#include "umma_conv.cuh"
#include "packed.cuh"
#include "prof.cuh"
#include "red_cluster.cuh"
#include "red_tc.cuh"
namespace cg = cooperative_groups;

namespace satmvs {

constexpr float kGnEps = 1e-5f;   // nn.GroupNorm(1, C, 1e-5, True), modules/module.py:15-20

// Programmatic dependent launch for the per-plane chain P1 -> E1 -> P2 -> E2 -> P1 ...: every kernel of the chain
// releases its successor at entry (griddepcontrol.launch_dependents), does the work that does not depend on its
// predecessor (staging the conv weights, index arithmetic) and only then waits for the predecessor's memory
// (griddepcontrol.wait).  Launch latency and the weight staging of kernel n+1 hide under kernel n.  Completion stays
// serial, so every kernel still sees the results of all earlier ones after its wait.
__device__ __forceinline__ void pdl_release() { asm volatile("griddepcontrol.launch_dependents;"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

template <class Kern, class... Args>
static void launch_chain(Kern kern, int grid, int block, size_t smem, cudaStream_t st, bool pdl, const Args&... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(grid); cfg.blockDim = dim3(block); cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr; cfg.numAttrs = pdl ? 1 : 0;
  cudaLaunchKernelEx(&cfg, kern, args...);
}

struct RedLevel {
  int ch, h, w;          // hidden channels, spatial size
  int cx;                // channels of the x input of this level's GRU
  float* gx;             // [2ch][D][h][w]   gates (x-half, then full gates in place)
  float* ox;             // [ch][D][h][w]    output conv (x-half, then full in place)
  float* rh;             // [ch][1][h][w]
  float* ub;             // [2][ch][h][w]   update gate of the current / next plane (cluster recurrence)
  float* s;              // [ch][D+1][h][w]  state history; slot 0 = initial state
};

struct RedPlan {
  int C, D, H, W;
  float* e[3];           // E1 [16][D][H/2][W/2], E2 [32][D][H/4][W/4], E3 [64][D][H/8][W/8]
  float* u[3];           // U1 [8][D+1][H][W], U2 [16][D+1][H/2][W/2], U3 [32][D+1][H/4][W/4]
  RedLevel lv[4];
  double* stats;         // [D][4][3][2]
  char* wpack[4]; size_t wpack_bytes[4];
  int* umma_err;
  int* cl_flags;         // [4][2][kClFlagStride] plane counters exchanged by the two clusters of a level
  char* tcpack[4];       // packed hidden-state filters of the tensor-core recurrence (red_tc.cuh)
  int* ready;            // [4][D] per-(level, plane) completion counters of the batched x-half convs (overlapped flow)
  size_t bytes;
  size_t pack_bytes;     // bytes used in the external pack region (red_plan with pack_base)
};

static RedPlan red_plan(int C, int D, int H, int W, char* base, char* pack_base = nullptr) {
  RedPlan p{};
  p.C = C; p.D = D; p.H = H; p.W = W;
  size_t off = 0, poff = 0;
  auto take = [&](size_t nfloats) {
    float* r = reinterpret_cast<float*>(base + off);
    off += ((nfloats * sizeof(float) + 255) / 256) * 256;
    return r;
  };
  const int chs[4] = {8, 16, 32, 64};
  for (int l = 0; l < 4; ++l) {
    RedLevel& L = p.lv[l];
    L.ch = chs[l]; L.h = H >> l; L.w = W >> l;
    L.cx = (l == 0) ? C : chs[l];          // GRU_l's x input: -cost, E1 (16), E2 (32), E3 (64)
    const size_t px = (size_t)L.h * L.w;
    L.gx = take(2 * L.ch * D * px);
    L.ox = take(L.ch * D * px);
    L.rh = take(L.ch * px);
    L.ub = take(2 * L.ch * px);
    L.s = take(L.ch * (size_t)(D + 1) * px);
  }
  for (int i = 0; i < 3; ++i) {
    p.e[i] = take((size_t)chs[i + 1] * D * (H >> (i + 1)) * (W >> (i + 1)));
    p.u[i] = take((size_t)chs[i] * (D + 1) * (H >> i) * (W >> i));
  }
  p.stats = reinterpret_cast<double*>(base + off);
  off += ((size_t)D * 4 * 3 * 2 * sizeof(double) + 64 * sizeof(double) + 255) / 256 * 256;   // + 64 debug counters
  for (int l = 0; l < 4; ++l) {   // packed (raw, lo) x-half weights of the tensor-core convs, per level
    p.wpack_bytes[l] = (size_t)(p.lv[l].cx / 8 + 1) * 2 * 9 * 2 * ((5 * p.lv[l].ch + 15) / 16 * 16) * 16;   // gates 2ch + output ch + encoder 2ch
    if (pack_base) { p.wpack[l] = pack_base + poff; poff += (p.wpack_bytes[l] + 255) / 256 * 256; }
    else { p.wpack[l] = base + off; off += (p.wpack_bytes[l] + 255) / 256 * 256; }
  }
  p.umma_err = reinterpret_cast<int*>(base + off);
  off += 256;
  p.cl_flags = reinterpret_cast<int*>(base + off);
  off += 4 * 2 * 32 * sizeof(int) + 8 * 16 * sizeof(unsigned long long);   // + debug counters
  off = (off + 255) / 256 * 256;
  for (int l = 0; l < 4; ++l) {
    if (pack_base) { p.tcpack[l] = pack_base + poff; poff += (tc_pack_bytes(chs[l]) + 255) / 256 * 256; }
    else { p.tcpack[l] = base + off; off += (tc_pack_bytes(chs[l]) + 255) / 256 * 256; }
  }
  p.pack_bytes = poff;
  p.ready = reinterpret_cast<int*>(base + off);
  off += ((size_t)4 * D * sizeof(int) + 255) / 256 * 256;
  p.bytes = off;
  return p;
}

// ---- elementwise GRU kernels over the 4 levels of one plane ----
struct GruLevelArgs {
  const float* g;        // gates of this plane: [2ch][.][h][w] base already at plane d; channel stride gcs
  const float* o;        // output conv of this plane; channel stride ocs
  const float* hprev;    // [ch] planes, channel stride scs
  float* hnext;
  float* rh;             // [ch][h][w]
  const float *rn_w, *rn_b, *un_w, *un_b, *on_w, *on_b;
  const double* stats;   // [3][2] of this (plane, level): r, u, o
  long long gcs, ocs, scs;
  double inv_n;          // 1 / (ch * px)
  int ch, px;
  int begin;             // first flat element index of this level in the launch
};
struct GruArgs { GruLevelArgs l[4]; int total; };

// y = (x - mean) * rstd * gamma + beta = x*a + b.  Sums are fp64; mean/var formed in fp64 (no
// cancellation problem), rstd in fp32 like ATen's GroupNorm.
__device__ __forceinline__ void gn_coeff(const double* st, double inv_n, float gamma, float beta, float& a, float& b) {
  const double mean = st[0] * inv_n;
  const float var = (float)fmax(st[1] * inv_n - mean * mean, 0.0);
  const float rstd = rsqrtf(var + kGnEps);
  a = gamma * rstd;
  b = beta - (float)mean * a;
}

__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + expf(-x)); }

__global__ void gru_reset_kernel(const __grid_constant__ GruArgs a) {
  pdl_release();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  pdl_wait();
  if (i >= a.total) return;
  int li = 0;
#pragma unroll
  for (int k = 1; k < 4; ++k) if (i >= a.l[k].begin) li = k;
  const GruLevelArgs& L = a.l[li];
  const int e = i - L.begin, c = e / L.px, p = e - c * L.px;
  float ga, gb;
  gn_coeff(L.stats, L.inv_n, __ldg(L.rn_w + c), __ldg(L.rn_b + c), ga, gb);
  // read-only path is safe here: nothing reads gates / states through L1 before their final value is written
  // (the conv kernels, which update gates in place, use L2-coherent loads: see gru_conv_kernel)
  const float r = sigmoidf_(fmaf(__ldg(L.g + c * L.gcs + p), ga, gb));
  L.rh[e] = r * __ldg(L.hprev + c * L.scs + p);
}

__global__ void gru_update_kernel(const __grid_constant__ GruArgs a) {
  pdl_release();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  pdl_wait();
  if (i >= a.total) return;
  int li = 0;
#pragma unroll
  for (int k = 1; k < 4; ++k) if (i >= a.l[k].begin) li = k;
  const GruLevelArgs& L = a.l[li];
  const int e = i - L.begin, c = e / L.px, p = e - c * L.px;
  float ua, ub, oa, ob;
  gn_coeff(L.stats + 2, L.inv_n, __ldg(L.un_w + c), __ldg(L.un_b + c), ua, ub);
  gn_coeff(L.stats + 4, L.inv_n, __ldg(L.on_w + c), __ldg(L.on_b + c), oa, ob);
  const float u = sigmoidf_(fmaf(__ldg(L.g + (c + L.ch) * L.gcs + p), ua, ub));
  const float y = tanhf(fmaf(__ldg(L.o + c * L.ocs + p), oa, ob));
  const float h = __ldg(L.hprev + c * L.scs + p);
  L.hnext[c * L.scs + p] = u * h + (1.0f - u) * y;       // module.py:57
}


// ---------------------------------------------------------------------------------------------
// recurrent step convolution: out[co, p] = pre[co, p] + sum_{ci, 3x3} w[co, ci, tap] * in[ci, p + tap]
// for the hidden-state halves of the four GRUs in ONE launch.  These problems are tiny (21 M MAC per
// level) and latency-bound, so the decomposition maximises resident warps instead of tile reuse:
// one warp = 8 output channels x 128 pixels (4 per lane) x 4 (8 at Cin 64) input channels; the
// warps that share an output tile reduce through shared memory in a fixed order (deterministic).
// Every CTA has 8 busy warps whatever the level: 4 tiles x 2 k-parts (Cin 8) ... 1 tile x 8 k-parts
// (Cin >= 32); 264 CTAs x 8 warps at 96x192, one wave at two CTAs per SM.
// ---------------------------------------------------------------------------------------------
struct GruConvLevel {
  const float* in;  long long in_cs;        // [cin] planes of h*w
  const float* w;   long long w_co;         // w[co * w_co + ci * 9 + tap]   (h-half of the conv weight)
  const float* pre; float* out; long long out_cs;
  double* stats;                            // (sum, sum^2) per group of stats_group output channels
  int cin, cout, h, w_, stats_group;
  int ksplit;                               // warps sharing one output tile: min(cin / 4, 8)
  int ci_per_warp;                          // cin / ksplit (4 or 8)
  int px_groups;                            // CTAs per output-channel chunk
  int cta_begin;
};
struct GruConvArgs { GruConvLevel l[4]; };

constexpr int kGcWarps = 8, kGcCo = 8, kGcPx = 4, kGcTilePx = 32 * kGcPx, kGcCi = 4;
constexpr int kGcThreads = kGcWarps * 32;

// kAligned: every level's width is a multiple of 4, so a lane's 4 pixels sit in one row at a
// 16-byte aligned address: 3 vector + 6 scalar loads per input channel instead of 36 predicated ones.
// kCoherent: inputs / addends were written earlier in the SAME kernel by other CTAs (persistent
// recurrence): they are read with ld.global.cg (L2, coherent) instead of the read-only path.
template <bool kCoherent> __device__ __forceinline__ float ldf(const float* p) { return kCoherent ? __ldcg(p) : __ldg(p); }
template <bool kCoherent> __device__ __forceinline__ float4 ldf4(const float4* p) { return kCoherent ? __ldcg(p) : __ldg(p); }

template <bool kAligned, bool kCoherent>
__device__ __forceinline__ void gru_conv_unit(const GruConvLevel& L, int cta, float* wsm,
                                              float (*part)[kGcCo][kGcTilePx], double (*red)[kGcWarps], bool stage_weights) {
  const int co0 = (cta / L.px_groups) * kGcCo;
  const int tiles_per_cta = kGcWarps / L.ksplit;
  const int tile0 = (cta % L.px_groups) * tiles_per_cta;
  const int npx = L.h * L.w_;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  // weights of this CTA's 8 output channels -> smem [ci][tap][co].  All loads are issued before the
  // first store (one round trip to L2 instead of one per loop iteration).
  if (stage_weights) {
    const int kr = L.cin * 9;                                   // contiguous run per output channel
    const int per_co = (kr + kGcWarps * 32 - 1) / (kGcWarps * 32);   // <= 5 (cin 64)
    constexpr int kMaxPerCo = (64 * 9 + kGcWarps * 32 - 1) / (kGcWarps * 32);
    float t[kGcCo][kMaxPerCo];
#pragma unroll
    for (int co = 0; co < kGcCo; ++co) {
      const float* wp = L.w + (long long)(co0 + co) * L.w_co;
#pragma unroll
      for (int k = 0; k < kMaxPerCo; ++k) {
        const int r = tid + k * (kGcWarps * 32);
        t[co][k] = (k < per_co && r < kr) ? __ldg(wp + r) : 0.0f;
      }
    }
#pragma unroll
    for (int co = 0; co < kGcCo; ++co)
#pragma unroll
      for (int k = 0; k < kMaxPerCo; ++k) {
        const int r = tid + k * (kGcWarps * 32);
        if (k < per_co && r < kr) wsm[r * kGcCo + co] = t[co][k];
      }
  }

  const int tile = tile0 + warp / L.ksplit, kpart = warp % L.ksplit;
  int base[kGcPx]; unsigned mask[kGcPx];
#pragma unroll
  for (int j = 0; j < kGcPx; ++j) {
    const int p = tile * kGcTilePx + lane * kGcPx + j;
    const bool ok = p < npx;
    const int y = ok ? p / L.w_ : 0, x = ok ? p - y * L.w_ : 0;
    base[j] = y * L.w_ + x;
    unsigned m = 0;
#pragma unroll
    for (int t = 0; t < 9; ++t) {
      const int iy = y + t / 3 - 1, ix = x + t % 3 - 1;
      if (ok && (unsigned)iy < (unsigned)L.h && (unsigned)ix < (unsigned)L.w_) m |= 1u << t;
    }
    mask[j] = m;
  }
  float acc[kGcCo][kGcPx];
#pragma unroll
  for (int i = 0; i < kGcCo; ++i)
#pragma unroll
    for (int j = 0; j < kGcPx; ++j) acc[i][j] = 0.0f;
  pdl_wait();                 // inputs, addends and GroupNorm sums of the predecessor kernel are visible from here on
  __syncthreads();

  // input taps of one input channel for this lane's 4 pixels (predicated: zero padding)
  auto load_taps = [&](int ci, float (&v)[kGcPx][9]) {
    const float* ip = L.in + ci * L.in_cs;
#pragma unroll
    for (int j = 0; j < kGcPx; ++j)
#pragma unroll
      for (int t = 0; t < 9; ++t)
        v[j][t] = ((mask[j] >> t) & 1u) ? ldf<kCoherent>(ip + base[j] + (t / 3 - 1) * L.w_ + (t % 3 - 1)) : 0.0f;
  };
  auto fma_taps = [&](int ci, const float (&v)[kGcPx][9]) {
#pragma unroll
    for (int t = 0; t < 9; ++t) {
      const float4 w0 = *reinterpret_cast<const float4*>(&wsm[(ci * 9 + t) * kGcCo]);
      const float4 w1 = *reinterpret_cast<const float4*>(&wsm[(ci * 9 + t) * kGcCo + 4]);
      const float wv[kGcCo] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
#pragma unroll
      for (int i = 0; i < kGcCo; ++i)
#pragma unroll
        for (int j = 0; j < kGcPx; ++j) acc[i][j] = fmaf(wv[i], v[j][t], acc[i][j]);
    }
  };
  if constexpr (kAligned) {
    const int pl = tile * kGcTilePx + lane * kGcPx;
    const bool ok = pl < npx;
    const int y = ok ? pl / L.w_ : 0, x = ok ? pl - y * L.w_ : 0;
    const bool rowok[3] = {ok && y > 0, ok, ok && y + 1 < L.h};
    const bool lok = x > 0, rok = x + kGcPx < L.w_;
    const int ci0 = kpart * L.ci_per_warp;
    // rows[dy][0..5] = in[y+dy-1][x-1 .. x+4]
    auto load_rows = [&](int ci, float (&r)[3][6]) {
      const float* ip = L.in + ci * L.in_cs + (y - 1) * L.w_ + x;
#pragma unroll
      for (int dy = 0; dy < 3; ++dy) {
        const float* rp = ip + dy * L.w_;
        float4 m = make_float4(0.f, 0.f, 0.f, 0.f);
        float lft = 0.f, rgt = 0.f;
        if (rowok[dy]) {
          m = ldf4<kCoherent>(reinterpret_cast<const float4*>(rp));
          if (lok) lft = ldf<kCoherent>(rp - 1);
          if (rok) rgt = ldf<kCoherent>(rp + kGcPx);
        }
        r[dy][0] = lft; r[dy][1] = m.x; r[dy][2] = m.y; r[dy][3] = m.z; r[dy][4] = m.w; r[dy][5] = rgt;
      }
    };
    // packed accumulation: (co, co+1) pairs per FFMA2, weight pairs straight from the 128-bit smem loads
    u64 acc2[kGcCo / 2][kGcPx];
#pragma unroll
    for (int i = 0; i < kGcCo / 2; ++i)
#pragma unroll
      for (int j = 0; j < kGcPx; ++j) acc2[i][j] = 0ULL;
    auto fma_rows = [&](int ci, const float (&r)[3][6]) {
      u64 rr[3][6];
#pragma unroll
      for (int dy = 0; dy < 3; ++dy)
#pragma unroll
        for (int c = 0; c < 6; ++c) rr[dy][c] = pk(r[dy][c], r[dy][c]);
#pragma unroll
      for (int t = 0; t < 9; ++t) {
        const ulonglong2 w0 = *reinterpret_cast<const ulonglong2*>(&wsm[(ci * 9 + t) * kGcCo]);
        const ulonglong2 w1 = *reinterpret_cast<const ulonglong2*>(&wsm[(ci * 9 + t) * kGcCo + 4]);
        const u64 wv[kGcCo / 2] = {w0.x, w0.y, w1.x, w1.y};
#pragma unroll
        for (int i = 0; i < kGcCo / 2; ++i)
#pragma unroll
          for (int j = 0; j < kGcPx; ++j) acc2[i][j] = ffma2(wv[i], rr[t / 3][j + t % 3], acc2[i][j]);
      }
    };
    // 3-slot register ring: the rows of channels c+1, c+2 are in flight while channel c's 288 FMAs issue
    float ring[3][3][6];
#pragma unroll
    for (int c = 0; c < 2; ++c)
      if (c < L.ci_per_warp) load_rows(ci0 + c, ring[c]);
#pragma unroll 1
    for (int c = 0; c < L.ci_per_warp; c += 3) {
#pragma unroll
      for (int u = 0; u < 3; ++u) {
        if (c + u + 2 < L.ci_per_warp) load_rows(ci0 + c + u + 2, ring[(u + 2) % 3]);
        if (c + u < L.ci_per_warp) fma_rows(ci0 + c + u, ring[u]);
      }
    }
#pragma unroll
    for (int i = 0; i < kGcCo / 2; ++i)
#pragma unroll
      for (int j = 0; j < kGcPx; ++j) upk(acc2[i][j], acc[2 * i][j], acc[2 * i + 1][j]);
  } else {
  // two input channels per iteration, ping-pong registers: the next channel's taps are in flight
  // while the current channel's 288 FMAs issue
  float va[kGcPx][9], vb[kGcPx][9];
  const int ci0 = kpart * L.ci_per_warp;
  load_taps(ci0, va);
  for (int c = 0; c < L.ci_per_warp; c += 2) {
    load_taps(ci0 + c + 1, vb);
    fma_taps(ci0 + c, va);
    if (c + 2 < L.ci_per_warp) load_taps(ci0 + c + 2, va);
    fma_taps(ci0 + c + 1, vb);
  }
  }

  // This thread's share of the CTA's outputs: output o = tid + NT*r is (tile-in-CTA q/8, channel q%8,
  // pixel tid%128) with q = r*NT/128 + tid/128, so the index math is compile time up to tid.
  // The addends are fetched now, so the loads fly during the partial-sum exchange.
  constexpr int kRows = kGcThreads / kGcTilePx;              // (tile, channel) rows covered per pass: 2
  constexpr int kMaxOut = kGcWarps * kGcCo / kRows;          // outputs per thread when ksplit == 1: 32
  const int nrows = tiles_per_cta * kGcCo;
  const int px_l = tid & (kGcTilePx - 1), row_l = tid / kGcTilePx;
  float pre[kMaxOut / 2];                                     // ksplit >= 2 in every configuration
#pragma unroll
  for (int r = 0; r < kMaxOut / 2; ++r) {
    const int q = r * kRows + row_l;
    const int p = (tile0 + (q >> 3)) * kGcTilePx + px_l;
    pre[r] = (q < nrows && p < npx) ? ldf<kCoherent>(L.pre + (long long)(co0 + (q & 7)) * L.out_cs + p) : 0.0f;
  }
#pragma unroll
  for (int i = 0; i < kGcCo; ++i)
    *reinterpret_cast<float4*>(&part[warp][i][lane * kGcPx]) = make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
  __syncthreads();

  // fixed-order reduction over the k-parts, epilogue, GroupNorm sums
  float ssum = 0.0f, ssq = 0.0f;
  auto epilogue = [&](auto ks_tag) {
    constexpr int KS = decltype(ks_tag)::value;
#pragma unroll
    for (int r = 0; r < kMaxOut / KS; ++r) {
      const int q = r * kRows + row_l;
      const int ts = q >> 3, i = q & 7;
      const int p = (tile0 + ts) * kGcTilePx + px_l;
      if (p < npx) {
        float sum = 0.0f;
#pragma unroll
        for (int kp = 0; kp < KS; ++kp) sum += part[ts * KS + kp][i][px_l];
        const float val = sum + pre[r];
        L.out[(long long)(co0 + i) * L.out_cs + p] = val;
        ssum += val; ssq += val * val;
      }
    }
  };
  if (L.ksplit == 2) epilogue(std::integral_constant<int, 2>{});
  else if (L.ksplit == 4) epilogue(std::integral_constant<int, 4>{});
  else epilogue(std::integral_constant<int, 8>{});
  double ds = ssum, dq = ssq;
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) { ds += __shfl_xor_sync(0xffffffffu, ds, off); dq += __shfl_xor_sync(0xffffffffu, dq, off); }
  if (lane == 0) { red[0][warp] = ds; red[1][warp] = dq; }
  __syncthreads();
  if (tid == 0) {
    double s = 0.0, q = 0.0;
    for (int i = 0; i < kGcWarps; ++i) { s += red[0][i]; q += red[1][i]; }
    const int grp = co0 / L.stats_group;
    atomicAdd(L.stats + 2 * grp, s);
    atomicAdd(L.stats + 2 * grp + 1, q);
  }

}

template <bool kAligned>
__global__ void __launch_bounds__(kGcThreads, 2)
gru_conv_kernel(const __grid_constant__ GruConvArgs a) {
  extern __shared__ __align__(16) float gc_smem[];
  float* wsm = gc_smem;                                                                  // [ci][tap][co]  (<= 18 KB)
  float (*part)[kGcCo][kGcTilePx] = reinterpret_cast<float (*)[kGcCo][kGcTilePx]>(gc_smem + 64 * 9 * kGcCo);   // 32 KB
  __shared__ double red[2][kGcWarps];
  pdl_release();
  int li = 0;
#pragma unroll
  for (int k = 1; k < 4; ++k) if ((int)blockIdx.x >= a.l[k].cta_begin) li = k;
  // Coherent (L2) loads for everything the predecessor kernel wrote: under programmatic dependent launch this
  // kernel's CTAs become resident (L1 invalidated) BEFORE the predecessor finishes, and the predecessor's own
  // read-then-overwrite of the same addresses (gates / output conv are updated in place) can leave stale lines in
  // the L1 of a shared SM, which ld.global.nc would hit after the wait.
  gru_conv_unit<kAligned, true>(a.l[li], blockIdx.x - a.l[li].cta_begin, wsm, part, red, true);
}


static int gru_conv_launch(const GruConvArgs& c, int ctas, cudaStream_t st, const char* what, bool pdl) {
  bool aligned = true;
  for (int l = 0; l < 4; ++l) aligned = aligned && (c.l[l].w_ % kGcPx == 0);
  constexpr size_t smem = (64 * 9 * kGcCo + kGcWarps * kGcCo * kGcTilePx) * sizeof(float);
  static thread_local int ready_dev = -1;
  int dev = 0;
  cudaGetDevice(&dev);
  if (ready_dev != dev) {
    cudaFuncSetAttribute(gru_conv_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaFuncSetAttribute(gru_conv_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    ready_dev = dev;
  }
  if (aligned) launch_chain(gru_conv_kernel<true>, ctas, kGcThreads, smem, st, pdl, c);
  else launch_chain(gru_conv_kernel<false>, ctas, kGcThreads, smem, st, pdl, c);
  return check_launch(what);
}

static int gru_conv_fill(GruConvLevel& g, int cta_begin) {
  g.ksplit = g.cin / kGcCi < kGcWarps ? g.cin / kGcCi : kGcWarps;   // cin 8 -> 2, 16 -> 4, >= 32 -> 8
  g.ci_per_warp = g.cin / g.ksplit;
  const int tiles = (g.h * g.w_ + kGcTilePx - 1) / kGcTilePx;
  const int tiles_per_cta = kGcWarps / g.ksplit;
  g.px_groups = (tiles + tiles_per_cta - 1) / tiles_per_cta;
  g.cta_begin = cta_begin;
  return cta_begin + g.px_groups * (g.cout / kGcCo);
}

// conv problem over [Cin][Di][Hi][Wi] -> [Cout][Do][Ho][Wo], 2-D 3x3 taps applied per plane
static ConvProblem plane_conv(const float* in, int Cin, int Di, int Hi, int Wi, const float* w, long long w_co, long long w_ci,
                              float* out, int Cout, int Do, int Ho, int Wo, int stride) {
  ConvProblem p;
  conv_problem_defaults(p);
  p.in = in; p.w = w; p.out = out;
  p.Cin = Cin; p.Cout = Cout;
  p.Di = Di; p.Hi = Hi; p.Wi = Wi; p.Do = Do; p.Ho = Ho; p.Wo = Wo;
  p.w_co_stride = w_co; p.w_ci_stride = w_ci;
  p.q2i_mul[0] = 1; p.q2i_mul[1] = stride; p.q2i_mul[2] = stride;
  p.q2i_add[0] = 0; p.q2i_add[1] = -1; p.q2i_add[2] = -1;
  conv_taps_dense(p, false);
  return p;
}

template <class T>
static int launch_one(ConvProblem p, cudaStream_t st, const char* what) {
  conv_finalize(p);
  ConvGroup g{};
  g.p[0] = p; g.n = 1;
  return conv_launch<T>(g, st, what);
}

// Batched 2-D convs: the register-tiled direct kernel when rows are 16-byte aligned, else the engine.
static int launch_plane_conv(const ConvProblem& p, int stride, cudaStream_t st, const char* what);

static int launch_by_cout(ConvProblem p, cudaStream_t st, const char* what) {
  if (p.Cout >= 64) return launch_one<Tile64>(p, st, what);
  if (p.Cout >= 32) return launch_one<Tile32>(p, st, what);
  if (p.Cout >= 16) return launch_one<Tile16>(p, st, what);
  return launch_one<Tile8>(p, st, what);
}

static int launch_plane_conv(const ConvProblem& p, int stride, cudaStream_t st, const char* what) {
  DirectConv d{};
  d.in = p.in + (long long)p.in_c_off * p.Di * p.Hi * p.Wi; d.w = p.w; d.scale = p.scale; d.shift = p.shift;
  d.post_add = p.post_add; d.out = p.out;
  d.Cin = p.Cin; d.Cout = p.Cout; d.Di = p.Di; d.Hi = p.Hi; d.Wi = p.Wi; d.Do = p.Do; d.Ho = p.Ho; d.Wo = p.Wo;
  d.w_co = p.w_co_stride; d.w_ci = p.w_ci_stride; d.acc_scale = p.acc_scale; d.relu = p.relu;
  static const bool no_direct = getenv("SATMVS_NO_DIRECT_CONV") != nullptr;
  if (!no_direct && p.pre_add == nullptr && p.stats == nullptr && p.out_c_off == 0 && direct_conv_supported(d, 1, stride))
    return direct_conv_launch(d, 1, stride, st, what);
  return launch_by_cout(p, st, what);
}

// Side streams of the overlapped flow (per host thread and device; created once, live as long as the library): `rec` and
// `lv[1..3]` run the recurrences of levels 0 and 1..3 (highest priority: their 16-CTA clusters must become resident before the
// producers fill the machine); every level has a stream of its own so that it starts as soon as ITS producers are done -- the
// levels' producers form a chain (level l reads the encoder output of level l-1), and level 0, the slowest recurrence, is ready first.
constexpr int kRedMaxChunks = 32;
struct RedSideStream {
  cudaStream_t rec = nullptr, lv[4] = {nullptr, nullptr, nullptr, nullptr};
  cudaEvent_t fork = nullptr, join[4] = {nullptr, nullptr, nullptr, nullptr};
  cudaEvent_t chunk[8][kRedMaxChunks];     // [0..3]: producers of level l, chunk c done; [4..7]: recurrence of level l over chunk c done
  int dev = -1; bool ok = false;
};
static RedSideStream& red_side_stream() {
  static thread_local RedSideStream s[16];
  int dev = 0;
  cudaGetDevice(&dev);
  RedSideStream& r = s[dev & 15];
  if (r.dev != dev) {
    int lo = 0, hi = 0;
    cudaDeviceGetStreamPriorityRange(&lo, &hi);
    bool ok = cudaStreamCreateWithPriority(&r.rec, cudaStreamNonBlocking, hi) == cudaSuccess &&
              cudaEventCreateWithFlags(&r.fork, cudaEventDisableTiming) == cudaSuccess;
    for (int i = 0; i < 4 && ok; ++i) ok = cudaEventCreateWithFlags(&r.join[i], cudaEventDisableTiming) == cudaSuccess;
    for (int i = 1; i < 4 && ok; ++i) ok = cudaStreamCreateWithPriority(&r.lv[i], cudaStreamNonBlocking, hi) == cudaSuccess;
    for (int i = 0; i < 8 && ok; ++i)
      for (int c = 0; c < kRedMaxChunks && ok; ++c) ok = cudaEventCreateWithFlags(&r.chunk[i][c], cudaEventDisableTiming) == cudaSuccess;
    if (!ok) cudaGetLastError();
    r.ok = ok; r.dev = dev;
  }
  return r;
}

}  // namespace satmvs

using namespace satmvs;

static thread_local int g_red_last_path = -1;

extern "C" {

int satmvs_red_last_path(void) { return g_red_last_path; }

int satmvs_red_workspace_layout(int C, int D, int H, int W, size_t* offsets10) {
  SATMVS_REQUIRE(offsets10 && C >= 1 && D >= 1 && H >= 8 && W >= 8 && H % 8 == 0 && W % 8 == 0);
  char* base = nullptr;
  const RedPlan p = red_plan(C, D, H, W, base);
  for (int l = 0; l < 4; ++l) offsets10[l] = (size_t)(reinterpret_cast<char*>(p.lv[l].s) - base);
  for (int i = 0; i < 3; ++i) offsets10[4 + i] = (size_t)(reinterpret_cast<char*>(p.e[i]) - base);
  for (int i = 0; i < 3; ++i) offsets10[7 + i] = (size_t)(reinterpret_cast<char*>(p.u[i]) - base);
  return SATMVS_OK;
}

size_t satmvs_red_workspace_bytes(int C, int D, int H, int W) {
  if (C < 1 || D < 1 || H < 8 || W < 8 || (H % 8) || (W % 8)) return 0;
  return red_plan(C, D, H, W, nullptr).bytes;
}

size_t satmvs_red_pack_bytes(int C) {
  if (C < 1) return 0;
  static char dummy[1];
  return red_plan(C, 1, 8, 8, nullptr, dummy).pack_bytes;
}

int satmvs_red_forward(const satmvs_red_weights* wt, const float* volume, int C, int D, int H, int W,
                       const float* const* state_in, float* const* state_out, float* logits,
                       void* workspace, size_t workspace_bytes, void* stream) {
  return satmvs_red_forward_packed(wt, volume, C, D, H, W, state_in, state_out, logits, workspace, workspace_bytes,
                                   nullptr, 0, nullptr, stream);
}

int satmvs_red_forward_packed(const satmvs_red_weights* wt, const float* volume, int C, int D, int H, int W,
                              const float* const* state_in, float* const* state_out, float* logits,
                              void* workspace, size_t workspace_bytes, void* pack, size_t pack_bytes,
                              unsigned long long* pack_tag, void* stream) {
  SATMVS_REQUIRE(wt && volume && logits && workspace);
  SATMVS_REQUIRE(C >= 1 && D >= 1 && H >= 8 && W >= 8 && H % 8 == 0 && W % 8 == 0);
  SATMVS_REQUIRE((pack == nullptr) == (pack_tag == nullptr));
  SATMVS_CHECK_ASYNC();
  cudaStream_t st = (cudaStream_t)stream;
  RedPlan P = red_plan(C, D, H, W, reinterpret_cast<char*>(workspace), reinterpret_cast<char*>(pack));
  if (int* ae = async_error_devptr()) P.umma_err = ae;
  SATMVS_REQUIRE(workspace_bytes >= P.bytes);
  SATMVS_REQUIRE(pack == nullptr || (pack_bytes >= P.pack_bytes && (reinterpret_cast<uintptr_t>(pack) & 255) == 0));
  int rc;
#define RUN(x) do { rc = (x); if (rc) return rc; } while (0)

  cudaMemsetAsync(P.stats, 0, (size_t)D * 4 * 3 * 2 * sizeof(double), st);
  for (int l = 0; l < 4; ++l) {   // slot 0 of the state history = the initial hidden state (zeros, module.py:617-620)
    RedLevel& L = P.lv[l];
    const size_t px = (size_t)L.h * L.w;
    if (state_in && state_in[l])
      cudaMemcpy2DAsync(L.s, (size_t)(D + 1) * px * 4, state_in[l], px * 4, px * 4, L.ch, cudaMemcpyDeviceToDevice, st);
    else
      cudaMemset2DAsync(L.s, (size_t)(D + 1) * px * 4, 0, px * 4, L.ch, st);
  }

  // ---- A. batched over all planes ----
  const int ech[4] = {C, 16, 32, 64};
  const float* xin[4] = {volume, P.e[0], P.e[1], P.e[2]};
  // encoder i: ConvReLU stride 2 (module.py:627-629); conv1 sees -cost.  Run by the direct kernel unless the level's
  // tensor-core launch below takes it as a third head.
  auto run_encoder = [&](int i) -> int {
    ConvProblem p = plane_conv(xin[i], ech[i], D, H >> i, W >> i, wt->conv_w[i], (long long)ech[i] * 9, 9,
                               P.e[i], ech[i + 1], D, H >> (i + 1), W >> (i + 1), 2);
    p.Qd = D; p.Qh = H >> (i + 1); p.Qw = W >> (i + 1);
    p.relu = 1;
    p.acc_scale = (i == 0) ? -1.0f : 1.0f;
    ProfScope prof(kProfConvBatched, st);
    return launch_plane_conv(p, 2, st, "red encoder");
  };
  static const bool no_umma = getenv("SATMVS_NO_UMMA") != nullptr;
  // x-halves of the GRU convolutions, bias folded in (module.py:29-30, :44-45).  Gate and output x-halves share their input
  // (and so does the stride-2 encoder that feeds the next level): one tensor-core launch with two or three heads
  // (umma_conv.cuh); level 4's fused N does not fit (192 columns of packed weights) and takes one launch per head.
  struct XPlan { UmmaConvPlan u[3]; int n; bool sig[3]; bool enc_done; const char* what[3]; } xp[4];
  bool all_umma = !no_umma;
  for (int l = 0; l < 4; ++l) {
    RedLevel& L = P.lv[l];
    XPlan& X = xp[l];
    X.n = 0; X.enc_done = false;
    if (no_umma) continue;
    const long long kin = (long long)(L.cx + L.ch) * 9;
    const float sgn = (l == 0) ? -1.0f : 1.0f;
    const long long pl = (long long)D * L.h * L.w;
    // heads: 0 = gates, 1 = output, 2 = the stride-2 encoder of the next level (levels 0..2)
    UmmaPackHead wh[3] = {{wt->gate_w[l], kin, 9, 2 * L.ch, 0}, {wt->out_w[l], kin, 9, L.ch, 0}, {nullptr, 0, 0, 0, 0}};
    UmmaHead oh[3] = {{nullptr, wt->gate_b[l], L.gx, 2 * L.ch, 0, sgn, 0, 1}, {nullptr, wt->out_b[l], L.ox, L.ch, 0, sgn, 0, 1}, {}};
    const int nh = l < 3 ? 3 : 2;
    if (l < 3) {
      wh[2] = UmmaPackHead{wt->conv_w[l], (long long)ech[l] * 9, 9, ech[l + 1], 0};
      oh[2] = UmmaHead{nullptr, nullptr, P.e[l], ech[l + 1], 0, sgn, 1, 2};
    }
    // group the heads into as few launches as fit: [G O E], else [G O] [E], else [G] [O] [E]; packed weights back to back
    const int splits[3][3] = {{nh, 0, 0}, {2, nh - 2, 0}, {1, 1, nh - 2}};
    const char* names[3][3] = {{"red gate/output x-halves + encoder (tcgen05)", "", ""},
                               {"red gate/output x-halves (tcgen05)", "red encoder (tcgen05)", ""},
                               {"red gate x-half (tcgen05)", "red output x-half (tcgen05)", "red encoder (tcgen05)"}};
    for (int t = 0; t < 3 && X.n == 0; ++t) {
      size_t off = 0;
      int h0 = 0, n = 0;
      bool ok = true;
      for (int g = 0; g < 3 && ok; ++g) {
        const int cnt = splits[t][g];
        if (cnt == 0) continue;
        ok = off < P.wpack_bytes[l] &&
             umma_conv_plan(X.u[n], xin[l], pl, L.cx, D, L.h, L.w, cnt, wh + h0, oh + h0, P.wpack[l] + off, P.wpack_bytes[l] - off,
                            1, /*perf_rules=*/h0 < 2);
        if (ok) {
          X.sig[n] = h0 < 2;                                 // launches that write gate / output x-halves count towards ready[]
          X.what[n] = names[t][g];
          off += (X.u[n].wpack_bytes + 255) / 256 * 256;
          h0 += cnt; ++n;
        }
      }
      if (ok) { X.n = n; X.enc_done = l < 3; }
    }
    if (X.n == 0) all_umma = false;
  }
  // Packed tensor-core weights (umma x-half packs + recurrence packs) are a pure function of the weights and of the plan: with a
  // caller-owned pack region they are written once and reused while *pack_tag matches the plan's signature (the caller zeroes
  // the tag when the weights change)
  unsigned long long sig = 1469598103934665603ULL;
  {
    auto mix = [&](unsigned long long v) { sig ^= v; sig *= 1099511628211ULL; };
    mix((unsigned)C); mix((unsigned)D); mix((unsigned)H); mix((unsigned)W);
    for (int l = 0; l < 4; ++l) {
      mix((unsigned)xp[l].n);
      for (int i = 0; i < xp[l].n; ++i) { mix((unsigned)xp[l].u[i].conv.NP); mix((unsigned)xp[l].u[i].conv.nheads); }
    }
    mix(getenv("SATMVS_RED_NO_TC") ? 2u : 3u);
    if (sig == 0) sig = 1;
  }
  const bool do_pack = !(pack_tag != nullptr && *pack_tag == sig);
  // launches the batched convs of level l for planes [d0, d0 + np) (np = 0: all planes, weights packed on the way)
  auto run_xhalf = [&](int l, int d0, int np, bool pack, int* ready, cudaStream_t xs) -> int {
    RedLevel& L = P.lv[l];
    XPlan& X = xp[l];
    if (l < 3 && !X.enc_done) { int r0 = run_encoder(l); if (r0) return r0; }
    if (X.n > 0) {
      ProfScope prof(kProfConvBatched, xs);
      for (int i = 0; i < X.n; ++i) {
        int r0 = umma_conv_launch(X.u[i], P.umma_err, xs, X.what[i], pack, d0, np, X.sig[i] ? ready : nullptr);
        if (r0) return r0;
      }
      return SATMVS_OK;
    }
    const long long kin = (long long)(L.cx + L.ch) * 9;
    ConvProblem g = plane_conv(xin[l], L.cx, D, L.h, L.w, wt->gate_w[l], kin, 9, L.gx, 2 * L.ch, D, L.h, L.w, 1);
    g.Qd = D; g.Qh = L.h; g.Qw = L.w;
    g.shift = wt->gate_b[l];
    g.acc_scale = (l == 0) ? -1.0f : 1.0f;
    { ProfScope prof(kProfConvBatched, st); int r0 = launch_plane_conv(g, 1, st, "red gate x-half"); if (r0) return r0; }
    ConvProblem o = plane_conv(xin[l], L.cx, D, L.h, L.w, wt->out_w[l], kin, 9, L.ox, L.ch, D, L.h, L.w, 1);
    o.Qd = D; o.Qh = L.h; o.Qw = L.w;
    o.shift = wt->out_b[l];
    o.acc_scale = (l == 0) ? -1.0f : 1.0f;
    { ProfScope prof(kProfConvBatched, st); int r0 = launch_plane_conv(o, 1, st, "red output x-half"); if (r0) return r0; }
    return SATMVS_OK;
  };

  // ---- A + B overlapped (default): the tensor-core recurrence (64 SMs) is launched FIRST on a side stream and trails the
  // batched convs, which run chunk by chunk on the remaining SMs and count finished planes in P.ready ----
  bool persistent = false;
  const bool no_cluster = getenv("SATMVS_RED_NO_CLUSTER") != nullptr;   // read per call: tests toggle it
  const bool no_tc = getenv("SATMVS_RED_NO_TC") != nullptr;
  const bool no_overlap = getenv("SATMVS_RED_NO_OVERLAP") != nullptr;
  static const bool tc_dbg = getenv("SATMVS_RED_DEBUG") != nullptr;
  long long* dbg = tc_dbg ? reinterpret_cast<long long*>(P.cl_flags + 4 * 2 * 32) : nullptr;
  auto tc_args = [&](bool with_ready) {
    TcArgs ta{};
    ta.d_begin = 0; ta.d_end = D;
    for (int l = 0; l < 4; ++l) {
      RedLevel& L = P.lv[l];
      TcLevel& R = ta.l[l];
      const long long px = (long long)L.h * L.w;
      R.s = L.s; R.s_cs = (long long)(D + 1) * px;
      R.gx = L.gx; R.g_cs = (long long)D * px;
      R.ox = L.ox; R.o_cs = (long long)D * px;
      R.rn_w = wt->rn_w[l]; R.rn_b = wt->rn_b[l]; R.un_w = wt->un_w[l]; R.un_b = wt->un_b[l];
      R.on_w = wt->on_w[l]; R.on_b = wt->on_b[l];
      R.inv_n = 1.0 / ((double)L.ch * (double)px);
      R.ch = L.ch; R.h = L.h; R.w = L.w; R.px = (int)px;
      if (with_ready) {
        R.ready = P.ready + (size_t)l * D;
        R.expected = 0;
        for (int i = 0; i < xp[l].n; ++i) if (xp[l].sig[i]) R.expected += (int)xp[l].u[i].grid.x;
      }
    }
    return ta;
  };
  const float* gwh[4]; const float* owh[4]; long long wco[4];
  for (int l = 0; l < 4; ++l) {
    gwh[l] = wt->gate_w[l] + (size_t)P.lv[l].cx * 9; owh[l] = wt->out_w[l] + (size_t)P.lv[l].cx * 9;
    wco[l] = (long long)(P.lv[l].cx + P.lv[l].ch) * 9;
  }
  auto tc_report = [&]() {
    if (!tc_dbg) return;
    long long h[4 * 20] = {};
    cudaStreamSynchronize(st);
    cudaMemcpy(h, dbg, sizeof(h), cudaMemcpyDeviceToHost);
    for (int c = 0; c < 4; ++c) {
      fprintf(stderr, "red_tc: level %d kcycles per phase slot:", c);
      for (int i = 0; i < 20; ++i) fprintf(stderr, " %.0f", h[c * 20 + i] * 1e-3);
      fprintf(stderr, "\n");
    }
  };
  // ---- C. decoder over planes [d0, d0 + np): U_l = relu(convT_s2(U_{l+1})) + S_l  (module.py:633-642) ----
  auto run_decoder = [&](int d0, int np, cudaStream_t st) -> int {
  const float* up_in = P.lv[3].s;
  for (int l = 2; l >= 0; --l) {
    RedLevel& L = P.lv[l];          // output level
    RedLevel& Lin = P.lv[l + 1];
    ConvGroup g{};
    int n = 0;
    for (int py = 0; py < 2; ++py)
      for (int px = 0; px < 2; ++px) {
        ConvProblem p;
        conv_problem_defaults(p);
        p.in = up_in; p.w = wt->upconv_w[l]; p.out = P.u[l];
        p.Cin = Lin.ch; p.Cout = L.ch;
        p.Di = D + 1; p.Hi = Lin.h; p.Wi = Lin.w; p.Do = D + 1; p.Ho = L.h; p.Wo = L.w;
        p.w_ci_stride = (long long)L.ch * 9; p.w_co_stride = 9;       // ConvTranspose2d weight [Cin][Cout][3][3]
        p.Qd = np; p.Qh = Lin.h; p.Qw = Lin.w;
        p.q2i_add[0] = 1 + d0; p.q2o_add[0] = 1 + d0;                  // slots d0+1..d0+np
        p.q2o_mul[1] = 2; p.q2o_mul[2] = 2; p.q2o_add[1] = py; p.q2o_add[2] = px;
        conv_taps_deconv_class(p, false, 0, py, px);
        p.relu = 1;
        p.post_add = L.s;
        conv_finalize(p);
        g.p[n++] = p;
      }
    g.n = n;
    {
      ProfScope prof(kProfDecoder, st);
      // slots d0+1..d0+np of the [C][D+1][h][w] tensors are np contiguous planes starting 1+d0 planes into each channel
      DirectDeconv dd{};
      const long long pin = (long long)Lin.h * Lin.w, pout = (long long)L.h * L.w;
      dd.in = up_in + (1 + d0) * pin; dd.in_cs = (long long)(D + 1) * pin;
      dd.w = wt->upconv_w[l]; dd.post_add = L.s + (1 + d0) * pout; dd.out = P.u[l] + (1 + d0) * pout; dd.out_cs = (long long)(D + 1) * pout;
      dd.Cin = Lin.ch; dd.Cout = L.ch; dd.Dn = np; dd.Hi = Lin.h; dd.Wi = Lin.w; dd.relu = 1;
      static const bool no_direct = getenv("SATMVS_NO_DIRECT_CONV") != nullptr;
      if (!no_direct && direct_deconv_supported(dd)) {
        RUN(direct_deconv_launch(dd, st, "red upconv (direct)"));
      } else {
        if (L.ch >= 32) RUN(conv_launch<Tile32>(g, st, "red upconv")); else if (L.ch >= 16) RUN(conv_launch<Tile16>(g, st, "red upconv"));
        else RUN(conv_launch<Tile8>(g, st, "red upconv"));
      }
    }
    up_in = P.u[l];
  }
  {  // upconv2d: ConvTranspose2d(8, 1, k3, stride 1, pad 1) with bias (module.py:610, :643): out[o] = sum_k in[o+1-k] w[k]
    ConvProblem p;
    conv_problem_defaults(p);
    p.in = P.u[0]; p.w = wt->upconv2d_w; p.out = logits + (size_t)d0 * H * W;
    p.Cin = 8; p.Cout = 1;
    p.Di = D + 1; p.Hi = H; p.Wi = W; p.Do = np; p.Ho = H; p.Wo = W;
    p.w_ci_stride = 9; p.w_co_stride = 9;
    p.Qd = np; p.Qh = H; p.Qw = W;
    p.q2i_add[0] = 1 + d0;
    int n = 0;
    for (int ky = 0; ky < 3; ++ky)
      for (int kx = 0; kx < 3; ++kx) { p.tap_dz[n] = 0; p.tap_dy[n] = 1 - ky; p.tap_dx[n] = 1 - kx; p.tap_w[n] = ky * 3 + kx; ++n; }
    p.ntaps = n;
    p.shift = wt->upconv2d_b;
    // out[o] = sum_k in[o+1-k] w[k] = sum_k' in[o-1+k'] w[8-k']: a 3x3 correlation with mirrored taps
    DirectConv dc{};
    dc.in = P.u[0] + (size_t)(1 + d0) * H * W;        // slot 1 + d0 of the [8][D+1][H][W] decoder tensor
    dc.w = wt->upconv2d_w; dc.shift = wt->upconv2d_b; dc.out = logits + (size_t)d0 * H * W;
    dc.Cin = 8; dc.Cout = 1; dc.Di = np; dc.Hi = H; dc.Wi = W; dc.Do = np; dc.Ho = H; dc.Wo = W;
    dc.w_co = 9; dc.w_ci = 9; dc.acc_scale = 1.0f; dc.relu = 0; dc.flip = 1;
    static const bool no_direct = getenv("SATMVS_NO_DIRECT_CONV") != nullptr;
    ProfScope prof(kProfDecoder, st);
    dc.in_cs = (long long)(D + 1) * H * W;      // the input tensor has D+1 planes per channel
    if (!no_direct && direct_conv3d_c1_supported(dc)) {
      RUN(direct_conv3d_c1_launch(dc, st, "red upconv2d (direct, 1 channel)", 1));
    } else if (!no_direct && direct_conv_supported(dc, 1, 1)) {
      RUN(direct_conv_launch_cs(dc, (long long)(D + 1) * H * W, st, "red upconv2d (direct)"));
    } else {
      RUN(launch_one<Tile8>(p, st, "red upconv2d"));
    }
  }
  return SATMVS_OK;
  };
  bool decoded = false;
  bool xhalf_done = false, xpacked = false;
  if (all_umma && !no_cluster && !no_tc && !no_overlap) {
    RedSideStream& side = red_side_stream();
    // Event-ordered pipeline (no kernel ever spins on another one): the caller's stream produces the x-halves chunk by chunk,
    // the side stream runs the recurrence over chunk c as soon as the producers of chunk c are done (state carried through
    // the history slots), i.e. concurrently with the producers of chunk c + 1 on the 84 SMs the recurrence leaves free.
    // Chunk schedule: the producers of a chunk cost ~125 us of launch latencies + ~7.5 us per plane, the recurrence ~21 us per
    // plane + ~30 us per launch; with one recurrence stream per level (below) equal chunks of 16 planes are best: a big last chunk
    // makes the deepest level wait for the very last producer launch
    // (measured at cfg-2, profiles/r02_red_overlap_notes.md: [8]x8 1.99 ms, [16]x4 1.95 ms per forward against 2.03 sequential)
    static const int kChunk = getenv("SATMVS_RED_CHUNK") ? atoi(getenv("SATMVS_RED_CHUNK")) : 16;
    static const bool chunked_decoder = getenv("SATMVS_RED_NO_CHUNKED_DECODER") == nullptr;
    int cstart[kRedMaxChunks + 1], nchunks = 0;
    cstart[0] = 0;
    if (const char* sched = getenv("SATMVS_RED_SCHED")) {          // tuning: "8,16,16" = chunk sizes, the rest in a last chunk
      for (const char* q = sched; *q && nchunks < kRedMaxChunks - 1 && cstart[nchunks] < D;) {
        const int take = atoi(q);
        if (take <= 0) break;
        cstart[nchunks + 1] = cstart[nchunks] + (take < D - cstart[nchunks] ? take : D - cstart[nchunks]);
        ++nchunks;
        while (*q && *q != ',') ++q;
        if (*q == ',') ++q;
      }
      if (cstart[nchunks] < D) { cstart[nchunks + 1] = D; ++nchunks; }
    } else if (kChunk > 0 && D >= 2 * kChunk) {
      while (cstart[nchunks] < D && nchunks < kRedMaxChunks) {
        const int left = D - cstart[nchunks];
        const int take = (left < 2 * kChunk || nchunks == kRedMaxChunks - 1) ? left : kChunk;
        cstart[nchunks + 1] = cstart[nchunks] + take;
        ++nchunks;
      }
    }
    if (side.ok && nchunks > 1 && cstart[nchunks] == D) {
      if (do_pack)
        for (int l = 0; l < 4; ++l)
          for (int i = 0; i < xp[l].n; ++i) umma_conv_pack(xp[l].u[i], st);
      xpacked = true;
      cudaStream_t rs[4] = {side.rec, side.lv[1], side.lv[2], side.lv[3]};
      static const bool per_level = getenv("SATMVS_RED_ONE_LAUNCH") == nullptr;   // off: the four levels in one 64-CTA grid
      cudaEventRecord(side.fork, st);
      for (int l = 0; l < (per_level ? 4 : 1); ++l) cudaStreamWaitEvent(rs[l], side.fork, 0);
      TcArgs ta = tc_args(false);
      for (int c = 0; c < nchunks; ++c) {
        const int d0 = cstart[c], np = cstart[c + 1] - cstart[c];
        if (c == 0 || persistent) {
          ta.d_begin = d0; ta.d_end = d0 + np;
          bool ran = false;
          if (per_level) {
            for (int l = 0; l < 4; ++l) {
              RUN(run_xhalf(l, d0, np, false, nullptr, st));
              cudaEventRecord(side.chunk[l][c], st);
              cudaStreamWaitEvent(rs[l], side.chunk[l][c], 0);
              ProfScope prof(kProfGruGate, rs[l]);
              RUN(red_tc_launch(ta, gwh, owh, wco, P.tcpack, P.umma_err, dbg, rs[l], &ran, c == 0 && do_pack, l));
              if (!ran) break;
              cudaEventRecord(side.chunk[4 + l][c], rs[l]);
            }
          } else {
            for (int l = 0; l < 4; ++l) RUN(run_xhalf(l, d0, np, false, nullptr, st));
            cudaEventRecord(side.chunk[0][c], st);
            cudaStreamWaitEvent(side.rec, side.chunk[0][c], 0);
            ProfScope prof(kProfGruGate, side.rec);
            RUN(red_tc_launch(ta, gwh, owh, wco, P.tcpack, P.umma_err, dbg, side.rec, &ran, c == 0 && do_pack));
            if (ran) cudaEventRecord(side.chunk[4][c], side.rec);
          }
          if (c == 0) persistent = ran;
          if (!ran) break;                                                 // shape not taken: the sequential flow below finishes the job
        }
      }
      if (persistent && chunked_decoder) {
        // the decoder of chunk c follows the producers of all chunks on the caller's stream and starts when the recurrences have
        // left chunk c behind: only the last chunk's decoder stays on the critical path
        for (int c = 0; c < nchunks; ++c) {
          for (int l = 0; l < (per_level ? 4 : 1); ++l) cudaStreamWaitEvent(st, side.chunk[4 + l][c], 0);
          RUN(run_decoder(cstart[c], cstart[c + 1] - cstart[c], st));
        }
        decoded = true;
      }
      for (int l = 0; l < (per_level ? 4 : 1); ++l) { cudaEventRecord(side.join[l], rs[l]); cudaStreamWaitEvent(st, side.join[l], 0); }
      if (persistent) { g_red_last_path = 3; xhalf_done = true; tc_report(); }
    }
  }
  if (!xhalf_done)
    for (int l = 0; l < 4; ++l) RUN(run_xhalf(l, 0, 0, !xpacked && do_pack, nullptr, st));

  // ---- B. recurrence over planes (when it did not run overlapped above) ----
  {
    if (!persistent && !no_cluster && !no_tc) {
      // the recurrence on the tensor cores, one 16-CTA cluster per level (red_tc.cuh), after the batched convs
      TcArgs ta = tc_args(false);
      ProfScope prof(kProfGruGate, st);
      RUN(red_tc_launch(ta, gwh, owh, wco, P.tcpack, P.umma_err, dbg, st, &persistent, do_pack));
      if (persistent) { g_red_last_path = 2; tc_report(); }
    }
    if (!persistent && !no_cluster) {
      ClArgs ca{};
      ca.D = D;
      for (int l = 0; l < 4; ++l) {
        RedLevel& L = P.lv[l];
        ClLevel& R = ca.l[l];
        const long long px = (long long)L.h * L.w;
        R.s = L.s; R.s_cs = (long long)(D + 1) * px;
        R.gx = L.gx; R.g_cs = (long long)D * px;
        R.ox = L.ox; R.o_cs = (long long)D * px;
        R.rh = L.rh; R.ub = L.ub;
        R.gate_w = wt->gate_w[l] + (size_t)L.cx * 9; R.out_w = wt->out_w[l] + (size_t)L.cx * 9;
        R.w_co = (long long)(L.cx + L.ch) * 9;
        R.rn_w = wt->rn_w[l]; R.rn_b = wt->rn_b[l]; R.un_w = wt->un_w[l]; R.un_b = wt->un_b[l];
        R.on_w = wt->on_w[l]; R.on_b = wt->on_b[l];
        R.inv_n = 1.0 / ((double)L.ch * (double)px);
        R.ch = L.ch; R.h = L.h; R.w = L.w; R.px = (int)px;
      }
      ProfScope prof(kProfGruGate, st);     // one class: the cluster kernel has no per-phase boundary
      RUN(red_cluster_launch(ca, P.cl_flags, st, &persistent));
      if (persistent) g_red_last_path = 1;
    }
  }
  if (!persistent) g_red_last_path = 0;
  static const bool pdl = getenv("SATMVS_RED_NO_PDL") == nullptr;
  for (int d = 0; d < D && !persistent; ++d) {
    GruConvArgs c1{}, c2{};
    GruArgs ga{};
    int total = 0, ctas1 = 0, ctas2 = 0;
    for (int l = 0; l < 4; ++l) {
      RedLevel& L = P.lv[l];
      const size_t px = (size_t)L.h * L.w;
      const long long kin = (long long)(L.cx + L.ch) * 9;
      double* stats = P.stats + ((size_t)d * 4 + l) * 6;
      // P1: gates[d] += conv(h_prev; h-half of gate_conv.weight), GroupNorm sums of the r and u halves
      GruConvLevel& a = c1.l[l];
      a.in = L.s + (size_t)d * px; a.in_cs = (long long)(D + 1) * px;
      a.w = wt->gate_w[l] + (size_t)L.cx * 9; a.w_co = kin;
      a.pre = a.out = L.gx + (size_t)d * px; a.out_cs = (long long)D * px;
      a.stats = stats; a.stats_group = L.ch;
      a.cin = L.ch; a.cout = 2 * L.ch; a.h = L.h; a.w_ = L.w;
      ctas1 = gru_conv_fill(a, ctas1);
      // P2: out[d] += conv(r*h; h-half of output_conv.weight), GroupNorm sums
      GruConvLevel& b = c2.l[l];
      b.in = L.rh; b.in_cs = (long long)px;
      b.w = wt->out_w[l] + (size_t)L.cx * 9; b.w_co = kin;
      b.pre = b.out = L.ox + (size_t)d * px; b.out_cs = (long long)D * px;
      b.stats = stats + 4; b.stats_group = L.ch;
      b.cin = L.ch; b.cout = L.ch; b.h = L.h; b.w_ = L.w;
      ctas2 = gru_conv_fill(b, ctas2);
      GruLevelArgs& e = ga.l[l];
      e.g = L.gx + (size_t)d * px; e.gcs = (long long)D * px;
      e.o = L.ox + (size_t)d * px; e.ocs = (long long)D * px;
      e.hprev = L.s + (size_t)d * px; e.hnext = L.s + (size_t)(d + 1) * px; e.scs = (long long)(D + 1) * px;
      e.rh = L.rh;
      e.rn_w = wt->rn_w[l]; e.rn_b = wt->rn_b[l]; e.un_w = wt->un_w[l]; e.un_b = wt->un_b[l];
      e.on_w = wt->on_w[l]; e.on_b = wt->on_b[l];
      e.stats = stats; e.ch = L.ch; e.px = (int)px; e.begin = total; e.inv_n = 1.0 / ((double)L.ch * (double)px);
      total += L.ch * (int)px;
    }
    ga.total = total;
    // the first gate conv follows the batched launches (plain stream order); everything after it is chained
    { ProfScope prof(kProfGruGate, st); RUN(gru_conv_launch(c1, ctas1, st, "gru_conv_kernel (gates)", pdl && d > 0)); }
    { ProfScope prof(kProfGruPointwise, st);
      launch_chain(gru_reset_kernel, ceil_div(total, 256), 256, 0, st, pdl, ga);
      RUN(check_launch("gru_reset_kernel")); }
    { ProfScope prof(kProfGruOutput, st); RUN(gru_conv_launch(c2, ctas2, st, "gru_conv_kernel (output)", pdl)); }
    { ProfScope prof(kProfGruPointwise, st);
      launch_chain(gru_update_kernel, ceil_div(total, 256), 256, 0, st, pdl, ga);
      RUN(check_launch("gru_update_kernel")); }
  }

  if (!decoded) RUN(run_decoder(0, D, st));
  if (state_out)
    for (int l = 0; l < 4; ++l)
      if (state_out[l]) {
        RedLevel& L = P.lv[l];
        const size_t px = (size_t)L.h * L.w;
        cudaMemcpy2DAsync(state_out[l], px * 4, L.s + (size_t)D * px, (size_t)(D + 1) * px * 4, px * 4, L.ch,
                          cudaMemcpyDeviceToDevice, st);
      }
#undef RUN
  if (pack_tag != nullptr) *pack_tag = sig;
  return check_launch("satmvs_red_forward");
}

}  // extern "C"
