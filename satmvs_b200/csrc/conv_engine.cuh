// conv_engine.cuh — one fp32 implicit-GEMM convolution engine for every convolution on the path:
//   * 2-D 3x3 convs of the RED regulariser (ConvReLU / ConvGRUCell2 / ConvTransReLU,
//     modules/module.py:6-58, :178-215, :595-693), applied to ALL depth planes of a [C,D,h,w]
//     tensor at once (taps with dz = 0) or to a single plane (z pinned by q2i_add / q2o_add);
//   * 3-D 3x3x3 convs and stride-2 transposed convs of CostRegNet (modules/module.py:324-410, :546-577).
//
// out[co, p] = post( sum_k W[co, k] * X[k, p] ),  k = (ci, tap),  p = (qd, qh, qw) an iteration point.
// A "problem" describes the gather: input coordinate = q * q2i_mul + q2i_add + tap offset (zero outside
// the tensor), output coordinate = q * q2o_mul + q2o_add.  A stride-2 transposed conv is 4 (2-D) or
// 8 (3-D) problems, one per output parity class, each with only the taps that land on real inputs,
// so no multiplications by structural zeros are issued.  Several problems are launched together
// (grouped launch) so the small, latency-bound recurrent steps cost one launch.
//
// CTA tile BM (co) x BN (points), K step 16, register micro-tile TM x TN, operands staged in shared
// memory with register prefetch of the next K step.  fp32 FFMA throughout: the regularisers feed a
// softmax over D, and north_star's 1e-3 bound on depth leaves no room for tf32 here.
#pragma once
#include "common.cuh"

namespace satmvs {

constexpr int kMaxTaps = 27;
constexpr int kMaxGroup = 8;

struct ConvProblem {
  const float* in;         // [in_c_total][Di][Hi][Wi]
  const float* w;          // w[co * w_co_stride + ci * w_ci_stride + tap_w[t]]
  const float* scale;      // [Cout] or null (1)
  const float* shift;      // [Cout] or null (0)
  const float* pre_add;    // indexed like out, added before the activation, or null
  const float* post_add;   // indexed like out, added after the activation, or null
  float* out;              // [out_c_total][Do][Ho][Wo]
  double* stats;           // [Cout / stats_group][2] running (sum, sum of squares) of the outputs, or null
  int Cin, Cout;
  int Di, Hi, Wi;
  int Do, Ho, Wo;
  int Qd, Qh, Qw;          // iteration space
  int in_c_off, out_c_off;
  long long w_co_stride, w_ci_stride;
  int q2i_mul[3], q2i_add[3];
  int q2o_mul[3], q2o_add[3];
  int ntaps;
  unsigned ntaps_magic;    // ceil(2^32 / ntaps): k / ntaps == __umulhi(k, magic) for k * ntaps < 2^32
  signed char tap_dz[kMaxTaps], tap_dy[kMaxTaps], tap_dx[kMaxTaps], tap_w[kMaxTaps];
  int tap_off[kMaxTaps];   // dz*Hi*Wi + dy*Wi + dx, filled by conv_finalize()
  int relu;
  float acc_scale;         // multiplies the accumulator first (-1 implements conv(-x))
  int stats_group;         // output channels per statistics group (>= the CTA's BM)
};

struct ConvGroup {
  ConvProblem p[kMaxGroup];
  int tile_begin[kMaxGroup + 1];   // CTA index range of each problem
  int n;
};

// k / ntaps without an integer division (ntaps == 1 has no 32-bit magic)
__device__ __forceinline__ int fast_div(int k, const ConvProblem& P) {
  return P.ntaps == 1 ? k : (int)__umulhi((unsigned)k, P.ntaps_magic);
}

template <int BM, int BN, int TM, int TN>
struct ConvTile {
  static constexpr int kBM = BM, kBN = BN, kTM = TM, kTN = TN, kBK = 16;
  static constexpr int kThreads = (BM / TM) * (BN / TN);
  static constexpr int kBLoads = kBK * BN / kThreads;           // B elements per thread per K step
  static constexpr int kBRowStep = kThreads / BN;               // >= 1 (kThreads >= BN in every config)
  static constexpr int kALoads = (BM * kBK + kThreads - 1) / kThreads;
  static_assert(kThreads % BN == 0 || BN % kThreads == 0, "loader mapping");
};

template <class T>
__global__ void __launch_bounds__(T::kThreads)
conv_gemm_kernel(const __grid_constant__ ConvGroup g) {
  constexpr int BM = T::kBM, BN = T::kBN, BK = T::kBK, TM = T::kTM, TN = T::kTN, NT = T::kThreads;
  __shared__ __align__(16) float As[BK][BM + 4];
  __shared__ __align__(16) float Bs[BK][BN];
  __shared__ double red[2][NT / 32];

  // which problem / tile
  int pi = 0;
#pragma unroll
  for (int i = 1; i < kMaxGroup; ++i)
    if (i < g.n && (int)blockIdx.x >= g.tile_begin[i]) pi = i;
  const ConvProblem& P = g.p[pi];
  const int tile = blockIdx.x - g.tile_begin[pi];
  const long long npts = (long long)P.Qd * P.Qh * P.Qw;
  const int tiles_n = (int)((npts + BN - 1) / BN);
  const int m0 = (tile / tiles_n) * BM;
  const long long n0 = (long long)(tile % tiles_n) * BN;
  const int K = P.Cin * P.ntaps;
  const int tid = threadIdx.x;

  // ---- B loader: this thread always gathers for one iteration point (column) ----
  constexpr int BCOLS_PER_PASS = (NT >= BN) ? BN : NT;
  const int bcol = tid % BCOLS_PER_PASS;
  const int brow0 = tid / BCOLS_PER_PASS;
  constexpr int BROWSTEP = (NT >= BN) ? NT / BN : 1;
  constexpr int BCOLPASSES = (NT >= BN) ? 1 : BN / NT;
  constexpr int BROWS = (NT >= BN) ? BK / BROWSTEP : BK;
  long long boff[BCOLPASSES];          // offset of the column's input base coordinate inside a channel
  unsigned bmask[BCOLPASSES];          // bit t set <=> tap t of this column lands inside the tensor
#pragma unroll
  for (int cp = 0; cp < BCOLPASSES; ++cp) {
    long long p = n0 + bcol + cp * NT;
    const bool ok = p < npts;
    if (!ok) p = 0;
    const int qw = (int)(p % P.Qw); const long long t = p / P.Qw;
    const int qh = (int)(t % P.Qh); const int qd = (int)(t / P.Qh);
    const int z = qd * P.q2i_mul[0] + P.q2i_add[0];
    const int y = qh * P.q2i_mul[1] + P.q2i_add[1];
    const int x = qw * P.q2i_mul[2] + P.q2i_add[2];
    boff[cp] = ((long long)z * P.Hi + y) * P.Wi + x;
    unsigned m = 0;
    for (int tt = 0; tt < P.ntaps; ++tt) {
      const int iz = z + P.tap_dz[tt], iy = y + P.tap_dy[tt], ix = x + P.tap_dx[tt];
      if (ok && (unsigned)iz < (unsigned)P.Di && (unsigned)iy < (unsigned)P.Hi && (unsigned)ix < (unsigned)P.Wi) m |= 1u << tt;
    }
    bmask[cp] = m;
  }
  const long long in_cs = (long long)P.Di * P.Hi * P.Wi;
  const float* in_base = P.in + (long long)P.in_c_off * in_cs;

  // ---- A loader ----
  constexpr int AL = T::kALoads;

  float breg[BCOLPASSES][BROWS];
  float areg[AL];

  auto load_tiles = [&](int k0) {
#pragma unroll
    for (int i = 0; i < AL; ++i) {
      const int e = tid + i * NT;           // element of the BM x BK tile, k fastest
      float v = 0.0f;
      if (e < BM * BK) {
        const int kk = e % BK, mm = e / BK;
        const int k = k0 + kk, co = m0 + mm;
        if (k < K && co < P.Cout) {
          const int ci = fast_div(k, P), t = k - ci * P.ntaps;
          v = __ldg(P.w + co * P.w_co_stride + ci * P.w_ci_stride + P.tap_w[t]);
        }
      }
      areg[i] = v;
    }
#pragma unroll
    for (int r = 0; r < BROWS; ++r) {
      const int k = k0 + brow0 + r * BROWSTEP;
      const bool kok = k < K;
      const int ci = kok ? fast_div(k, P) : 0;
      const int t = kok ? k - ci * P.ntaps : 0;
      const long long koff = ci * in_cs + P.tap_off[t];
#pragma unroll
      for (int cp = 0; cp < BCOLPASSES; ++cp) {
        float v = 0.0f;
        if (kok && ((bmask[cp] >> t) & 1u)) v = __ldg(in_base + boff[cp] + koff);
        breg[cp][r] = v;
      }
    }
  };
  auto store_tiles = [&]() {
#pragma unroll
    for (int i = 0; i < AL; ++i) {
      const int e = tid + i * NT;
      if (e < BM * BK) As[e % BK][e / BK] = areg[i];
    }
#pragma unroll
    for (int r = 0; r < BROWS; ++r)
#pragma unroll
      for (int cp = 0; cp < BCOLPASSES; ++cp) Bs[brow0 + r * BROWSTEP][bcol + cp * NT] = breg[cp][r];
  };

  // ---- compute mapping ----
  const int tx = tid % (BN / TN), ty = tid / (BN / TN);
  float acc[TM][TN];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.0f;

  load_tiles(0);
  for (int k0 = 0; k0 < K; k0 += BK) {
    store_tiles();
    __syncthreads();
    if (k0 + BK < K) load_tiles(k0 + BK);
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      float a[TM], b[TN];
#pragma unroll
      for (int i = 0; i < TM; i += 4) {
        if constexpr (TM % 4 == 0) {
          const float4 v = *reinterpret_cast<const float4*>(&As[kk][ty * TM + i]);
          a[i] = v.x; a[i + 1] = v.y; a[i + 2] = v.z; a[i + 3] = v.w;
        }
      }
      if constexpr (TM % 4 != 0) {
#pragma unroll
        for (int i = 0; i < TM; ++i) a[i] = As[kk][ty * TM + i];
      }
#pragma unroll
      for (int j = 0; j < TN; j += 4) {
        const float4 v = *reinterpret_cast<const float4*>(&Bs[kk][tx * TN + j]);
        b[j] = v.x; b[j + 1] = v.y; b[j + 2] = v.z; b[j + 3] = v.w;
      }
#pragma unroll
      for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }

  // ---- epilogue ----
  const long long out_cs = (long long)P.Do * P.Ho * P.Wo;
  float ssum = 0.0f, ssq = 0.0f;
#pragma unroll
  for (int j = 0; j < TN; ++j) {
    const long long p = n0 + tx * TN + j;
    if (p >= npts) continue;
    const int qw = (int)(p % P.Qw); const long long t = p / P.Qw;
    const int qh = (int)(t % P.Qh); const int qd = (int)(t / P.Qh);
    const long long sp = ((long long)(qd * P.q2o_mul[0] + P.q2o_add[0]) * P.Ho + (qh * P.q2o_mul[1] + P.q2o_add[1])) * P.Wo +
                         (qw * P.q2o_mul[2] + P.q2o_add[2]);
#pragma unroll
    for (int i = 0; i < TM; ++i) {
      const int co = m0 + ty * TM + i;
      if (co >= P.Cout) continue;
      const long long idx = (long long)(P.out_c_off + co) * out_cs + sp;
      float v = acc[i][j] * P.acc_scale;
      if (P.scale) v *= __ldg(P.scale + co);
      if (P.shift) v += __ldg(P.shift + co);
      if (P.pre_add) v += __ldg(P.pre_add + idx);
      if (P.relu) v = fmaxf(v, 0.0f);
      if (P.post_add) v += __ldg(P.post_add + idx);
      P.out[idx] = v;
      ssum += v; ssq += v * v;
    }
  }
  if (P.stats) {   // GroupNorm(1, C) statistics: (sum, sum^2) of this CTA's outputs -> fp64 atomics
    double ds = ssum, dq = ssq;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { ds += __shfl_xor_sync(0xffffffffu, ds, o); dq += __shfl_xor_sync(0xffffffffu, dq, o); }
    if ((tid & 31) == 0) { red[0][tid >> 5] = ds; red[1][tid >> 5] = dq; }
    __syncthreads();
    if (tid == 0) {
      double s = 0.0, q = 0.0;
      for (int i = 0; i < NT / 32; ++i) { s += red[0][i]; q += red[1][i]; }
      const int grp = m0 / P.stats_group;
      atomicAdd(P.stats + 2 * grp, s);
      atomicAdd(P.stats + 2 * grp + 1, q);
    }
  }
}

// ---------------------------------------------------------------------------------------------
// host helpers
// ---------------------------------------------------------------------------------------------
using Tile64 = ConvTile<64, 128, 8, 4>;   // Cout >= 64
using Tile32 = ConvTile<32, 128, 4, 4>;
using Tile16 = ConvTile<16, 256, 4, 4>;
using Tile8 = ConvTile<8, 512, 4, 4>;     // Cout <= 8 (and the 1-channel heads)
using Tile32s = ConvTile<32, 32, 4, 4>;   // small-image variants for the recurrent steps: more CTAs
using Tile16s = ConvTile<16, 64, 4, 4>;
using Tile8s = ConvTile<8, 128, 4, 4>;

inline void conv_problem_defaults(ConvProblem& p) {
  p = ConvProblem{};
  p.acc_scale = 1.0f;
  p.stats_group = 1 << 30;
  for (int i = 0; i < 3; ++i) { p.q2i_mul[i] = 1; p.q2o_mul[i] = 1; }
}

// taps of an ordinary 3x3 (2-D, applied per plane) or 3x3x3 conv with padding 1
inline void conv_taps_dense(ConvProblem& p, bool three_d) {
  int n = 0;
  for (int kz = 0; kz < (three_d ? 3 : 1); ++kz)
    for (int ky = 0; ky < 3; ++ky)
      for (int kx = 0; kx < 3; ++kx) {
        p.tap_dz[n] = three_d ? kz : 0; p.tap_dy[n] = ky; p.tap_dx[n] = kx;
        p.tap_w[n] = (three_d ? kz * 9 : 0) + ky * 3 + kx;
        ++n;
      }
  p.ntaps = n;
}

// taps of one output-parity class of ConvTranspose(k=3, stride 2, padding 1, output_padding 1):
// even output o = 2q uses k = 1 at input q; odd output o = 2q+1 uses k = 0 at input q+1 and k = 2 at q.
inline void conv_taps_deconv_class(ConvProblem& p, bool three_d, int pz, int py, int px) {
  const int koff[2][2] = {{0, 1}, {1, 0}};   // parity -> input offsets
  const int kidx[2][2] = {{1, 1}, {0, 2}};   // parity -> kernel index
  const int cnt[2] = {1, 2};
  int n = 0;
  for (int a = 0; a < (three_d ? cnt[pz] : 1); ++a)
    for (int b = 0; b < cnt[py]; ++b)
      for (int c = 0; c < cnt[px]; ++c) {
        p.tap_dz[n] = three_d ? koff[pz][a] : 0; p.tap_dy[n] = koff[py][b]; p.tap_dx[n] = koff[px][c];
        p.tap_w[n] = (three_d ? kidx[pz][a] * 9 : 0) + kidx[py][b] * 3 + kidx[px][c];
        ++n;
      }
  p.ntaps = n;
}

// derived fields; call after the geometry and taps of a problem are set
inline void conv_finalize(ConvProblem& p) {
  p.ntaps_magic = p.ntaps > 1 ? (unsigned)(((1ULL << 32) + p.ntaps - 1) / p.ntaps) : 0u;
  for (int t = 0; t < p.ntaps; ++t) p.tap_off[t] = (p.tap_dz[t] * p.Hi + p.tap_dy[t]) * p.Wi + p.tap_dx[t];
}

template <class T>
inline int conv_tiles(const ConvProblem& p) {
  const long long npts = (long long)p.Qd * p.Qh * p.Qw;
  return (int)(((npts + T::kBN - 1) / T::kBN) * ((p.Cout + T::kBM - 1) / T::kBM));
}

template <class T>
inline int conv_launch(ConvGroup& g, cudaStream_t st, const char* what) {
  int total = 0;
  for (int i = 0; i < g.n; ++i) { g.tile_begin[i] = total; total += conv_tiles<T>(g.p[i]); }
  for (int i = g.n; i <= kMaxGroup; ++i) g.tile_begin[i] = total;
  if (total == 0) return SATMVS_OK;
  conv_gemm_kernel<T><<<total, T::kThreads, 0, st>>>(g);
  return check_launch(what);
}

}  // namespace satmvs
