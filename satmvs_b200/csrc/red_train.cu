// red_train.cu — pointwise / normalisation kernels of the BACKWARD of the RED regulariser (RED_Regularization.forward,
// modules/module.py:614-649; ConvGRUCell2 :6-58), used by satmvs_b200/training.py together with satmvs_conv3d_raw /
// satmvs_conv3d_wgrad (train.cu).  train.py:267-287 calls loss.backward() through this network.
//
// Tensors are channel-major with the depth planes as the second axis: t[c][d][px] (px = h*w), channel stride D*px, the layout
// of the forward's state history.  What is parallel over planes runs batched (grid.y = plane); what is sequential in depth
// (the gradient of the hidden state travelling from plane d+1 to plane d) runs one plane per launch:
//
//   forward of a cell at plane d (h = state before, h' = state after):
//     G = conv([x, h]) + b_g;  r = sig(GN_r(G[:ch])), u = sig(GN_u(G[ch:]));  O = conv([x, r*h]) + b_o;  y = tanh(GN_o(O))
//     h' = u*h + (1-u)*y
//   backward given dh':
//     (1) gru_bwd_out:   dYn = dh' (1-u) (1-y^2);  dUn = dh' (h-y) u (1-u);  carry = dh' u;   sums for GN_o' and GN_u'
//     (2) gn_bwd_apply:  dO = GN_o'(dYn)                    -> conv^T with the h-half of the output filters -> dRH
//     (3) gru_bwd_reset: dRn = dRH h r (1-r);  carry += dRH r;                                 sums for GN_r'
//     (4) gn_bwd_apply:  dG = [GN_r'(dRn), GN_u'(dUn)]      -> conv^T with the h-half of the gate filters   -> dHg
//     (5) folded into step (1) of plane d-1: dh'(d-1) = carry + dHg + (gradient reaching h'(d-1) from the decoder)
#include "common.cuh"
#include "direct_conv.cuh"
#include "prof.cuh"

namespace satmvs {

__device__ __forceinline__ void block_atomic_add2(double a, double b, double* dst) {
  __shared__ double red[2][32];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) { a += __shfl_xor_sync(0xffffffffu, a, o); b += __shfl_xor_sync(0xffffffffu, b, o); }
  const int w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  if ((threadIdx.x & 31) == 0) { red[0][w] = a; red[1][w] = b; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0.0, t = 0.0;
    for (int i = 0; i < nw; ++i) { s += red[0][i]; t += red[1][i]; }
    atomicAdd(dst, s);
    atomicAdd(dst + 1, t);
  }
  __syncthreads();
}

// ---- GroupNorm(1, ch) + activation over all planes: statistics ----
// pre [groups*ch][D][px] (+ bias per channel); acc [D][groups][2] doubles (sum, sum of squares), zeroed by the caller
__global__ void __launch_bounds__(256) gn_stats_kernel(const float* __restrict__ pre, const float* __restrict__ bias, int ch, int groups,
                                                       int D, int px, double* __restrict__ acc) {
  const int d = blockIdx.y, g = blockIdx.z;
  const long long cs = (long long)D * px;
  double s = 0.0, q = 0.0;
  const int n = ch * px;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const int c = g * ch + i / px, p = i - (i / px) * px;
    const float v = __ldg(pre + c * cs + (long long)d * px + p) + (bias ? __ldg(bias + c) : 0.0f);
    s += v; q += (double)v * v;
  }
  block_atomic_add2(s, q, acc + ((size_t)d * groups + g) * 2);
}

// out = act(gamma_c * (v - mean) * rstd + beta_c); stats [D][groups][2] floats (mean, rstd) written by block 0
__global__ void __launch_bounds__(256) gn_apply_kernel(const float* __restrict__ pre, const float* __restrict__ bias,
                                                       const float* __restrict__ gamma, const float* __restrict__ beta, int ch, int groups,
                                                       int D, int px, const double* __restrict__ acc, float eps, int act,
                                                       float* __restrict__ out, float* __restrict__ stats) {
  const int d = blockIdx.y, g = blockIdx.z;
  const long long cs = (long long)D * px;
  const int n = ch * px;
  const double* a = acc + ((size_t)d * groups + g) * 2;
  const double mean_d = a[0] / n;
  double var = a[1] / n - mean_d * mean_d;
  if (var < 0.0) var = 0.0;
  const float mean = (float)mean_d, rstd = (float)(1.0 / sqrt(var + (double)eps));
  if (blockIdx.x == 0 && threadIdx.x == 0) { stats[((size_t)d * groups + g) * 2] = mean; stats[((size_t)d * groups + g) * 2 + 1] = rstd; }
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const int c = g * ch + i / px, p = i - (i / px) * px;
    const long long idx = c * cs + (long long)d * px + p;
    const float v = __ldg(pre + idx) + (bias ? __ldg(bias + c) : 0.0f);
    const float t = (v - mean) * rstd * __ldg(gamma + c) + __ldg(beta + c);
    out[idx] = act == 0 ? 1.0f / (1.0f + expf(-t)) : tanhf(t);
  }
}

// ---- generic elementwise over [C][D][px] tensors with per-tensor plane offsets: out = scale * (a [+ b]) [* (m1 [- m2] > 0)] ----
struct EwArgs {
  const float* a; const float* b; const float* m1; const float* m2; const float* mul;   // b, m1, m2, mul may be null
  float* out;
  long long a_cs, b_cs, m1_cs, m2_cs, mul_cs, out_cs;   // channel strides (floats); plane offsets folded into the pointers
  long long n_per_c;                                     // D * px elements per channel
  int C;
  float scale;
};

__global__ void __launch_bounds__(256) ew_kernel(const EwArgs e) {
  const int c = blockIdx.y;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < e.n_per_c; i += (long long)gridDim.x * blockDim.x) {
    float v = __ldg(e.a + c * e.a_cs + i);
    if (e.b) v += __ldg(e.b + c * e.b_cs + i);
    if (e.mul) v *= __ldg(e.mul + c * e.mul_cs + i);
    if (e.m1) {
      float m = __ldg(e.m1 + c * e.m1_cs + i);
      if (e.m2) m -= __ldg(e.m2 + c * e.m2_cs + i);
      if (!(m > 0.0f)) v = 0.0f;
    }
    e.out[c * e.out_cs + i] = v * e.scale;
  }
}

// sum over (d, px) per channel: out[c] = sum t[c][:][:]   (bias gradients)
__global__ void __launch_bounds__(256) channel_sum_kernel(const float* __restrict__ t, long long n_per_c, double* __restrict__ acc) {
  const int c = blockIdx.y;
  double s = 0.0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n_per_c; i += (long long)gridDim.x * blockDim.x)
    s += __ldg(t + c * n_per_c + i);
  block_atomic_add2(s, 0.0, acc + 2 * (size_t)c);
}

// GroupNorm affine gradients over all planes: dgamma_c = sum dout * xhat, dbeta_c = sum dout
__global__ void __launch_bounds__(256) gn_param_grad_kernel(const float* __restrict__ dout, const float* __restrict__ pre,
                                                            const float* __restrict__ bias, const float* __restrict__ stats, int ch,
                                                            int groups, int D, int px, double* __restrict__ acc) {
  const int c = blockIdx.y;                  // 0 .. groups*ch - 1
  const int g = c / ch;
  const long long cs = (long long)D * px;
  double s1 = 0.0, s2 = 0.0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < cs; i += (long long)gridDim.x * blockDim.x) {
    const int d = (int)(i / px);
    const float mean = __ldg(stats + ((size_t)d * groups + g) * 2), rstd = __ldg(stats + ((size_t)d * groups + g) * 2 + 1);
    const float v = __ldg(pre + c * cs + i) + (bias ? __ldg(bias + c) : 0.0f);
    const float go = __ldg(dout + c * cs + i);
    s1 += go * ((v - mean) * rstd);
    s2 += go;
  }
  block_atomic_add2(s1, s2, acc + 2 * (size_t)c);
}

__global__ void acc_to_float_kernel(const double* __restrict__ acc, int n, int stride, int offset, float* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = (float)acc[(size_t)i * stride + offset];
}

// ---- sequential part, one plane of one level per launch ----
struct GruBwdArgs {
  // per-plane views (already offset to plane d); cs = channel stride of the [C][D][px] tensors
  const float* dh;      // null at the last plane (only the decoder reaches h'(D-1)), any non-null value below it
  const float* h;       // state before the cell, channel stride s_cs
  const float* ru;      // [2ch] r then u, channel stride cs
  const float* y;       // [ch], cs
  const float* opre;    // [ch], cs   output-conv result without bias
  const float* gpre;    // [2ch], cs  gate-conv result without bias
  const float* ob; const float* gb;                 // conv biases [ch], [2ch]
  const float* on_w; const float* rn_w; const float* un_w;   // GroupNorm weights
  const float* ostat; const float* gstat;           // (mean, rstd) of this plane: [2], [2][2]
  float* dyn;           // [ch], cs    dL/d(GN_o output)           (kept for the affine gradients)
  float* dgn;           // [2ch], cs   dL/d(GN_r / GN_u outputs)
  float* dO;            // [ch], cs    dL/dO
  float* dG;            // [2ch], cs
  float* dO_p;          // [ch][px] contiguous copies feeding the transposed convs
  float* dG_p;          // [2ch][px]
  const float* dRH;     // [ch][px] contiguous: conv^T(dO) restricted to the hidden channels
  const float* dHg;     // [ch][px] contiguous: conv^T(dG) restricted to the hidden channels
  float* carry;         // [ch][px] contiguous
  const float* dec;     // gradient reaching h'(d) from the decoder, channel stride dec_cs
  double* red;          // [6] of this plane, zeroed: (a1, a2) output norm, (b1, b2) update norm, (c1, c2) reset norm
  long long cs, s_cs, dec_cs;
  int ch, px;
};

__global__ void __launch_bounds__(256) gru_bwd_out_kernel(const GruBwdArgs a) {
  const int n = a.ch * a.px;
  const float om = a.ostat[0], orstd = a.ostat[1], um = a.gstat[2], urstd = a.gstat[3];
  double a1 = 0.0, a2 = 0.0, b1 = 0.0, b2 = 0.0;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const int c = i / a.px, p = i - c * a.px;
    // gradient at h'(d): from the decoder, plus (below the last plane) what plane d+1 left in carry and its gate-conv gradient
    const float dh = __ldg(a.dec + c * a.dec_cs + p) + (a.dh ? a.carry[i] + __ldg(a.dHg + i) : 0.0f);
    const float h = __ldg(a.h + c * a.s_cs + p), u = __ldg(a.ru + (a.ch + c) * a.cs + p), y = __ldg(a.y + c * a.cs + p);
    const float dyn = dh * (1.0f - u) * (1.0f - y * y);
    const float dun = dh * (h - y) * u * (1.0f - u);
    a.dyn[c * a.cs + p] = dyn;
    a.dgn[(a.ch + c) * a.cs + p] = dun;
    a.carry[i] = dh * u;
    const float xo = (__ldg(a.opre + c * a.cs + p) + __ldg(a.ob + c) - om) * orstd;
    const float xu = (__ldg(a.gpre + (a.ch + c) * a.cs + p) + __ldg(a.gb + a.ch + c) - um) * urstd;
    const float go = dyn * __ldg(a.on_w + c), gu = dun * __ldg(a.un_w + c);
    a1 += go; a2 += go * xo; b1 += gu; b2 += gu * xu;
  }
  block_atomic_add2(a1, a2, a.red);
  block_atomic_add2(b1, b2, a.red + 2);
}

// dO = rstd * (dYn * gamma - m1 - xhat * m2)
__global__ void __launch_bounds__(256) gru_bwd_dO_kernel(const GruBwdArgs a) {
  const int n = a.ch * a.px;
  const float om = a.ostat[0], orstd = a.ostat[1];
  const float m1 = (float)(a.red[0] / n), m2 = (float)(a.red[1] / n);
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const int c = i / a.px, p = i - c * a.px;
    const float xo = (__ldg(a.opre + c * a.cs + p) + __ldg(a.ob + c) - om) * orstd;
    const float v = orstd * (a.dyn[c * a.cs + p] * __ldg(a.on_w + c) - m1 - xo * m2);
    a.dO[c * a.cs + p] = v;
    a.dO_p[i] = v;
  }
}

__global__ void __launch_bounds__(256) gru_bwd_reset_kernel(const GruBwdArgs a) {
  const int n = a.ch * a.px;
  const float rm = a.gstat[0], rrstd = a.gstat[1];
  double c1 = 0.0, c2 = 0.0;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const int c = i / a.px, p = i - c * a.px;
    const float drh = __ldg(a.dRH + i), h = __ldg(a.h + c * a.s_cs + p), r = __ldg(a.ru + c * a.cs + p);
    const float drn = drh * h * r * (1.0f - r);
    a.dgn[c * a.cs + p] = drn;
    a.carry[i] += drh * r;
    const float xr = (__ldg(a.gpre + c * a.cs + p) + __ldg(a.gb + c) - rm) * rrstd;
    const float gr = drn * __ldg(a.rn_w + c);
    c1 += gr; c2 += gr * xr;
  }
  block_atomic_add2(c1, c2, a.red + 4);
}

__global__ void __launch_bounds__(256) gru_bwd_dG_kernel(const GruBwdArgs a) {
  const int n = 2 * a.ch * a.px, half = a.ch * a.px;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const int c = i / a.px, p = i - c * a.px;          // c in [0, 2ch)
    const int g = c >= a.ch;
    const float mean = a.gstat[2 * g], rstd = a.gstat[2 * g + 1];
    const float m1 = (float)(a.red[g ? 2 : 4] / half), m2 = (float)(a.red[g ? 3 : 5] / half);
    const float gamma = g ? __ldg(a.un_w + c - a.ch) : __ldg(a.rn_w + c);
    const float xh = (__ldg(a.gpre + c * a.cs + p) + __ldg(a.gb + c) - mean) * rstd;
    const float v = rstd * (a.dgn[c * a.cs + p] * gamma - m1 - xh * m2);
    a.dG[c * a.cs + p] = v;
    a.dG_p[i] = v;
  }
}

static inline int grid_for(long long n) {
  long long b = (n + 255) / 256;
  const long long cap = 4LL * kNumSMs;
  return (int)(b < 1 ? 1 : (b > cap ? cap : b));
}

// conv^T restricted to the hidden input channels, one plane: a 3x3 conv with mirrored taps over `cin` gradient channels.  The deep
// levels are tiny (288 pixels x 64 channels at level 3 of a 96x192 plane): their input channels are split over CTAs (ksplit) so the
// per-plane latency is a few microseconds instead of one long serial loop on 8 CTAs.
constexpr int kMaxKsplit = 8;
static int hidden_dgrad(const float* in, int cin, const satmvs_gru_bwd_level& L, const float* w_h, float* out, float* partial,
                        cudaStream_t st) {
  DirectConv d{};
  d.in = in; d.w = w_h; d.out = out;
  d.Cin = cin; d.Cout = L.ch; d.Di = 1; d.Hi = L.h; d.Wi = L.w; d.Do = 1; d.Ho = L.h; d.Wo = L.w;
  d.w_co = 9; d.w_ci = L.w_ci; d.acc_scale = 1.0f; d.flip = 1;
  if (!direct_conv_supported(d, 1, 1))
    return satmvs_conv3d_raw(in, cin, 1, L.h, L.w, w_h, 9, L.w_ci, 1, 2, out, L.ch, st);
  const long long tiles = (long long)ceil_div((long long)L.h * L.w, kDcWarps * 32 * kDcPx) * ceil_div(L.ch, kDcCo);
  int ks = 1;
  while (ks < kMaxKsplit && tiles * ks < kNumSMs && cin / (2 * ks) >= kDcCiChunk) ks *= 2;
  if (ks > 1) { d.ksplit = ks; d.partial = partial; }
  ProfScope prof(kProfTrainConv, st);
  return direct_conv_launch(d, 1, 1, st, "satmvs_red_recurrence_bwd (conv^T)");
}

}  // namespace satmvs

using namespace satmvs;

extern "C" {

int satmvs_gn_act_fwd(const float* pre, const float* bias, const float* gamma, const float* beta, int ch, int groups, int D, int px,
                      float eps, int act, float* out, float* stats, double* acc, void* stream) {
  SATMVS_REQUIRE(pre && gamma && beta && out && stats && acc && ch >= 1 && groups >= 1 && D >= 1 && px >= 1 && (act == 0 || act == 1));
  cudaStream_t st = (cudaStream_t)stream;
  ProfScope prof(kProfTrainNorm, st);
  cudaMemsetAsync(acc, 0, (size_t)D * groups * 2 * sizeof(double), st);
  int bx = grid_for((long long)ch * px);
  if ((long long)bx * D * groups > 16LL * kNumSMs) { bx = (int)(16LL * kNumSMs / ((long long)D * groups)); if (bx < 1) bx = 1; }
  gn_stats_kernel<<<dim3(bx, D, groups), 256, 0, st>>>(pre, bias, ch, groups, D, px, acc);
  gn_apply_kernel<<<dim3(bx, D, groups), 256, 0, st>>>(pre, bias, gamma, beta, ch, groups, D, px, acc, eps, act, out, stats);
  return check_launch("satmvs_gn_act_fwd");
}

int satmvs_elementwise(const float* a, long long a_cs, const float* b, long long b_cs, const float* mul, long long mul_cs,
                       const float* m1, long long m1_cs, const float* m2, long long m2_cs, float scale, float* out, long long out_cs,
                       int C, long long n_per_c, void* stream) {
  SATMVS_REQUIRE(a && out && C >= 1 && n_per_c >= 1);
  cudaStream_t st = (cudaStream_t)stream;
  ProfScope prof(kProfTrainNorm, st);
  EwArgs e{};
  e.a = a; e.b = b; e.m1 = m1; e.m2 = m2; e.mul = mul; e.out = out;
  e.a_cs = a_cs; e.b_cs = b_cs; e.m1_cs = m1_cs; e.m2_cs = m2_cs; e.mul_cs = mul_cs; e.out_cs = out_cs;
  e.n_per_c = n_per_c; e.C = C; e.scale = scale;
  int bx = grid_for(n_per_c);
  if ((long long)bx * C > 16LL * kNumSMs) { bx = (int)(16LL * kNumSMs / C); if (bx < 1) bx = 1; }
  ew_kernel<<<dim3(bx, C), 256, 0, st>>>(e);
  return check_launch("satmvs_elementwise");
}

int satmvs_channel_sum(const float* t, int C, long long n_per_c, float* out, double* acc, void* stream) {
  SATMVS_REQUIRE(t && out && acc && C >= 1 && n_per_c >= 1);
  cudaStream_t st = (cudaStream_t)stream;
  ProfScope prof(kProfTrainNorm, st);
  cudaMemsetAsync(acc, 0, (size_t)C * 2 * sizeof(double), st);
  int bx = grid_for(n_per_c);
  if ((long long)bx * C > 16LL * kNumSMs) { bx = (int)(16LL * kNumSMs / C); if (bx < 1) bx = 1; }
  channel_sum_kernel<<<dim3(bx, C), 256, 0, st>>>(t, n_per_c, acc);
  acc_to_float_kernel<<<ceil_div(C, 128), 128, 0, st>>>(acc, C, 2, 0, out);
  return check_launch("satmvs_channel_sum");
}

int satmvs_gn_param_grad(const float* dout, const float* pre, const float* bias, const float* stats, int ch, int groups, int D, int px,
                         float* dgamma, float* dbeta, double* acc, void* stream) {
  SATMVS_REQUIRE(dout && pre && stats && dgamma && dbeta && acc && ch >= 1 && groups >= 1 && D >= 1 && px >= 1);
  cudaStream_t st = (cudaStream_t)stream;
  ProfScope prof(kProfTrainNorm, st);
  const int C = ch * groups;
  cudaMemsetAsync(acc, 0, (size_t)C * 2 * sizeof(double), st);
  int bx = grid_for((long long)D * px);
  if ((long long)bx * C > 16LL * kNumSMs) { bx = (int)(16LL * kNumSMs / C); if (bx < 1) bx = 1; }
  gn_param_grad_kernel<<<dim3(bx, C), 256, 0, st>>>(dout, pre, bias, stats, ch, groups, D, px, acc);
  acc_to_float_kernel<<<ceil_div(C, 128), 128, 0, st>>>(acc, C, 2, 0, dgamma);
  acc_to_float_kernel<<<ceil_div(C, 128), 128, 0, st>>>(acc, C, 2, 1, dbeta);
  return check_launch("satmvs_gn_param_grad");
}

// The sequential part of the backward for up to 4 levels (independent recurrences): planes D-1 .. 0, per plane the five steps of
// the header comment around the two transposed convolutions (satmvs_conv3d_raw, mirrored taps, hidden input channels only).
// Every level runs on a stream of its own between a fork from / join into the caller's stream.
struct TrainStreams {
  cudaStream_t s[4] = {nullptr, nullptr, nullptr, nullptr};
  cudaEvent_t fork = nullptr, join[4] = {nullptr, nullptr, nullptr, nullptr};
  int dev = -1;
  bool ok = false;
};
static TrainStreams& train_streams() {
  static thread_local TrainStreams t[16];
  int dev = 0;
  cudaGetDevice(&dev);
  TrainStreams& r = t[dev & 15];
  if (r.dev != dev) {
    r.dev = dev;
    r.ok = cudaEventCreateWithFlags(&r.fork, cudaEventDisableTiming) == cudaSuccess;
    for (int i = 0; i < 4 && r.ok; ++i)
      r.ok = cudaStreamCreateWithFlags(&r.s[i], cudaStreamNonBlocking) == cudaSuccess &&
             cudaEventCreateWithFlags(&r.join[i], cudaEventDisableTiming) == cudaSuccess;
    if (!r.ok) cudaGetLastError();
  }
  return r;
}

int satmvs_red_recurrence_bwd(const satmvs_gru_bwd_level* lv, int nlevels, void* stream) {
  SATMVS_CHECK_ASYNC();
  SATMVS_REQUIRE(lv && nlevels >= 1 && nlevels <= 4);
  cudaStream_t st = (cudaStream_t)stream;
  TrainStreams& ts = train_streams();
  cudaStream_t ls[4];
  for (int l = 0; l < nlevels; ++l) ls[l] = ts.ok ? ts.s[l] : st;
  if (ts.ok) {
    cudaEventRecord(ts.fork, st);
    for (int l = 0; l < nlevels; ++l) cudaStreamWaitEvent(ls[l], ts.fork, 0);
  }
  GruBwdArgs a[4];
  float* dRH[4]; float* dHg[4]; float* ksp[4]; double* red[4];
  int Dmax = 0;
  for (int l = 0; l < nlevels; ++l) {
    const satmvs_gru_bwd_level& L = lv[l];
    SATMVS_REQUIRE(L.S && L.ru && L.y && L.opre && L.gpre && L.dec && L.wo_h && L.wg_h && L.dyn && L.dgn && L.dO && L.dG && L.scratch);
    SATMVS_REQUIRE(L.ch >= 1 && L.h >= 1 && L.w >= 1 && L.D >= 1);
    const int px = L.h * L.w;
    const size_t n = (size_t)L.ch * px;
    GruBwdArgs& g = a[l];
    g = GruBwdArgs{};
    g.ob = L.ob; g.gb = L.gb; g.on_w = L.on_w; g.rn_w = L.rn_w; g.un_w = L.un_w;
    float* sc = L.scratch;
    red[l] = reinterpret_cast<double*>(sc); sc += 12 * (size_t)L.D + 4;      // 6 doubles per plane, zeroed once
    g.dO_p = sc; sc += n;
    g.dG_p = sc; sc += 2 * n;
    g.carry = sc; sc += n;
    dRH[l] = sc; sc += n;
    dHg[l] = sc; sc += n;
    ksp[l] = sc; sc += kMaxKsplit * n;
    g.cs = (long long)L.D * px; g.s_cs = (long long)(L.D + 1) * px; g.dec_cs = g.cs; g.ch = L.ch; g.px = px;
    cudaMemsetAsync(red[l], 0, (size_t)L.D * 6 * sizeof(double), ls[l]);
    if (L.D > Dmax) Dmax = L.D;
  }
  int rc = check_launch("satmvs_red_recurrence_bwd (init)");
  if (rc) return rc;
  for (int step = 0; step < Dmax; ++step) {
    for (int l = 0; l < nlevels; ++l) {
      const satmvs_gru_bwd_level& L = lv[l];
      const int d = L.D - 1 - step;
      if (d < 0) continue;
      GruBwdArgs& g = a[l];
      const int px = g.px;
      const size_t o = (size_t)d * px;
      g.dh = step == 0 ? nullptr : dHg[l];       // flag: below the last plane the carry / gate-conv gradient of plane d+1 are added
      g.h = L.S + o; g.ru = L.ru + o; g.y = L.y + o; g.opre = L.opre + o; g.gpre = L.gpre + o;
      g.ostat = L.ostat + (size_t)d * 2; g.gstat = L.gstat + (size_t)d * 4;
      g.dyn = L.dyn + o; g.dgn = L.dgn + o; g.dO = L.dO + o; g.dG = L.dG + o;
      g.dRH = dRH[l]; g.dHg = dHg[l];
      g.dec = L.dec + o;
      g.red = red[l] + (size_t)d * 6;
      const int n = g.ch * px, bx = grid_for(n);
      gru_bwd_out_kernel<<<bx, 256, 0, ls[l]>>>(g);
      gru_bwd_dO_kernel<<<bx, 256, 0, ls[l]>>>(g);
      rc = hidden_dgrad(g.dO_p, g.ch, L, L.wo_h, dRH[l], ksp[l], ls[l]);
      if (rc) return rc;
      gru_bwd_reset_kernel<<<bx, 256, 0, ls[l]>>>(g);
      gru_bwd_dG_kernel<<<grid_for(2LL * n), 256, 0, ls[l]>>>(g);
      rc = hidden_dgrad(g.dG_p, 2 * g.ch, L, L.wg_h, dHg[l], ksp[l], ls[l]);
      if (rc) return rc;
    }
  }
  if (ts.ok)
    for (int l = 0; l < nlevels; ++l) { cudaEventRecord(ts.join[l], ls[l]); cudaStreamWaitEvent(st, ts.join[l], 0); }
  return check_launch("satmvs_red_recurrence_bwd");
}

}  // extern "C"
