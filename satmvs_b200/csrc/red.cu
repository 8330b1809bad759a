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
#include "conv_engine.cuh"

namespace satmvs {

constexpr float kGnEps = 1e-5f;   // nn.GroupNorm(1, C, 1e-5, True), modules/module.py:15-20

struct RedLevel {
  int ch, h, w;          // hidden channels, spatial size
  int cx;                // channels of the x input of this level's GRU
  float* gx;             // [2ch][D][h][w]   gates (x-half, then full gates in place)
  float* ox;             // [ch][D][h][w]    output conv (x-half, then full in place)
  float* rh;             // [ch][1][h][w]
  float* s;              // [ch][D+1][h][w]  state history; slot 0 = initial state
};

struct RedPlan {
  int C, D, H, W;
  float* e[3];           // E1 [16][D][H/2][W/2], E2 [32][D][H/4][W/4], E3 [64][D][H/8][W/8]
  float* u[3];           // U1 [8][D+1][H][W], U2 [16][D+1][H/2][W/2], U3 [32][D+1][H/4][W/4]
  RedLevel lv[4];
  double* stats;         // [D][4][3][2]
  size_t bytes;
};

static RedPlan red_plan(int C, int D, int H, int W, char* base) {
  RedPlan p{};
  p.C = C; p.D = D; p.H = H; p.W = W;
  size_t off = 0;
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
    L.s = take(L.ch * (size_t)(D + 1) * px);
  }
  for (int i = 0; i < 3; ++i) {
    p.e[i] = take((size_t)chs[i + 1] * D * (H >> (i + 1)) * (W >> (i + 1)));
    p.u[i] = take((size_t)chs[i] * (D + 1) * (H >> i) * (W >> i));
  }
  p.stats = reinterpret_cast<double*>(base + off);
  off += ((size_t)D * 4 * 3 * 2 * sizeof(double) + 255) / 256 * 256;
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
  int ch, px;
  int begin;             // first flat element index of this level in the launch
};
struct GruArgs { GruLevelArgs l[4]; int total; };

__device__ __forceinline__ void gn_coeff(const double* st, double n, float gamma, float beta, float& a, float& b) {
  const double mean = st[0] / n;
  const double var = fmax(st[1] / n - mean * mean, 0.0);
  const float rstd = (float)(1.0 / sqrt(var + (double)kGnEps));
  a = gamma * rstd;
  b = beta - (float)mean * a;
}

__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + expf(-x)); }

__global__ void gru_reset_kernel(const __grid_constant__ GruArgs a) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= a.total) return;
  int li = 0;
#pragma unroll
  for (int k = 1; k < 4; ++k) if (i >= a.l[k].begin) li = k;
  const GruLevelArgs& L = a.l[li];
  const int e = i - L.begin, c = e / L.px, p = e - c * L.px;
  float ga, gb;
  gn_coeff(L.stats, (double)L.ch * L.px, __ldg(L.rn_w + c), __ldg(L.rn_b + c), ga, gb);
  const float r = sigmoidf_(fmaf(__ldg(L.g + c * L.gcs + p), ga, gb));
  L.rh[e] = r * __ldg(L.hprev + c * L.scs + p);
}

__global__ void gru_update_kernel(const __grid_constant__ GruArgs a) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= a.total) return;
  int li = 0;
#pragma unroll
  for (int k = 1; k < 4; ++k) if (i >= a.l[k].begin) li = k;
  const GruLevelArgs& L = a.l[li];
  const int e = i - L.begin, c = e / L.px, p = e - c * L.px;
  const double n = (double)L.ch * L.px;
  float ua, ub, oa, ob;
  gn_coeff(L.stats + 2, n, __ldg(L.un_w + c), __ldg(L.un_b + c), ua, ub);
  gn_coeff(L.stats + 4, n, __ldg(L.on_w + c), __ldg(L.on_b + c), oa, ob);
  const float u = sigmoidf_(fmaf(__ldg(L.g + (c + L.ch) * L.gcs + p), ua, ub));
  const float y = tanhf(fmaf(__ldg(L.o + c * L.ocs + p), oa, ob));
  const float h = __ldg(L.hprev + c * L.scs + p);
  L.hnext[c * L.scs + p] = u * h + (1.0f - u) * y;       // module.py:57
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

static int launch_by_cout(ConvProblem p, cudaStream_t st, const char* what) {
  if (p.Cout >= 64) return launch_one<Tile64>(p, st, what);
  if (p.Cout >= 32) return launch_one<Tile32>(p, st, what);
  if (p.Cout >= 16) return launch_one<Tile16>(p, st, what);
  return launch_one<Tile8>(p, st, what);
}

}  // namespace satmvs

using namespace satmvs;

extern "C" {

size_t satmvs_red_workspace_bytes(int C, int D, int H, int W) {
  if (C < 1 || D < 1 || H < 8 || W < 8 || (H % 8) || (W % 8)) return 0;
  return red_plan(C, D, H, W, nullptr).bytes;
}

int satmvs_red_forward(const satmvs_red_weights* wt, const float* volume, int C, int D, int H, int W,
                       const float* const* state_in, float* const* state_out, float* logits,
                       void* workspace, size_t workspace_bytes, void* stream) {
  SATMVS_REQUIRE(wt && volume && logits && workspace);
  SATMVS_REQUIRE(C >= 1 && D >= 1 && H >= 8 && W >= 8 && H % 8 == 0 && W % 8 == 0);
  cudaStream_t st = (cudaStream_t)stream;
  RedPlan P = red_plan(C, D, H, W, reinterpret_cast<char*>(workspace));
  SATMVS_REQUIRE(workspace_bytes >= P.bytes);
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
  for (int i = 0; i < 3; ++i) {   // ConvReLU stride 2 (module.py:627-629); conv1 sees -cost
    ConvProblem p = plane_conv(xin[i], ech[i], D, H >> i, W >> i, wt->conv_w[i], (long long)ech[i] * 9, 9,
                               P.e[i], ech[i + 1], D, H >> (i + 1), W >> (i + 1), 2);
    p.Qd = D; p.Qh = H >> (i + 1); p.Qw = W >> (i + 1);
    p.relu = 1;
    p.acc_scale = (i == 0) ? -1.0f : 1.0f;
    RUN(launch_by_cout(p, st, "red encoder"));
  }
  for (int l = 0; l < 4; ++l) {   // x-halves of the GRU convolutions, bias folded in (module.py:29-30, :44-45)
    RedLevel& L = P.lv[l];
    const long long kin = (long long)(L.cx + L.ch) * 9;
    ConvProblem g = plane_conv(xin[l], L.cx, D, L.h, L.w, wt->gate_w[l], kin, 9, L.gx, 2 * L.ch, D, L.h, L.w, 1);
    g.Qd = D; g.Qh = L.h; g.Qw = L.w;
    g.shift = wt->gate_b[l];
    g.acc_scale = (l == 0) ? -1.0f : 1.0f;
    RUN(launch_by_cout(g, st, "red gate x-half"));
    ConvProblem o = plane_conv(xin[l], L.cx, D, L.h, L.w, wt->out_w[l], kin, 9, L.ox, L.ch, D, L.h, L.w, 1);
    o.Qd = D; o.Qh = L.h; o.Qw = L.w;
    o.shift = wt->out_b[l];
    o.acc_scale = (l == 0) ? -1.0f : 1.0f;
    RUN(launch_by_cout(o, st, "red output x-half"));
  }

  // ---- B. recurrence over planes ----
  for (int d = 0; d < D; ++d) {
    ConvGroup g1{}, g2{};
    GruArgs ga{};
    int total = 0;
    for (int l = 0; l < 4; ++l) {
      RedLevel& L = P.lv[l];
      const size_t px = (size_t)L.h * L.w;
      const long long kin = (long long)(L.cx + L.ch) * 9;
      double* stats = P.stats + ((size_t)d * 4 + l) * 6;
      // P1: gates += conv(h_prev; h-half of gate_conv.weight)
      ConvProblem a = plane_conv(L.s, L.ch, D + 1, L.h, L.w, wt->gate_w[l] + (size_t)L.cx * 9, kin, 9,
                                 L.gx, 2 * L.ch, D, L.h, L.w, 1);
      a.Qd = 1; a.Qh = L.h; a.Qw = L.w;
      a.q2i_add[0] = d; a.q2o_add[0] = d;
      a.pre_add = L.gx;
      a.stats = stats; a.stats_group = L.ch;
      conv_finalize(a);
      g1.p[l] = a;
      // P2: out += conv(r*h; h-half of output_conv.weight)
      ConvProblem b = plane_conv(L.rh, L.ch, 1, L.h, L.w, wt->out_w[l] + (size_t)L.cx * 9, kin, 9,
                                 L.ox, L.ch, D, L.h, L.w, 1);
      b.Qd = 1; b.Qh = L.h; b.Qw = L.w;
      b.q2o_add[0] = d;
      b.pre_add = L.ox;
      b.stats = stats + 4; b.stats_group = L.ch;
      conv_finalize(b);
      g2.p[l] = b;
      GruLevelArgs& e = ga.l[l];
      e.g = L.gx + (size_t)d * px; e.gcs = (long long)D * px;
      e.o = L.ox + (size_t)d * px; e.ocs = (long long)D * px;
      e.hprev = L.s + (size_t)d * px; e.hnext = L.s + (size_t)(d + 1) * px; e.scs = (long long)(D + 1) * px;
      e.rh = L.rh;
      e.rn_w = wt->rn_w[l]; e.rn_b = wt->rn_b[l]; e.un_w = wt->un_w[l]; e.un_b = wt->un_b[l];
      e.on_w = wt->on_w[l]; e.on_b = wt->on_b[l];
      e.stats = stats; e.ch = L.ch; e.px = (int)px; e.begin = total;
      total += L.ch * (int)px;
    }
    g1.n = g2.n = 4; ga.total = total;
    RUN(conv_launch<Tile8s>(g1, st, "red gate h-half"));
    gru_reset_kernel<<<ceil_div(total, 256), 256, 0, st>>>(ga);
    RUN(check_launch("gru_reset_kernel"));
    RUN(conv_launch<Tile8s>(g2, st, "red output h-half"));
    gru_update_kernel<<<ceil_div(total, 256), 256, 0, st>>>(ga);
    RUN(check_launch("gru_update_kernel"));
  }

  // ---- C. decoder over all planes: U_l = relu(convT_s2(U_{l+1})) + S_l  (module.py:633-642) ----
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
        p.Qd = D; p.Qh = Lin.h; p.Qw = Lin.w;
        p.q2i_add[0] = 1; p.q2o_add[0] = 1;                            // slots 1..D
        p.q2o_mul[1] = 2; p.q2o_mul[2] = 2; p.q2o_add[1] = py; p.q2o_add[2] = px;
        conv_taps_deconv_class(p, false, 0, py, px);
        p.relu = 1;
        p.post_add = L.s;
        conv_finalize(p);
        g.p[n++] = p;
      }
    g.n = n;
    if (L.ch >= 32) RUN(conv_launch<Tile32>(g, st, "red upconv")); else if (L.ch >= 16) RUN(conv_launch<Tile16>(g, st, "red upconv"));
    else RUN(conv_launch<Tile8>(g, st, "red upconv"));
    up_in = P.u[l];
  }
  {  // upconv2d: ConvTranspose2d(8, 1, k3, stride 1, pad 1) with bias (module.py:610, :643): out[o] = sum_k in[o+1-k] w[k]
    ConvProblem p;
    conv_problem_defaults(p);
    p.in = P.u[0]; p.w = wt->upconv2d_w; p.out = logits;
    p.Cin = 8; p.Cout = 1;
    p.Di = D + 1; p.Hi = H; p.Wi = W; p.Do = D; p.Ho = H; p.Wo = W;
    p.w_ci_stride = 9; p.w_co_stride = 9;
    p.Qd = D; p.Qh = H; p.Qw = W;
    p.q2i_add[0] = 1;
    int n = 0;
    for (int ky = 0; ky < 3; ++ky)
      for (int kx = 0; kx < 3; ++kx) { p.tap_dz[n] = 0; p.tap_dy[n] = 1 - ky; p.tap_dx[n] = 1 - kx; p.tap_w[n] = ky * 3 + kx; ++n; }
    p.ntaps = n;
    p.shift = wt->upconv2d_b;
    RUN(launch_one<Tile8>(p, st, "red upconv2d"));
  }
  if (state_out)
    for (int l = 0; l < 4; ++l)
      if (state_out[l]) {
        RedLevel& L = P.lv[l];
        const size_t px = (size_t)L.h * L.w;
        cudaMemcpy2DAsync(state_out[l], px * 4, L.s + (size_t)D * px, (size_t)(D + 1) * px * 4, px * 4, L.ch,
                          cudaMemcpyDeviceToDevice, st);
      }
#undef RUN
  return check_launch("satmvs_red_forward");
}

}  // extern "C"
