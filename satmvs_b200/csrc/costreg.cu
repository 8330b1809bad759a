// costreg.cu — CostRegNet, the 3-D conv UNet regulariser of CasMVSNet / UCS-Net, on the conv engine.
//
// Reference: CostRegNet.forward (modules/module.py:546-577) with blocks Conv3d (:324-366) and
// Deconv3d (:369-410): every block is conv(bias=False) + BatchNorm3d + ReLU; the three transposed
// blocks add a skip tensor AFTER the ReLU (x = conv4 + conv7(x), module.py:573-575); `prob` is a bare
// 3x3x3 conv to one channel.  BatchNorm is applied in inference form (running statistics) folded to
// a per-channel scale/shift by the host wrapper.
//
// 11 launches: 7 dense 27-tap convs (3 of them stride 2), 3 transposed convs (each one grouped launch
// of its 8 output-parity classes, so no multiply ever touches a structural zero), 1 head conv.
#include "conv_engine.cuh"
#include "direct_conv.cuh"
#include "umma_conv.cuh"
#include "prof.cuh"

namespace satmvs {

static ConvProblem conv3d_problem(const float* in, int Cin, int Di, int Hi, int Wi, const float* w,
                                  float* out, int Cout, int stride) {
  ConvProblem p;
  conv_problem_defaults(p);
  p.in = in; p.w = w; p.out = out;
  p.Cin = Cin; p.Cout = Cout;
  p.Di = Di; p.Hi = Hi; p.Wi = Wi;
  p.Do = Di / stride; p.Ho = Hi / stride; p.Wo = Wi / stride;
  p.Qd = p.Do; p.Qh = p.Ho; p.Qw = p.Wo;
  p.w_co_stride = (long long)Cin * 27; p.w_ci_stride = 27;          // nn.Conv3d weight [Cout][Cin][3][3][3]
  for (int i = 0; i < 3; ++i) { p.q2i_mul[i] = stride; p.q2i_add[i] = -1; }
  conv_taps_dense(p, true);
  p.relu = 1;
  return p;
}

template <class T>
static int run_conv(ConvProblem p, cudaStream_t st, const char* what) {
  conv_finalize(p);
  ConvGroup g{};
  g.p[0] = p; g.n = 1;
  return conv_launch<T>(g, st, what);
}

static int run_by_cout(const ConvProblem& p, cudaStream_t st, const char* what, float* ksplit_buf = nullptr, size_t ksplit_floats = 0) {
  {  // dense 3x3x3 layers: register-tiled direct kernel when rows are 16-byte aligned
    DirectConv d{};
    d.in = p.in; d.w = p.w; d.scale = p.scale; d.shift = p.shift; d.post_add = p.post_add; d.out = p.out;
    d.Cin = p.Cin; d.Cout = p.Cout; d.Di = p.Di; d.Hi = p.Hi; d.Wi = p.Wi; d.Do = p.Do; d.Ho = p.Ho; d.Wo = p.Wo;
    d.w_co = p.w_co_stride; d.w_ci = p.w_ci_stride; d.acc_scale = p.acc_scale; d.relu = p.relu;
    const int stride = p.q2i_mul[0];
    static const bool no_direct = getenv("SATMVS_NO_DIRECT_CONV") != nullptr;
    if (!no_direct && p.ntaps == 27 && stride == 1 && direct_conv3d_c1_supported(d)) return direct_conv3d_c1_launch(d, st, what);
    if (!no_direct && p.ntaps == 27 && direct_conv_supported(d, 3, stride)) {
      // small deep layers (conv5 / conv6: 40 output tiles of 8 channels x 512 voxels): split the input channels over grid.z
      const long long tiles = (long long)ceil_div((long long)p.Do * p.Ho * p.Wo, kDcWarps * 32 * kDcPx) * ceil_div(p.Cout, kDcCo);
      int ks = 1;
      while (ks < 4 && tiles * ks < kNumSMs && p.Cin / (2 * ks) >= kDcCiChunk) ks *= 2;
      if (ks > 1 && ksplit_buf && (size_t)ks * p.Cout * p.Do * p.Ho * p.Wo <= ksplit_floats) { d.ksplit = ks; d.partial = ksplit_buf; }
      return direct_conv_launch(d, 3, stride, st, what);
    }
  }
  if (p.Cout >= 64) return run_conv<Tile64>(p, st, what);
  if (p.Cout >= 32) return run_conv<Tile32>(p, st, what);
  if (p.Cout >= 16) return run_conv<Tile16>(p, st, what);
  return run_conv<Tile8>(p, st, what);
}

// ConvTranspose3d(k=3, stride 2, padding 1, output_padding 1): 8 parity classes in one launch
static int run_deconv3d(const float* in, int Cin, int Di, int Hi, int Wi, const float* w, const float* scale,
                        const float* shift, const float* skip, float* out, int Cout, cudaStream_t st) {
  {  // register-tiled direct kernel (direct_conv.cuh) when rows are 16-byte aligned; the implicit-GEMM engine otherwise
    DirectDeconv3d d{};
    d.in = in; d.w = w; d.w_ci = (long long)Cout * 27; d.w_co = 27; d.scale = scale; d.shift = shift; d.post_add = skip; d.out = out;
    d.Cin = Cin; d.Cout = Cout; d.Di = Di; d.Hi = Hi; d.Wi = Wi; d.relu = 1;
    static const bool no_direct = getenv("SATMVS_NO_DIRECT_CONV") != nullptr;
    if (!no_direct && direct_deconv3d_supported(d)) return direct_deconv3d_launch(d, st, "costreg deconv (direct)");
  }
  ConvGroup g{};
  int n = 0;
  for (int pz = 0; pz < 2; ++pz)
    for (int py = 0; py < 2; ++py)
      for (int px = 0; px < 2; ++px) {
        ConvProblem p;
        conv_problem_defaults(p);
        p.in = in; p.w = w; p.out = out;
        p.Cin = Cin; p.Cout = Cout;
        p.Di = Di; p.Hi = Hi; p.Wi = Wi; p.Do = 2 * Di; p.Ho = 2 * Hi; p.Wo = 2 * Wi;
        p.Qd = Di; p.Qh = Hi; p.Qw = Wi;
        p.w_ci_stride = (long long)Cout * 27; p.w_co_stride = 27;   // nn.ConvTranspose3d weight [Cin][Cout][3][3][3]
        for (int i = 0; i < 3; ++i) p.q2o_mul[i] = 2;
        p.q2o_add[0] = pz; p.q2o_add[1] = py; p.q2o_add[2] = px;
        conv_taps_deconv_class(p, true, pz, py, px);
        p.scale = scale; p.shift = shift; p.relu = 1; p.post_add = skip;
        conv_finalize(p);
        g.p[n++] = p;
      }
  g.n = n;
  if (Cout >= 32) return conv_launch<Tile32>(g, st, "costreg deconv");
  if (Cout >= 16) return conv_launch<Tile16>(g, st, "costreg deconv");
  return conv_launch<Tile8>(g, st, "costreg deconv");
}

struct CostRegPlan {
  float* c[7]; float* x7; float* x9; float* x11;
  float* ksplit; size_t ksplit_floats;         // partial sums of the layers whose input channels are split over CTAs
  char* wpack[5]; size_t wpack_bytes[5];       // packed (raw, lo) weights of the tensor-core layers: conv0, conv2, conv4, conv6, prob
  int* umma_err;
  size_t bytes;
};

static CostRegPlan costreg_plan(int base, int D, int H, int W, char* mem) {
  CostRegPlan p{};
  size_t off = 0;
  auto take = [&](size_t n) { float* r = reinterpret_cast<float*>(mem + off); off += (n * 4 + 255) / 256 * 256; return r; };
  const size_t v0 = (size_t)D * H * W, v1 = v0 / 8, v2 = v1 / 8, v3 = v2 / 8;
  p.c[0] = take(base * v0);
  p.c[1] = take(2 * base * v1); p.c[2] = take(2 * base * v1);
  p.c[3] = take(4 * base * v2); p.c[4] = take(4 * base * v2);
  p.c[5] = take(8 * base * v3); p.c[6] = take(8 * base * v3);
  p.x7 = take(4 * base * v2); p.x9 = take(2 * base * v1); p.x11 = take(base * v0);
  p.ksplit_floats = 4 * 8 * base * v3 > 2 * 4 * base * v2 ? 4 * 8 * base * v3 : 2 * 4 * base * v2;   // conv5 / conv6 (x4), conv4 (x2)
  p.ksplit = take(p.ksplit_floats);
  // (Cin / 8) x 3 planes x (raw, lo) x 9 taps x 2 quads x N (padded to 16) float4; conv0's Cin is bounded by 64 here
  const int wcin[5] = {64, 2 * base, 4 * base, 8 * base, base}, wn[5] = {base, 2 * base, 4 * base, 8 * base, 1};
  for (int i = 0; i < 5; ++i) {
    p.wpack_bytes[i] = (size_t)((wcin[i] + 7) / 8) * 3 * 2 * 9 * 2 * ((wn[i] + 15) / 16 * 16) * 16;
    p.wpack[i] = mem + off;
    off += (p.wpack_bytes[i] + 255) / 256 * 256;
  }
  p.umma_err = reinterpret_cast<int*>(mem + off);
  off += 256;
  p.bytes = off;
  return p;
}

}  // namespace satmvs

using namespace satmvs;

extern "C" {

size_t satmvs_costreg_workspace_bytes(int base, int D, int H, int W) {
  if (base < 1 || D < 8 || H < 8 || W < 8 || (D % 8) || (H % 8) || (W % 8)) return 0;
  return costreg_plan(base, D, H, W, nullptr).bytes;
}

int satmvs_costreg_forward(const satmvs_costreg_weights* wt, const float* x, int Cin, int base, int D, int H, int W,
                           float* out, void* workspace, size_t workspace_bytes, void* stream) {
  SATMVS_CHECK_ASYNC();
  SATMVS_REQUIRE(wt && x && out && workspace);
  SATMVS_REQUIRE(Cin >= 1 && base >= 1 && D >= 8 && H >= 8 && W >= 8 && D % 8 == 0 && H % 8 == 0 && W % 8 == 0);
  cudaStream_t st = (cudaStream_t)stream;
  CostRegPlan P = costreg_plan(base, D, H, W, reinterpret_cast<char*>(workspace));
  SATMVS_REQUIRE(workspace_bytes >= P.bytes);
  int rc;
  ProfScope prof(kProfCostReg, st);
#define RUN(e) do { rc = (e); if (rc) return rc; } while (0)
  // conv0..conv6 (module.py:569-572): channel and stride schedule
  const int cin[7] = {Cin, base, 2 * base, 2 * base, 4 * base, 4 * base, 8 * base};
  const int cout[7] = {base, 2 * base, 2 * base, 4 * base, 4 * base, 8 * base, 8 * base};
  const int stride[7] = {1, 2, 1, 2, 1, 2, 1};
  const float* cur = x;
  int d = D, h = H, w = W;
  static const bool no_umma = getenv("SATMVS_NO_UMMA") != nullptr;
  // dense stride-1 3x3x3 layer on the tensor cores (umma_conv.cuh); false when the shape does not fit
  auto try_umma = [&](int slot, const float* in, int ci, int dd, int hh, int ww, const float* wgt, const float* scale,
                      const float* shift, int relu, float* o, int co) -> int {
    if (no_umma) return -1;
    // Small-N 27-tap layers (conv0: N 8, prob: N 1): with shifted descriptors they are bound by the 27 x 3 re-reads of the
    // A operand from shared memory (335 us against 284 us for the direct FFMA kernel at cfg-2), so their nine in-plane
    // taps ride in the N dimension instead (umma_conv_tn_kernel: one pass over A, shifted row sum in the epilogue).
    // Measured at cfg-2 (64 planes of 96x192): conv0 (Cin 32, Cout 8) 322 us against 453 us on the FFMA kernel; prob (Cin 8,
    // Cout 1) 137 us against 115 us -- the fixed cost per CTA (TMEM allocation, nine-tap epilogue) needs >= 16 input channels
    // to pay off (profiles/r01_umma_conv_notes.md).
    if (co <= kTnCo) {
      if (ci < 16) return -1;
      UmmaConvTnPlan tp;
      if (!umma_conv_tn_plan(tp, in, (long long)dd * hh * ww, ci, dd, hh, ww, wgt, (long long)ci * 27, 27, scale, shift, o, co, 3,
                             relu, 1.0f, P.wpack[slot], P.wpack_bytes[slot]) || (long long)tp.grid.x * tp.grid.y < 96) return -1;
      return umma_conv_tn_launch(tp, P.umma_err, st, "costreg conv (tcgen05, taps in N)");
    }
    if (co < 16) return -1;
    UmmaPackHead wh{wgt, (long long)ci * 27, 27, co, 0};
    UmmaHead oh{scale, shift, o, co, 0, 1.0f, relu, 1};
    UmmaConvPlan up;
    if (!umma_conv_plan(up, in, (long long)dd * hh * ww, ci, dd, hh, ww, 1, &wh, &oh, P.wpack[slot], P.wpack_bytes[slot], 3)) return -1;
    return umma_conv_launch(up, P.umma_err, st, "costreg conv (tcgen05)");
  };
  for (int i = 0; i < 7; ++i) {
    if (stride[i] == 1) {
      const int r = try_umma(i / 2, cur, cin[i], d, h, w, wt->conv_w[i], wt->bn_scale[i], wt->bn_shift[i], 1, P.c[i], cout[i]);
      if (r > 0) return r;
      if (r == 0) { cur = P.c[i]; continue; }
    }
    ConvProblem p = conv3d_problem(cur, cin[i], d, h, w, wt->conv_w[i], P.c[i], cout[i], stride[i]);
    p.scale = wt->bn_scale[i]; p.shift = wt->bn_shift[i];
    RUN(run_by_cout(p, st, "costreg conv", P.ksplit, P.ksplit_floats));
    cur = P.c[i];
    d /= stride[i]; h /= stride[i]; w /= stride[i];
  }
  // conv7 / conv9 / conv11: transposed, BN + ReLU, then + skip (module.py:573-575)
  RUN(run_deconv3d(P.c[6], 8 * base, d, h, w, wt->conv_w[7], wt->bn_scale[7], wt->bn_shift[7], P.c[4], P.x7, 4 * base, st));
  RUN(run_deconv3d(P.x7, 4 * base, 2 * d, 2 * h, 2 * w, wt->conv_w[8], wt->bn_scale[8], wt->bn_shift[8], P.c[2], P.x9, 2 * base, st));
  RUN(run_deconv3d(P.x9, 2 * base, 4 * d, 4 * h, 4 * w, wt->conv_w[9], wt->bn_scale[9], wt->bn_shift[9], P.c[0], P.x11, base, st));
  {  // prob: bare Conv3d(base, 1, 3, padding=1, bias=False) (module.py:566, :576)
    const int r = try_umma(4, P.x11, base, D, H, W, wt->prob_w, nullptr, nullptr, 0, out, 1);
    if (r > 0) return r;
    if (r < 0) {
      ConvProblem p = conv3d_problem(P.x11, base, D, H, W, wt->prob_w, out, 1, 1);
      p.relu = 0;
      RUN(run_by_cout(p, st, "costreg prob"));
    }
  }
#undef RUN
  return SATMVS_OK;
}

}  // extern "C"
