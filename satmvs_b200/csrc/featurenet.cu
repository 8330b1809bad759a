// featurenet.cu — FeatureNet (2-D UNet, arch_mode "unet", 3 stages), the step directly in front of the plane sweep
// (SURVEY.md §8f-2).  Reference: modules/module.py:442-543 (FeatureNet), :78-159 (Conv2d / Deconv2d blocks: conv + BatchNorm
// + ReLU), :303-321 (DeConv2dFuse: stride-2 transposed conv, concatenation with the skip tensor, 3x3 conv); called once per
// view by networks/casred.py:116-119, :288-290.
//
// All V views of a stack go through every layer in ONE launch: tensors are [C][V][h][w] and the 2-D taps are applied per
// plane (the conv engine's "depth" axis carries the views).  Inference-mode BatchNorm arrives folded to per-channel
// scale / shift.  The two concatenations never happen: the skip tensors (conv0, conv1 outputs) are produced directly in the
// upper channel halves of the concatenation buffers and the transposed convs write the lower halves.
// Kernels: the 3x3 stride-1 layers with >= 8 input channels (7 of the 15 layers, 60 % of the FLOPs) run on the tensor cores
// (umma_conv.cuh: tcgen05 kind::tf32, 3-way split, fp32-faithful); the 3-channel first layer, the 5x5 stride-2 convs, the transposed
// convs and the 1x1 heads on register-tiled fp32 kernels (direct_conv.cuh); the implicit-GEMM engine (conv_engine.cuh) is the
// fallback for unaligned widths.
#include "conv_engine.cuh"
#include "direct_conv.cuh"
#include "umma_conv.cuh"
#include "prof.cuh"
#include <cstdlib>

namespace satmvs {

namespace {

struct FnTensor { float* p; int C, h, w; };

template <class T>
int fn_launch1(ConvProblem& p, cudaStream_t st, const char* what) {
  conv_finalize(p);
  ConvGroup g{};
  g.p[0] = p; g.n = 1;
  return conv_launch<T>(g, st, what);
}

int fn_launch_group(ConvGroup& g, int Cout, cudaStream_t st, const char* what) {
  if (Cout >= 32) return conv_launch<Tile32>(g, st, what);
  if (Cout >= 16) return conv_launch<Tile16>(g, st, what);
  return conv_launch<Tile8>(g, st, what);
}

int fn_launch(ConvProblem& p, cudaStream_t st, const char* what) {
  if (p.Cout >= 32) return fn_launch1<Tile32>(p, st, what);
  if (p.Cout >= 16) return fn_launch1<Tile16>(p, st, what);
  return fn_launch1<Tile8>(p, st, what);
}

// k x k convolution (k = 1, 3, 5; padding k/2) with stride s over every view plane; input channels [in_off, in_off + Cin) of
// `in`, output channels [out_off, out_off + Cout) of `out`
ConvProblem fn_conv(const FnTensor& in, int in_off, int Cin, const float* w, const float* scale, const float* shift, int relu,
                    const FnTensor& out, int out_off, int Cout, int k, int s, int V) {
  ConvProblem p;
  conv_problem_defaults(p);
  p.in = in.p; p.w = w; p.scale = scale; p.shift = shift; p.out = out.p; p.relu = relu;
  p.Cin = Cin; p.Cout = Cout; p.in_c_off = in_off; p.out_c_off = out_off;
  p.Di = V; p.Hi = in.h; p.Wi = in.w; p.Do = V; p.Ho = out.h; p.Wo = out.w;
  p.Qd = V; p.Qh = out.h; p.Qw = out.w;
  p.w_co_stride = (long long)Cin * k * k; p.w_ci_stride = k * k;
  p.q2i_mul[1] = s; p.q2i_mul[2] = s; p.q2i_add[1] = -(k / 2); p.q2i_add[2] = -(k / 2);
  int n = 0;
  for (int ky = 0; ky < k; ++ky)
    for (int kx = 0; kx < k; ++kx) { p.tap_dz[n] = 0; p.tap_dy[n] = ky; p.tap_dx[n] = kx; p.tap_w[n] = ky * k + kx; ++n; }
  p.ntaps = n;
  return p;
}

// ConvTranspose2d(k 3, stride 2, padding 1, output_padding 1) + folded BN + ReLU (Deconv2d, module.py:121-159): four
// output-parity problems in one grouped launch; weight layout [Cin][Cout][3][3]
int fn_deconv(const FnTensor& in, int Cin, const float* w, const float* scale, const float* shift, const FnTensor& out, int out_off,
              int Cout, int V, cudaStream_t st, const char* what) {
  {  // register-tiled direct kernel (direct_conv.cuh) when rows are 16-byte aligned
    DirectDeconv d{};
    const long long ics = (long long)V * in.h * in.w, ocs = (long long)V * out.h * out.w;
    d.in = in.p; d.in_cs = ics; d.w = w; d.scale = scale; d.shift = shift; d.out = out.p + (size_t)out_off * ocs; d.out_cs = ocs;
    d.Cin = Cin; d.Cout = Cout; d.Dn = V; d.Hi = in.h; d.Wi = in.w; d.relu = 1;
    static const bool no_direct = getenv("SATMVS_NO_DIRECT_CONV") != nullptr;
    if (!no_direct && direct_deconv_supported(d)) return direct_deconv_launch(d, st, what);
  }
  ConvGroup g{};
  int n = 0;
  for (int py = 0; py < 2; ++py)
    for (int px = 0; px < 2; ++px) {
      ConvProblem p;
      conv_problem_defaults(p);
      p.in = in.p; p.w = w; p.scale = scale; p.shift = shift; p.out = out.p; p.relu = 1;
      p.Cin = Cin; p.Cout = Cout; p.out_c_off = out_off;
      p.Di = V; p.Hi = in.h; p.Wi = in.w; p.Do = V; p.Ho = out.h; p.Wo = out.w;
      p.Qd = V; p.Qh = in.h; p.Qw = in.w;
      p.w_ci_stride = (long long)Cout * 9; p.w_co_stride = 9;
      p.q2o_mul[1] = 2; p.q2o_mul[2] = 2; p.q2o_add[1] = py; p.q2o_add[2] = px;
      conv_taps_deconv_class(p, false, 0, py, px);
      conv_finalize(p);
      g.p[n++] = p;
    }
  g.n = n;
  return fn_launch_group(g, Cout, st, what);
}

}  // namespace

}  // namespace satmvs

using namespace satmvs;

extern "C" {

static size_t fn_pack_bytes(int base) {   // packed tcgen05 weights of the seven tensor-core layers (Cin, Cout <= 4b), 256-byte slots
  const size_t one = (size_t)(4 * base / 8 + 1) * 2 * 9 * 2 * ((4 * base + 15) / 16 * 16) * 16 + 256;
  return 7 * one;
}

size_t satmvs_featurenet_workspace_bytes(int base, int V, int H, int W) {
  if (base < 1 || V < 1 || H < 4 || W < 4 || (H % 4) || (W % 4)) return 0;
  const size_t px = (size_t)V * H * W;
  // t0a b, cat2 2b, f2 b at full resolution; t1a 2b, t1b 2b, cat1 4b, f1 2b at 1/2; t2a, t2b, t2c 4b each at 1/4
  const size_t floats = (size_t)base * (4 * px + 10 * (px / 4) + 12 * (px / 16));
  return floats * sizeof(float) + 16 * 256 + fn_pack_bytes(base) + 512;
}

int satmvs_featurenet_forward(const satmvs_featurenet_weights* wt, const float* images, int base, int V, int H, int W,
                              float* out1, float* out2, float* out3, void* workspace, size_t workspace_bytes, void* stream) {
  SATMVS_CHECK_ASYNC();
  SATMVS_REQUIRE(wt && images && out1 && out2 && out3 && workspace);
  SATMVS_REQUIRE(base >= 1 && V >= 1 && H >= 4 && W >= 4 && H % 4 == 0 && W % 4 == 0);
  SATMVS_REQUIRE(workspace_bytes >= satmvs_featurenet_workspace_bytes(base, V, H, W));
  cudaStream_t st = (cudaStream_t)stream;
  const int b = base, H2 = H / 2, W2 = W / 2, H4 = H / 4, W4 = W / 4;
  char* cur = reinterpret_cast<char*>(workspace);
  auto take = [&](int C, int h, int w) {
    FnTensor t{reinterpret_cast<float*>(cur), C, h, w};
    cur += (((size_t)C * V * h * w * sizeof(float)) + 255) / 256 * 256;
    return t;
  };
  const FnTensor img{const_cast<float*>(images), 3, H, W};
  FnTensor t0a = take(b, H, W), cat2 = take(2 * b, H, W), f2 = take(b, H, W);
  FnTensor t1a = take(2 * b, H2, W2), t1b = take(2 * b, H2, W2), cat1 = take(4 * b, H2, W2), f1 = take(2 * b, H2, W2);
  FnTensor t2a = take(4 * b, H4, W4), t2b = take(4 * b, H4, W4), t2c = take(4 * b, H4, W4);
  const FnTensor o1{out1, 4 * b, H4, W4}, o2{out2, 2 * b, H2, W2}, o3{out3, b, H, W};
  cur = reinterpret_cast<char*>((reinterpret_cast<uintptr_t>(cur) + 255) / 256 * 256);
  int* tc_err = async_error_devptr() ? async_error_devptr() : reinterpret_cast<int*>(cur);
  cur += 256;
  const size_t pack_slot = fn_pack_bytes(b) / 7;
  static const bool no_umma = getenv("SATMVS_NO_UMMA") != nullptr;
  int rc;
#define RUN(x) do { rc = (x); if (rc) return rc; } while (0)
  ProfScope prof(kProfFeature, st);
  const satmvs_conv_bn* L = wt->block;
  int tc_slot = 0;
  auto conv = [&](int i, const FnTensor& in, int in_off, int Cin, const FnTensor& out, int out_off, int Cout, int k, int s, const char* what) {
    if (k == 3 && s == 1 && Cin % 8 == 0 && !no_umma && tc_slot < 7) {      // tensor cores: one head = this layer (folded BN + ReLU in the epilogue)
      const long long cs = (long long)V * in.h * in.w;
      UmmaPackHead wh{L[i].w, (long long)Cin * 9, 9, Cout, 0};
      UmmaHead oh{L[i].scale, L[i].shift, out.p + (size_t)out_off * cs, Cout, 0, 1.0f, 1, 1};
      UmmaConvPlan up;
      char* pk = cur + (size_t)tc_slot * pack_slot;
      if (umma_conv_plan(up, in.p + (size_t)in_off * cs, cs, Cin, V, in.h, in.w, 1, &wh, &oh, pk, pack_slot, 1, /*perf_rules=*/false)) {
        ++tc_slot;
        return umma_conv_launch(up, tc_err, st, what);
      }
    }
    static const bool no_direct = getenv("SATMVS_NO_DIRECT_CONV") != nullptr;
    if (k == 3 && s == 1 && !no_direct) {        // the 3-channel first layer: register-tiled direct kernel
      const long long ics = (long long)V * in.h * in.w, ocs = (long long)V * out.h * out.w;
      DirectConv d{};
      d.in = in.p + (size_t)in_off * ics; d.w = L[i].w; d.scale = L[i].scale; d.shift = L[i].shift; d.out = out.p + (size_t)out_off * ocs;
      d.Cin = Cin; d.Cout = Cout; d.Di = V; d.Hi = in.h; d.Wi = in.w; d.Do = V; d.Ho = out.h; d.Wo = out.w;
      d.w_co = (long long)Cin * 9; d.w_ci = 9; d.acc_scale = 1.0f; d.relu = 1;
      if (direct_conv_supported(d, 1, 1)) return direct_conv_launch(d, 1, 1, st, what);
    }
    if (k == 5 && s == 2 && !no_direct) {        // the two down-sampling layers: register-tiled direct kernel
      DirectConv5 d{};
      const long long ics = (long long)V * in.h * in.w, ocs = (long long)V * out.h * out.w;
      d.in = in.p + (size_t)in_off * ics; d.in_cs = ics; d.w = L[i].w; d.scale = L[i].scale; d.shift = L[i].shift;
      d.out = out.p + (size_t)out_off * ocs; d.out_cs = ocs; d.Cin = Cin; d.Cout = Cout; d.N = V; d.Hi = in.h; d.Wi = in.w; d.relu = 1;
      if (direct_conv5_supported(d)) return direct_conv5_launch(d, st, what);
    }
    ConvProblem p = fn_conv(in, in_off, Cin, L[i].w, L[i].scale, L[i].shift, 1, out, out_off, Cout, k, s, V);
    return fn_launch(p, st, what);
  };
  // bare 1x1 conv, no bias; output VIEW-major [V][C][h][w]: every view's feature map is a contiguous [C][h][w] tensor
  auto head = [&](const FnTensor& in, int Cin, const float* w, const FnTensor& out, const char* what) {
    const long long hw = (long long)in.h * in.w;
    Conv1x1 c{in.p, w, out.p, (long long)V * hw, Cin, Cin, hw};
    static const bool no_direct = getenv("SATMVS_NO_DIRECT_CONV") != nullptr;
    if (!no_direct && conv1x1_supported(c)) return conv1x1_launch(c, st, what);
    for (int v = 0; v < V; ++v) {      // implicit-GEMM engine: one launch per view, the input plane pinned
      const FnTensor ov{out.p + (size_t)v * Cin * hw, Cin, in.h, in.w};
      ConvProblem p = fn_conv(in, 0, Cin, w, nullptr, nullptr, 0, ov, 0, Cin, 1, 1, V);
      p.Do = 1; p.Qd = 1; p.q2i_add[0] = v;
      const int r = fn_launch(p, st, what);
      if (r) return r;
    }
    return (int)SATMVS_OK;
  };
  // conv0 (module.py:452-455); its output is the skip tensor of deconv2: upper half of cat2
  RUN(conv(0, img, 0, 3, t0a, 0, b, 3, 1, "featurenet conv0.0"));
  RUN(conv(1, t0a, 0, b, cat2, b, b, 3, 1, "featurenet conv0.1"));
  // conv1 (:457-461): 5x5 stride 2, then two 3x3; output = skip tensor of deconv1: upper half of cat1
  RUN(conv(2, cat2, b, b, t1a, 0, 2 * b, 5, 2, "featurenet conv1.0"));
  RUN(conv(3, t1a, 0, 2 * b, t1b, 0, 2 * b, 3, 1, "featurenet conv1.1"));
  RUN(conv(4, t1b, 0, 2 * b, cat1, 2 * b, 2 * b, 3, 1, "featurenet conv1.2"));
  // conv2 (:463-467)
  RUN(conv(5, cat1, 2 * b, 2 * b, t2a, 0, 4 * b, 5, 2, "featurenet conv2.0"));
  RUN(conv(6, t2a, 0, 4 * b, t2b, 0, 4 * b, 3, 1, "featurenet conv2.1"));
  RUN(conv(7, t2b, 0, 4 * b, t2c, 0, 4 * b, 3, 1, "featurenet conv2.2"));
  RUN(head(t2c, 4 * b, wt->out_w[0], o1, "featurenet out1"));       // out1: bare 1x1 conv, no bias (:469, :511-512)
  // deconv1 = DeConv2dFuse(4b -> 2b) (:474, :515): transposed conv into the lower half of cat1, 3x3 conv over both halves
  RUN(fn_deconv(t2c, 4 * b, L[8].w, L[8].scale, L[8].shift, cat1, 0, 2 * b, V, st, "featurenet deconv1.deconv"));
  RUN(conv(9, cat1, 0, 4 * b, f1, 0, 2 * b, 3, 1, "featurenet deconv1.conv"));
  RUN(head(f1, 2 * b, wt->out_w[1], o2, "featurenet out2"));
  // deconv2 = DeConv2dFuse(2b -> b) (:475, :519)
  RUN(fn_deconv(f1, 2 * b, L[10].w, L[10].scale, L[10].shift, cat2, 0, b, V, st, "featurenet deconv2.deconv"));
  RUN(conv(11, cat2, 0, 2 * b, f2, 0, b, 3, 1, "featurenet deconv2.conv"));
  RUN(head(f2, b, wt->out_w[2], o3, "featurenet out3"));
#undef RUN
  return check_launch("satmvs_featurenet_forward");
}

}  // extern "C"
