// convop.cu — one convolution block of the regularisers as a C-ABI operator: ConvReLU (modules/module.py:178-186, 3x3
// applied per depth plane) and Conv3d (modules/module.py:324-366, 3x3x3 + folded BatchNorm + ReLU), padding 1.
// The regularisers call the same kernels internally; this entry exists so that the tensor-core path (umma_conv.cuh) and
// the fp32 FFMA path (direct_conv.cuh) can be driven and compared on arbitrary shapes (tests/test_gpu_convop.py).
#include "direct_conv.cuh"
#include "umma_conv.cuh"

using namespace satmvs;

extern "C" {

size_t satmvs_conv_workspace_bytes(int Cin, int Cout, int NZ) {
  if (Cin < 1 || Cout < 1 || (NZ != 1 && NZ != 3)) return 0;
  const size_t shifted = (size_t)((Cin + 7) / 8) * NZ * 2 * 9 * 2 * (((Cout + 7) / 8 * 8 + 15) / 16 * 16) * 16;
  const size_t taps_in_n = (size_t)((Cin + 7) / 8) * NZ * 2 * 2 * kTnNP * 16;
  return (shifted > taps_in_n ? shifted : taps_in_n) + 256;
}

int satmvs_conv_forward(const float* in, int Cin, int D, int H, int W, const float* w, const float* scale, const float* shift,
                        int Cout, int NZ, int stride, int relu, float acc_scale, float* out, int engine,
                        void* workspace, size_t workspace_bytes, void* stream) {
  SATMVS_CHECK_ASYNC();
  SATMVS_REQUIRE(in && w && out);
  SATMVS_REQUIRE(Cin >= 1 && Cout >= 1 && D >= 1 && H >= 1 && W >= 1 && (NZ == 1 || NZ == 3) && (stride == 1 || stride == 2));
  SATMVS_REQUIRE(engine >= 0 && engine <= 3);
  if (stride == 2) SATMVS_REQUIRE(H % 2 == 0 && W % 2 == 0 && (NZ == 1 || D % 2 == 0));
  cudaStream_t st = (cudaStream_t)stream;
  const int taps = NZ * 9;
  if ((engine == 3 || engine == 0) && workspace && stride == 1 && Cout <= kTnCo) {
    // few output channels: the nine in-plane taps ride in the N dimension (one pass over the A operand)
    char* ws = static_cast<char*>(workspace);
    UmmaConvTnPlan tp;
    const size_t cap = workspace_bytes > 256 ? workspace_bytes - 256 : 0;
    if (umma_conv_tn_plan(tp, in, (long long)D * H * W, Cin, D, H, W, w, (long long)Cin * taps, taps, scale, shift, out, Cout, NZ,
                          relu, acc_scale, ws + 256, cap) && (engine == 3 || (long long)tp.grid.x * tp.grid.y >= 96))
      return umma_conv_tn_launch(tp, reinterpret_cast<int*>(ws), st, "satmvs_conv_forward (tcgen05, taps in N)");
  }
  if (engine == 3) return fail_invalid("shape does not fit the taps-in-N tcgen05 convolution (stride 1, Cout <= 8, Cin % 8) or no workspace");
  if (engine != 2 && workspace && (NZ == 1 || stride == 1)) {
    char* ws = static_cast<char*>(workspace);
    UmmaPackHead wh{w, (long long)Cin * taps, taps, Cout, 0};
    UmmaHead oh{scale, shift, out, Cout, 0, acc_scale, relu, stride};
    UmmaConvPlan up;
    const size_t cap = workspace_bytes > 256 ? workspace_bytes - 256 : 0;
    const bool fits = umma_conv_plan(up, in, (long long)D * H * W, Cin, D, H, W, 1, &wh, &oh, ws + 256, cap, NZ, engine != 1);
    // automatic mode keeps small-N 27-tap layers on the FFMA kernel (costreg.cu)
    if (fits && (engine == 1 || NZ == 1 || Cout >= 16))
      return umma_conv_launch(up, reinterpret_cast<int*>(ws), st, "satmvs_conv_forward (tcgen05)");
  }
  if (engine == 1) return fail_invalid("shape does not fit the tcgen05 convolution (Cin % 8, shared memory, grid size) or no workspace");
  DirectConv d{};
  d.in = in; d.w = w; d.scale = scale; d.shift = shift; d.out = out;
  d.Cin = Cin; d.Cout = Cout; d.Di = D; d.Hi = H; d.Wi = W;
  d.Do = (NZ == 3) ? D / stride : D; d.Ho = H / stride; d.Wo = W / stride;
  d.w_co = (long long)Cin * taps; d.w_ci = taps; d.acc_scale = acc_scale; d.relu = relu;
  if (!direct_conv_supported(d, NZ, stride)) return fail_invalid("direct convolution needs widths that are multiples of 4 and 16-byte aligned tensors");
  return direct_conv_launch(d, NZ, stride, st, "satmvs_conv_forward (direct)");
}

}  // extern "C"
