// remap.cu — bilinear gather of a height map at projected positions, for the RPC geometric-consistency filter.
//
// Reference: tools/rpc_filter.py:29-30 calls cv2.remap(depth_src, x_src.astype(float32), y_src.astype(float32),
// INTER_LINEAR, BORDER_CONSTANT, borderValue=-999).  OpenCV's remap (third-party, imgproc/src/imgwarp.cpp remap() +
// remapBilinear<float>) is restated here from its published algorithm: coordinates are rounded to 1/32 pixel
// (sx = cvRound(x * 32), round-half-even), the four weights are products of the two 1-D weights (1 - f/32, f/32)
// formed in fp32, out-of-range taps take the border value one by one, and the result is
// ((v00*w00 + v01*w01) + v10*w10) + v11*w11 in fp32 without contraction.
#include "common.cuh"

namespace satmvs {

__global__ void remap_bilinear_kernel(const float* __restrict__ src, int Hs, int Ws, const float* __restrict__ mapx,
                                      const float* __restrict__ mapy, long long n, float border, float* __restrict__ out) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  // cvRound: nearest, ties to even; the product with 32 is exact unless it overflows
  const float fx = __fmul_rn(__ldg(mapx + i), 32.0f), fy = __fmul_rn(__ldg(mapy + i), 32.0f);
  // non-finite or huge coordinates land far outside, like saturate_cast<short> of the integer part
  const bool sane = fabsf(fx) < 1.0e9f && fabsf(fy) < 1.0e9f;
  const int sx = sane ? __float2int_rn(fx) : (1 << 30), sy = sane ? __float2int_rn(fy) : (1 << 30);
  int x0 = sx >> 5, y0 = sy >> 5;
  x0 = max(-32768, min(32767, x0)); y0 = max(-32768, min(32767, y0));
  const float tx1 = __fmul_rn((float)(sx & 31), 1.0f / 32.0f), ty1 = __fmul_rn((float)(sy & 31), 1.0f / 32.0f);
  const float tx0 = __fsub_rn(1.0f, tx1), ty0 = __fsub_rn(1.0f, ty1);
  const float w00 = __fmul_rn(ty0, tx0), w01 = __fmul_rn(ty0, tx1), w10 = __fmul_rn(ty1, tx0), w11 = __fmul_rn(ty1, tx1);
  const bool xin0 = (unsigned)x0 < (unsigned)Ws, xin1 = (unsigned)(x0 + 1) < (unsigned)Ws;
  const bool yin0 = (unsigned)y0 < (unsigned)Hs, yin1 = (unsigned)(y0 + 1) < (unsigned)Hs;
  float r = border;
  if ((xin0 || xin1) && (yin0 || yin1)) {
    const float v00 = (xin0 && yin0) ? __ldg(src + (long long)y0 * Ws + x0) : border;
    const float v01 = (xin1 && yin0) ? __ldg(src + (long long)y0 * Ws + x0 + 1) : border;
    const float v10 = (xin0 && yin1) ? __ldg(src + (long long)(y0 + 1) * Ws + x0) : border;
    const float v11 = (xin1 && yin1) ? __ldg(src + (long long)(y0 + 1) * Ws + x0 + 1) : border;
    r = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(v00, w00), __fmul_rn(v01, w01)), __fmul_rn(v10, w10)), __fmul_rn(v11, w11));
  }
  out[i] = r;
}

}  // namespace satmvs

extern "C" int satmvs_remap_bilinear(const float* src, int Hs, int Ws, const float* mapx, const float* mapy, int64_t n,
                                     float border, float* out, void* stream) {
  using namespace satmvs;
  SATMVS_REQUIRE(src && mapx && mapy && out);
  SATMVS_REQUIRE(Hs >= 1 && Ws >= 1 && n >= 0);
  if (n == 0) return SATMVS_OK;
  remap_bilinear_kernel<<<ceil_div(n, 256), 256, 0, (cudaStream_t)stream>>>(src, Hs, Ws, mapx, mapy, n, border, out);
  return check_launch("remap_bilinear_kernel");
}
