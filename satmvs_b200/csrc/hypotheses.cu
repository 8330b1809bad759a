// hypotheses.cu — depth-hypothesis generation fused with the cascade's two resampling steps.
//
// Reference glue (networks/casred.py:132-145, == casmvs.py:148-163):
//   cur   = F.interpolate(prev_depth, [Himg, Wimg], bilinear, align_corners=False)        (stage >= 2)
//   samp  = get_depth_range_samples(cur, D, interval)   -> [D, Himg, Wimg]   (modules/depth_range.py:4-42)
//   dv    = F.interpolate(samp, [D, h, w], trilinear, align_corners=False)   -> [D, h, w]
// The reference materialises the full-resolution [D, Himg, Wimg] volume; here every output hypothesis
// is computed directly from the <= 4 image-resolution pixels it averages, each of which is a bilinear
// read of the previous stage's depth map.  The depth axis is not resampled (D -> D, source index = d).
// ATen's index/weight formulas (area_pixel_compute_source_index, align_corners=False) are restated.
#include "common.cuh"

namespace satmvs {

struct Lin1 { int i0, i1; float w0, w1; };

// linear resampling coefficients of output index `dst` for in_size -> out_size, align_corners=False
__device__ __forceinline__ Lin1 lin_coeff(int dst, int in_size, int out_size) {
  const float scale = (float)in_size / (float)out_size;
  float src = scale * ((float)dst + 0.5f) - 0.5f;
  src = fmaxf(src, 0.0f);
  Lin1 c;
  c.i0 = min((int)src, in_size - 1);
  c.i1 = min(c.i0 + 1, in_size - 1);
  c.w1 = src - (float)c.i0;
  c.w0 = 1.0f - c.w1;
  return c;
}

__global__ void depth_hypotheses_kernel(const float* __restrict__ prev, int hp, int wp,
                                        const float* __restrict__ range, int n_range,
                                        int D, float interval, int Himg, int Wimg, int h, int w,
                                        float* __restrict__ out) {
  const int pix = blockIdx.x * blockDim.x + threadIdx.x;
  if (pix >= h * w) return;
  const int oy = pix / w, ox = pix - oy * w;
  if (prev == nullptr) {
    // first stage: D uniform planes between range[0] and range[n_range-1] inclusive (depth_range.py:27-36)
    const float lo = __ldg(range), hi = __ldg(range + n_range - 1);
    const float step = __fdiv_rn(__fsub_rn(hi, lo), (float)(D - 1));
    for (int d = 0; d < D; ++d) out[(size_t)d * h * w + pix] = __fadd_rn(lo, __fmul_rn((float)d, step));
    return;
  }
  const Lin1 cy = lin_coeff(oy, Himg, h), cx = lin_coeff(ox, Wimg, w);     // stage grid <- image grid
  const int ys[2] = {cy.i0, cy.i1}, xs[2] = {cx.i0, cx.i1};
  float lo[2][2], step[2][2];
#pragma unroll
  for (int a = 0; a < 2; ++a)
#pragma unroll
    for (int b = 0; b < 2; ++b) {
      // cur = bilinear up-sample of the previous depth at image pixel (ys[a], xs[b])
      const Lin1 py = lin_coeff(ys[a], hp, Himg), px = lin_coeff(xs[b], wp, Wimg);
      const float v00 = __ldg(prev + py.i0 * wp + px.i0), v01 = __ldg(prev + py.i0 * wp + px.i1);
      const float v10 = __ldg(prev + py.i1 * wp + px.i0), v11 = __ldg(prev + py.i1 * wp + px.i1);
      const float cur = py.w0 * (px.w0 * v00 + px.w1 * v01) + py.w1 * (px.w0 * v10 + px.w1 * v11);
      // min = cur - D/2*I, max = cur + D/2*I, step = (max - min)/(D-1)   (depth_range.py:8-14)
      const float half = (float)D / 2.0f * interval;
      const float mn = __fsub_rn(cur, half), mx = __fadd_rn(cur, half);
      lo[a][b] = mn;
      step[a][b] = __fdiv_rn(__fsub_rn(mx, mn), (float)(D - 1));
    }
  for (int d = 0; d < D; ++d) {
    float s[2][2];
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
      for (int b = 0; b < 2; ++b) s[a][b] = __fadd_rn(lo[a][b], __fmul_rn((float)d, step[a][b]));
    out[(size_t)d * h * w + pix] = cy.w0 * (cx.w0 * s[0][0] + cx.w1 * s[0][1]) + cy.w1 * (cx.w0 * s[1][0] + cx.w1 * s[1][1]);
  }
}

// F.interpolate(x, [ho, wo], mode="bilinear", align_corners=False) over N planes (depth_regression resizes the
// hypotheses to the probability volume with it, modules/module.py:437)
__global__ void resize_bilinear_kernel(const float* __restrict__ in, int N, int hi, int wi, int ho, int wo, float* __restrict__ out) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)N * ho * wo) return;
  const int ox = (int)(i % wo), oy = (int)((i / wo) % ho), n = (int)(i / ((long long)wo * ho));
  const Lin1 cy = lin_coeff(oy, hi, ho), cx = lin_coeff(ox, wi, wo);
  const float* p = in + (size_t)n * hi * wi;
  out[i] = cy.w0 * (cx.w0 * __ldg(p + cy.i0 * wi + cx.i0) + cx.w1 * __ldg(p + cy.i0 * wi + cx.i1)) +
           cy.w1 * (cx.w0 * __ldg(p + cy.i1 * wi + cx.i0) + cx.w1 * __ldg(p + cy.i1 * wi + cx.i1));
}

}  // namespace satmvs

using namespace satmvs;

extern "C" int satmvs_resize_bilinear(const float* in, int N, int hi, int wi, int ho, int wo, float* out, void* stream) {
  SATMVS_REQUIRE(in && out && N >= 1 && hi >= 1 && wi >= 1 && ho >= 1 && wo >= 1);
  resize_bilinear_kernel<<<ceil_div((int64_t)N * ho * wo, 256), 256, 0, (cudaStream_t)stream>>>(in, N, hi, wi, ho, wo, out);
  return check_launch("resize_bilinear_kernel");
}

extern "C" int satmvs_depth_hypotheses(const float* prev_depth, int hp, int wp, const float* depth_range, int n_range,
                                       int D, float interval, int Himg, int Wimg, int h, int w, float* out, void* stream) {
  SATMVS_REQUIRE(out && D >= 2 && Himg >= 1 && Wimg >= 1 && h >= 1 && w >= 1);
  SATMVS_REQUIRE((prev_depth && hp >= 1 && wp >= 1) || (depth_range && n_range >= 2));
  depth_hypotheses_kernel<<<ceil_div((int64_t)h * w, 128), 128, 0, (cudaStream_t)stream>>>(
      prev_depth, hp, wp, depth_range, n_range, D, interval, Himg, Wimg, h, w, out);
  return check_launch("depth_hypotheses_kernel");
}
