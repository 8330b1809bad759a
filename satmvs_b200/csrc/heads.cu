// heads.cu — soft-argmin depth regression heads.
//
//   mode 0  RED train head  (networks/casred.py:58-62):   p = softmax_d(logits); depth = sum p*d; conf = max p
//   mode 1  CasMVS head     (networks/casmvs.py:66-74):   conf = sum of the 4 probabilities around
//           the regressed plane index (pad 1 before / 2 after, index = trunc(sum p*k) clamped)
//   mode 2  depth_regression (modules/module.py:433-439) on given probabilities: depth = sum p*d only
//   streaming fp64 head of the inference net (networks/casred.py:182-184, :218-236)
//
// One thread per pixel, planes strided by H*W so every load is 128-byte coalesced across the warp.
// HBM-bound: (4 + 4) bytes per (plane, pixel) read once (logits are re-read from L2 for the second pass).
#include "common.cuh"
#include "prof.cuh"

namespace satmvs {

__global__ void softargmin_kernel(const float* __restrict__ logits, const float* __restrict__ depth,
                                  int depth_per_pixel, int mode, int D, int HW,
                                  float* __restrict__ out_depth, float* __restrict__ out_conf) {
  const int pix = blockIdx.x * blockDim.x + threadIdx.x;
  if (pix >= HW) return;
  if (mode == 2) {   // depth_regression on given probabilities (modules/module.py:433-439): sum_d p*d, nothing else
    float acc = 0.0f;
    for (int d = 0; d < D; ++d) {
      const float dv = depth_per_pixel ? __ldg(depth + (size_t)d * HW + pix) : __ldg(depth + d);
      acc += __fmul_rn(__ldg(logits + (size_t)d * HW + pix), dv);
    }
    out_depth[pix] = acc;
    return;
  }
  // pass 1: max over planes (F.softmax subtracts the max)
  float mx = -INFINITY;
  for (int d = 0; d < D; ++d) mx = fmaxf(mx, __ldg(logits + (size_t)d * HW + pix));
  // pass 2: sum of exp
  float se = 0.0f;
  for (int d = 0; d < D; ++d) se += expf(__ldg(logits + (size_t)d * HW + pix) - mx);
  // pass 3: probabilities -> expectation, index expectation, max prob
  float acc = 0.0f, idx = 0.0f, pmax = 0.0f;
  for (int d = 0; d < D; ++d) {
    const float p = expf(__ldg(logits + (size_t)d * HW + pix) - mx) / se;
    const float dv = depth_per_pixel ? __ldg(depth + (size_t)d * HW + pix) : __ldg(depth + d);
    acc += __fmul_rn(p, dv);
    idx += __fmul_rn(p, (float)d);
    pmax = fmaxf(pmax, p);
  }
  out_depth[pix] = acc;
  if (mode == 0) {
    out_conf[pix] = pmax;
  } else {
    int k = (int)idx;                     // .long() truncates (casmvs.py:71)
    k = min(max(k, 0), D - 1);
    float s4 = 0.0f;                      // 4 * avg_pool3d over planes k-1 .. k+2 of the zero-padded volume
    for (int j = k - 1; j <= k + 2; ++j)
      if (j >= 0 && j < D) s4 += expf(__ldg(logits + (size_t)j * HW + pix) - mx) / se;
    out_conf[pix] = s4;
  }
}

__global__ void softargmin_stream_update_kernel(const float* __restrict__ reg, const float* __restrict__ depth_plane,
                                                int depth_per_pixel, int HW, double* __restrict__ state) {
  const int pix = blockIdx.x * blockDim.x + threadIdx.x;
  if (pix >= HW) return;
  const double e = exp((double)__ldg(reg + pix));           // reg_cost.double().exp(), casred.py:218-219
  const double dv = (double)(depth_per_pixel ? __ldg(depth_plane + pix) : __ldg(depth_plane));
  double* es = state, *da = state + HW, *me = state + 2 * (size_t)HW;
  es[pix] = es[pix] + e;
  da[pix] = dv * e + da[pix];                               // casred.py:227 (product rounded, then added)
  if (me[pix] < e) me[pix] = e;
}

// K planes in their order, same arithmetic per plane as the single-plane update: the chunked plane-streaming stage
// (satmvs_b200/stages.py) feeds the regularised planes of a whole chunk at once
__global__ void softargmin_stream_update_planes_kernel(const float* __restrict__ reg, const float* __restrict__ depth,
                                                       int depth_per_pixel, int K, int HW, double* __restrict__ state) {
  const int pix = blockIdx.x * blockDim.x + threadIdx.x;
  if (pix >= HW) return;
  double es = state[pix], da = state[HW + pix], me = state[2 * (size_t)HW + pix];
  for (int k = 0; k < K; ++k) {
    const double e = exp((double)__ldg(reg + (size_t)k * HW + pix));
    const double dv = (double)(depth_per_pixel ? __ldg(depth + (size_t)k * HW + pix) : __ldg(depth + k));
    es = es + e;
    da = dv * e + da;
    if (me < e) me = e;
  }
  state[pix] = es; state[HW + pix] = da; state[2 * (size_t)HW + pix] = me;
}

// Plane sweep without a regulariser: the matching cost of plane k is scale * mean_c var[c,k] (scale < 0: low variance = good match);
// the planes of a [C,K,H,W] variance slab are folded into the fp64 running sums in their order, so a depth-sharded sweep only
// has to exchange the three [H,W] sums.  The slab is read once, coalesced across pixels.
__global__ void softargmin_stream_update_volume_kernel(const float* __restrict__ var, const float* __restrict__ depth,
                                                       int depth_per_pixel, float scale, int C, int K, int HW,
                                                       double* __restrict__ state) {
  const int pix = blockIdx.x * blockDim.x + threadIdx.x;
  if (pix >= HW) return;
  double es = state[pix], da = state[HW + pix], me = state[2 * (size_t)HW + pix];
  const float inv_c = 1.0f / (float)C;
  for (int k = 0; k < K; ++k) {
    float s = 0.0f;
    for (int c = 0; c < C; ++c) s += __ldg(var + ((size_t)c * K + k) * HW + pix);
    const float reg = scale * (s * inv_c);
    const double e = exp((double)reg);
    const double dv = (double)(depth_per_pixel ? __ldg(depth + (size_t)k * HW + pix) : __ldg(depth + k));
    es = es + e;
    da = dv * e + da;
    if (me < e) me = e;
  }
  state[pix] = es; state[HW + pix] = da; state[2 * (size_t)HW + pix] = me;
}

__global__ void softargmin_stream_finish_kernel(const double* __restrict__ state, int HW,
                                                float* __restrict__ out_depth, float* __restrict__ out_conf) {
  const int pix = blockIdx.x * blockDim.x + threadIdx.x;
  if (pix >= HW) return;
  const double tot = state[pix] + 1e-10;                    // casred.py:234
  out_depth[pix] = (float)(state[HW + pix] / tot);
  out_conf[pix] = (float)(state[2 * (size_t)HW + pix] / tot);
}

// gradient of depth = sum_d p_d * dv_d (p = softmax over planes) to the logits: dL/dlogit_d = g * p_d * (dv_d - depth)
__global__ void softargmin_bwd_kernel(const float* __restrict__ logits, const float* __restrict__ depth, int depth_per_pixel,
                                      int D, int HW, const float* __restrict__ grad_depth, float* __restrict__ grad_logits) {
  const int pix = blockIdx.x * blockDim.x + threadIdx.x;
  if (pix >= HW) return;
  float mx = -INFINITY;
  for (int d = 0; d < D; ++d) mx = fmaxf(mx, __ldg(logits + (size_t)d * HW + pix));
  float se = 0.0f, acc = 0.0f;
  for (int d = 0; d < D; ++d) {
    const float e = expf(__ldg(logits + (size_t)d * HW + pix) - mx);
    const float dv = depth_per_pixel ? __ldg(depth + (size_t)d * HW + pix) : __ldg(depth + d);
    se += e; acc += e * dv;
  }
  const float inv = 1.0f / se, mean = acc * inv, g = __ldg(grad_depth + pix);
  for (int d = 0; d < D; ++d) {
    const float p = expf(__ldg(logits + (size_t)d * HW + pix) - mx) * inv;
    const float dv = depth_per_pixel ? __ldg(depth + (size_t)d * HW + pix) : __ldg(depth + d);
    grad_logits[(size_t)d * HW + pix] = g * p * (dv - mean);
  }
}

}  // namespace satmvs

using namespace satmvs;

extern "C" {

int satmvs_softargmin_fwd(const float* logits, const float* depth, int depth_per_pixel, int mode,
                          int D, int H, int W, float* out_depth, float* out_conf, void* stream) {
  SATMVS_REQUIRE(logits && depth && out_depth && (out_conf || mode == 2));
  SATMVS_REQUIRE(D >= 1 && H >= 1 && W >= 1 && mode >= 0 && mode <= 2);
  const int HW = H * W;
  ProfScope prof(kProfHead, (cudaStream_t)stream);
  softargmin_kernel<<<ceil_div(HW, 128), 128, 0, (cudaStream_t)stream>>>(logits, depth, depth_per_pixel, mode, D, HW,
                                                                         out_depth, out_conf);
  return check_launch("softargmin_kernel");
}

int satmvs_softargmin_stream_update(const float* reg, const float* depth_plane, int depth_per_pixel,
                                    int H, int W, double* state, void* stream) {
  SATMVS_REQUIRE(reg && depth_plane && state && H >= 1 && W >= 1);
  const int HW = H * W;
  softargmin_stream_update_kernel<<<ceil_div(HW, 128), 128, 0, (cudaStream_t)stream>>>(reg, depth_plane, depth_per_pixel,
                                                                                       HW, state);
  return check_launch("softargmin_stream_update_kernel");
}

int satmvs_softargmin_stream_update_planes(const float* reg, const float* depth, int depth_per_pixel, int K,
                                           int H, int W, double* state, void* stream) {
  SATMVS_REQUIRE(reg && depth && state && K >= 1 && H >= 1 && W >= 1);
  const int HW = H * W;
  softargmin_stream_update_planes_kernel<<<ceil_div(HW, 128), 128, 0, (cudaStream_t)stream>>>(reg, depth, depth_per_pixel, K, HW, state);
  return check_launch("softargmin_stream_update_planes_kernel");
}

int satmvs_softargmin_stream_update_volume(const float* var, const float* depth, int depth_per_pixel, float scale, int C, int K,
                                           int H, int W, double* state, void* stream) {
  SATMVS_REQUIRE(var && depth && state && C >= 1 && K >= 1 && H >= 1 && W >= 1);
  const int HW = H * W;
  ProfScope prof(kProfHead, (cudaStream_t)stream);
  softargmin_stream_update_volume_kernel<<<ceil_div(HW, 64), 64, 0, (cudaStream_t)stream>>>(var, depth, depth_per_pixel, scale, C, K,
                                                                                            HW, state);
  return check_launch("softargmin_stream_update_volume_kernel");
}

int satmvs_softargmin_stream_finish(const double* state, int H, int W,
                                    float* out_depth, float* out_conf, void* stream) {
  SATMVS_REQUIRE(state && out_depth && out_conf && H >= 1 && W >= 1);
  const int HW = H * W;
  softargmin_stream_finish_kernel<<<ceil_div(HW, 128), 128, 0, (cudaStream_t)stream>>>(state, HW, out_depth, out_conf);
  return check_launch("softargmin_stream_finish_kernel");
}

int satmvs_softargmin_bwd(const float* logits, const float* depth, int depth_per_pixel, int D, int H, int W,
                          const float* grad_depth, float* grad_logits, void* stream) {
  SATMVS_REQUIRE(logits && depth && grad_depth && grad_logits && D >= 1 && H >= 1 && W >= 1);
  const int HW = H * W;
  softargmin_bwd_kernel<<<ceil_div(HW, 128), 128, 0, (cudaStream_t)stream>>>(logits, depth, depth_per_pixel, D, HW, grad_depth,
                                                                             grad_logits);
  return check_launch("softargmin_bwd_kernel");
}

}  // extern "C"
