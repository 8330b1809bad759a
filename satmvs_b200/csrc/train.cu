// train.cu — the training form of the Conv3d / Deconv3d blocks of CostRegNet (modules/module.py:324-410, :546-577) as C-ABI
// primitives: train.py:267-287 runs the network in train() mode (BatchNorm3d on batch statistics) and calls loss.backward().
//
//   forward of a block    y = conv(x, W)                       satmvs_conv3d_raw     (modes 0 / 1: conv stride 1 / 2, 3: transposed)
//                         z = relu(BN_batch(y)) [+ skip]       satmvs_bn_train_fwd
//   backward of a block   dy = BN'(relu'(dz))                  satmvs_bn_train_bwd   (also d gamma, d beta)
//                         dx = conv^T(dy, W)                   satmvs_conv3d_raw     (the data gradient of a stride-1 conv is the conv
//                                                              with mirrored taps and swapped channel strides (mode 2); of a stride-2 conv
//                                                              the transposed conv reading the SAME weight tensor (mode 3); of the
//                                                              transposed conv the stride-2 conv reading the same weight tensor (mode 1))
//                         dW = sum_o dy[o] x[s o + k - 1]      satmvs_conv3d_wgrad
//
// The data-path convolutions reuse the fp32 engines of the inference path (direct_conv.cuh, conv_engine.cuh); the weight gradient
// is a contraction over all positions: every CTA walks output rows, keeps the nine (kz, ky) input rows of a few input channels and
// the output-gradient row of all output channels in shared memory, accumulates (co, ci, kz, ky) x 3 kx partial filters in registers,
// and a second kernel adds the per-CTA partials in a fixed order (deterministic, no float atomics).
#include "conv_engine.cuh"
#include "direct_conv.cuh"
#include "prof.cuh"

namespace satmvs {

// ----------------------------------------------------------------------------------------------------------------------------
// BatchNorm (training form) on [B][C][n] tensors
// ----------------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) bn_stats_kernel(const float* __restrict__ y, int B, int C, long long n, double* __restrict__ acc) {
  const int c = blockIdx.y;
  double s = 0.0, q = 0.0;
  const long long n4 = (n & 3) ? 0 : (n >> 2);          // rows of a multiple of 4 floats are read as float4; anything else scalar
  for (int b = 0; b < B; ++b) {
    const float* pb = y + ((long long)b * C + c) * n;
    const float4* p = reinterpret_cast<const float4*>(pb);
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
      const float4 v = __ldg(p + i);
      const float ls = (v.x + v.y) + (v.z + v.w);
      const float lq = (v.x * v.x + v.y * v.y) + (v.z * v.z + v.w * v.w);
      s += ls; q += lq;
    }
    for (long long i = 4 * n4 + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
      const float v = __ldg(pb + i);
      s += v; q += v * v;
    }
  }
  __shared__ double red[2][8];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) { s += __shfl_xor_sync(0xffffffffu, s, o); q += __shfl_xor_sync(0xffffffffu, q, o); }
  if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = s; red[1][threadIdx.x >> 5] = q; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double ts = 0.0, tq = 0.0;
    for (int i = 0; i < 8; ++i) { ts += red[0][i]; tq += red[1][i]; }
    atomicAdd(acc + 2 * c, ts);
    atomicAdd(acc + 2 * c + 1, tq);
  }
}

__global__ void bn_finalize_kernel(const double* __restrict__ acc, int C, double count, float* __restrict__ mean, float* __restrict__ var) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const double m = acc[2 * c] / count;
  double v = acc[2 * c + 1] / count - m * m;
  if (v < 0.0) v = 0.0;
  mean[c] = (float)m;
  var[c] = (float)v;      // biased, as used for normalisation; the caller derives the unbiased running_var update
}

__global__ void __launch_bounds__(256) bn_apply_kernel(const float* __restrict__ y, int B, int C, long long n, const float* __restrict__ gamma,
                                                       const float* __restrict__ beta, const float* __restrict__ mean,
                                                       const float* __restrict__ var, float eps, int relu,
                                                       const float* __restrict__ post_add, float* __restrict__ z) {
  const int c = blockIdx.y;
  const float istd = rsqrtf(var[c] + eps);
  const float g = (gamma ? gamma[c] : 1.0f) * istd, sh = (beta ? beta[c] : 0.0f) - mean[c] * g;
  const long long n4 = (n & 3) ? 0 : (n >> 2);
  for (int b = 0; b < B; ++b) {
    const long long base = ((long long)b * C + c) * n;
    const float4* p = reinterpret_cast<const float4*>(y + base);
    const float4* a = post_add ? reinterpret_cast<const float4*>(post_add + base) : nullptr;
    float4* o = reinterpret_cast<float4*>(z + base);
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
      float4 v = __ldg(p + i);
      v.x = fmaf(v.x, g, sh); v.y = fmaf(v.y, g, sh); v.z = fmaf(v.z, g, sh); v.w = fmaf(v.w, g, sh);
      if (relu) { v.x = fmaxf(v.x, 0.0f); v.y = fmaxf(v.y, 0.0f); v.z = fmaxf(v.z, 0.0f); v.w = fmaxf(v.w, 0.0f); }
      if (a) { const float4 s = __ldg(a + i); v.x += s.x; v.y += s.y; v.z += s.z; v.w += s.w; }
      o[i] = v;
    }
    for (long long i = 4 * n4 + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
      float v = fmaf(__ldg(y + base + i), g, sh);
      if (relu) v = fmaxf(v, 0.0f);
      if (post_add) v += __ldg(post_add + base + i);
      z[base + i] = v;
    }
  }
}

// sums of the masked output gradient and of (masked gradient x normalised activation) per channel
__global__ void __launch_bounds__(256) bn_bwd_reduce_kernel(const float* __restrict__ dz, const float* __restrict__ dz2, const float* __restrict__ y, int B, int C, long long n,
                                                            const float* __restrict__ gamma, const float* __restrict__ beta,
                                                            const float* __restrict__ mean, const float* __restrict__ var, float eps,
                                                            int relu, double* __restrict__ acc) {
  const int c = blockIdx.y;
  const float istd = rsqrtf(var[c] + eps), m = mean[c];
  const float g = (gamma ? gamma[c] : 1.0f) * istd, sh = (beta ? beta[c] : 0.0f) - m * g;
  double s1 = 0.0, s2 = 0.0;
  const long long n4 = (n & 3) ? 0 : (n >> 2);
  for (int b = 0; b < B; ++b) {
    const long long base = ((long long)b * C + c) * n;
    const float4* py = reinterpret_cast<const float4*>(y + base);
    const float4* pd = reinterpret_cast<const float4*>(dz + base);
    const float4* pe = dz2 ? reinterpret_cast<const float4*>(dz2 + base) : nullptr;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
      const float4 v = __ldg(py + i);
      float4 d = __ldg(pd + i);
      if (pe) { const float4 e = __ldg(pe + i); d.x += e.x; d.y += e.y; d.z += e.z; d.w += e.w; }
      if (relu) {
        if (fmaf(v.x, g, sh) <= 0.0f) d.x = 0.0f;
        if (fmaf(v.y, g, sh) <= 0.0f) d.y = 0.0f;
        if (fmaf(v.z, g, sh) <= 0.0f) d.z = 0.0f;
        if (fmaf(v.w, g, sh) <= 0.0f) d.w = 0.0f;
      }
      const float l1 = (d.x + d.y) + (d.z + d.w);
      const float l2 = (d.x * ((v.x - m) * istd) + d.y * ((v.y - m) * istd)) + (d.z * ((v.z - m) * istd) + d.w * ((v.w - m) * istd));
      s1 += l1; s2 += l2;
    }
    for (long long i = 4 * n4 + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
      const float v = __ldg(y + base + i);
      float d = __ldg(dz + base + i) + (dz2 ? __ldg(dz2 + base + i) : 0.0f);
      if (relu && fmaf(v, g, sh) <= 0.0f) d = 0.0f;
      s1 += d; s2 += d * ((v - m) * istd);
    }
  }
  __shared__ double red[2][8];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) { s1 += __shfl_xor_sync(0xffffffffu, s1, o); s2 += __shfl_xor_sync(0xffffffffu, s2, o); }
  if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = s1; red[1][threadIdx.x >> 5] = s2; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double a = 0.0, b2 = 0.0;
    for (int i = 0; i < 8; ++i) { a += red[0][i]; b2 += red[1][i]; }
    atomicAdd(acc + 2 * c, a);
    atomicAdd(acc + 2 * c + 1, b2);
  }
}

__global__ void __launch_bounds__(256) bn_bwd_apply_kernel(const float* __restrict__ dz, const float* __restrict__ dz2, const float* __restrict__ y, int B, int C, long long n,
                                                           const float* __restrict__ gamma, const float* __restrict__ beta,
                                                           const float* __restrict__ mean, const float* __restrict__ var, float eps,
                                                           int relu, const double* __restrict__ acc, double count, float* __restrict__ dy,
                                                           float* __restrict__ dgamma, float* __restrict__ dbeta) {
  const int c = blockIdx.y;
  const float istd = rsqrtf(var[c] + eps), m = mean[c];
  const float gm = gamma ? gamma[c] : 1.0f;
  const float g = gm * istd, sh = (beta ? beta[c] : 0.0f) - m * g;
  const float k1 = (float)(acc[2 * c] / count), k2 = (float)(acc[2 * c + 1] / count);
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    if (dgamma) dgamma[c] = (float)acc[2 * c + 1];
    if (dbeta) dbeta[c] = (float)acc[2 * c];
  }
  const long long n4 = (n & 3) ? 0 : (n >> 2);
  for (int b = 0; b < B; ++b) {
    const long long base = ((long long)b * C + c) * n;
    const float4* py = reinterpret_cast<const float4*>(y + base);
    const float4* pd = reinterpret_cast<const float4*>(dz + base);
    const float4* pe = dz2 ? reinterpret_cast<const float4*>(dz2 + base) : nullptr;
    float4* po = reinterpret_cast<float4*>(dy + base);
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
      const float4 v = __ldg(py + i);
      float4 d = __ldg(pd + i);
      if (pe) { const float4 e = __ldg(pe + i); d.x += e.x; d.y += e.y; d.z += e.z; d.w += e.w; }
      if (relu) {
        if (fmaf(v.x, g, sh) <= 0.0f) d.x = 0.0f;
        if (fmaf(v.y, g, sh) <= 0.0f) d.y = 0.0f;
        if (fmaf(v.z, g, sh) <= 0.0f) d.z = 0.0f;
        if (fmaf(v.w, g, sh) <= 0.0f) d.w = 0.0f;
      }
      float4 r;
      r.x = g * (d.x - k1 - (v.x - m) * istd * k2);
      r.y = g * (d.y - k1 - (v.y - m) * istd * k2);
      r.z = g * (d.z - k1 - (v.z - m) * istd * k2);
      r.w = g * (d.w - k1 - (v.w - m) * istd * k2);
      po[i] = r;
    }
    for (long long i = 4 * n4 + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
      const float v = __ldg(y + base + i);
      float d = __ldg(dz + base + i) + (dz2 ? __ldg(dz2 + base + i) : 0.0f);
      if (relu && fmaf(v, g, sh) <= 0.0f) d = 0.0f;
      dy[base + i] = g * (d - k1 - (v - m) * istd * k2);
    }
  }
}

// ----------------------------------------------------------------------------------------------------------------------------
// weight gradient of a 3x3x3 (NZ 3) or 3x3 (NZ 1) convolution with padding 1 and stride S
// ----------------------------------------------------------------------------------------------------------------------------
struct WgradArgs {
  const float* x;      // [Cin][Di][Hi][Wi]
  const float* dy;     // [Cout][Do][Ho][Wo]
  float* partial;      // [gridDim.x * split][Cout][Cin][NZ * 9]
  int Cin, Cout, Di, Hi, Wi, Do, Ho, Wo, NZ;
  int ci_tile;         // input channels per CTA (grid.y = Cin / ci_tile, rounded up)
  int rb;              // consecutive output rows per strip: their S * (rb - 1) + 3 input rows are staged once
  int rsx, rsy;        // shared-memory row strides (floats): rs / 4 odd, so that lanes on different rows hit different banks
  int split;           // threads sharing one (co pair, ci, kz, ky) item, each on its own range of output columns
  int groups_per_split;  // 4-column groups per such range
};

constexpr int kWgThreads = 512;
constexpr int kWgMaxItems = 4;    // work units (item x column range) per thread

__device__ __forceinline__ void cp_async4(float* dst_smem, const float* src, bool valid) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(dst_smem);
  const int bytes = valid ? 4 : 0;                                     // src-size 0: the destination is zero-filled
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(d), "l"(src), "r"(bytes) : "memory");
}

template <int S>
__global__ void __launch_bounds__(kWgThreads, 2) wgrad_kernel(const WgradArgs a) {
  extern __shared__ __align__(16) float smem[];
  const int KZ = a.NZ == 3 ? 3 : 1, nkk = KZ * 3;
  const int NR = S * (a.rb - 1) + 3;                        // input rows of a strip (per kz plane)
  float* xs = smem;                                        // [ci_tile][KZ][NR][rsx]; index p of a row <-> input column p - 1
  float* dys = smem + (size_t)a.ci_tile * KZ * NR * a.rsx;  // [Cout][rb][rsy]
  const int ci0 = blockIdx.y * a.ci_tile;
  const int nci = min(a.ci_tile, a.Cin - ci0);
  const int cop = (a.Cout + 1) >> 1;                        // output-channel pairs: one staged input window feeds two filters
  const int items = cop * nci * nkk;
  const int units = items * a.split;
  float acc[kWgMaxItems][2][3];
#pragma unroll
  for (int i = 0; i < kWgMaxItems; ++i)
#pragma unroll
    for (int h = 0; h < 2; ++h) acc[i][h][0] = acc[i][h][1] = acc[i][h][2] = 0.0f;
  const int strips_y = (a.Ho + a.rb - 1) / a.rb;
  const int strips = a.Do * strips_y;
  const int wo4 = (a.Wo + 3) & ~3;
  const long long in_cs = (long long)a.Di * a.Hi * a.Wi, out_cs = (long long)a.Do * a.Ho * a.Wo;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  // the zero padding of every staged row (column -1 and everything from the row's end on) is written once: the copies below only
  // touch the interior
  for (int row = warp; row < a.Cout * a.rb; row += nwarps)
    for (int ox = a.Wo + lane; ox < a.rsy; ox += 32) dys[(size_t)row * a.rsy + ox] = 0.0f;
  for (int row = warp; row < nci * KZ * NR; row += nwarps) {
    float* dst = xs + (size_t)row * a.rsx;
    if (lane == 0) dst[0] = 0.0f;
    for (int p = a.Wi + 1 + lane; p < a.rsx; p += 32) dst[p] = 0.0f;
  }
  for (int sidx = blockIdx.x; sidx < strips; sidx += gridDim.x) {
    const int oz = sidx / strips_y, oy0 = (sidx - oz * strips_y) * a.rb;
    __syncthreads();
    // staging by asynchronous 4-byte copies (every element of the strip in flight at once): one warp per shared-memory row, the row
    // decode is warp-uniform, lanes walk the columns; rows outside the tensors are zero-filled with plain stores
    for (int row = warp; row < a.Cout * a.rb; row += nwarps) {   // output-gradient rows
      const int co = row / a.rb, r = row - co * a.rb;
      float* dst = dys + (size_t)row * a.rsy;
      if (oy0 + r < a.Ho) {
        const float* src = a.dy + co * out_cs + ((long long)oz * a.Ho + oy0 + r) * a.Wo;
        for (int ox = lane; ox < a.Wo; ox += 32) cp_async4(dst + ox, src + ox, true);
      } else {
        for (int ox = lane; ox < a.Wo; ox += 32) dst[ox] = 0.0f;
      }
    }
    for (int row = warp; row < nci * KZ * NR; row += nwarps) {   // input rows
      const int cl = row / (KZ * NR), rem = row - cl * (KZ * NR), kz = rem / NR, j = rem - kz * NR;
      const int iz = a.NZ == 3 ? S * oz + kz - 1 : oz, iy = S * oy0 - 1 + j;
      float* dst = xs + (size_t)row * a.rsx + 1;
      if (iz >= 0 && iz < a.Di && iy >= 0 && iy < a.Hi) {
        const float* src = a.x + (ci0 + cl) * in_cs + ((long long)iz * a.Hi + iy) * a.Wi;
        for (int p = lane; p < a.Wi; p += 32) cp_async4(dst + p, src + p, true);
      } else {
        for (int p = lane; p < a.Wi; p += 32) dst[p] = 0.0f;
      }
    }
    asm volatile("cp.async.wait_all;" ::: "memory");
    __syncthreads();
#pragma unroll
    for (int it = 0; it < kWgMaxItems; ++it) {
      const int unit = threadIdx.x + it * blockDim.x;
      if (unit < units) {
        const int sp = unit / items, id = unit - sp * items;      // items fastest: the lanes of a warp read different rows (fewer bank conflicts)
        const int cp = id / (nci * nkk), rem = id - cp * (nci * nkk);
        const int cl = rem / nkk, kk = rem - cl * nkk, kz = kk / 3, ky = kk - kz * 3;
        const int co0 = 2 * cp, co1 = min(2 * cp + 1, a.Cout - 1);       // an odd Cout computes its last filter twice (stored once)
        const int ox_lo = sp * a.groups_per_split * 4, ox_hi = min(wo4, ox_lo + a.groups_per_split * 4);
        float a0 = acc[it][0][0], a1 = acc[it][0][1], a2 = acc[it][0][2];
        float b0 = acc[it][1][0], b1 = acc[it][1][1], b2 = acc[it][1][2];
        for (int r = 0; r < a.rb; ++r) {
          const float* xr = xs + (size_t)((cl * KZ + kz) * NR + S * r + ky) * a.rsx;
          const float* dr = dys + (size_t)(co0 * a.rb + r) * a.rsy;
          const float* er = dys + (size_t)(co1 * a.rb + r) * a.rsy;
#pragma unroll 4
          for (int ox = ox_lo; ox < ox_hi; ox += 4) {
            const float4 d = *reinterpret_cast<const float4*>(dr + ox);
            const float4 e = *reinterpret_cast<const float4*>(er + ox);
            float x0, x1, x2, x3, x4, x5, x6, x7, x8;
            if (S == 1) {
              const float4 u = *reinterpret_cast<const float4*>(xr + ox);
              const float2 v = *reinterpret_cast<const float2*>(xr + ox + 4);
              x0 = u.x; x1 = u.y; x2 = u.z; x3 = u.w; x4 = v.x; x5 = v.y; x6 = x7 = x8 = 0.0f;
              a0 = fmaf(d.x, x0, a0); a0 = fmaf(d.y, x1, a0); a0 = fmaf(d.z, x2, a0); a0 = fmaf(d.w, x3, a0);
              a1 = fmaf(d.x, x1, a1); a1 = fmaf(d.y, x2, a1); a1 = fmaf(d.z, x3, a1); a1 = fmaf(d.w, x4, a1);
              a2 = fmaf(d.x, x2, a2); a2 = fmaf(d.y, x3, a2); a2 = fmaf(d.z, x4, a2); a2 = fmaf(d.w, x5, a2);
              b0 = fmaf(e.x, x0, b0); b0 = fmaf(e.y, x1, b0); b0 = fmaf(e.z, x2, b0); b0 = fmaf(e.w, x3, b0);
              b1 = fmaf(e.x, x1, b1); b1 = fmaf(e.y, x2, b1); b1 = fmaf(e.z, x3, b1); b1 = fmaf(e.w, x4, b1);
              b2 = fmaf(e.x, x2, b2); b2 = fmaf(e.y, x3, b2); b2 = fmaf(e.z, x4, b2); b2 = fmaf(e.w, x5, b2);
            } else {
              const float4 u = *reinterpret_cast<const float4*>(xr + 2 * ox);
              const float4 v = *reinterpret_cast<const float4*>(xr + 2 * ox + 4);
              x0 = u.x; x1 = u.y; x2 = u.z; x3 = u.w; x4 = v.x; x5 = v.y; x6 = v.z; x7 = v.w; x8 = xr[2 * ox + 8];
              a0 = fmaf(d.x, x0, a0); a0 = fmaf(d.y, x2, a0); a0 = fmaf(d.z, x4, a0); a0 = fmaf(d.w, x6, a0);
              a1 = fmaf(d.x, x1, a1); a1 = fmaf(d.y, x3, a1); a1 = fmaf(d.z, x5, a1); a1 = fmaf(d.w, x7, a1);
              a2 = fmaf(d.x, x2, a2); a2 = fmaf(d.y, x4, a2); a2 = fmaf(d.z, x6, a2); a2 = fmaf(d.w, x8, a2);
              b0 = fmaf(e.x, x0, b0); b0 = fmaf(e.y, x2, b0); b0 = fmaf(e.z, x4, b0); b0 = fmaf(e.w, x6, b0);
              b1 = fmaf(e.x, x1, b1); b1 = fmaf(e.y, x3, b1); b1 = fmaf(e.z, x5, b1); b1 = fmaf(e.w, x7, b1);
              b2 = fmaf(e.x, x2, b2); b2 = fmaf(e.y, x4, b2); b2 = fmaf(e.z, x6, b2); b2 = fmaf(e.w, x8, b2);
            }
          }
        }
        acc[it][0][0] = a0; acc[it][0][1] = a1; acc[it][0][2] = a2;
        acc[it][1][0] = b0; acc[it][1][1] = b1; acc[it][1][2] = b2;
      }
    }
  }
  const int taps = a.NZ * 9;
#pragma unroll
  for (int it = 0; it < kWgMaxItems; ++it) {
    const int unit = threadIdx.x + it * blockDim.x;
    if (unit < units) {
      const int sp = unit / items, id = unit - sp * items;      // items fastest: the lanes of a warp read different rows (fewer bank conflicts)
      const int cp = id / (nci * nkk), rem = id - cp * (nci * nkk);
      const int cl = rem / nkk, kk = rem - cl * nkk;
      float* out = a.partial + ((size_t)blockIdx.x * a.split + sp) * a.Cout * a.Cin * taps;
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int co = 2 * cp + h;
        if (co < a.Cout) {
          float* o = out + ((size_t)co * a.Cin + ci0 + cl) * taps + kk * 3;
          o[0] = acc[it][h][0]; o[1] = acc[it][h][1]; o[2] = acc[it][h][2];
        }
      }
    }
  }
}

__global__ void wgrad_reduce_kernel(const float* __restrict__ partial, int chunks, int Cout, int Cin, int taps, float* __restrict__ dw,
                                    long long dw_co, long long dw_ci, int accumulate) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int total = Cout * Cin * taps;
  if (i >= total) return;
  float s = 0.0f;
  for (int c = 0; c < chunks; ++c) s += partial[(size_t)c * total + i];
  const int co = i / (Cin * taps), rem = i - co * (Cin * taps), ci = rem / taps, t = rem - ci * taps;
  float* o = dw + co * dw_co + ci * dw_ci + t;
  *o = accumulate ? *o + s : s;
}

struct WgradPlan { int ci_tile, rb, rsx, rsy, chunks, groups, threads, split, groups_per_split; size_t smem, ws_bytes; bool ok; };

static WgradPlan wgrad_plan(int Cin, int Cout, int Wi, int Do, int Ho, int Wo, int NZ, int stride) {
  WgradPlan p{};
  const int wo4 = (Wo + 3) & ~3;
  auto odd4 = [](int n) { n = (n + 3) & ~3; if (((n >> 2) & 1) == 0) n += 4; return n; };
  p.rsy = odd4(wo4);
  const int need = stride * wo4 + 12;                 // the furthest float the vector loads of the last group touch
  p.rsx = odd4(need > Wi + 2 ? need : Wi + 2);
  const int KZ = NZ == 3 ? 3 : 1, nkk = KZ * 3;
  p.ok = false;
  const int cand[8][2] = {{4, 8}, {4, 4}, {2, 8}, {2, 4}, {1, 8}, {1, 4}, {1, 2}, {1, 1}};   // (rows per strip, channels per CTA)
  static const int rb_cap = getenv("SATMVS_WGRAD_RB") ? atoi(getenv("SATMVS_WGRAD_RB")) : 2;
  static const int ci_cap = getenv("SATMVS_WGRAD_CI") ? atoi(getenv("SATMVS_WGRAD_CI")) : 8;
  for (int c = 0; c < 8 && !p.ok; ++c) {
    if (cand[c][0] > rb_cap || cand[c][1] > ci_cap) continue;
    const int rb = cand[c][0] < Ho ? cand[c][0] : Ho, t = cand[c][1];
    const int NR = stride * (rb - 1) + 3;
    const size_t sm = ((size_t)t * KZ * NR * p.rsx + (size_t)Cout * rb * p.rsy) * 4;
    const int items = (Cout + 1) / 2 * t * nkk;
    if (sm <= 100 * 1024 && items <= kWgMaxItems * kWgThreads) { p.ci_tile = t; p.rb = rb; p.smem = sm; p.ok = true; }
  }
  if (!p.ok) return p;
  {  // threads: every (co pair, ci, kz, ky) item is shared by `split` threads on disjoint column ranges until the CTA has ~512
     // work units: the staged strip is the same, the thread-level parallelism is what hides the shared-memory latency
    const int items = (Cout + 1) / 2 * (Cin < p.ci_tile ? Cin : p.ci_tile) * nkk;
    const int groups4 = wo4 / 4;
    static const int target = getenv("SATMVS_WGRAD_UNITS") ? atoi(getenv("SATMVS_WGRAD_UNITS")) : 512;
    int split = 1;
    while (items * split * 2 <= target && split * 2 <= groups4) split *= 2;
    p.split = split;
    p.groups_per_split = (groups4 + split - 1) / split;
    const int units = items * split;
    int threads = units < kWgThreads ? (units + 31) / 32 * 32 : kWgThreads;
    const int passes = (units + threads - 1) / threads;
    threads = ((units + passes - 1) / passes + 31) / 32 * 32;          // even out the passes
    p.threads = threads;
    if (passes > kWgMaxItems) { p.ok = false; return p; }
  }
  p.groups = (Cin + p.ci_tile - 1) / p.ci_tile;
  const int strips = Do * ((Ho + p.rb - 1) / p.rb);
  int chunks = (2 * kNumSMs + p.groups - 1) / p.groups;
  if (chunks > strips) chunks = strips;
  if (chunks < 1) chunks = 1;
  p.chunks = chunks;
  p.ws_bytes = (size_t)chunks * p.split * Cout * Cin * NZ * 9 * 4;
  return p;
}

// ----------------------------------------------------------------------------------------------------------------------------
// weight gradient of a K x K (K = 1 or 5) per-plane convolution with padding K/2 and stride S: the small layers of FeatureNet
// (5x5 stride-2 and 1x1 convs, modules/module.py:456-470).  One thread per filter element and position chunk; partials as above.
// ----------------------------------------------------------------------------------------------------------------------------
struct WgradGenArgs {
  const float* x; const float* dy; float* partial;
  int Cin, Cout, N, Hi, Wi, Ho, Wo, K, S, chunk_len;
};

__global__ void __launch_bounds__(256) wgrad_generic_kernel(const WgradGenArgs a) {
  const int taps = a.K * a.K, total = a.Cout * a.Cin * taps;
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= total) return;
  const int co = e / (a.Cin * taps), rem = e - co * (a.Cin * taps), ci = rem / taps, t = rem - ci * taps;
  const int ky = t / a.K, kx = t - ky * a.K, pad = a.K / 2;
  const long long npos = (long long)a.N * a.Ho * a.Wo;
  const long long o0 = (long long)blockIdx.y * a.chunk_len;
  long long o1 = o0 + a.chunk_len;
  if (o1 > npos) o1 = npos;
  const float* dyc = a.dy + (long long)co * npos;
  const float* xc = a.x + (long long)ci * a.N * a.Hi * a.Wi;
  float acc = 0.0f;
  for (long long o = o0; o < o1; ++o) {
    const int n = (int)(o / (a.Ho * a.Wo)), r = (int)(o - (long long)n * a.Ho * a.Wo), oy = r / a.Wo, ox = r - oy * a.Wo;
    const int iy = a.S * oy + ky - pad, ix = a.S * ox + kx - pad;
    if (iy >= 0 && iy < a.Hi && ix >= 0 && ix < a.Wi)
      acc = fmaf(__ldg(dyc + o), __ldg(xc + ((long long)n * a.Hi + iy) * a.Wi + ix), acc);
  }
  a.partial[(size_t)blockIdx.y * total + e] = acc;
}

}  // namespace satmvs

using namespace satmvs;

extern "C" {

int satmvs_conv3d_raw(const float* in, int Cin, int Di, int Hi, int Wi, const float* w, long long w_co, long long w_ci,
                      int NZ, int mode, float* out, int Cout, void* stream) {
  SATMVS_CHECK_ASYNC();
  SATMVS_REQUIRE(in && w && out && Cin >= 1 && Cout >= 1 && Di >= 1 && Hi >= 1 && Wi >= 1);
  SATMVS_REQUIRE((NZ == 1 || NZ == 3) && mode >= 0 && mode <= 3);
  cudaStream_t st = (cudaStream_t)stream;
  ProfScope prof(kProfTrainConv, st);
  const bool three = NZ == 3;
  if (mode == 3 && three) {   // register-tiled direct kernel when rows are 16-byte aligned
    DirectDeconv3d d{};
    d.in = in; d.w = w; d.w_ci = w_ci; d.w_co = w_co; d.out = out; d.Cin = Cin; d.Cout = Cout; d.Di = Di; d.Hi = Hi; d.Wi = Wi;
    if (direct_deconv3d_supported(d)) return direct_deconv3d_launch(d, st, "satmvs_conv3d_raw (transposed, direct)");
  }
  if (mode == 3) {   // ConvTranspose(k 3, stride 2, padding 1, output_padding 1): one problem per output parity class
    ConvGroup g{};
    int n = 0;
    for (int pz = 0; pz < (three ? 2 : 1); ++pz)
      for (int py = 0; py < 2; ++py)
        for (int px = 0; px < 2; ++px) {
          ConvProblem p;
          conv_problem_defaults(p);
          p.in = in; p.w = w; p.out = out;
          p.Cin = Cin; p.Cout = Cout;
          p.Di = Di; p.Hi = Hi; p.Wi = Wi; p.Do = three ? 2 * Di : Di; p.Ho = 2 * Hi; p.Wo = 2 * Wi;
          p.Qd = Di; p.Qh = Hi; p.Qw = Wi;
          p.w_ci_stride = w_ci; p.w_co_stride = w_co;
          if (three) { p.q2o_mul[0] = 2; p.q2o_add[0] = pz; }
          p.q2o_mul[1] = 2; p.q2o_mul[2] = 2; p.q2o_add[1] = py; p.q2o_add[2] = px;
          conv_taps_deconv_class(p, three, pz, py, px);
          conv_finalize(p);
          g.p[n++] = p;
        }
    g.n = n;
    if (Cout >= 32) return conv_launch<Tile32>(g, st, "satmvs_conv3d_raw (transposed)");
    if (Cout >= 16) return conv_launch<Tile16>(g, st, "satmvs_conv3d_raw (transposed)");
    return conv_launch<Tile8>(g, st, "satmvs_conv3d_raw (transposed)");
  }
  const int stride = mode == 1 ? 2 : 1;
  if (stride == 2) SATMVS_REQUIRE(Hi % 2 == 0 && Wi % 2 == 0 && (!three || Di % 2 == 0));
  const int Do = three ? Di / stride : Di, Ho = Hi / stride, Wo = Wi / stride;
  {
    DirectConv d{};
    d.in = in; d.w = w; d.out = out;
    d.Cin = Cin; d.Cout = Cout; d.Di = Di; d.Hi = Hi; d.Wi = Wi; d.Do = Do; d.Ho = Ho; d.Wo = Wo;
    d.w_co = w_co; d.w_ci = w_ci; d.acc_scale = 1.0f; d.relu = 0; d.flip = mode == 2;
    if (NZ == 3 && stride == 1 && direct_conv3d_c1_supported(d)) return direct_conv3d_c1_launch(d, st, "satmvs_conv3d_raw (direct, 1 channel)");
    if (direct_conv_supported(d, NZ, stride)) return direct_conv_launch(d, NZ, stride, st, "satmvs_conv3d_raw (direct)");
  }
  ConvProblem p;
  conv_problem_defaults(p);
  p.in = in; p.w = w; p.out = out;
  p.Cin = Cin; p.Cout = Cout; p.Di = Di; p.Hi = Hi; p.Wi = Wi; p.Do = Do; p.Ho = Ho; p.Wo = Wo;
  p.Qd = Do; p.Qh = Ho; p.Qw = Wo;
  p.w_co_stride = w_co; p.w_ci_stride = w_ci;
  for (int i = 0; i < 3; ++i) { p.q2i_mul[i] = stride; p.q2i_add[i] = -1; }
  if (!three) { p.q2i_mul[0] = 1; p.q2i_add[0] = 0; }
  conv_taps_dense(p, three);
  if (mode == 2)
    for (int t = 0; t < p.ntaps; ++t) p.tap_w[t] = (signed char)(p.ntaps - 1 - p.tap_w[t]);
  conv_finalize(p);
  ConvGroup g{};
  g.p[0] = p; g.n = 1;
  if (Cout >= 64) return conv_launch<Tile64>(g, st, "satmvs_conv3d_raw");
  if (Cout >= 32) return conv_launch<Tile32>(g, st, "satmvs_conv3d_raw");
  if (Cout >= 16) return conv_launch<Tile16>(g, st, "satmvs_conv3d_raw");
  return conv_launch<Tile8>(g, st, "satmvs_conv3d_raw");
}

size_t satmvs_conv3d_wgrad_workspace_bytes(int Cin, int Cout, int Di, int Hi, int Wi, int NZ, int stride) {
  if (Cin < 1 || Cout < 1 || Di < 1 || Hi < 1 || Wi < 1 || (NZ != 1 && NZ != 3) || (stride != 1 && stride != 2)) return 0;
  const WgradPlan p = wgrad_plan(Cin, Cout, Wi, NZ == 3 ? Di / stride : Di, Hi / stride, Wi / stride, NZ, stride);
  return p.ok ? p.ws_bytes + 256 : 0;
}

int satmvs_conv3d_wgrad(const float* x, int Cin, int Di, int Hi, int Wi, const float* dy, int Cout, int NZ, int stride,
                        float* dw, long long dw_co, long long dw_ci, int accumulate, void* workspace, size_t workspace_bytes,
                        void* stream) {
  SATMVS_CHECK_ASYNC();
  SATMVS_REQUIRE(x && dy && dw && workspace && Cin >= 1 && Cout >= 1 && Di >= 1 && Hi >= 1 && Wi >= 1);
  SATMVS_REQUIRE((NZ == 1 || NZ == 3) && (stride == 1 || stride == 2));
  if (stride == 2) SATMVS_REQUIRE(Hi % 2 == 0 && Wi % 2 == 0 && (NZ == 1 || Di % 2 == 0));
  const int Do = NZ == 3 ? Di / stride : Di, Ho = Hi / stride, Wo = Wi / stride;
  const WgradPlan p = wgrad_plan(Cin, Cout, Wi, Do, Ho, Wo, NZ, stride);
  if (!p.ok) return fail_invalid("satmvs_conv3d_wgrad: a row of the tensors does not fit shared memory (Cout x Wo, Wi)");
  SATMVS_REQUIRE(workspace_bytes >= p.ws_bytes);
  SATMVS_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 15) == 0);
  cudaStream_t st = (cudaStream_t)stream;
  ProfScope prof(kProfTrainWgrad, st);
  WgradArgs a{};
  a.x = x; a.dy = dy; a.partial = static_cast<float*>(workspace);
  a.Cin = Cin; a.Cout = Cout; a.Di = Di; a.Hi = Hi; a.Wi = Wi; a.Do = Do; a.Ho = Ho; a.Wo = Wo; a.NZ = NZ;
  a.ci_tile = p.ci_tile; a.rb = p.rb; a.rsx = p.rsx; a.rsy = p.rsy; a.split = p.split; a.groups_per_split = p.groups_per_split;
  const dim3 grid(p.chunks, p.groups);
  if (stride == 1) {
    static const cudaError_t e1 = cudaFuncSetAttribute(wgrad_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    (void)e1;
    wgrad_kernel<1><<<grid, p.threads, p.smem, st>>>(a);
  } else {
    static const cudaError_t e2 = cudaFuncSetAttribute(wgrad_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    (void)e2;
    wgrad_kernel<2><<<grid, p.threads, p.smem, st>>>(a);
  }
  int rc = check_launch("wgrad_kernel");
  if (rc) return rc;
  const int total = Cout * Cin * NZ * 9;
  wgrad_reduce_kernel<<<ceil_div(total, 256), 256, 0, st>>>(a.partial, p.chunks * p.split, Cout, Cin, NZ * 9, dw, dw_co, dw_ci, accumulate);
  return check_launch("wgrad_reduce_kernel");
}

// Per-plane K x K convolution (K = 1, 3, 5; padding K/2) of a [Cin][N][H][W] tensor in the arrangements of satmvs_conv3d_raw:
// mode 0 stride 1, 1 stride 2, 2 stride 1 with mirrored taps, 3 transposed stride 2 (output_padding 1).  K = 3 is satmvs_conv3d_raw
// with NZ = 1; the others (FeatureNet's 5x5 stride-2 and 1x1 layers and their data gradients) run on the implicit-GEMM engine.
int satmvs_conv2d_raw(const float* in, int Cin, int N, int Hi, int Wi, const float* w, long long w_co, long long w_ci, int K, int mode,
                      float* out, int Cout, void* stream) {
  if (K == 3) return satmvs_conv3d_raw(in, Cin, N, Hi, Wi, w, w_co, w_ci, 1, mode, out, Cout, stream);
  SATMVS_CHECK_ASYNC();
  SATMVS_REQUIRE(in && w && out && Cin >= 1 && Cout >= 1 && N >= 1 && Hi >= 1 && Wi >= 1 && (K == 1 || K == 5) && mode >= 0 && mode <= 3);
  SATMVS_REQUIRE(K * K <= kMaxTaps);
  cudaStream_t st = (cudaStream_t)stream;
  ProfScope prof(kProfTrainConv, st);
  const int pad = K / 2;
  ConvGroup g{};
  int n = 0;
  if (mode == 3) {
    for (int py = 0; py < 2; ++py)
      for (int px = 0; px < 2; ++px) {
        ConvProblem p;
        conv_problem_defaults(p);
        p.in = in; p.w = w; p.out = out; p.Cin = Cin; p.Cout = Cout;
        p.Di = N; p.Hi = Hi; p.Wi = Wi; p.Do = N; p.Ho = 2 * Hi; p.Wo = 2 * Wi;
        p.Qd = N; p.Qh = Hi; p.Qw = Wi;
        p.w_ci_stride = w_ci; p.w_co_stride = w_co;
        p.q2o_mul[1] = 2; p.q2o_mul[2] = 2; p.q2o_add[1] = py; p.q2o_add[2] = px;
        p.q2i_add[1] = -1; p.q2i_add[2] = -1;
        int t = 0;
        for (int ky = 0; ky < K; ++ky) {
          if (((ky - py - pad) & 1) != 0) continue;              // i = 2 o - pad + k: k has the parity of i + pad
          for (int kx = 0; kx < K; ++kx) {
            if (((kx - px - pad) & 1) != 0) continue;
            p.tap_dz[t] = 0; p.tap_dy[t] = (signed char)((py + pad - ky) / 2 + 1); p.tap_dx[t] = (signed char)((px + pad - kx) / 2 + 1);
            p.tap_w[t] = (signed char)(ky * K + kx);
            ++t;
          }
        }
        if (t == 0) continue;
        p.ntaps = t;
        conv_finalize(p);
        g.p[n++] = p;
      }
  } else {
    const int stride = mode == 1 ? 2 : 1;
    if (stride == 2) SATMVS_REQUIRE(Hi % 2 == 0 && Wi % 2 == 0);
    ConvProblem p;
    conv_problem_defaults(p);
    p.in = in; p.w = w; p.out = out; p.Cin = Cin; p.Cout = Cout;
    p.Di = N; p.Hi = Hi; p.Wi = Wi; p.Do = N; p.Ho = Hi / stride; p.Wo = Wi / stride;
    p.Qd = N; p.Qh = p.Ho; p.Qw = p.Wo;
    p.w_co_stride = w_co; p.w_ci_stride = w_ci;
    p.q2i_mul[1] = stride; p.q2i_mul[2] = stride; p.q2i_add[1] = -pad; p.q2i_add[2] = -pad;
    int t = 0;
    for (int ky = 0; ky < K; ++ky)
      for (int kx = 0; kx < K; ++kx) {
        p.tap_dz[t] = 0; p.tap_dy[t] = (signed char)ky; p.tap_dx[t] = (signed char)kx;
        p.tap_w[t] = (signed char)(mode == 2 ? K * K - 1 - (ky * K + kx) : ky * K + kx);
        ++t;
      }
    p.ntaps = t;
    conv_finalize(p);
    g.p[n++] = p;
  }
  g.n = n;
  if (Cout >= 32) return conv_launch<Tile32>(g, st, "satmvs_conv2d_raw");
  if (Cout >= 16) return conv_launch<Tile16>(g, st, "satmvs_conv2d_raw");
  return conv_launch<Tile8>(g, st, "satmvs_conv2d_raw");
}

size_t satmvs_conv2d_wgrad_workspace_bytes(int Cin, int Cout, int N, int Hi, int Wi, int K, int stride) {
  if (K == 3) return satmvs_conv3d_wgrad_workspace_bytes(Cin, Cout, N, Hi, Wi, 1, stride);
  if (Cin < 1 || Cout < 1 || N < 1 || Hi < 1 || Wi < 1 || (K != 1 && K != 5) || (stride != 1 && stride != 2)) return 0;
  return (size_t)256 * Cout * Cin * K * K * 4 + 256;
}

// dw[co * dw_co + ci * dw_ci + ky * K + kx] (+)= sum over planes and positions of dy[co, o] * x[ci, stride * o + k - K/2]
int satmvs_conv2d_wgrad(const float* x, int Cin, int N, int Hi, int Wi, const float* dy, int Cout, int K, int stride,
                        float* dw, long long dw_co, long long dw_ci, int accumulate, void* workspace, size_t workspace_bytes,
                        void* stream) {
  if (K == 3) return satmvs_conv3d_wgrad(x, Cin, N, Hi, Wi, dy, Cout, 1, stride, dw, dw_co, dw_ci, accumulate, workspace,
                                         workspace_bytes, stream);
  SATMVS_CHECK_ASYNC();
  SATMVS_REQUIRE(x && dy && dw && workspace && Cin >= 1 && Cout >= 1 && N >= 1 && Hi >= 1 && Wi >= 1);
  SATMVS_REQUIRE((K == 1 || K == 5) && (stride == 1 || stride == 2));
  if (stride == 2) SATMVS_REQUIRE(Hi % 2 == 0 && Wi % 2 == 0);
  cudaStream_t st = (cudaStream_t)stream;
  ProfScope prof(kProfTrainWgrad, st);
  WgradGenArgs a{};
  a.x = x; a.dy = dy; a.partial = static_cast<float*>(workspace);
  a.Cin = Cin; a.Cout = Cout; a.N = N; a.Hi = Hi; a.Wi = Wi; a.Ho = Hi / stride; a.Wo = Wi / stride; a.K = K; a.S = stride;
  const int total = Cout * Cin * K * K;
  const long long npos = (long long)N * a.Ho * a.Wo;
  int chunks = 256;
  if (chunks > npos) chunks = (int)npos;
  a.chunk_len = (int)((npos + chunks - 1) / chunks);
  chunks = (int)((npos + a.chunk_len - 1) / a.chunk_len);
  SATMVS_REQUIRE(workspace_bytes >= (size_t)chunks * total * 4);
  wgrad_generic_kernel<<<dim3(ceil_div(total, 256), chunks), 256, 0, st>>>(a);
  int rc = check_launch("wgrad_generic_kernel");
  if (rc) return rc;
  wgrad_reduce_kernel<<<ceil_div(total, 256), 256, 0, st>>>(a.partial, chunks, Cout, Cin, K * K, dw, dw_co, dw_ci, accumulate);
  return check_launch("wgrad_reduce_kernel");
}

// z = [relu](gamma (y - mean_batch) / sqrt(var_batch + eps) + beta) [+ post_add]; mean / var (biased) are outputs.
// y, z, post_add: [B][C][n] with n a multiple of 4; acc: 2 C doubles of scratch.
int satmvs_bn_train_fwd(const float* y, int B, int C, long long n, const float* gamma, const float* beta, float eps, int relu,
                        const float* post_add, float* z, float* mean, float* var, double* acc, void* stream) {
  SATMVS_CHECK_ASYNC();
  SATMVS_REQUIRE(y && z && mean && var && acc && B >= 1 && C >= 1 && n >= 1);
  cudaStream_t st = (cudaStream_t)stream;
  ProfScope prof(kProfTrainNorm, st);
  cudaMemsetAsync(acc, 0, (size_t)C * 2 * sizeof(double), st);
  int bx = (int)(((n + 3) / 4 + 255) / 256);
  const int cap = (8 * kNumSMs + C - 1) / C;
  if (bx > cap) bx = cap;
  bn_stats_kernel<<<dim3(bx, C), 256, 0, st>>>(y, B, C, n, acc);
  bn_finalize_kernel<<<ceil_div(C, 128), 128, 0, st>>>(acc, C, (double)B * (double)n, mean, var);
  bn_apply_kernel<<<dim3(bx, C), 256, 0, st>>>(y, B, C, n, gamma, beta, mean, var, eps, relu, post_add, z);
  return check_launch("satmvs_bn_train_fwd");
}

// dy (gradient at the conv output) from dz (+ dz2 when the block output feeds two consumers: the next block and a skip);
// dgamma / dbeta [C] outputs (may be null).
int satmvs_bn_train_bwd(const float* dz, const float* dz2, const float* y, int B, int C, long long n, const float* gamma, const float* beta,
                        const float* mean, const float* var, float eps, int relu, float* dy, float* dgamma, float* dbeta,
                        double* acc, void* stream) {
  SATMVS_CHECK_ASYNC();
  SATMVS_REQUIRE(dz && y && mean && var && dy && acc && B >= 1 && C >= 1 && n >= 1);
  cudaStream_t st = (cudaStream_t)stream;
  ProfScope prof(kProfTrainNorm, st);
  cudaMemsetAsync(acc, 0, (size_t)C * 2 * sizeof(double), st);
  int bx = (int)(((n + 3) / 4 + 255) / 256);
  const int cap = (8 * kNumSMs + C - 1) / C;
  if (bx > cap) bx = cap;
  bn_bwd_reduce_kernel<<<dim3(bx, C), 256, 0, st>>>(dz, dz2, y, B, C, n, gamma, beta, mean, var, eps, relu, acc);
  bn_bwd_apply_kernel<<<dim3(bx, C), 256, 0, st>>>(dz, dz2, y, B, C, n, gamma, beta, mean, var, eps, relu, acc, (double)B * (double)n, dy,
                                                   dgamma, dbeta);
  return check_launch("satmvs_bn_train_bwd");
}

}  // extern "C"
