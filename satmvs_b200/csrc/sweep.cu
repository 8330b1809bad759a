// sweep.cu — the fused plane-sweep kernels: per-hypothesis camera geometry (fp64) + bilinear
// source-feature gather + variance reduction over views, one pass, one write of the cost volume.
//
// Replaces, for one batch element, the ~140-launch ATen chain of rpc_warping
// (modules/warping.py:310-365) x (V-1) views plus the ~10 elementwise passes of
// networks/casred.py:26-53 (== casmvs.py:30-59).  The same kernel template serves the single-view
// operators rpc_warping / homo_warping (kVariance = false) and the pin-hole branch (Geo = HomoSweep).
//
// Work decomposition (v1): one thread = one reference pixel x DK consecutive depth planes.
//   phase 1  geometry: the reference-view localisation is evaluated once per (pixel, plane) and
//            shared by all source views; every source projection yields a 5-register tap record.
//   phase 2  channel loop: ref feature read once per channel and reused for the DK planes; 4 taps
//            per (view, plane, channel) through L1 (__ldg); S and Q accumulated in registers;
//            one coalesced streaming store per (channel, plane).
// Lanes map to consecutive pixels (row-major over H*W), so ref loads and volume stores are 128-byte
// coalesced and the gathers of a warp fall into 1-2 cache lines.
#include "geometry.cuh"
#include "prof.cuh"

#ifndef SATMVS_DK2
#define SATMVS_DK2 8     // planes per thread with <= 2 source views (tuning knob, see profiles/)
#endif
#ifndef SATMVS_MIN_BLOCKS
#define SATMVS_MIN_BLOCKS 1
#endif

namespace satmvs {

constexpr int kSweepThreads = 128;

template <class Geo>
struct SweepArgs {
  const float* ref_fea;                       // [C,H,W] (variance mode) or nullptr
  const float* src_fea[Geo::kNumSrc];         // each [C,H,W]
  const float* depth;                         // [D] or [D,H,W]
  float* out[SATMVS_MAX_PEERS];               // each [C,out_D,H,W]; every buffer receives the same planes
  int n_out;                                  // 1, or the number of peer GPUs written over NVLink (fused all-gather)
  int out_D, out_d0;                          // planes of the output tensor, first plane written by this launch
  int C, D, H, W;                             // D = planes swept by this launch
  int depth_per_pixel;
  int n_src;                                  // live source views (<= Geo::kNumSrc; the rest carry zero weights)
  float half_w, half_h;                       // W/2, H/2 (ATen un-normalise)
  float num_views, inv_num_views;             // V as fp32 (div_(num_views), casred.py:53) and RN(1/V)
  Geo geo;
};

template <class Geo, int DK, bool kVariance>
__global__ void __launch_bounds__(kSweepThreads, SATMVS_MIN_BLOCKS)
sweep_fwd_kernel(const __grid_constant__ SweepArgs<Geo> a) {
  constexpr int NSRC = Geo::kNumSrc;
  const int HW = a.H * a.W;
  const int pix = blockIdx.x * kSweepThreads + threadIdx.x;
  const bool active = pix < HW;
  const int pixc = active ? pix : HW - 1;
  const int y = pixc / a.W, x = pixc - y * a.W;
  const int d0 = blockIdx.y * DK;

  Tap taps[DK][NSRC];
  {
    const typename Geo::Pixel px = a.geo.pixel(x, y);
#pragma unroll
    for (int k = 0; k < DK; ++k) {
      const int d = min(d0 + k, a.D - 1);
      const float h = a.depth_per_pixel ? __ldg(a.depth + (size_t)d * HW + pixc) : __ldg(a.depth + d);
      const typename Geo::Plane pl = a.geo.plane(px, h);
#pragma unroll
      for (int v = 0; v < NSRC; ++v) {
        float gx, gy;
        a.geo.project(v, px, pl, gx, gy);
        taps[k][v] = make_tap(gx, gy, a.H, a.W, a.half_w, a.half_h);
        if (v >= a.n_src) { taps[k][v].w00 = taps[k][v].w01 = taps[k][v].w10 = taps[k][v].w11 = 0.0f; taps[k][v].off = 0; }
      }
    }
  }

  const size_t plane_stride = (size_t)HW;
  for (int c = 0; c < a.C; ++c) {
    float r = 0.0f;
    if (kVariance) r = __ldg(a.ref_fea + (size_t)c * HW + pixc);
    const float r2 = __fmul_rn(r, r);
    const size_t oidx = ((size_t)c * a.out_D + a.out_d0 + d0) * plane_stride + pix;
#pragma unroll
    for (int k = 0; k < DK; ++k) {
      if (d0 + k < a.D) {
        float s = r, q = r2, val = 0.0f;
#pragma unroll
        for (int v = 0; v < NSRC; ++v) {
          const float* f0 = a.src_fea[v] + (size_t)c * HW;
          val = tap_fetch(f0, f0 + a.W, taps[k][v]);
          if (kVariance) {
            // volume_sum + warped ; volume_sq_sum + warped**2  (casred.py:47-48): separate roundings
            s = __fadd_rn(s, val);
            q = __fadd_rn(q, __fmul_rn(val, val));
          }
        }
        float res = val;
        if (kVariance) {
          // volume_sq_sum.div_(V).sub_(volume_sum.div_(V).pow_(2))  (casred.py:53)
          const float m = div_const(s, a.num_views, a.inv_num_views);
          res = __fsub_rn(div_const(q, a.num_views, a.inv_num_views), __fmul_rn(m, m));
        }
        if (active) {
          // one store per destination: the local volume, or every peer's volume (NVLink posted writes)
          for (int o = 0; o < a.n_out; ++o) __stcs(a.out[o] + oidx + (size_t)k * plane_stride, res);
        }
      }
    }
  }
}

// ------------------------------------------------------------------------------------------
// backward: gradients reach the feature maps only (the grid is built under no_grad,
// warping.py:322-356).  Same geometry, scatter with float atomics (RED.ADD.F32 in L2).
//   warp  : grad_src[c, tap] += w_tap * g[c,d,pix]
//   var   : var = Q/V - (S/V)^2  =>  d var / d x_v = (2/V) * (x_v - S/V) for every view's sample x_v
//           (x_0 = the reference feature itself); S is recomputed from the taps.
// ------------------------------------------------------------------------------------------
template <class Geo>
struct SweepBwdArgs {
  const float* grad_out;                      // [C,D,H,W]
  const float* ref_fea;                       // variance mode
  const float* src_fea[Geo::kNumSrc];         // variance mode
  const float* depth;
  float* grad_ref;                            // [C,H,W] variance mode
  float* grad_src[Geo::kNumSrc];              // [C,H,W]
  int C, D, H, W;
  int depth_per_pixel;
  int n_src;
  float half_w, half_h;
  float num_views, inv_num_views;
  Geo geo;
};

__device__ __forceinline__ void tap_scatter(float* __restrict__ g, const Tap& t, int W, float v) {
  float* p = g + t.off;
  if (t.w00 != 0.0f) atomicAdd(p, t.w00 * v);
  if (t.w01 != 0.0f) atomicAdd(p + 1, t.w01 * v);
  if (t.w10 != 0.0f) atomicAdd(p + W, t.w10 * v);
  if (t.w11 != 0.0f) atomicAdd(p + W + 1, t.w11 * v);
}

template <class Geo, int DK, bool kVariance>
__global__ void __launch_bounds__(kSweepThreads)
sweep_bwd_kernel(const __grid_constant__ SweepBwdArgs<Geo> a) {
  constexpr int NSRC = Geo::kNumSrc;
  const int HW = a.H * a.W;
  const int pix = blockIdx.x * kSweepThreads + threadIdx.x;
  if (pix >= HW) return;
  const int y = pix / a.W, x = pix - y * a.W;
  const int d0 = blockIdx.y * DK;

  Tap taps[DK][NSRC];
  {
    const typename Geo::Pixel px = a.geo.pixel(x, y);
#pragma unroll
    for (int k = 0; k < DK; ++k) {
      const int d = min(d0 + k, a.D - 1);
      const float h = a.depth_per_pixel ? __ldg(a.depth + (size_t)d * HW + pix) : __ldg(a.depth + d);
      const typename Geo::Plane pl = a.geo.plane(px, h);
#pragma unroll
      for (int v = 0; v < NSRC; ++v) {
        float gx, gy;
        a.geo.project(v, px, pl, gx, gy);
        taps[k][v] = make_tap(gx, gy, a.H, a.W, a.half_w, a.half_h);
        if (v >= a.n_src) { taps[k][v].w00 = taps[k][v].w01 = taps[k][v].w10 = taps[k][v].w11 = 0.0f; taps[k][v].off = 0; }
      }
    }
  }

  const float two_over_v = 2.0f / a.num_views;
  for (int c = 0; c < a.C; ++c) {
    const float* gc = a.grad_out + ((size_t)c * a.D + d0) * HW + pix;
    float r = 0.0f, gref = 0.0f;
    if (kVariance) r = __ldg(a.ref_fea + (size_t)c * HW + pix);
#pragma unroll
    for (int k = 0; k < DK; ++k) {
      if (d0 + k < a.D) {
        const float g = __ldg(gc + (size_t)k * HW);
        if (kVariance) {
          float vals[NSRC];
          float s = r;
#pragma unroll
          for (int v = 0; v < NSRC; ++v) {
            const float* f0 = a.src_fea[v] + (size_t)c * HW;
            vals[v] = tap_fetch(f0, f0 + a.W, taps[k][v]);
            s += vals[v];
          }
          const float mean = s / a.num_views;
          gref += g * two_over_v * (r - mean);
#pragma unroll
          for (int v = 0; v < NSRC; ++v)
            if (v < a.n_src) tap_scatter(a.grad_src[v] + (size_t)c * HW, taps[k][v], a.W, g * two_over_v * (vals[v] - mean));
        } else {
          tap_scatter(a.grad_src[0] + (size_t)c * HW, taps[k][0], a.W, g);
        }
      }
    }
    if (kVariance) atomicAdd(a.grad_ref + (size_t)c * HW + pix, gref);
  }
}

// ------------------------------------------------------------------------------------------
// point-list RPC ops (fp64 in, fp64 out): RPC_Photo2Obj / RPC_Obj2Photo and their
// tools/rpc_tensor.py twins.  Plain 20-term evaluation, one point per thread.
// ------------------------------------------------------------------------------------------
__global__ void rpc_localise_kernel(const __grid_constant__ RpcRefPack r, const double* __restrict__ samp,
                                    const double* __restrict__ line, const double* __restrict__ hei,
                                    int64_t n, double* __restrict__ lat, double* __restrict__ lon) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double s = (samp[i] - r.samp_off) * r.samp_iscale;
  double l = (line[i] - r.line_off) * r.line_iscale;
  double h = (hei[i] - r.hei_off) * r.hei_iscale;
  double la = poly20(r.lat_num, l, s, h) / poly20(r.lat_den, l, s, h);
  double lo = poly20(r.lon_num, l, s, h) / poly20(r.lon_den, l, s, h);
  lat[i] = fma(la, r.lat_scale, r.lat_off);
  lon[i] = fma(lo, r.lon_scale, r.lon_off);
}

__global__ void rpc_project_kernel(const __grid_constant__ RpcSrcPack r, const double* __restrict__ lat,
                                   const double* __restrict__ lon, const double* __restrict__ hei,
                                   int64_t n, double* __restrict__ samp, double* __restrict__ line) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double P = (lat[i] - r.lat_off) * r.lat_iscale;
  double L = (lon[i] - r.lon_off) * r.lon_iscale;
  double H = fma(hei[i], r.h_a, r.h_b);
  double sn = poly20(r.samp_num, L, P, H) / poly20(r.samp_den, L, P, H);
  double ln = poly20(r.line_num, L, P, H) / poly20(r.line_den, L, P, H);
  samp[i] = fma(sn, r.samp_scale, r.samp_off);
  line[i] = fma(ln, r.line_scale, r.line_off);
}

// ------------------------------------------------------------------------------------------
// host-side dispatch
// ------------------------------------------------------------------------------------------
template <int NSRC> struct PlanesPerThread { static constexpr int value = NSRC <= 2 ? SATMVS_DK2 : (NSRC <= 4 ? 4 : 2); };

static int check_dims(int n_src, int C, int D, int H, int W) {
  SATMVS_REQUIRE(n_src >= 1 && n_src <= SATMVS_MAX_SRC_VIEWS);
  SATMVS_REQUIRE(C >= 1 && D >= 1 && H >= 2 && W >= 2);
  SATMVS_REQUIRE((int64_t)H * W < (1LL << 30));
  return SATMVS_OK;
}

template <class Geo, bool kVariance>
static int launch_fwd(SweepArgs<Geo>& a, cudaStream_t st) {
  constexpr int DK = PlanesPerThread<Geo::kNumSrc>::value;
  dim3 grid(ceil_div((int64_t)a.H * a.W, kSweepThreads), ceil_div(a.D, DK));
  ProfScope prof(kProfSweep, st);
  sweep_fwd_kernel<Geo, DK, kVariance><<<grid, kSweepThreads, 0, st>>>(a);
  return check_launch("sweep_fwd_kernel");
}

template <class Geo, bool kVariance>
static int launch_bwd(SweepBwdArgs<Geo>& a, cudaStream_t st) {
  constexpr int DK = PlanesPerThread<Geo::kNumSrc>::value;
  dim3 grid(ceil_div((int64_t)a.H * a.W, kSweepThreads), ceil_div(a.D, DK));
  sweep_bwd_kernel<Geo, DK, kVariance><<<grid, kSweepThreads, 0, st>>>(a);
  return check_launch("sweep_bwd_kernel");
}

template <int NSRC>
static void fill_rpc_geo(RpcSweep<NSRC>& g, int n_src, const double* ref_rpc, const double* src_rpcs, int H, int W) {
  g.ref = make_rpc_ref_pack(ref_rpc);
  for (int v = 0; v < NSRC; ++v) g.src[v] = make_rpc_src_pack(src_rpcs + (size_t)(v < n_src ? v : 0) * SATMVS_RPC_LEN, ref_rpc);
  g.half_wm1 = (float)((W - 1) / 2.0);
  g.half_hm1 = (float)((H - 1) / 2.0);
  g.inv_half_wm1 = 1.0f / g.half_wm1;
  g.inv_half_hm1 = 1.0f / g.half_hm1;
}

template <int NSRC>
static int fill_homo_geo(HomoSweep<NSRC>& g, int n_src, const double* ref_proj, const double* src_projs, int H, int W) {
  for (int v = 0; v < NSRC; ++v)
    if (!make_homo_src_pack(g.src[v], src_projs + (size_t)(v < n_src ? v : 0) * 16, ref_proj)) {
      set_error("ref_proj is singular");
      return SATMVS_EINVAL;
    }
  g.inv_half_wm1 = 1.0 / ((W - 1) / 2.0);
  g.inv_half_hm1 = 1.0 / ((H - 1) / 2.0);
  return SATMVS_OK;
}

template <class Args>
static void fill_common(Args& a, const float* depth, int depth_per_pixel, int n_src, int C, int D, int H, int W) {
  a.depth = depth; a.depth_per_pixel = depth_per_pixel; a.n_src = n_src;
  a.C = C; a.D = D; a.H = H; a.W = W;
  a.half_w = (float)(W / 2.0); a.half_h = (float)(H / 2.0);
  a.num_views = (float)(n_src + 1);
  a.inv_num_views = 1.0f / a.num_views;
}

// template slot count for a runtime number of source views
static int slot_for(int n_src) { return n_src <= 4 ? n_src : (n_src <= 6 ? 6 : 8); }

#define SATMVS_DISPATCH_NSRC(n_src, ...)                      \
  switch (slot_for(n_src)) {                                   \
    case 1: { constexpr int NSRC = 1; __VA_ARGS__ } break;            \
    case 2: { constexpr int NSRC = 2; __VA_ARGS__ } break;            \
    case 3: { constexpr int NSRC = 3; __VA_ARGS__ } break;            \
    case 4: { constexpr int NSRC = 4; __VA_ARGS__ } break;            \
    case 6: { constexpr int NSRC = 6; __VA_ARGS__ } break;            \
    default: { constexpr int NSRC = 8; __VA_ARGS__ } break;           \
  }

}  // namespace satmvs

using namespace satmvs;

extern "C" {

int satmvs_cost_volume_rpc_fwd(const float* ref_fea, const float* const* src_feas, int n_src,
                               const double* ref_rpc, const double* src_rpcs,
                               const float* depth, int depth_per_pixel,
                               int C, int D, int H, int W, float* out_var, void* stream) {
  float* outs[1] = {out_var};
  return satmvs_cost_volume_rpc_fwd_sharded(ref_fea, src_feas, n_src, ref_rpc, src_rpcs, depth, depth_per_pixel,
                                            C, D, H, W, 0, D, outs, 1, stream);
}

int satmvs_cost_volume_rpc_fwd_sharded(const float* ref_fea, const float* const* src_feas, int n_src,
                                       const double* ref_rpc, const double* src_rpcs,
                                       const float* depth, int depth_per_pixel,
                                       int C, int D, int H, int W, int d0, int D_total,
                                       float* const* outs, int n_outs, void* stream) {
  if (int e = check_dims(n_src, C, D, H, W)) return e;
  SATMVS_REQUIRE(ref_fea && src_feas && ref_rpc && src_rpcs && depth && outs);
  SATMVS_REQUIRE(n_outs >= 1 && n_outs <= SATMVS_MAX_PEERS && d0 >= 0 && d0 + D <= D_total);
  for (int o = 0; o < n_outs; ++o) SATMVS_REQUIRE(outs[o] != nullptr);
  SATMVS_DISPATCH_NSRC(n_src, {
    SweepArgs<RpcSweep<NSRC>> a{};
    fill_common(a, depth, depth_per_pixel, n_src, C, D, H, W);
    a.ref_fea = ref_fea;
    for (int o = 0; o < n_outs; ++o) a.out[o] = outs[o];
    a.n_out = n_outs; a.out_D = D_total; a.out_d0 = d0;
    for (int v = 0; v < NSRC; ++v) a.src_fea[v] = src_feas[v < n_src ? v : 0];
    fill_rpc_geo(a.geo, n_src, ref_rpc, src_rpcs, H, W);
    return launch_fwd<RpcSweep<NSRC>, true>(a, (cudaStream_t)stream);
  })
  return SATMVS_OK;
}

int satmvs_cost_volume_homo_fwd(const float* ref_fea, const float* const* src_feas, int n_src,
                                const double* ref_proj, const double* src_projs,
                                const float* depth, int depth_per_pixel,
                                int C, int D, int H, int W, float* out_var, void* stream) {
  float* outs[1] = {out_var};
  return satmvs_cost_volume_homo_fwd_sharded(ref_fea, src_feas, n_src, ref_proj, src_projs, depth, depth_per_pixel,
                                             C, D, H, W, 0, D, outs, 1, stream);
}

int satmvs_cost_volume_homo_fwd_sharded(const float* ref_fea, const float* const* src_feas, int n_src,
                                        const double* ref_proj, const double* src_projs,
                                        const float* depth, int depth_per_pixel,
                                        int C, int D, int H, int W, int d0, int D_total,
                                        float* const* outs, int n_outs, void* stream) {
  if (int e = check_dims(n_src, C, D, H, W)) return e;
  SATMVS_REQUIRE(ref_fea && src_feas && ref_proj && src_projs && depth && outs);
  SATMVS_REQUIRE(n_outs >= 1 && n_outs <= SATMVS_MAX_PEERS && d0 >= 0 && d0 + D <= D_total);
  for (int o = 0; o < n_outs; ++o) SATMVS_REQUIRE(outs[o] != nullptr);
  SATMVS_DISPATCH_NSRC(n_src, {
    SweepArgs<HomoSweep<NSRC>> a{};
    fill_common(a, depth, depth_per_pixel, n_src, C, D, H, W);
    a.ref_fea = ref_fea;
    for (int o = 0; o < n_outs; ++o) a.out[o] = outs[o];
    a.n_out = n_outs; a.out_D = D_total; a.out_d0 = d0;
    for (int v = 0; v < NSRC; ++v) a.src_fea[v] = src_feas[v < n_src ? v : 0];
    if (int e = fill_homo_geo(a.geo, n_src, ref_proj, src_projs, H, W)) return e;
    return launch_fwd<HomoSweep<NSRC>, true>(a, (cudaStream_t)stream);
  })
  return SATMVS_OK;
}

int satmvs_rpc_warp_fwd(const float* src_fea, const double* src_rpc, const double* ref_rpc,
                        const float* depth, int depth_per_pixel,
                        int C, int D, int H, int W, float* out, void* stream) {
  if (int e = check_dims(1, C, D, H, W)) return e;
  SATMVS_REQUIRE(src_fea && src_rpc && ref_rpc && depth && out);
  SweepArgs<RpcSweep<1>> a{};
  fill_common(a, depth, depth_per_pixel, 1, C, D, H, W);
  a.ref_fea = nullptr; a.out[0] = out; a.n_out = 1; a.out_D = D; a.out_d0 = 0; a.src_fea[0] = src_fea;
  fill_rpc_geo(a.geo, 1, ref_rpc, src_rpc, H, W);
  return launch_fwd<RpcSweep<1>, false>(a, (cudaStream_t)stream);
}

int satmvs_homo_warp_fwd(const float* src_fea, const double* src_proj, const double* ref_proj,
                         const float* depth, int depth_per_pixel,
                         int C, int D, int H, int W, float* out, void* stream) {
  if (int e = check_dims(1, C, D, H, W)) return e;
  SATMVS_REQUIRE(src_fea && src_proj && ref_proj && depth && out);
  SweepArgs<HomoSweep<1>> a{};
  fill_common(a, depth, depth_per_pixel, 1, C, D, H, W);
  a.ref_fea = nullptr; a.out[0] = out; a.n_out = 1; a.out_D = D; a.out_d0 = 0; a.src_fea[0] = src_fea;
  if (int e = fill_homo_geo(a.geo, 1, ref_proj, src_proj, H, W)) return e;
  return launch_fwd<HomoSweep<1>, false>(a, (cudaStream_t)stream);
}

int satmvs_rpc_warp_bwd(const float* grad_out, const double* src_rpc, const double* ref_rpc,
                        const float* depth, int depth_per_pixel,
                        int C, int D, int H, int W, float* grad_src, void* stream) {
  if (int e = check_dims(1, C, D, H, W)) return e;
  SATMVS_REQUIRE(grad_out && src_rpc && ref_rpc && depth && grad_src);
  SweepBwdArgs<RpcSweep<1>> a{};
  fill_common(a, depth, depth_per_pixel, 1, C, D, H, W);
  a.grad_out = grad_out; a.grad_src[0] = grad_src;
  fill_rpc_geo(a.geo, 1, ref_rpc, src_rpc, H, W);
  return launch_bwd<RpcSweep<1>, false>(a, (cudaStream_t)stream);
}

int satmvs_homo_warp_bwd(const float* grad_out, const double* src_proj, const double* ref_proj,
                         const float* depth, int depth_per_pixel,
                         int C, int D, int H, int W, float* grad_src, void* stream) {
  if (int e = check_dims(1, C, D, H, W)) return e;
  SATMVS_REQUIRE(grad_out && src_proj && ref_proj && depth && grad_src);
  SweepBwdArgs<HomoSweep<1>> a{};
  fill_common(a, depth, depth_per_pixel, 1, C, D, H, W);
  a.grad_out = grad_out; a.grad_src[0] = grad_src;
  if (int e = fill_homo_geo(a.geo, 1, ref_proj, src_proj, H, W)) return e;
  return launch_bwd<HomoSweep<1>, false>(a, (cudaStream_t)stream);
}

int satmvs_cost_volume_rpc_bwd(const float* grad_var, const float* ref_fea, const float* const* src_feas,
                               int n_src, const double* ref_rpc, const double* src_rpcs,
                               const float* depth, int depth_per_pixel, int C, int D, int H, int W,
                               float* grad_ref, float* const* grad_srcs, void* stream) {
  if (int e = check_dims(n_src, C, D, H, W)) return e;
  SATMVS_REQUIRE(grad_var && ref_fea && src_feas && ref_rpc && src_rpcs && depth && grad_ref && grad_srcs);
  SATMVS_DISPATCH_NSRC(n_src, {
    SweepBwdArgs<RpcSweep<NSRC>> a{};
    fill_common(a, depth, depth_per_pixel, n_src, C, D, H, W);
    a.grad_out = grad_var; a.ref_fea = ref_fea; a.grad_ref = grad_ref;
    for (int v = 0; v < NSRC; ++v) { a.src_fea[v] = src_feas[v < n_src ? v : 0]; a.grad_src[v] = grad_srcs[v < n_src ? v : 0]; }
    fill_rpc_geo(a.geo, n_src, ref_rpc, src_rpcs, H, W);
    return launch_bwd<RpcSweep<NSRC>, true>(a, (cudaStream_t)stream);
  })
  return SATMVS_OK;
}

int satmvs_cost_volume_homo_bwd(const float* grad_var, const float* ref_fea, const float* const* src_feas,
                                int n_src, const double* ref_proj, const double* src_projs,
                                const float* depth, int depth_per_pixel, int C, int D, int H, int W,
                                float* grad_ref, float* const* grad_srcs, void* stream) {
  if (int e = check_dims(n_src, C, D, H, W)) return e;
  SATMVS_REQUIRE(grad_var && ref_fea && src_feas && ref_proj && src_projs && depth && grad_ref && grad_srcs);
  SATMVS_DISPATCH_NSRC(n_src, {
    SweepBwdArgs<HomoSweep<NSRC>> a{};
    fill_common(a, depth, depth_per_pixel, n_src, C, D, H, W);
    a.grad_out = grad_var; a.ref_fea = ref_fea; a.grad_ref = grad_ref;
    for (int v = 0; v < NSRC; ++v) { a.src_fea[v] = src_feas[v < n_src ? v : 0]; a.grad_src[v] = grad_srcs[v < n_src ? v : 0]; }
    if (int e = fill_homo_geo(a.geo, n_src, ref_proj, src_projs, H, W)) return e;
    return launch_bwd<HomoSweep<NSRC>, true>(a, (cudaStream_t)stream);
  })
  return SATMVS_OK;
}

int satmvs_rpc_localise(const double* rpc, const double* samp, const double* line, const double* hei,
                        int64_t n, double* lat, double* lon, void* stream) {
  SATMVS_REQUIRE(rpc && n >= 0);
  if (n == 0) return SATMVS_OK;
  SATMVS_REQUIRE(samp && line && hei && lat && lon);
  RpcRefPack r = make_rpc_ref_pack(rpc);
  rpc_localise_kernel<<<ceil_div(n, 256), 256, 0, (cudaStream_t)stream>>>(r, samp, line, hei, n, lat, lon);
  return check_launch("rpc_localise_kernel");
}

int satmvs_rpc_project(const double* rpc, const double* lat, const double* lon, const double* hei,
                       int64_t n, double* samp, double* line, void* stream) {
  SATMVS_REQUIRE(rpc && n >= 0);
  if (n == 0) return SATMVS_OK;
  SATMVS_REQUIRE(samp && line && hei && lat && lon);
  RpcSrcPack r = make_rpc_src_pack(rpc, rpc);
  rpc_project_kernel<<<ceil_div(n, 256), 256, 0, (cudaStream_t)stream>>>(r, lat, lon, hei, n, samp, line);
  return check_launch("rpc_project_kernel");
}

}  // extern "C"
