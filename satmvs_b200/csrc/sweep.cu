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
#include "packed.cuh"
#include "prof.cuh"
#include <cstdlib>
#include <type_traits>

#ifndef SATMVS_DK2
#define SATMVS_DK2 8     // planes per thread with <= 2 source views (tuning knob, see profiles/)
#endif
#ifndef SATMVS_MIN_BLOCKS
#define SATMVS_MIN_BLOCKS 1
#endif
#ifndef SATMVS_NP
#define SATMVS_NP 4       // hypothesis planes evaluated in lock step in the v3 geometry phase
#endif
#ifndef SATMVS_CH
#define SATMVS_CH 16      // channels carried per gather pass in the v3 kernel
#endif

namespace satmvs {

constexpr int kSweepThreads = 128;

// store through an NVLink-switch multicast mapping (symmetric memory): the switch replicates the write to every GPU
__device__ __forceinline__ void mc_store(float* p, float v) {
  asm volatile("multimem.st.relaxed.sys.global.f32 [%0], %1;" :: "l"(p), "f"(v) : "memory");
}

template <class Geo>
struct SweepArgs {
  const float* ref_fea;                       // [C,H,W] (variance mode) or nullptr
  const float* src_fea[Geo::kNumSrc];         // each [C,H,W]
  const float4* src_v4[Geo::kNumSrc];         // each [C/4,H,W] of float4 (4 consecutive channels per pixel), or null
  const float* depth;                         // [D] or [D,H,W]
  float* out[SATMVS_MAX_PEERS];               // each [C,out_D,H,W]; every buffer receives the same planes
  int n_out;                                  // 1, or the number of peer GPUs written over NVLink (fused all-gather)
  int multicast;                              // out[0] is an NVLS multicast address: ONE multimem.st per value reaches every GPU
  int out_D, out_d0;                          // planes of the output tensor, first plane written by this launch
  int C, D, H, W;                             // D = planes swept by this launch
  int depth_per_pixel;
  int n_src;                                  // live source views (<= Geo::kNumSrc; the rest carry zero weights)
  float half_w, half_h;                       // W/2, H/2 (ATen un-normalise)
  float num_views, inv_num_views;             // V as fp32 (div_(num_views), casred.py:53) and RN(1/V)
  float neg_zero;                             // -0.0f, opaque to the compiler: fma(x, x, neg_zero) == RN(x*x) as one instruction
  int packed_now;                             // src_v4 was written by the launch directly in front of this one
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
          if (a.multicast) mc_store(a.out[0] + oidx + (size_t)k * plane_stride, res);
          else for (int o = 0; o < a.n_out; ++o) __stcs(a.out[o] + oidx + (size_t)k * plane_stride, res);
        }
      }
    }
  }
}


// ------------------------------------------------------------------------------------------
// v2 of the forward sweep: source features re-packed once per call to [C/4][H][W] float4 (four
// consecutive channels of a pixel in one 16-byte word, pack_vec4_kernel).  A tap is then ONE
// 128-bit load per 4 channels, lanes of a warp read consecutive pixels = consecutive 16-byte words
// (full 128-byte lines, no partial sectors), and the channel loop runs over C/4 quads:
// 4x fewer load instructions, 4 channels of S / Q / variance per iteration.  Geometry phase,
// tap records, op order and therefore results are identical to sweep_fwd_kernel.
// ------------------------------------------------------------------------------------------
struct PackArgs { const float* in[SATMVS_MAX_SRC_VIEWS]; float4* out[SATMVS_MAX_SRC_VIEWS]; int HW; };

__global__ void pack_vec4_kernel(const __grid_constant__ PackArgs a) {
  asm volatile("griddepcontrol.launch_dependents;");     // the sweep's geometry phase may start while this runs
  const int i = blockIdx.x * blockDim.x + threadIdx.x;   // pixel
  const int q = blockIdx.y;                               // channel quad
  const int HW = a.HW;
  if (i >= HW) return;
  const float* p = a.in[blockIdx.z] + (size_t)(4 * q) * HW + i;   // blockIdx.z = source view
  a.out[blockIdx.z][(size_t)q * HW + i] = make_float4(__ldg(p), __ldg(p + HW), __ldg(p + 2 * (size_t)HW), __ldg(p + 3 * (size_t)HW));
}

__device__ __forceinline__ float4 tap_fetch4(const float4* __restrict__ f0, const float4* __restrict__ f1, const Tap& t) {
  const float4 a = __ldg(f0 + t.off), b = __ldg(f0 + t.off + 1), c = __ldg(f1 + t.off), d = __ldg(f1 + t.off + 1);
  float4 r;
  r.x = __fmaf_rn(d.x, t.w11, __fmaf_rn(c.x, t.w10, __fmaf_rn(b.x, t.w01, __fmul_rn(a.x, t.w00))));
  r.y = __fmaf_rn(d.y, t.w11, __fmaf_rn(c.y, t.w10, __fmaf_rn(b.y, t.w01, __fmul_rn(a.y, t.w00))));
  r.z = __fmaf_rn(d.z, t.w11, __fmaf_rn(c.z, t.w10, __fmaf_rn(b.z, t.w01, __fmul_rn(a.z, t.w00))));
  r.w = __fmaf_rn(d.w, t.w11, __fmaf_rn(c.w, t.w10, __fmaf_rn(b.w, t.w01, __fmul_rn(a.w, t.w00))));
  return r;
}

// ------------------------------------------------------------------------------------------
// v3 of the forward sweep (the default when a workspace is given and C % 4 == 0).
//   phase A  geometry for the CTA's 128 pixels x DK planes, hypothesis planes evaluated in lock step
//            (poly20_many: one constant-bank fetch per coefficient per 4 planes); tap records go to
//            shared memory, so the fp64 state is dead before the gather starts.
//   phase B  gather: for every (plane, view) the record is read back once and its weights duplicated
//            into 2-lane operands; the channel loop issues one LDG.128 per tap per 4 channels from the
//            [C/4][H][W] float4 re-pack and does ALL fp32 arithmetic with Blackwell's packed
//            fma/mul/add.f32x2 (two IEEE-rounded results per instruction, so the op sequence and the
//            bits are those of the scalar kernels).  CH channels (16) are carried per pass to keep
//            ~20 warps resident.
// ------------------------------------------------------------------------------------------
#ifndef SATMVS_PACKED_MASK
#define SATMVS_PACKED_MASK 15   // debug knob: bit0 fma2, bit1 mul2, bit2 add2, bit3 sub2 use the packed instruction
#endif
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) {
  if (SATMVS_PACKED_MASK & 1) { u64 d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
  float a0, a1, b0, b1, c0, c1; upk(a, a0, a1); upk(b, b0, b1); upk(c, c0, c1);
  return pk(__fmaf_rn(a0, b0, c0), __fmaf_rn(a1, b1, c1));
}
__device__ __forceinline__ u64 mul2(u64 a, u64 b) {
  if (SATMVS_PACKED_MASK & 2) { u64 d; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
  float a0, a1, b0, b1; upk(a, a0, a1); upk(b, b0, b1);
  return pk(__fmul_rn(a0, b0), __fmul_rn(a1, b1));
}
__device__ __forceinline__ u64 add2(u64 a, u64 b) {
  if (SATMVS_PACKED_MASK & 4) { u64 d; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
  float a0, a1, b0, b1; upk(a, a0, a1); upk(b, b0, b1);
  return pk(__fadd_rn(a0, b0), __fadd_rn(a1, b1));
}
__device__ __forceinline__ u64 sub2(u64 a, u64 b) {
  if (SATMVS_PACKED_MASK & 8) { u64 d; asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
  float a0, a1, b0, b1; upk(a, a0, a1); upk(b, b0, b1);
  return pk(__fsub_rn(a0, b0), __fsub_rn(a1, b1));
}
// A product that feeds an add/sub and must be rounded on its own.  ptxas contracts mul.rn.f32x2 + add/sub.rn.f32x2
// into one FFMA2 even with explicit .rn (and folds fma(a,b,-0) back to a multiply first), which shows up as
// 1-ulp differences against the scalar kernels (profiles/r01_sweep_v3_notes.md).  Two scalar FMULs cannot be
// merged into a packed add, so these few products stay scalar.
__device__ __forceinline__ u64 mul2_rounded(u64 a, u64 b) {
  float a0, a1, b0, b1; upk(a, a0, a1); upk(b, b0, b1);
  return pk(__fmul_rn(a0, b0), __fmul_rn(a1, b1));
}

struct __align__(16) TapW { float w00, w01, w10, w11; };

template <class Geo, int DK, int CH, bool kVariance, bool kMultiOut>
__global__ void __launch_bounds__(kSweepThreads, SATMVS_MIN_BLOCKS)
sweep_fwd_v3_kernel(const __grid_constant__ SweepArgs<Geo> a) {
  constexpr int NSRC = Geo::kNumSrc;
  constexpr int NP = DK < SATMVS_NP ? DK : SATMVS_NP;   // planes evaluated in lock step
  __shared__ int rec_off[DK * NSRC][kSweepThreads];
  __shared__ TapW rec_w[DK * NSRC][kSweepThreads];

  const int HW = a.H * a.W;
  const int tid = threadIdx.x;
  const int pix = blockIdx.x * kSweepThreads + tid;
  const bool active = pix < HW;
  const int pixc = active ? pix : HW - 1;
  const int d0 = blockIdx.y * DK;

  {  // ---- phase A ----
    const int y = pixc / a.W, x = pixc - y * a.W;
    const typename Geo::Pixel px = a.geo.pixel(x, y);
#pragma unroll 1
    for (int g = 0; g < DK; g += NP) {
      float h[NP];
#pragma unroll
      for (int k = 0; k < NP; ++k) {
        const int d = min(d0 + g + k, a.D - 1);
        h[k] = a.depth_per_pixel ? __ldg(a.depth + (size_t)d * HW + pixc) : __ldg(a.depth + d);
      }
      a.geo.template grid_coords<NP>(px, h, [&](int k, int v, float gx, float gy) {
        Tap t = make_tap(gx, gy, a.H, a.W, a.half_w, a.half_h);
        if (v >= a.n_src) { t.w00 = t.w01 = t.w10 = t.w11 = 0.0f; t.off = 0; }
        rec_off[(g + k) * NSRC + v][tid] = t.off;
        rec_w[(g + k) * NSRC + v][tid] = TapW{t.w00, t.w01, t.w10, t.w11};
      });
    }
  }
  // each thread only reads back its own records: no barrier needed

  // ---- phase B ----
  // addresses = warp-uniform 64-bit base + 32-bit per-thread byte offset: no per-thread 64-bit arithmetic
  const unsigned upix = (unsigned)pix * 4u;
  const size_t quad_bytes = (size_t)HW * sizeof(float4);
  const size_t row_bytes = (size_t)a.W * sizeof(float4);
  const size_t plane_bytes = (size_t)HW * sizeof(float);
  const u64 vinv = pk(a.inv_num_views, a.inv_num_views), vneg = pk(-a.num_views, -a.num_views);
  for (int c0 = 0; c0 < a.C; c0 += CH) {
    u64 r[CH / 2];
#pragma unroll
    for (int j = 0; j < CH / 2; ++j) {
      float lo = 0.f, hi = 0.f;
      if (kVariance) {
        const char* rb = reinterpret_cast<const char*>(a.ref_fea) + (size_t)(c0 + 2 * j) * plane_bytes;
        lo = __ldg(reinterpret_cast<const float*>(rb + (unsigned)pixc * 4u));
        hi = __ldg(reinterpret_cast<const float*>(rb + plane_bytes + (unsigned)pixc * 4u));
      }
      r[j] = pk(lo, hi);
    }
#pragma unroll 1
    for (int k = 0; k < DK; ++k) {
      if (d0 + k >= a.D) break;
      u64 s[CH / 2], q[CH / 2];
#pragma unroll
      for (int j = 0; j < CH / 2; ++j) { s[j] = r[j]; q[j] = mul2_rounded(r[j], r[j]); }
#pragma unroll
      for (int v = 0; v < NSRC; ++v) {
        const unsigned offb = (unsigned)rec_off[k * NSRC + v][tid] * (unsigned)sizeof(float4);
        const TapW w = rec_w[k * NSRC + v][tid];
        const u64 w00 = pk(w.w00, w.w00), w01 = pk(w.w01, w.w01), w10 = pk(w.w10, w.w10), w11 = pk(w.w11, w.w11);
        const char* vb = reinterpret_cast<const char*>(a.src_v4[v]) + (size_t)(c0 >> 2) * quad_bytes;
#pragma unroll
        for (int qd = 0; qd < CH / 4; ++qd) {
          const char* b0 = vb + (size_t)qd * quad_bytes;          // uniform
          const char* b1 = b0 + row_bytes;                         // uniform
          const float4 A = __ldg(reinterpret_cast<const float4*>(b0 + offb));
          const float4 B = __ldg(reinterpret_cast<const float4*>(b0 + offb + 16u));
          const float4 Cc = __ldg(reinterpret_cast<const float4*>(b1 + offb));
          const float4 Dd = __ldg(reinterpret_cast<const float4*>(b1 + offb + 16u));
          // nw, ne, sw, se accumulated with FMAs (ATen order), two channels per instruction
          u64 lo = fma2(pk(Dd.x, Dd.y), w11, fma2(pk(Cc.x, Cc.y), w10, fma2(pk(B.x, B.y), w01, mul2(pk(A.x, A.y), w00))));
          u64 hi = fma2(pk(Dd.z, Dd.w), w11, fma2(pk(Cc.z, Cc.w), w10, fma2(pk(B.z, B.w), w01, mul2(pk(A.z, A.w), w00))));
          if (kVariance) {
            s[2 * qd] = add2(s[2 * qd], lo);         q[2 * qd] = add2(q[2 * qd], mul2_rounded(lo, lo));
            s[2 * qd + 1] = add2(s[2 * qd + 1], hi); q[2 * qd + 1] = add2(q[2 * qd + 1], mul2_rounded(hi, hi));
          } else {
            s[2 * qd] = lo; s[2 * qd + 1] = hi;
          }
        }
      }
      const size_t obase = ((size_t)c0 * a.out_D + a.out_d0 + d0 + k) * plane_bytes;      // uniform
      const size_t ostride = (size_t)a.out_D * plane_bytes;                                 // uniform, per channel
#pragma unroll
      for (int j = 0; j < CH / 2; ++j) {
        u64 res = s[j];
        if (kVariance) {
          // x / V as a correctly rounded constant division (div_const), packed
          u64 m = mul2(s[j], vinv);  m = fma2(fma2(vneg, m, s[j]), vinv, m);
          u64 e = mul2(q[j], vinv);  e = fma2(fma2(vneg, e, q[j]), vinv, e);
          res = sub2(e, mul2_rounded(m, m));
        }
        float lo, hi;
        upk(res, lo, hi);
        if (active) {
          if (kMultiOut) {
            if (a.multicast) {
              char* ob = reinterpret_cast<char*>(a.out[0]) + obase + (size_t)(2 * j) * ostride;
              mc_store(reinterpret_cast<float*>(ob + upix), lo);
              mc_store(reinterpret_cast<float*>(ob + ostride + upix), hi);
            } else {
              for (int o = 0; o < a.n_out; ++o) {
                char* ob = reinterpret_cast<char*>(a.out[o]) + obase + (size_t)(2 * j) * ostride;
                __stcs(reinterpret_cast<float*>(ob + upix), lo);
                __stcs(reinterpret_cast<float*>(ob + ostride + upix), hi);
              }
            }
          } else {
            char* ob = reinterpret_cast<char*>(a.out[0]) + obase + (size_t)(2 * j) * ostride;
            __stcs(reinterpret_cast<float*>(ob + upix), lo);
            __stcs(reinterpret_cast<float*>(ob + ostride + upix), hi);
          }
        }
      }
    }
  }
}

// ------------------------------------------------------------------------------------------
// v5 of the forward sweep (default): 2-D reference tiles + TMA-staged, double-buffered source windows.
//   CTA      = 32 x TH reference pixels (one warp per tile row) x DK planes; 3-4 CTAs per SM.
//   phase A  fp64 geometry (lock-step planes) -> tap records in shared memory + per-view bounding box
//            of the clamped tap origins (packed u16x2 min/max, warp redux, shared atomics).
//   fix-up   each thread rewrites its own records into window-relative offsets (or global offsets
//            when a view's box does not fit the staging buffer: steep geometry falls back to L1/L2).
//   phase B  passes of CH (8) channels.  Warp 0 issues one cp.async.bulk per (view, quad, row) of the
//            [C/4][H][W] float4 re-pack into stage (p+1)&1 BEFORE the CTA gathers pass p from stage p&1
//            (completion on one mbarrier per stage), so the copy latency hides under the gather.  The
//            pitch is padded to 8 pixels so that row breaks inside a quarter-warp stay conflict-free;
//            every tap is an LDS.128 at base + immediate (no per-load address arithmetic).
// The kernel is launched as a programmatic dependent of pack_vec4_kernel: its geometry phase overlaps
// the re-pack, griddepcontrol.wait sits in front of the first read of the packed features.
// Arithmetic and op order are those of v1/v3: results are bit-identical.
// ------------------------------------------------------------------------------------------
// mbarrier / TMA bulk-copy helpers of the staged kernel
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE_%=;\n\t"
      "bra WAIT_%=;\n\t"
      "DONE_%=:\n\t}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
// global -> shared bulk copy through the TMA engine (UBLKCP); bytes and both addresses multiples of 16
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, unsigned bytes, unsigned long long* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}


#ifndef SATMVS_V5_KUNROLL
#define SATMVS_V5_KUNROLL 2
#endif
constexpr int kV5PlaneUnroll = SATMVS_V5_KUNROLL;   // planes gathered in one loop body (ILP for the 4 warps of a CTA)

template <int NSRC, int TH, int DK, int CH, int WINPX>
constexpr size_t v5_smem_bytes() {
  return (size_t)2 * NSRC * (CH / 4) * WINPX * 16 + (size_t)DK * NSRC * (TH * 32) * (4 + 16);
}

__device__ __forceinline__ unsigned min_u16x2(unsigned a, unsigned b) { unsigned r; asm("min.u16x2 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
__device__ __forceinline__ unsigned max_u16x2(unsigned a, unsigned b) { unsigned r; asm("max.u16x2 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }

template <class Geo, int TH, int DK, int CH, int WINPX, int MINB, bool kVariance, bool kMultiOut>
__global__ void __launch_bounds__(TH * 32, MINB)
sweep_fwd_v5_kernel(const __grid_constant__ SweepArgs<Geo> a, int tiles_x) {
  constexpr int NSRC = Geo::kNumSrc;
  constexpr int NT = TH * 32;
  constexpr int NP = DK < SATMVS_NP ? DK : SATMVS_NP;
  constexpr int NQ = CH / 4;
  constexpr int STAGE = NSRC * NQ * WINPX;                                                  // float4 per stage
  extern __shared__ __align__(128) unsigned char v5_smem[];
  float4* win = reinterpret_cast<float4*>(v5_smem);                                         // [2][NSRC][NQ][WINPX]
  int (*rec_off)[NT] = reinterpret_cast<int (*)[NT]>(win + 2 * STAGE);                      // [DK*NSRC][NT]
  TapW (*rec_w)[NT] = reinterpret_cast<TapW (*)[NT]>(rec_off + DK * NSRC);                  // [DK*NSRC][NT]
  __shared__ int bbox[NSRC][4];                 // xmin, xmax, ymin, ymax of the clamped origins of live taps
  __shared__ __align__(8) unsigned long long mbar[2];

  const int HW = a.H * a.W;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int ty = blockIdx.x / tiles_x, tx = blockIdx.x - ty * tiles_x;
  const int x = tx * 32 + lane, y = ty * TH + warp;
  const bool active = x < a.W && y < a.H;
  const int pix = min(y, a.H - 1) * a.W + min(x, a.W - 1);
  const int d0 = blockIdx.y * DK;

  if (tid < NSRC) { bbox[tid][0] = 1 << 30; bbox[tid][1] = -1; bbox[tid][2] = 1 << 30; bbox[tid][3] = -1; }
  if (tid == 0) { mbar_init(&mbar[0], 1); mbar_init(&mbar[1], 1); }
  __syncthreads();

  {  // ---- phase A: geometry, records, bounding boxes ----
    unsigned bmin[NSRC], bmax[NSRC];            // packed (y << 16 | x), reduced per 16-bit half
#pragma unroll
    for (int v = 0; v < NSRC; ++v) { bmin[v] = 0xffffffffu; bmax[v] = 0u; }
    const typename Geo::Pixel px = a.geo.pixel(min(x, a.W - 1), min(y, a.H - 1));
#pragma unroll 1
    for (int g = 0; g < DK; g += NP) {
      float h[NP];
#pragma unroll
      for (int k = 0; k < NP; ++k) {
        const int d = min(d0 + g + k, a.D - 1);
        h[k] = a.depth_per_pixel ? __ldg(a.depth + (size_t)d * HW + pix) : __ldg(a.depth + d);
      }
      a.geo.template grid_coords<NP>(px, h, [&](int k, int v, float gx, float gy) {
        TapXY t = make_tap_xy(gx, gy, a.H, a.W, a.half_w, a.half_h);
        const bool live = t.live && active && v < a.n_src && d0 + g + k < a.D;
        if (!live) { t.w00 = t.w01 = t.w10 = t.w11 = 0.0f; }
        const unsigned xy = ((unsigned)t.yc << 16) | (unsigned)t.xc;
#pragma unroll
        for (int vv = 0; vv < NSRC; ++vv)
          if (vv == v) { bmin[vv] = min_u16x2(bmin[vv], live ? xy : 0xffffffffu); bmax[vv] = max_u16x2(bmax[vv], live ? xy : 0u); }
        rec_off[(g + k) * NSRC + v][tid] = live ? (int)xy : -1;
        rec_w[(g + k) * NSRC + v][tid] = TapW{t.w00, t.w01, t.w10, t.w11};
      });
    }
#pragma unroll
    for (int v = 0; v < NSRC; ++v) {
      const unsigned x0 = __reduce_min_sync(0xffffffffu, bmin[v] & 0xffffu), y0 = __reduce_min_sync(0xffffffffu, bmin[v] >> 16);
      const unsigned x1 = __reduce_max_sync(0xffffffffu, bmax[v] & 0xffffu), y1 = __reduce_max_sync(0xffffffffu, bmax[v] >> 16);
      if (lane == 0 && x0 != 0xffffu) {         // this warp has at least one live tap in view v
        atomicMin(&bbox[v][0], (int)x0); atomicMax(&bbox[v][1], (int)x1); atomicMin(&bbox[v][2], (int)y0); atomicMax(&bbox[v][3], (int)y1);
      }
    }
  }
  __syncthreads();
  asm volatile("griddepcontrol.wait;" ::: "memory");   // the packed features (previous kernel) are complete from here on

  // window geometry per view (uniform): origin, copied width, padded pitch, rows
  int xs[NSRC], ys[NSRC], wc[NSRC], pitch[NSRC], rows[NSRC];
  bool staged = true;
#pragma unroll
  for (int v = 0; v < NSRC; ++v) {
    if (bbox[v][1] < 0) { xs[v] = 0; ys[v] = 0; wc[v] = 2; rows[v] = 2; }             // no live tap at all
    else { xs[v] = bbox[v][0]; ys[v] = bbox[v][2]; wc[v] = bbox[v][1] - bbox[v][0] + 2; rows[v] = bbox[v][3] - bbox[v][2] + 2; }
    pitch[v] = (wc[v] + 7) & ~7;
    staged = staged && (pitch[v] * rows[v] <= WINPX);
  }

  // one bulk copy per (view, quad, row) of pass p into stage p & 1, issued by warp 0
  auto issue = [&](int p) {
    if (warp == 0) {
      unsigned long long* bar = &mbar[p & 1];
      float4* dst0 = win + (p & 1) * STAGE;
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic reads of this stage -> async-proxy writes
      if (lane == 0) {
        unsigned total = 0;
#pragma unroll
        for (int v = 0; v < NSRC; ++v) total += (unsigned)(NQ * rows[v] * wc[v]) * 16u;
        mbar_expect_tx(bar, total);
      }
      __syncwarp();
      const int q0 = p * NQ;
#pragma unroll
      for (int v = 0; v < NSRC; ++v) {
        const int n = NQ * rows[v];
        for (int i = lane; i < n; i += 32) {
          const int qd = i / rows[v], r = i - qd * rows[v];
          const float4* src = a.src_v4[v] + ((size_t)(q0 + qd) * a.H + ys[v] + r) * a.W + xs[v];
          bulk_g2s(dst0 + (v * NQ + qd) * WINPX + r * pitch[v], src, (unsigned)wc[v] * 16u, bar);
        }
      }
    }
  };
  const int P = a.C / CH;
  if (staged) issue(0);

  // fix-up: packed (y, x) -> float4 index relative to the view's window (staged) or to the quad plane (global)
#pragma unroll
  for (int k = 0; k < DK; ++k)
#pragma unroll
    for (int v = 0; v < NSRC; ++v) {
      const int xy = rec_off[k * NSRC + v][tid];
      const int xc = xy & 0xffff, yc = xy >> 16;
      int off = 0;
      if (xy >= 0) off = staged ? (yc - ys[v]) * pitch[v] + (xc - xs[v]) : yc * a.W + xc;
      rec_off[k * NSRC + v][tid] = off;
    }

  const unsigned upix = (unsigned)pix * 4u;
  const size_t plane_bytes = (size_t)HW * sizeof(float);
  const size_t ostride = (size_t)a.out_D * plane_bytes;                                    // per channel
  const u64 vinv = pk(a.inv_num_views, a.inv_num_views), vneg = pk(-a.num_views, -a.num_views);
  // x*x rounded on its own as fma(x, x, -0): the addend arrives as a kernel argument, so ptxas can neither fold
  // it away nor contract the product into the following add (profiles/r01_sweep_v3_notes.md)
  const u64 nz = pk(a.neg_zero, a.neg_zero);
#pragma unroll 1
  for (int p = 0; p < P; ++p) {
    const int c0 = p * CH;
    if (staged && p + 1 < P) {
      if (p >= 1) __syncthreads();                // every thread is done with pass p-1, whose stage is refilled now
      issue(p + 1);
    }
    u64 r[CH / 2];
#pragma unroll
    for (int j = 0; j < CH / 2; ++j) {
      float lo = 0.f, hi = 0.f;
      if (kVariance) {
        const char* rb = reinterpret_cast<const char*>(a.ref_fea) + (size_t)(c0 + 2 * j) * plane_bytes;
        lo = __ldg(reinterpret_cast<const float*>(rb + upix));
        hi = __ldg(reinterpret_cast<const float*>(rb + plane_bytes + upix));
      }
      r[j] = pk(lo, hi);
    }
    u64 r2[CH / 2];
#pragma unroll
    for (int j = 0; j < CH / 2; ++j) r2[j] = fma2(r[j], r[j], nz);
    if (staged) mbar_wait(&mbar[p & 1], (unsigned)(p >> 1) & 1u);
    const float4* wst = win + (p & 1) * STAGE;

    // one plane: gather all views from `fetch`, then the variance epilogue.  Planes past D (last chunk) carry dead
    // records (zero weights, offset 0) and are computed but not stored, so the loop body has no exits and the
    // compiler can interleave consecutive planes.
    auto plane = [&](int k, auto fetch) {
      u64 s[CH / 2], q[CH / 2];
#pragma unroll
      for (int v = 0; v < NSRC; ++v) {
        const int off = rec_off[k * NSRC + v][tid];
        const TapW w = rec_w[k * NSRC + v][tid];
        const u64 w00 = pk(w.w00, w.w00), w01 = pk(w.w01, w.w01), w10 = pk(w.w10, w.w10), w11 = pk(w.w11, w.w11);
#pragma unroll
        for (int qd = 0; qd < NQ; ++qd) {
          float4 A, B, Cc, Dd;
          fetch(v, qd, off, A, B, Cc, Dd);
          // nw, ne, sw, se accumulated with FMAs (ATen order), two channels per instruction
          u64 lo = fma2(pk(Dd.x, Dd.y), w11, fma2(pk(Cc.x, Cc.y), w10, fma2(pk(B.x, B.y), w01, mul2(pk(A.x, A.y), w00))));
          u64 hi = fma2(pk(Dd.z, Dd.w), w11, fma2(pk(Cc.z, Cc.w), w10, fma2(pk(B.z, B.w), w01, mul2(pk(A.z, A.w), w00))));
          if (kVariance) {
            // the first view accumulates onto the reference terms directly (no per-plane copies of r / r2)
            s[2 * qd] = add2(v == 0 ? r[2 * qd] : s[2 * qd], lo);
            q[2 * qd] = add2(v == 0 ? r2[2 * qd] : q[2 * qd], fma2(lo, lo, nz));
            s[2 * qd + 1] = add2(v == 0 ? r[2 * qd + 1] : s[2 * qd + 1], hi);
            q[2 * qd + 1] = add2(v == 0 ? r2[2 * qd + 1] : q[2 * qd + 1], fma2(hi, hi, nz));
          } else {
            s[2 * qd] = lo; s[2 * qd + 1] = hi;
          }
        }
      }
      const bool store = active && d0 + k < a.D;
      const size_t obase = ((size_t)c0 * a.out_D + a.out_d0 + d0 + k) * plane_bytes + upix;
      char* op = reinterpret_cast<char*>(a.out[0]) + obase;
#pragma unroll
      for (int j = 0; j < CH / 2; ++j) {
        u64 res = s[j];
        if (kVariance) {
          // x / V as a correctly rounded constant division (div_const), packed
          u64 m = mul2(s[j], vinv);  m = fma2(fma2(vneg, m, s[j]), vinv, m);
          u64 e = mul2(q[j], vinv);  e = fma2(fma2(vneg, e, q[j]), vinv, e);
          res = sub2(e, fma2(m, m, nz));
        }
        float lo, hi;
        upk(res, lo, hi);
        if (store) {
          if (kMultiOut) {
            if (a.multicast) {
              char* ob = reinterpret_cast<char*>(a.out[0]) + obase + (size_t)(2 * j) * ostride;
              mc_store(reinterpret_cast<float*>(ob), lo);
              mc_store(reinterpret_cast<float*>(ob + ostride), hi);
            } else {
              for (int o = 0; o < a.n_out; ++o) {
                char* ob = reinterpret_cast<char*>(a.out[o]) + obase + (size_t)(2 * j) * ostride;
                __stcs(reinterpret_cast<float*>(ob), lo);
                __stcs(reinterpret_cast<float*>(ob + ostride), hi);
              }
            }
          } else {
            __stcs(reinterpret_cast<float*>(op), lo);
            __stcs(reinterpret_cast<float*>(op + ostride), hi);
            op += 2 * ostride;
          }
        }
      }
    };
    if (staged) {
      int pt[NSRC];
#pragma unroll
      for (int v = 0; v < NSRC; ++v) pt[v] = pitch[v];
      auto fetch = [&](int v, int qd, int off, float4& A, float4& B, float4& Cc, float4& Dd) {
        const float4* b0 = wst + v * (NQ * WINPX) + qd * WINPX + off;      // shared: LDS.128 at base + immediate
        A = b0[0]; B = b0[1]; Cc = b0[pt[v]]; Dd = b0[pt[v] + 1];
      };
#pragma unroll kV5PlaneUnroll
      for (int k = 0; k < DK; ++k) plane(k, fetch);
    } else {
      auto fetch = [&](int v, int qd, int off, float4& A, float4& B, float4& Cc, float4& Dd) {
        const float4* b0 = a.src_v4[v] + (size_t)(p * NQ + qd) * HW + off;
        A = __ldg(b0); B = __ldg(b0 + 1); Cc = __ldg(b0 + a.W); Dd = __ldg(b0 + a.W + 1);
      };
#pragma unroll 1
      for (int k = 0; k < DK; ++k) plane(k, fetch);
    }
  }
}

// opt in to > 48 KB of dynamic shared memory once per (kernel, device, host thread)
template <auto Kern>
static void allow_dynamic_smem(size_t smem) {
  static thread_local int ready_dev = -1;
  int dev = 0;
  cudaGetDevice(&dev);
  if (ready_dev != dev) { cudaFuncSetAttribute(Kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); ready_dev = dev; }
}

template <class Geo, int TH, int DK, int CH, int WINPX, int MINB, bool kVariance>
static int launch_v5(const SweepArgs<Geo>& a, cudaStream_t st, bool after_pack) {
  constexpr size_t smem = v5_smem_bytes<Geo::kNumSrc, TH, DK, CH, WINPX>();
  static_assert(smem <= 220 * 1024, "v5 configuration exceeds shared memory");
  const int tiles_x = ceil_div(a.W, 32), tiles_y = ceil_div(a.H, TH);
  auto run = [&](auto kern) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(tiles_x * tiles_y, ceil_div(a.D, DK));
    cfg.blockDim = dim3(TH * 32);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = after_pack ? 1 : 0;          // only directly behind pack_vec4_kernel (which waited for all earlier work)
    cudaLaunchKernelEx(&cfg, kern, a, tiles_x);
  };
  // (both instantiations have the same function type, so the opt-in is keyed on the kernel itself)
  if (a.n_out > 1 || a.multicast) {
    allow_dynamic_smem<sweep_fwd_v5_kernel<Geo, TH, DK, CH, WINPX, MINB, kVariance, true>>(smem);
    run(sweep_fwd_v5_kernel<Geo, TH, DK, CH, WINPX, MINB, kVariance, true>);
  } else {
    allow_dynamic_smem<sweep_fwd_v5_kernel<Geo, TH, DK, CH, WINPX, MINB, kVariance, false>>(smem);
    run(sweep_fwd_v5_kernel<Geo, TH, DK, CH, WINPX, MINB, kVariance, false>);
  }
  return check_launch("sweep_fwd_v5_kernel");
}

// v5 applies to <= 2 source views (3-view stacks, every cascade stage).  With 4 source views the records and
// windows leave 2 CTAs per SM and the L1-gather kernel (v3) is faster (2.2 ms against 1.97 ms at cfg-4):
// profiles/r01_sweep_v5_notes.md.  Returns -1 when v5 does not apply.
template <class Geo, bool kVariance>
static int dispatch_v5(const SweepArgs<Geo>& a, cudaStream_t st, bool after_pack) {
  constexpr int NS = Geo::kNumSrc;
  if (a.W >= 32768 || a.H >= 32768 || (a.C & 7)) return -1;
  if constexpr (NS <= 2) {
    static const bool dk8 = getenv("SATMVS_SWEEP_DK8") != nullptr;     // tuning alternative: 8 planes per CTA, 3 CTAs per SM
    if (dk8) return launch_v5<Geo, 4, 8, 8, 256, 3, kVariance>(a, st, after_pack);
    return launch_v5<Geo, 4, 4, 8, 256, 4, kVariance>(a, st, after_pack);
  } else {
    return -1;
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
  a.neg_zero = -0.0f;
  if (a.src_v4[0] != nullptr) {
    static const bool no_v5 = getenv("SATMVS_SWEEP_V3") != nullptr;
    if (!no_v5) {
      const int rc = dispatch_v5<Geo, kVariance>(a, st, a.packed_now != 0 && !prof_state().on);
      if (rc >= 0) return rc;
    }
    if ((a.n_out > 1 || a.multicast) && a.C % SATMVS_CH == 0) sweep_fwd_v3_kernel<Geo, DK, SATMVS_CH, kVariance, true><<<grid, kSweepThreads, 0, st>>>(a);
    else if (a.n_out > 1 || a.multicast) sweep_fwd_v3_kernel<Geo, DK, 4, kVariance, true><<<grid, kSweepThreads, 0, st>>>(a);
    else if (a.C % SATMVS_CH == 0) sweep_fwd_v3_kernel<Geo, DK, SATMVS_CH, kVariance, false><<<grid, kSweepThreads, 0, st>>>(a);
    else sweep_fwd_v3_kernel<Geo, DK, 4, kVariance, false><<<grid, kSweepThreads, 0, st>>>(a);
    return check_launch("sweep_fwd_v3_kernel");
  }
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

// Re-pack the source feature maps into the caller's workspace ([n_src][C/4][H][W] float4) and point the
// kernel arguments at them.  Without a workspace (or when C is not a multiple of 4) the scalar path runs.
template <class Args>
static int pack_sources(Args& a, const float* const* src_feas, int n_src, int nslots, void* workspace, size_t workspace_bytes,
                        cudaStream_t st) {
  for (int v = 0; v < nslots; ++v) a.src_v4[v] = nullptr;
  const size_t per_view = (size_t)a.C * a.H * a.W * sizeof(float);
  if (workspace == nullptr || (a.C & 3) || workspace_bytes < per_view * n_src ||
      (reinterpret_cast<uintptr_t>(workspace) & 15)) return SATMVS_OK;
  const int HW = a.H * a.W;
  ProfScope prof(kProfSweep, st);
  PackArgs pa{};
  pa.HW = HW;
  for (int v = 0; v < n_src; ++v) {
    float4* dst = reinterpret_cast<float4*>(static_cast<char*>(workspace) + per_view * v);
    pa.in[v] = src_feas[v]; pa.out[v] = dst;
    a.src_v4[v] = dst;
  }
  pack_vec4_kernel<<<dim3(ceil_div(HW, 256), a.C / 4, n_src), 256, 0, st>>>(pa);
  a.packed_now = 1;
  for (int v = n_src; v < nslots; ++v) a.src_v4[v] = a.src_v4[0];
  return check_launch("pack_vec4_kernel");
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
                                            C, D, H, W, 0, D, outs, 1, nullptr, 0, stream);
}

int satmvs_cost_volume_rpc_fwd_sharded(const float* ref_fea, const float* const* src_feas, int n_src,
                                       const double* ref_rpc, const double* src_rpcs,
                                       const float* depth, int depth_per_pixel,
                                       int C, int D, int H, int W, int d0, int D_total,
                                       float* const* outs, int n_outs, void* workspace, size_t workspace_bytes, void* stream) {
  if (int e = check_dims(n_src, C, D, H, W)) return e;
  SATMVS_REQUIRE(ref_fea && src_feas && ref_rpc && src_rpcs && depth && outs);
  const int multicast = n_outs == -1;          // outs[0] is a multicast address
  if (multicast) n_outs = 1;
  SATMVS_REQUIRE(n_outs >= 1 && n_outs <= SATMVS_MAX_PEERS && d0 >= 0 && d0 + D <= D_total);
  for (int o = 0; o < n_outs; ++o) SATMVS_REQUIRE(outs[o] != nullptr);
  SATMVS_DISPATCH_NSRC(n_src, {
    SweepArgs<RpcSweep<NSRC>> a{};
    fill_common(a, depth, depth_per_pixel, n_src, C, D, H, W);
    a.ref_fea = ref_fea;
    for (int o = 0; o < n_outs; ++o) a.out[o] = outs[o];
    a.n_out = n_outs; a.multicast = multicast; a.out_D = D_total; a.out_d0 = d0;
    for (int v = 0; v < NSRC; ++v) a.src_fea[v] = src_feas[v < n_src ? v : 0];
    if (int e = pack_sources(a, src_feas, n_src, NSRC, workspace, workspace_bytes, (cudaStream_t)stream)) return e;
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
                                             C, D, H, W, 0, D, outs, 1, nullptr, 0, stream);
}

int satmvs_cost_volume_homo_fwd_sharded(const float* ref_fea, const float* const* src_feas, int n_src,
                                        const double* ref_proj, const double* src_projs,
                                        const float* depth, int depth_per_pixel,
                                        int C, int D, int H, int W, int d0, int D_total,
                                        float* const* outs, int n_outs, void* workspace, size_t workspace_bytes, void* stream) {
  if (int e = check_dims(n_src, C, D, H, W)) return e;
  SATMVS_REQUIRE(ref_fea && src_feas && ref_proj && src_projs && depth && outs);
  const int multicast = n_outs == -1;          // outs[0] is a multicast address
  if (multicast) n_outs = 1;
  SATMVS_REQUIRE(n_outs >= 1 && n_outs <= SATMVS_MAX_PEERS && d0 >= 0 && d0 + D <= D_total);
  for (int o = 0; o < n_outs; ++o) SATMVS_REQUIRE(outs[o] != nullptr);
  SATMVS_DISPATCH_NSRC(n_src, {
    SweepArgs<HomoSweep<NSRC>> a{};
    fill_common(a, depth, depth_per_pixel, n_src, C, D, H, W);
    a.ref_fea = ref_fea;
    for (int o = 0; o < n_outs; ++o) a.out[o] = outs[o];
    a.n_out = n_outs; a.multicast = multicast; a.out_D = D_total; a.out_d0 = d0;
    for (int v = 0; v < NSRC; ++v) a.src_fea[v] = src_feas[v < n_src ? v : 0];
    if (int e = pack_sources(a, src_feas, n_src, NSRC, workspace, workspace_bytes, (cudaStream_t)stream)) return e;
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
