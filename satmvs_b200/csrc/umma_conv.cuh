// umma_conv.cuh — 3x3 / 3x3x3 (stride 1, padding 1) convolution on the 5th-generation tensor cores (tcgen05),
// fp32-faithful through a 3-way TF32 split, for the batched (all depth planes at once) layers of RED and the dense
// stride-1 layers of CostRegNet (a 3x3x3 conv = three plane convs over z-1, z, z+1 accumulated in the same TMEM tile).
//
// Formulation.  A plane is cut into vertical strips of TW columns; a strip is addressed in PADDED-FLATTENED positions
// q = ry*Wp + rx, Wp = TW + 2 (one halo column / row on every side: the neighbouring pixels, or zeros outside the image),
// and a CTA owns a run of 128*MT consecutive output positions of its strip (runs need not start at a row boundary).
// For such a run the input of tap (dy, dx) is the same run shifted by dy*Wp + dx, so
// with the input window in shared memory as [channel quad][position] float4 (the sweep's re-pack layout, which IS the
// SWIZZLE_NONE K-major canonical layout of tcgen05: 8 positions x 16 bytes per core matrix, SBO = 128 B between
// 8-position groups, LBO = window pitch between channel quads) every tap is ONE shared-memory descriptor whose start
// address is moved by the shift: no im2col copy, each input element is staged once per CTA.
//   D[128 positions x N] += A_tap[128 x 8 channels] * W_tap[N x 8 channels]^T      (kind::tf32, M 128, K 8)
// over 9 taps x Cin/8 k-steps.  N = all output channels that share the input (several "heads": GRU gate and output
// convolutions of one level), padded to a multiple of 16; accumulators live in TMEM (MT tiles x N columns).
//
// Precision.  The tensor core TRUNCATES fp32 operands to TF32 (measured: tools/probes/umma_probe.cu), so the raw fp32
// tile is the "hi" operand for free; lo = x - trunc(x) is exact in fp32 and is staged next to it.  Three MMAs per step
// (hi*hi + hi*lo + lo*hi, fp32 accumulation in TMEM) leave ~2^-20 relative error per product, against 2^-11 for a
// single TF32 pass: the logits stay within the parity bound of the fp32 FFMA path (tests/test_gpu_red.py).
//
// Pipeline per CTA (256 threads, 2 CTAs per SM so one stages while the other multiplies): for every chunk of 8 input
// channels: wait for the previous chunk's MMAs (tcgen05.commit -> mbarrier), stage window + packed weights, proxy
// fence + barrier, ONE thread issues MT x 9 x 3 tcgen05.mma; after the last chunk the accumulators are read back with
// tcgen05.ld (32 lanes x 32 bit, one row = one position per thread) and the epilogue (scale, bias, ReLU) stores
// coalesced rows of the NCDHW outputs.
#pragma once
#include <cstdlib>
#include "common.cuh"

namespace satmvs {

constexpr int kUcThreads = 256, kUcKC = 8, kUcMaxHeads = 3, kUcMaxMT = 4;

struct UmmaHead {
  const float* scale;      // [Cout] or null (folded BatchNorm)
  const float* shift;      // [Cout] or null
  float* out;              // [Cout][D][H][W]
  int Cout;
  int n0;                  // first column of this head in the fused N dimension
  float acc_scale;
  int relu;
  int stride;              // 1, or 2: only even (y, x) are stored, into a [Cout][D][H/2][W/2] tensor (a stride-2 conv shares
                           // its input window with the stride-1 heads; the unused 3/4 of its columns cost no extra staging)
};

struct UmmaConv2d {
  const float* in;         // [Cin] channels of D planes of H x W
  long long in_cs;         // input channel stride in elements
  const float4* wpack;     // packed weights, see umma_pack_weights_kernel
  int Cin, D, H, W;
  int NZ;                  // 1: per-plane 3x3, 3: 3x3x3 over planes z-1, z, z+1 (zero padding in z)
  int NP;                  // fused output channels, padded to a multiple of 16
  int TW;                  // useful columns of a strip (window pitch Wp = TW + 2)
  int strips;              // strips per plane: blockIdx.x = run * strips + strip
  int MT;                  // 128-position tiles per CTA
  int PW;                  // window positions = 128*MT + 2*Wp + 2
  int nheads;
  UmmaHead head[kUcMaxHeads];
  int d0;                  // first plane of this launch (blockIdx.y counts from it): a conv over D planes can be launched in chunks
  int* ready;              // optional per-plane counters [D]: every CTA adds 1 to ready[d] once its outputs are visible device-wide
};

struct UmmaPackHead { const float* w; long long w_co, w_ci; int Cout, n0; };
struct UmmaPack { UmmaPackHead head[kUcMaxHeads]; int nheads, Cin, NP, NZ; float4* out; };

// Packed weights: [chunk = Cin/8][kz NZ][part: 0 raw, 1 lo][tap 9][kq 2][n NP] float4 (4 consecutive input channels);
// the weight of (co, ci, kz, tap) sits at w[co*w_co + ci*w_ci + kz*9 + tap].
// Column n of the fused N dimension belongs to the head with n0 <= n < n0 + Cout; padding columns are zero.
static __global__ void umma_pack_weights_kernel(const __grid_constant__ UmmaPack a) {
  const int total = (a.Cin / kUcKC) * a.NZ * 2 * 9 * 2 * a.NP;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int n = i % a.NP;
  int r = i / a.NP;
  const int kq = r % 2; r /= 2;
  int tap = r % 9; r /= 9;
  const int part = r % 2; r /= 2;
  const int kz = r % a.NZ;
  const int chunk = r / a.NZ;
  tap += kz * 9;
  float v[4] = {0.f, 0.f, 0.f, 0.f};
  for (int h = 0; h < a.nheads; ++h) {
    const UmmaPackHead& H = a.head[h];
    if (n >= H.n0 && n < H.n0 + H.Cout) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int ci = chunk * kUcKC + 4 * kq + j;
        const float w = __ldg(H.w + (long long)(n - H.n0) * H.w_co + (long long)ci * H.w_ci + tap);
        const float hi = __uint_as_float(__float_as_uint(w) & 0xffffe000u);
        v[j] = part ? (w - hi) : w;
      }
    }
  }
  a.out[i] = make_float4(v[0], v[1], v[2], v[3]);
}

__device__ __forceinline__ unsigned uc_smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

// SWIZZLE_NONE K-major shared-memory matrix descriptor (cute::UMMA::SmemDescriptor): start >> 4 at bit 0,
// leading (K-chunk) byte offset >> 4 at bit 16, stride (8-row group) byte offset >> 4 at bit 32, version 1 at bit 46
__device__ __forceinline__ unsigned long long uc_desc(unsigned saddr, unsigned lbo_bytes, unsigned sbo_bytes) {
  return (unsigned long long)((saddr >> 4) & 0x3fffu) | ((unsigned long long)((lbo_bytes >> 4) & 0x3fffu) << 16) |
         ((unsigned long long)((sbo_bytes >> 4) & 0x3fffu) << 32) | (1ULL << 46);
}

__device__ __forceinline__ void uc_mma_tf32(unsigned tmem_d, unsigned long long da, unsigned long long db, unsigned idesc, unsigned accumulate) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
               "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
               ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate) : "memory");
}

__device__ __forceinline__ bool uc_elect_one() {
  unsigned pred = 0;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}

// bounded spin on an mbarrier phase; returns false on timeout (the kernel then skips its stores and flags the error)
__device__ __forceinline__ bool uc_wait(unsigned long long* bar, unsigned parity) {
  unsigned ok = 0;
  for (int it = 0; it < (1 << 22) && !ok; ++it)
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(uc_smem_u32(bar)), "r"(parity) : "memory");
  return ok != 0;
}

__host__ __device__ constexpr int uc_tmem_cols(int need) { return need <= 32 ? 32 : need <= 64 ? 64 : need <= 128 ? 128 : need <= 256 ? 256 : 512; }

inline size_t umma_conv_smem_bytes(int NP, int PW) {
  return (size_t)2 * 2 * PW * 16 + (size_t)2 * 9 * 2 * NP * 16;      // window (raw, lo) x 2 quads + packed weights of one chunk
}

// NPOS: window positions per thread: 4 (PW <= 1024, 3 CTAs per SM) or 8 (PW <= 2048, 2 CTAs per SM); NZ: 1 or 3 (a.NZ)
template <int NPOS, int NZ>
__global__ void __launch_bounds__(kUcThreads, NPOS == 4 ? 3 : 2)
umma_conv2d_kernel(const __grid_constant__ UmmaConv2d a, int* error_flag) {
  extern __shared__ __align__(128) unsigned char uc_smem[];
  float4* win = reinterpret_cast<float4*>(uc_smem);                    // [part 2][kq 2][PW]
  float4* wts = win + 4 * a.PW;                                        // [part 2][tap 9][kq 2][NP]
  __shared__ unsigned tmem_base_s;
  __shared__ __align__(8) unsigned long long bar, wbar;                // MMAs of a chunk done / packed weights of a chunk landed

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int Wp = a.TW + 2, HW = a.H * a.W;
  const int d = blockIdx.y + a.d0;
  const int run = blockIdx.x / a.strips, sx = blockIdx.x - run * a.strips;
  const int x0 = sx * a.TW;                                            // first useful column of the strip
  const int q0 = Wp + run * (128 * a.MT);                              // first output position of this CTA (row y = 0 starts at Wp)
  const int tmem_cols = uc_tmem_cols(a.MT * a.NP);

  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(uc_smem_u32(&bar)));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(uc_smem_u32(&wbar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(uc_smem_u32(&tmem_base_s)), "r"(tmem_cols));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }

  // window positions owned by this thread (fixed over the chunks): source pixel offset, or -1 for the zero halo
  int src_off[NPOS];
#pragma unroll
  for (int j = 0; j < NPOS; ++j) {
    const int p = tid + j * kUcThreads;
    const int qs = q0 - Wp - 1 + p;                                    // strip-local flattened position of window index p
    const int ry = qs / Wp, rx = qs - ry * Wp;
    const int y = ry - 1, x = x0 - 1 + rx;
    const bool ok = p < a.PW && qs >= 0 && y >= 0 && y < a.H && x >= 0 && x < a.W;
    src_off[j] = ok ? y * a.W + x : -1;
  }
  // 8 channels of a chunk for this thread's positions, global -> registers (all loads in flight together)
  float v[NPOS][kUcKC];
  auto load_chunk = [&](int c, int zi) {
    const float* in_c = a.in + (long long)(c * kUcKC) * a.in_cs + (long long)zi * HW;
#pragma unroll
    for (int k = 0; k < kUcKC; ++k) {
      const float* in_k = in_c + k * a.in_cs;
#pragma unroll
      for (int j = 0; j < NPOS; ++j) v[j][k] = src_off[j] >= 0 ? __ldg(in_k + src_off[j]) : 0.0f;
    }
  };
  // steps = (channel chunk, kz); planes outside [0, D) are the zero padding in z and are skipped altogether
  constexpr int zoff = NZ >> 1;
  const int nsteps = (a.Cin / kUcKC) * NZ;
  auto next_step = [&](int s0) {
    if (NZ > 1)
      while (s0 < nsteps) { const int zi = d + (s0 % NZ) - zoff; if (zi >= 0 && zi < a.D) break; ++s0; }
    return s0;
  };
  int step = next_step(0);
  load_chunk(step / NZ, d + (step % NZ) - zoff);

  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const unsigned tmem = tmem_base_s;
  const unsigned idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((unsigned)(a.NP >> 3) << 17) | ((unsigned)(128 >> 4) << 24);

  const unsigned wts_bytes = (unsigned)(2 * 9 * 2 * a.NP) * 16u;      // packed weights of one step
  bool alive = true;
  int done = 0;                                                        // steps executed so far (mbarrier phases)
  while (step < nsteps) {
    const int c = done;                                                // phase index of this step
    if (c > 0) alive = uc_wait(&bar, (unsigned)(c - 1) & 1u) && alive;   // previous step's MMAs have read the window and the weights
    if (tid == 0) {   // packed weights of this step: one TMA bulk copy (async proxy -> async proxy, no generic fence needed)
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(uc_smem_u32(&wbar)), "r"(wts_bytes) : "memory");
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                   ::"r"(uc_smem_u32(wts)), "l"(reinterpret_cast<const char*>(a.wpack) + (size_t)step * wts_bytes), "r"(wts_bytes),
                     "r"(uc_smem_u32(&wbar)) : "memory");
    }
    // window of this chunk: registers -> shared memory, raw value and low part (x - trunc_tf32(x))
#pragma unroll
    for (int j = 0; j < NPOS; ++j) {
      const int p = tid + j * kUcThreads;
      if (p < a.PW) {
        float lo[kUcKC];
#pragma unroll
        for (int k = 0; k < kUcKC; ++k) lo[k] = v[j][k] - __uint_as_float(__float_as_uint(v[j][k]) & 0xffffe000u);
        win[0 * a.PW + p] = make_float4(v[j][0], v[j][1], v[j][2], v[j][3]);
        win[1 * a.PW + p] = make_float4(v[j][4], v[j][5], v[j][6], v[j][7]);
        win[2 * a.PW + p] = make_float4(lo[0], lo[1], lo[2], lo[3]);
        win[3 * a.PW + p] = make_float4(lo[4], lo[5], lo[6], lo[7]);
      }
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");       // generic-proxy stores -> tensor-core (async proxy) reads
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0 && uc_elect_one()) {   // one elected lane of a converged warp: the MMAs issue as straight uniform-datapath code
      alive = uc_wait(&wbar, (unsigned)c & 1u) && alive;
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      // descriptors differ only in their start-address field (low word): one add per MMA
      const unsigned lbo_a = (unsigned)a.PW * 16u, lbo_b = (unsigned)a.NP * 16u;
      const unsigned long long da0 = uc_desc(uc_smem_u32(win), lbo_a, 128), db0 = uc_desc(uc_smem_u32(wts), lbo_b, 128);
      const unsigned a_lo_off = 2u * (unsigned)a.PW, b_lo_off = (unsigned)(9 * 2 * a.NP), b_tap = (unsigned)(2 * a.NP);   // 16-byte units
      for (int mt = 0; mt < a.MT; ++mt) {
        const unsigned dcol = tmem + (unsigned)(mt * a.NP);
#pragma unroll
        for (int tap = 0; tap < 9; ++tap) {
          const unsigned shift = (unsigned)(128 * mt + (tap / 3) * Wp + (tap % 3));   // (dy+1)*Wp + (dx+1), 16-byte units
          const unsigned long long a_raw = da0 + shift, a_lo = a_raw + a_lo_off;
          const unsigned long long b_raw = db0 + (unsigned)tap * b_tap, b_lo = b_raw + b_lo_off;
          uc_mma_tf32(dcol, a_raw, b_raw, idesc, (c == 0 && tap == 0) ? 0u : 1u);
          uc_mma_tf32(dcol, a_raw, b_lo, idesc, 1u);
          uc_mma_tf32(dcol, a_lo, b_raw, idesc, 1u);
        }
      }
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(uc_smem_u32(&bar)) : "memory");
    }
    __syncwarp();
    step = next_step(step + 1);
    ++done;
    if (step < nsteps) load_chunk(step / NZ, d + (step % NZ) - zoff);   // the next step's global loads fly while the tensor core works
  }
  alive = uc_wait(&bar, (unsigned)(done - 1) & 1u) && alive;
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  if (!alive) {            // a tensor-core completion never arrived: fail loudly (sticky CUDA error) instead of storing garbage
    if (tid == 0) *reinterpret_cast<volatile int*>(error_flag) = 1;
    __trap();
  }

  // epilogue: warp w reads TMEM lanes 32*(w%4) .. +31 (one output position per thread), tiles mt = w/4, w/4 + 2, ...
  {
    const int quarter = warp & 3;
    for (int mt = warp >> 2; mt < a.MT; mt += kUcThreads / 128) {
      const int q = q0 + 128 * mt + 32 * quarter + lane;
      const int ry = q / Wp, rx = q - ry * Wp;
      const int y = ry - 1, x = x0 - 1 + rx;
      const bool ok1 = alive && y < a.H && rx >= 1 && rx <= a.TW && x < a.W;
      const long long opix1 = (long long)d * HW + (long long)y * a.W + x;
      const bool ok2 = ok1 && !(y & 1) && !(x & 1);
      const long long opix2 = (long long)d * (HW >> 2) + (long long)(y >> 1) * (a.W >> 1) + (x >> 1);
      for (int h = 0; h < a.nheads; ++h) {
        const UmmaHead& Hd = a.head[h];
        const bool ok = Hd.stride == 2 ? ok2 : ok1;
        float* op = Hd.out + (Hd.stride == 2 ? opix2 : opix1);
        const long long ocs = (long long)a.D * (Hd.stride == 2 ? (HW >> 2) : HW);
        for (int c0 = 0; c0 < Hd.Cout; c0 += 8) {
          unsigned r[8];
          const unsigned taddr = tmem + ((unsigned)(32 * quarter) << 16) + (unsigned)(mt * a.NP + Hd.n0 + c0);
          asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                       : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "r"(taddr));
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
          const int nj = min(8, Hd.Cout - c0);               // uniform: 8 except in the last group of a ragged head
          if (ok) {
            if (nj == 8 && Hd.scale == nullptr) {            // RED heads: bias only
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                float val = __uint_as_float(r[j]) * Hd.acc_scale + (Hd.shift ? __ldg(Hd.shift + c0 + j) : 0.0f);
                if (Hd.relu) val = fmaxf(val, 0.0f);
                op[(long long)(c0 + j) * ocs] = val;
              }
            } else {                                          // folded BatchNorm and / or fewer than 8 channels left
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                if (j < nj) {
                  float val = __uint_as_float(r[j]) * Hd.acc_scale;
                  if (Hd.scale) val *= __ldg(Hd.scale + c0 + j);
                  if (Hd.shift) val += __ldg(Hd.shift + c0 + j);
                  if (Hd.relu) val = fmaxf(val, 0.0f);
                  op[(long long)(c0 + j) * ocs] = val;
                }
              }
            }
          }
        }
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  if (a.ready != nullptr) __threadfence();                             // consumers on other SMs poll ready[d] (red_tc.cuh)
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(tmem_cols));
  if (a.ready != nullptr && tid == 32) { __threadfence(); atomicAdd(a.ready + d, 1); }
}

// Plan + launch.  Returns -1 when the layer does not fit this kernel (the caller then uses the direct FFMA kernel).
struct UmmaConvPlan { UmmaConv2d conv; UmmaPack pack; size_t smem; size_t wpack_bytes; dim3 grid; };

inline bool umma_conv_plan(UmmaConvPlan& P, const float* in, long long in_cs, int Cin, int D, int H, int W,
                           int nheads, const UmmaPackHead* wheads, const UmmaHead* oheads, void* wpack_buf, size_t wpack_cap,
                           int NZ = 1, bool perf_rules = true) {
  if (Cin % kUcKC || nheads < 1 || nheads > kUcMaxHeads || (NZ != 1 && NZ != 3)) return false;
  int n = 0;
  P = UmmaConvPlan{};
  for (int h = 0; h < nheads; ++h) {
    P.pack.head[h] = wheads[h]; P.pack.head[h].n0 = n;
    P.conv.head[h] = oheads[h]; P.conv.head[h].n0 = n;
    if (wheads[h].Cout != oheads[h].Cout || wheads[h].Cout < 1) return false;
    n += (wheads[h].Cout + 7) / 8 * 8;            // every head starts on a multiple of 8 columns
  }
  const int NP = (n + 15) / 16 * 16;
  if (NP > 256) return false;
  // Tile search over (strips, MT).  The MMAs are shared-memory-read bound and cost ~7x a staged window position per row,
  // so the cost is 7 * MMA rows issued + window positions staged, per plane.  Preference: 3 CTAs per SM (<= 74 KB of shared
  // memory, <= 128 TMEM columns, <= 1024 window positions) -- co-resident CTAs are what overlaps one CTA's staging with
  // another's MMAs -- then 2 CTAs per SM (<= 100 KB, <= 256 columns, <= 2048 positions).
  static const int want3 = getenv("SATMVS_UMMA_2CTA") ? 0 : 1;
  int TW = 0, MT = 0, PW = 0;
  for (int pass = want3 ? 0 : 1; pass < 2 && TW == 0; ++pass) {
    const int max_cols = pass == 0 ? 128 : 256, max_pos = (pass == 0 ? 4 : 8) * kUcThreads;
    const size_t max_smem = (pass == 0 ? 74 : 100) * 1024;
    double best = 1e30;
    for (int nstr = 1; nstr <= 64; ++nstr) {
      const int tw = (W + nstr - 1) / nstr, wp = tw + 2;
      if (tw < 16 && nstr > 1) break;
      for (int mt = 1; mt <= kUcMaxMT; ++mt) {
        const int pw = 128 * mt + 2 * wp + 2;
        if (uc_tmem_cols(mt * NP) > max_cols || pw > max_pos || umma_conv_smem_bytes(NP, pw) > max_smem) break;
        const double runs = (double)((H * wp + 128 * mt - 1) / (128 * mt));
        double cost = (7.0 * 128 * mt + pw) * runs * nstr;
        if (tw < 32) cost *= 1.15;                    // rows shorter than a 128-byte line
        if (cost < best) { best = cost; TW = tw; MT = mt; PW = pw; }
      }
    }
  }
  if (TW == 0) return false;
  if (PW > 8 * kUcThreads || (size_t)PW * 16 >= (1u << 18)) return false;
  P.wpack_bytes = (size_t)(Cin / kUcKC) * NZ * 2 * 9 * 2 * NP * 16;
  if (wpack_buf == nullptr || wpack_cap < P.wpack_bytes || (reinterpret_cast<uintptr_t>(wpack_buf) & 15)) return false;
  P.pack.nheads = nheads; P.pack.Cin = Cin; P.pack.NP = NP; P.pack.NZ = NZ; P.pack.out = static_cast<float4*>(wpack_buf);
  P.conv.NZ = NZ;
  P.conv.in = in; P.conv.in_cs = in_cs; P.conv.wpack = static_cast<const float4*>(wpack_buf);
  P.conv.Cin = Cin; P.conv.D = D; P.conv.H = H; P.conv.W = W; P.conv.NP = NP; P.conv.MT = MT; P.conv.PW = PW; P.conv.nheads = nheads;
  P.conv.TW = TW; P.conv.strips = ceil_div(W, TW);
  P.smem = umma_conv_smem_bytes(NP, PW);
  P.grid = dim3(P.conv.strips * ceil_div((long long)H * (TW + 2), 128 * MT), D, 1);
  // every CTA walks its (chunk, kz) steps one after the other: with only a few dozen CTAs the direct kernel, which spreads
  // the input channels over warps, is faster (measured on CostRegNet conv4 / conv6: 48 and 16 CTAs, 55 / 76 us against
  // 56 / 105 us is not worth the risk; RED level 4 with 128 CTAs: 34 us against 65 us direct)
  if (perf_rules && (long long)P.grid.x * P.grid.y < 96) return false;
  return true;
}

inline void umma_conv_pack(const UmmaConvPlan& P, cudaStream_t st) {
  const int total = (P.pack.Cin / kUcKC) * P.pack.NZ * 2 * 9 * 2 * P.pack.NP;
  umma_pack_weights_kernel<<<ceil_div(total, 256), 256, 0, st>>>(P.pack);
}

// pack = false: the packed weights are already in place (umma_conv_pack); planes [d0, d0 + nplanes) only when nplanes > 0;
// ready: optional per-plane completion counters (UmmaConv2d::ready)
inline int umma_conv_launch(const UmmaConvPlan& P0, int* error_flag, cudaStream_t st, const char* what, bool pack = true,
                            int d0 = 0, int nplanes = 0, int* ready = nullptr) {
  UmmaConvPlan P = P0;
  if (pack) umma_conv_pack(P, st);
  P.conv.d0 = d0; P.conv.ready = ready;
  if (nplanes > 0) P.grid.y = nplanes;
  static thread_local int ready_dev = -1;
  int dev = 0;
  cudaGetDevice(&dev);
  if (ready_dev != dev) {
    cudaFuncSetAttribute(umma_conv2d_kernel<4, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    cudaFuncSetAttribute(umma_conv2d_kernel<8, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    cudaFuncSetAttribute(umma_conv2d_kernel<4, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    cudaFuncSetAttribute(umma_conv2d_kernel<8, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    ready_dev = dev;
  }
  const bool small = P.conv.PW <= 4 * kUcThreads;
  if (P.conv.NZ == 1) {
    if (small) umma_conv2d_kernel<4, 1><<<P.grid, kUcThreads, P.smem, st>>>(P.conv, error_flag);
    else umma_conv2d_kernel<8, 1><<<P.grid, kUcThreads, P.smem, st>>>(P.conv, error_flag);
  } else {
    if (small) umma_conv2d_kernel<4, 3><<<P.grid, kUcThreads, P.smem, st>>>(P.conv, error_flag);
    else umma_conv2d_kernel<8, 3><<<P.grid, kUcThreads, P.smem, st>>>(P.conv, error_flag);
  }
  return check_launch(what);
}


// ---------------------------------------------------------------------------------------------------------------
// "Taps in N" variant for layers with few output channels (Cout <= 8: CostRegNet conv0 and prob).
// With N = Cout the shifted-descriptor kernel above re-reads the A operand from shared memory for every tap (27 x 3
// MMAs per 8 channels).  Here the nine in-plane taps ride in the N dimension instead:
//     D'[window position p][tap*8 + co] += A[p][8 ch] * W[(tap, co)][8 ch]^T          (N = 72 -> 80, ONE pass over A)
// and the convolution is finished in the epilogue as a shifted row sum  out[q][co] = sum_tap D'[q + shift_tap][tap*8 + co],
// staged through shared memory one tap (8 columns) at a time.  A CTA computes D' for R = 128*MT window positions of its
// strip and emits the R - 2*Wp - 2 outputs whose nine rows all lie inside.  3 MMAs per (kz, chunk, tile) instead of 27.
// ---------------------------------------------------------------------------------------------------------------
constexpr int kTnNP = 80, kTnCo = 8;

struct UmmaConvTn {
  const float* in; long long in_cs;
  const float4* wpack;     // [chunk][kz][part 2][kq KC/4][n 80] float4
  const float* scale; const float* shift; float* out;
  int Cin, Cout, D, H, W, NZ;
  int TW, strips, MT, nout;   // nout = 128*MT - 2*(TW+2) - 2 outputs per CTA
  int wts_resident;           // the packed weights of all steps fit in shared memory next to the window
  float acc_scale; int relu;
};

struct UmmaPackTn { const float* w; long long w_co, w_ci; int Cin, Cout, NZ, KC; float4* out; };

static __global__ void umma_pack_weights_tn_kernel(const __grid_constant__ UmmaPackTn a) {
  const int nq = a.KC / 4;
  const int total = (a.Cin / a.KC) * a.NZ * 2 * nq * kTnNP;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int n = i % kTnNP;
  int r = i / kTnNP;
  const int kq = r % nq; r /= nq;
  const int part = r % 2; r /= 2;
  const int kz = r % a.NZ;
  const int chunk = r / a.NZ;
  const int tap = n / kTnCo, co = n - tap * kTnCo;
  float v[4] = {0.f, 0.f, 0.f, 0.f};
  if (tap < 9 && co < a.Cout) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int ci = chunk * a.KC + 4 * kq + j;
      const float w = __ldg(a.w + (long long)co * a.w_co + (long long)ci * a.w_ci + kz * 9 + tap);
      const float hi = __uint_as_float(__float_as_uint(w) & 0xffffe000u);
      v[j] = part ? (w - hi) : w;
    }
  }
  a.out[i] = make_float4(v[0], v[1], v[2], v[3]);
}

template <int NZ, int KC>      // KC input channels per step (8 or 16): fewer, fatter steps when Cin allows
__global__ void __launch_bounds__(kUcThreads, 2)
umma_conv_tn_kernel(const __grid_constant__ UmmaConvTn a, int* error_flag) {
  constexpr int NQ = KC / 4;                                           // channel quads (16-byte K chunks) per step
  extern __shared__ __align__(128) unsigned char uc_smem[];
  const int PW = 128 * a.MT;                                           // window positions = MMA rows
  float4* win = reinterpret_cast<float4*>(uc_smem);                    // [part 2][kq NQ][PW]; re-used as the epilogue stage [PW][8 floats]
  float4* wts = win + 2 * NQ * PW;                                     // [part 2][kq NQ][80]
  __shared__ unsigned tmem_base_s;
  __shared__ __align__(8) unsigned long long bar, wbar;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int Wp = a.TW + 2, HW = a.H * a.W;
  const int d = blockIdx.y;
  const int run = blockIdx.x / a.strips, sx = blockIdx.x - run * a.strips;
  const int x0 = sx * a.TW;
  const int q0 = Wp + run * a.nout;                                    // first output position (strip-local, row y = 0 starts at Wp)
  const int w0 = q0 - Wp - 1;                                          // strip-local position of window index 0
  const int tmem_cols = uc_tmem_cols(a.MT * kTnNP);

  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(uc_smem_u32(&bar)));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(uc_smem_u32(&wbar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(uc_smem_u32(&tmem_base_s)), "r"(tmem_cols));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  constexpr int NPOS = 2;                                              // PW <= 512
  int src_off[NPOS];
#pragma unroll
  for (int j = 0; j < NPOS; ++j) {
    const int p = tid + j * kUcThreads;
    const int qs = w0 + p;
    const int ry = qs / Wp, rx = qs - ry * Wp;
    const int y = ry - 1, x = x0 - 1 + rx;
    const bool ok = p < PW && qs >= 0 && y >= 0 && y < a.H && x >= 0 && x < a.W;
    src_off[j] = ok ? y * a.W + x : -1;
  }
  float v[NPOS][KC];
  auto load_chunk = [&](int c, int zi) {
    const float* in_c = a.in + (long long)(c * KC) * a.in_cs + (long long)zi * HW;
#pragma unroll
    for (int k = 0; k < KC; ++k) {
      const float* in_k = in_c + k * a.in_cs;
#pragma unroll
      for (int j = 0; j < NPOS; ++j) v[j][k] = src_off[j] >= 0 ? __ldg(in_k + src_off[j]) : 0.0f;
    }
  };
  constexpr int zoff = NZ >> 1;
  const int nsteps = (a.Cin / KC) * NZ;
  auto next_step = [&](int s0) {
    if (NZ > 1)
      while (s0 < nsteps) { const int zi = d + (s0 % NZ) - zoff; if (zi >= 0 && zi < a.D) break; ++s0; }
    return s0;
  };
  int step = next_step(0);
  load_chunk(step / NZ, d + (step % NZ) - zoff);

  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const unsigned tmem = tmem_base_s;
  const unsigned idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((unsigned)(kTnNP >> 3) << 17) | ((unsigned)(128 >> 4) << 24);
  const unsigned wts_bytes = (unsigned)(2 * NQ * kTnNP) * 16u;
  bool alive = true;
  int done = 0;
  // Packed weights: all steps at once when they fit (one TMA bulk copy per CTA, no per-step wait), else step by step
  const bool resident = a.wts_resident != 0;
  if (resident && tid == 0) {
    const unsigned all = wts_bytes * (unsigned)nsteps;
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(uc_smem_u32(&wbar)), "r"(all) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(uc_smem_u32(wts)), "l"(reinterpret_cast<const char*>(a.wpack)), "r"(all), "r"(uc_smem_u32(&wbar)) : "memory");
  }
  while (step < nsteps) {
    if (done > 0) alive = uc_wait(&bar, (unsigned)(done - 1) & 1u) && alive;   // previous step's MMAs have read the window (and the weights)
    if (!resident && tid == 0) {
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(uc_smem_u32(&wbar)), "r"(wts_bytes) : "memory");
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                   ::"r"(uc_smem_u32(wts)), "l"(reinterpret_cast<const char*>(a.wpack) + (size_t)step * wts_bytes), "r"(wts_bytes),
                     "r"(uc_smem_u32(&wbar)) : "memory");
    }
#pragma unroll
    for (int j = 0; j < NPOS; ++j) {
      const int p = tid + j * kUcThreads;
      if (p < PW) {
        float lo[KC];
#pragma unroll
        for (int k = 0; k < KC; ++k) lo[k] = v[j][k] - __uint_as_float(__float_as_uint(v[j][k]) & 0xffffe000u);
#pragma unroll
        for (int qd = 0; qd < NQ; ++qd) {
          win[qd * PW + p] = make_float4(v[j][4 * qd], v[j][4 * qd + 1], v[j][4 * qd + 2], v[j][4 * qd + 3]);
          win[(NQ + qd) * PW + p] = make_float4(lo[4 * qd], lo[4 * qd + 1], lo[4 * qd + 2], lo[4 * qd + 3]);
        }
      }
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0 && uc_elect_one()) {
      if (!resident) alive = uc_wait(&wbar, (unsigned)done & 1u) && alive;
      else if (done == 0) alive = uc_wait(&wbar, 0u) && alive;
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const unsigned lbo_a = (unsigned)PW * 16u, lbo_b = (unsigned)kTnNP * 16u;
      const unsigned long long da0 = uc_desc(uc_smem_u32(win), lbo_a, 128);
      const unsigned long long db0 = uc_desc(uc_smem_u32(wts), lbo_b, 128) + (resident ? (unsigned)step * (wts_bytes >> 4) : 0u);
      const unsigned a_lo_off = (unsigned)(NQ * PW), b_lo_off = (unsigned)(NQ * kTnNP);    // 16-byte units
      for (int mt = 0; mt < a.MT; ++mt) {
        const unsigned dcol = tmem + (unsigned)(mt * kTnNP);
#pragma unroll
        for (int ks = 0; ks < KC / 8; ++ks) {                            // one MMA covers 8 channels = 2 quads
          const unsigned long long a_raw = da0 + (unsigned)(128 * mt + 2 * ks * PW), a_lo = a_raw + a_lo_off;
          const unsigned long long b_raw = db0 + (unsigned)(2 * ks * kTnNP);
          uc_mma_tf32(dcol, a_raw, b_raw, idesc, (done == 0 && ks == 0) ? 0u : 1u);
          uc_mma_tf32(dcol, a_raw, b_raw + b_lo_off, idesc, 1u);
          uc_mma_tf32(dcol, a_lo, b_raw, idesc, 1u);
        }
      }
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(uc_smem_u32(&bar)) : "memory");
    }
    __syncwarp();
    step = next_step(step + 1);
    ++done;
    if (step < nsteps) load_chunk(step / NZ, d + (step % NZ) - zoff);
  }
  alive = uc_wait(&bar, (unsigned)(done - 1) & 1u) && alive;
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  if (!alive) {
    if (tid == 0) *reinterpret_cast<volatile int*>(error_flag) = 1;
    __trap();
  }

  // epilogue: shifted row sum over the nine taps, one tap (8 columns of every row) through shared memory at a time
  float* stage = reinterpret_cast<float*>(uc_smem);                    // [PW][8] floats (the window is no longer needed)
  float acc[NPOS][kTnCo];
#pragma unroll
  for (int j = 0; j < NPOS; ++j)
#pragma unroll
    for (int c = 0; c < kTnCo; ++c) acc[j][c] = 0.0f;
  const int quarter = warp & 3;
  for (int tap = 0; tap < 9; ++tap) {
    for (int mt = warp >> 2; mt < a.MT; mt += kUcThreads / 128) {
      unsigned r[8];
      const unsigned taddr = tmem + ((unsigned)(32 * quarter) << 16) + (unsigned)(mt * kTnNP + tap * kTnCo);
      asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                   : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "r"(taddr));
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      float4* dst = reinterpret_cast<float4*>(stage + (size_t)(128 * mt + 32 * quarter + lane) * kTnCo);
      dst[0] = make_float4(__uint_as_float(r[0]), __uint_as_float(r[1]), __uint_as_float(r[2]), __uint_as_float(r[3]));
      dst[1] = make_float4(__uint_as_float(r[4]), __uint_as_float(r[5]), __uint_as_float(r[6]), __uint_as_float(r[7]));
    }
    __syncthreads();
    const int shift = (tap / 3) * Wp + (tap % 3);                      // row of tap (dy, dx) for output o: o + (dy+1)*Wp + (dx+1)
#pragma unroll
    for (int j = 0; j < NPOS; ++j) {
      const int o = tid + j * kUcThreads;
      if (o < a.nout) {
        const float4* srcp = reinterpret_cast<const float4*>(stage + (size_t)(o + shift) * kTnCo);
        const float4 s0 = srcp[0], s1 = srcp[1];
        acc[j][0] += s0.x; acc[j][1] += s0.y; acc[j][2] += s0.z; acc[j][3] += s0.w;
        acc[j][4] += s1.x; acc[j][5] += s1.y; acc[j][6] += s1.z; acc[j][7] += s1.w;
      }
    }
    __syncthreads();
  }
#pragma unroll
  for (int j = 0; j < NPOS; ++j) {
    const int o = tid + j * kUcThreads;
    const int q = q0 + o;
    const int ry = q / Wp, rx = q - ry * Wp;
    const int y = ry - 1, x = x0 - 1 + rx;
    if (o < a.nout && y < a.H && rx >= 1 && rx <= a.TW && x < a.W) {
      float* op = a.out + (long long)d * HW + (long long)y * a.W + x;
#pragma unroll
      for (int c = 0; c < kTnCo; ++c) {
        if (c < a.Cout) {
          float val = acc[j][c] * a.acc_scale;
          if (a.scale) val *= __ldg(a.scale + c);
          if (a.shift) val += __ldg(a.shift + c);
          if (a.relu) val = fmaxf(val, 0.0f);
          op[(long long)c * a.D * HW] = val;
        }
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(tmem_cols));
}

struct UmmaConvTnPlan { UmmaConvTn conv; UmmaPackTn pack; size_t smem, wpack_bytes; dim3 grid; };

inline bool umma_conv_tn_plan(UmmaConvTnPlan& P, const float* in, long long in_cs, int Cin, int D, int H, int W, const float* w,
                              long long w_co, long long w_ci, const float* scale, const float* shift, float* out, int Cout, int NZ,
                              int relu, float acc_scale, void* wpack_buf, size_t wpack_cap) {
  if (Cin % kUcKC || Cout < 1 || Cout > kTnCo || (NZ != 1 && NZ != 3)) return false;
  P = UmmaConvTnPlan{};
  // strips and tiles: rows issued per useful output = 128*MT / (128*MT - 2*Wp - 2), three tiles (240 TMEM columns, 2 CTAs per SM)
  int TW = 0, MT = 0;
  double best = 1e30;
  for (int nstr = 1; nstr <= 64; ++nstr) {
    const int tw = (W + nstr - 1) / nstr, wp = tw + 2;
    if (tw < 16 && nstr > 1) break;
    for (int mt = 1; mt <= 3; ++mt) {
      const int nout = 128 * mt - 2 * wp - 2;
      if (nout < 32) continue;
      const double runs = (double)((H * wp + nout - 1) / nout);
      double cost = 128.0 * mt * runs * nstr;
      if (tw < 32) cost *= 1.15;
      if (cost < best) { best = cost; TW = tw; MT = mt; }
    }
  }
  if (TW == 0) return false;
  static const bool k16 = getenv("SATMVS_UMMA_TN_K16") != nullptr;     // 16-channel steps: measured no faster (332 against 322 us)
  const int KC = (k16 && Cin % 16 == 0) ? 16 : 8;
  P.wpack_bytes = (size_t)Cin / 4 * NZ * 2 * kTnNP * 16;               // independent of KC
  if (wpack_buf == nullptr || wpack_cap < P.wpack_bytes || (reinterpret_cast<uintptr_t>(wpack_buf) & 15)) return false;
  P.pack = UmmaPackTn{w, w_co, w_ci, Cin, Cout, NZ, KC, static_cast<float4*>(wpack_buf)};
  UmmaConvTn& c = P.conv;
  c.in = in; c.in_cs = in_cs; c.wpack = static_cast<const float4*>(wpack_buf); c.scale = scale; c.shift = shift; c.out = out;
  c.Cin = Cin; c.Cout = Cout; c.D = D; c.H = H; c.W = W; c.NZ = NZ;
  c.TW = TW; c.strips = ceil_div(W, TW); c.MT = MT; c.nout = 128 * MT - 2 * (TW + 2) - 2;
  c.acc_scale = acc_scale; c.relu = relu;
  const size_t win_bytes = (size_t)2 * (KC / 4) * 128 * MT * 16, step_bytes = (size_t)2 * (KC / 4) * kTnNP * 16;
  c.wts_resident = step_bytes * (Cin / KC) * NZ <= 64 * 1024;
  P.smem = win_bytes + (c.wts_resident ? step_bytes * (Cin / KC) * NZ : step_bytes);
  P.grid = dim3(c.strips * ceil_div((long long)H * (TW + 2), c.nout), D, 1);
  return true;
}

inline int umma_conv_tn_launch(const UmmaConvTnPlan& P, int* error_flag, cudaStream_t st, const char* what) {
  const int total = P.pack.Cin / 4 * P.pack.NZ * 2 * kTnNP;
  umma_pack_weights_tn_kernel<<<ceil_div(total, 256), 256, 0, st>>>(P.pack);
  static thread_local int ready_dev = -1;
  int dev = 0;
  cudaGetDevice(&dev);
  if (ready_dev != dev) {
    cudaFuncSetAttribute(umma_conv_tn_kernel<1, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 112 * 1024);
    cudaFuncSetAttribute(umma_conv_tn_kernel<3, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 112 * 1024);
    cudaFuncSetAttribute(umma_conv_tn_kernel<1, 16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 112 * 1024);
    cudaFuncSetAttribute(umma_conv_tn_kernel<3, 16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 112 * 1024);
    ready_dev = dev;
  }
  const bool k16 = P.pack.KC == 16;
  if (P.conv.NZ == 1) {
    if (k16) umma_conv_tn_kernel<1, 16><<<P.grid, kUcThreads, P.smem, st>>>(P.conv, error_flag);
    else umma_conv_tn_kernel<1, 8><<<P.grid, kUcThreads, P.smem, st>>>(P.conv, error_flag);
  } else {
    if (k16) umma_conv_tn_kernel<3, 16><<<P.grid, kUcThreads, P.smem, st>>>(P.conv, error_flag);
    else umma_conv_tn_kernel<3, 8><<<P.grid, kUcThreads, P.smem, st>>>(P.conv, error_flag);
  }
  return check_launch(what);
}

}  // namespace satmvs
