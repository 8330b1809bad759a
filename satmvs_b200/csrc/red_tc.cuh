// red_tc.cuh — the RED depth recurrence (phase B of red.cu) on the 5th-generation tensor cores: ONE launch of four
// 16-CTA thread-block clusters, one cluster per UNet level, hidden state resident in shared memory for the whole sweep.
//
// Reference: the plane loop of RED_Regularization.forward (modules/module.py:625-644) through ConvGRUCell2.forward
// (:27-58).  conv([x, h]) = conv_x(x) + conv_h(h); the x-halves GX / OX (bias included) are batched over all planes by
// red.cu, this kernel does what is sequential in depth, per plane d and level:
//     G = GX[d] + conv(h; Wg_h)           r = sigmoid(GN_r(G_r)), u = sigmoid(GN_u(G_u))          (module.py:29-43)
//     O = OX[d] + conv(r*h; Wo_h)         h' = u*h + (1-u)*tanh(GN_o(O))                          (module.py:44-57)
//
// Partition (the round-1 kernel gave a CTA 8 output channels x a row strip: N = 8, FFMA only).  Here a CTA owns a strip of
// rows x ALL output channels x a K-group of CK = ch/KG input channels:
//     level        0 (8 ch)   1 (16 ch)   2 (32 ch)   3 (64 ch)
//     KG           1          1           2           8            K-groups (shared-memory capacity: 216*ch^2 bytes of filters)
//     strips       16         16          8           2
//     N gates/out  16 / 8     32 / 16     64 / 32     128 / 64
// The strip's state lives in shared memory as [part: raw, lo][channel quad][padded-flattened position] float4 -- the
// SWIZZLE_NONE K-major canonical layout -- so a conv tap is one shared-memory descriptor shifted by dy*Wp + dx
// (umma_conv.cuh).  Precision: the tensor core truncates fp32 to TF32; x*w = raw(x)*raw(w) + raw(x)*lo(w) + lo(x)*raw(w)
// with lo = v - trunc(v).  The first two products share their A operand, so the filter bank is stacked [raw | lo] along N:
// two MMAs per (tap, 8 channels) instead of three -- D[:, 0:N] += A_raw*[W_raw | W_lo], then D[:, 0:N] += A_lo*W_raw --
// and the epilogue adds the two column halves.  Accumulators in TMEM (MT tiles x 2N columns).
//   * K-split levels: every CTA holds partial sums over its CK input channels for all N columns; columns are owned by the
//     CTA whose K-group has the same channel index (so the r*h / h' it produces are exactly the input channels it needs
//     next: no all-gather), partials travel as float4 rows into the owner's shared memory (st.shared::cluster).
//   * halo rows of r*h and h' are written straight into the neighbouring strips' windows through distributed shared memory;
//   * GroupNorm sums cross the cluster through distributed shared memory in rank order (deterministic);
//   * the x-half pre-activations of the strip arrive by cp.async.bulk (TMA engine) one half-plane ahead;
//   * dependencies inside a plane are barrier.cluster pairs: 4 per plane (6 on the K-split levels).
#pragma once
#include "common.cuh"
#include "umma_conv.cuh"

namespace satmvs {

constexpr int kTcEpiWarps = 16, kTcThreads = 32 * kTcEpiWarps, kTcWarps = kTcThreads / 32, kTcCluster = 16;
constexpr int kTcIssuer = kTcEpiWarps - 1;   // warp that issues the MMAs and bulk copies: its tiles (3, 7, 11) are the fewest, and a 17th warp would cap registers at 96
constexpr int kTcKG[4] = {1, 1, 2, 8};        // K-groups per level
constexpr int kTcMaxP[4] = {3, 1, 1, 1};      // output positions per thread (register budget: CK * MAXP <= 24)

struct TcLevel {
  float* s; long long s_cs;                   // state history [ch][D+1][px]: channel stride; slot stride = px
  const float* gx; long long g_cs;            // gate x-halves (+ bias) [2ch][D][px]
  const float* ox; long long o_cs;            // output x-halves (+ bias) [ch][D][px]
  const float4* wpack;                        // packed hidden-state filters, see tc_pack_kernel
  const float *rn_w, *rn_b, *un_w, *un_b, *on_w, *on_b;
  double inv_n;                               // 1 / (ch * px)
  int ch, h, w, px;
  int R, MT, PWa, NPP;                        // rows per strip, 128-position tiles, window positions, positions padded to 32
  const int* ready; int expected;             // optional: ready[d] reaches `expected` when the x-halves of plane d are in memory
};                                            // (their producers run concurrently on the SMs this kernel leaves free)
struct TcArgs { TcLevel l[4]; int d_begin, d_end; int* err; long long* dbg; int level0; };   // planes [d_begin, d_end): state read from history slot d_begin; level0: level of cluster 0 of the grid

struct TcGeom { int R, MT, PWa, NPP; size_t smem; };
// geometry + dynamic shared memory of one level (host and device agree on the carve-up through this function)
__host__ __device__ inline TcGeom tc_geom(int ch, int KG, int h, int w) {
  TcGeom g;
  const int S = kTcCluster / KG, CK = ch / KG, Wp = w + 2;
  g.R = (h + S - 1) / S;
  g.MT = (g.R * Wp + 127) / 128;
  g.PWa = 128 * g.MT + 2 * Wp + 2;
  g.NPP = (g.R * Wp + 31) / 32 * 32;
  const size_t win = (size_t)2 * (CK / 4) * g.PWa * 16;
  const size_t wts = (size_t)9 * (CK / 8) * 2 * (6 * ch) * 16;           // gates [raw | lo] 4ch rows + output 2ch rows
  const size_t pre = (size_t)2 * CK * g.R * w * 4;
  const size_t recv = (size_t)(KG - 1) * (2 * CK / 4) * g.NPP * 16;
  const size_t send = KG == 2 ? recv : 0;                                // K-pair levels stage their outgoing partial sums for one bulk copy
  g.smem = win + wts + pre + recv + send;
  return g;
}

// Packed filters of one level: [kg][conv: gates, output][tap 9][ks CK/8][kq 2][n NB] float4 of 4 consecutive input
// channels (kg*CK + ks*8 + kq*4 + 0..3); rows n < N hold the fp32 weight (the tensor core truncates it), rows N <= n < 2N
// its low part w - trunc_tf32(w).  w[(row) * w_co + ci * 9 + tap] with the pointer already at the hidden-state half.
struct TcPack { const float* gate_w; const float* out_w; long long w_co; int ch, KG; float4* out; };
static __global__ void tc_pack_kernel(const __grid_constant__ TcPack a) {
  const int CK = a.ch / a.KG, KS = CK / 8, NG = 2 * a.ch, NO = a.ch;
  const int per_g = 9 * KS * 2 * 2 * NG, per_o = 9 * KS * 2 * 2 * NO, per_kg = per_g + per_o;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= a.KG * per_kg) return;
  const int kg = i / per_kg;
  int r = i - kg * per_kg;
  const bool is_out = r >= per_g;
  if (is_out) r -= per_g;
  const int N = is_out ? NO : NG, NB = 2 * N;
  const int n = r % NB; r /= NB;
  const int kq = r % 2; r /= 2;
  const int ks = r % KS;
  const int tap = r / KS;
  const bool lo = n >= N;
  const int row = lo ? n - N : n;
  const float* wsrc = is_out ? a.out_w : a.gate_w;
  float v[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int ci = kg * CK + ks * 8 + kq * 4 + j;
    const float wv = __ldg(wsrc + (long long)row * a.w_co + (long long)ci * 9 + tap);
    const float hi = __uint_as_float(__float_as_uint(wv) & 0xffffe000u);
    v[j] = lo ? (wv - hi) : wv;
  }
  a.out[i] = make_float4(v[0], v[1], v[2], v[3]);
}

__device__ __forceinline__ unsigned tc_cluster_ctarank() { unsigned r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void tc_cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ unsigned tc_mapa(unsigned saddr, unsigned rank) {
  unsigned r; asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(saddr), "r"(rank)); return r;
}
__device__ __forceinline__ void tc_st_remote(unsigned raddr, float4 v) {
  asm volatile("st.shared::cluster.v4.f32 [%0], {%1, %2, %3, %4};" :: "r"(raddr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ double tc_ld_remote_f64(unsigned raddr) {
  double v; asm volatile("ld.shared::cluster.f64 %0, [%1];" : "=d"(v) : "r"(raddr) : "memory"); return v;
}
__device__ __forceinline__ void tc_tmem_ld8(unsigned taddr, float (&v)[8]) {
  unsigned r[8];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "r"(taddr));
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ float tc_lo(float v) { return v - __uint_as_float(__float_as_uint(v) & 0xffffe000u); }
__device__ __forceinline__ float tc_sigmoid(float x) { return __fdividef(1.0f, 1.0f + __expf(-x)); }
__device__ __forceinline__ float tc_tanh(float x) { return fmaf(2.0f, __fdividef(1.0f, 1.0f + __expf(-2.0f * x)), -1.0f); }

struct TcShared {
  double stat_in[2][kTcCluster][4];           // (sum, sum^2) x {r, u} after the gate conv / {o} after the output conv, pushed by every rank of the cluster
  double red[4][kTcWarps];
  float coef[3][16][2];                       // GroupNorm scale / shift of this CTA's channels: r, u, o
  unsigned long long mbar_tile[16], mbar_pre;      // MMAs of tile mt complete; x-halves landed
  unsigned long long mbar_halo;                    // the neighbouring strips' halo rows have landed in the window (complete_tx)
  unsigned long long mbar_x;                       // K-pair levels: the partner's partial sums have landed in recv (complete_tx)
  unsigned tmem_base;
  long long t_prev, t_acc[20];                // SATMVS_RED_DEBUG: cycles thread 0 spends per phase slot
};

template <int CH, int KG, int MAXP, int STASH>   // STASH: 0 none, 1 in the dead half of the gate accumulators (KG == 1), 2 own TMEM columns
__device__ __forceinline__ void tc_level_run(const TcLevel& L, const int d_begin, const int D, int* err, long long* dbg, unsigned char* smem, TcShared& sh) {
  constexpr int CK = CH / KG, NQ = CK / 4, KS = CK / 8;
  constexpr int NG = 2 * CH, NBG = 2 * NG, NO = CH, NBO = 2 * NO;
  constexpr int N2O = NO >= 16 ? NO : NBO;                 // an M = 128 MMA needs N >= 16: level 0 multiplies lo(x) by [W | lo(W)] (the extra lo*lo term is exact anyway)
  static_assert(CK % 8 == 0 && CK <= 16 && CK * MAXP <= 24 && (STASH != 1 || KG == 1), "register budget");
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const unsigned rank = tc_cluster_ctarank();
  const int strip = (int)rank / KG, kg = (int)rank % KG;
  const int w = L.w, Wp = w + 2, R = L.R, PWa = L.PWa, NPP = L.NPP, Rw = R * w;
  const int y0 = strip * R;
  int nrows = L.h - y0; nrows = nrows > R ? R : nrows; nrows = nrows < 0 ? 0 : nrows;
  const int npos = nrows * Wp;
  const int MTa = (npos + 127) / 128;                      // tiles that hold an output position of this strip

  float4* win = reinterpret_cast<float4*>(smem);           // [part 2][quad NQ][PWa]
  float4* wg = win + 2 * NQ * PWa;                         // [tap 9][ks][kq 2][NBG]
  float4* wo = wg + 9 * KS * 2 * NBG;                      // [tap 9][ks][kq 2][NBO]
  float* pre = reinterpret_cast<float*>(wo + 9 * KS * 2 * NBO);   // [2*CK][R*w]: x-halves of the conv whose epilogue comes next
  float4* recv = reinterpret_cast<float4*>(pre + 2 * CK * Rw);    // [src KG-1][f4 2*CK/4][NPP]: partial sums pushed by the other K-groups
  constexpr bool BULK = (KG == 2);                                // one partner: its columns are staged locally and travel as ONE
  float4* send = recv + (KG - 1) * (2 * CK / 4) * NPP;            // cp.async.bulk into its recv (DSMEM through the TMA engine)

  // ---- once: barriers, TMEM, zero window, filters, initial state ----
  constexpr int kTmemCols = MAXP * 4 * NBG <= 512 ? 512 : 512;
  if (tid == 0) {
    for (int i = 0; i < 16; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(uc_smem_u32(&sh.mbar_tile[i])));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(uc_smem_u32(&sh.mbar_pre)));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(uc_smem_u32(&sh.mbar_x)));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(uc_smem_u32(&sh.mbar_halo)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(uc_smem_u32(&sh.tmem_base)), "r"(kTmemCols));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  for (int i = tid; i < 2 * NQ * PWa; i += kTcThreads) win[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  {
    constexpr int per_kg = 9 * KS * 2 * (NBG + NBO);
    const float4* src = L.wpack + (size_t)kg * per_kg;
    for (int i = tid; i < per_kg; i += kTcThreads) wg[i] = __ldg(src + i);
  }
  __syncthreads();
  if (nrows > 0) {   // h[0]: rows y0-1 .. y0+nrows of this CTA's channels (slot 0 of the state history)
    const int ylo = y0 > 0 ? y0 - 1 : 0, yhi = (y0 + nrows < L.h) ? y0 + nrows : L.h - 1;
    const int items = NQ * (yhi - ylo + 1) * w;
    for (int i = tid; i < items; i += kTcThreads) {
      const int x = i % w; int r = i / w;
      const int yy = ylo + r % (yhi - ylo + 1); const int qd = r / (yhi - ylo + 1);
      const float* sp = L.s + (long long)(kg * CK + 4 * qd) * L.s_cs + (long long)d_begin * L.px + (long long)yy * w + x;
      const float4 v = make_float4(sp[0], sp[L.s_cs], sp[2 * L.s_cs], sp[3 * L.s_cs]);
      const int wi = (yy - y0 + 1) * Wp + x + 1;
      win[qd * PWa + wi] = v;
      win[(NQ + qd) * PWa + wi] = make_float4(tc_lo(v.x), tc_lo(v.y), tc_lo(v.z), tc_lo(v.w));
    }
  }
  // x-halves of this strip by bulk copies: conv 0 = gates (r and u rows of this CTA's channels), conv 1 = output
  auto issue_pre = [&](int conv, int d) {
    if (warp != kTcIssuer) return;                                        // one channel per lane of the issuer warp
    const unsigned bar = uc_smem_u32(&sh.mbar_pre);
    const int nchan = conv == 0 ? 2 * CK : CK;
    const unsigned bytes = (unsigned)(nrows * w) * 4u;
    if (lane == 0) {
      if (L.ready != nullptr && nrows > 0) {     // the producers of this plane's x-halves may still be running
        const long long t0 = clock64();
        int v;
        do {
          asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(L.ready + d) : "memory");
          if (v < L.expected && clock64() - t0 > (1LL << 32)) { *reinterpret_cast<volatile int*>(err) = 2; break; }      // ~2 s: give up loudly, never hang
        } while (v < L.expected);
        asm volatile("fence.proxy.async;" ::: "memory");                 // acquired by the generic proxy, read by the TMA engine
      }
      if (nrows == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
      else asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes * (unsigned)nchan) : "memory");
    }
    __syncwarp();
    if (nrows > 0 && lane < nchan) {
      const int i = lane;
      const float* src = conv == 0
          ? L.gx + (long long)((i < CK ? 0 : CH) + kg * CK + (i < CK ? i : i - CK)) * L.g_cs + (long long)d * L.px + (long long)y0 * w
          : L.ox + (long long)(kg * CK + i) * L.o_cs + (long long)d * L.px + (long long)y0 * w;
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                   ::"r"(uc_smem_u32(pre + (size_t)i * Rw)), "l"(src), "r"(bytes), "r"(bar) : "memory");
    }
    __syncwarp();
  };
  issue_pre(0, d_begin);
  asm volatile("fence.proxy.async;" ::: "memory");          // generic-proxy stores (window, filters) -> tensor-core reads
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  tc_cluster_sync();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const unsigned tmem = sh.tmem_base;

  // ---- this thread's output positions: TMEM lane = 32*(warp%4) + lane of tile mt = warp/4 + 4j ----
  // fl_: bit 0 = a pixel of the strip, bit 1 = also the bottom halo (row R+1) of the strip above, bit 2 = also the top halo
  // (row 0) of the strip below
  int q_[MAXP], sp_[MAXP], fl_[MAXP];
#pragma unroll
  for (int j = 0; j < MAXP; ++j) {
    const int mt = (warp >> 2) + 4 * j;
    const int q = mt * 128 + (warp & 3) * 32 + lane;
    const int ly = q / Wp, x = q - ly * Wp;
    const bool ok = warp < kTcEpiWarps && q < npos && x < w;
    q_[j] = q; sp_[j] = ly * w + x;
    fl_[j] = ok ? 1 : 0;
    if (ok && ly == 0 && y0 > 0) fl_[j] |= 2;
    if (ok && ly == nrows - 1 && y0 + nrows < L.h) fl_[j] |= 4;
  }
  const unsigned win_s = uc_smem_u32(win), recv_s = uc_smem_u32(recv);
  const int ocol0 = STASH ? L.MT * NBG : 0;                                // first TMEM column of the output-conv accumulators
  const int scol0 = L.MT * (NBG + NBO);                                    // STASH == 2: first column of the parked u / h values

  auto idesc_of = [](int N) { return (1u << 4) | (2u << 7) | (2u << 10) | ((unsigned)(N >> 3) << 17) | ((unsigned)(128 >> 4) << 24); };
  // MMAs of tiles [t_lo, t_hi) of one convolution, issued by one elected lane of the issuer warp; every tile commits to
  // its own mbarrier so that its read-back overlaps the MMAs of the tiles behind it
  auto issue_tiles = [&](const float4* wts, int NB, int N2, int col0, int t_lo, int t_hi) {
    if (warp == kTcIssuer) {                                              // tcgen05.mma issue blocks while the queue is full: this warp's own read-back comes last
      if (uc_elect_one() && t_lo < t_hi) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const unsigned id1 = idesc_of(NB), id2 = idesc_of(N2);
        const unsigned long long da0 = uc_desc(win_s, (unsigned)PWa * 16u, 128), db0 = uc_desc(uc_smem_u32(wts), (unsigned)NB * 16u, 128);
        const unsigned a_lo_off = (unsigned)(NQ * PWa);                  // 16-byte units
        for (int mt = t_lo; mt < t_hi; ++mt) {
          const unsigned dcol = tmem + (unsigned)(col0 + mt * NB);
#pragma unroll
          for (int tap = 0; tap < 9; ++tap) {
#pragma unroll
            for (int ks = 0; ks < KS; ++ks) {
              const unsigned shift = (unsigned)(128 * mt + (tap / 3) * Wp + (tap % 3) + 2 * ks * PWa);
              const unsigned long long bb = db0 + (unsigned)((tap * KS + ks) * 2 * NB);
              uc_mma_tf32(dcol, da0 + shift, bb, id1, (tap == 0 && ks == 0) ? 0u : 1u);
              uc_mma_tf32(dcol, da0 + shift + a_lo_off, bb, id2, 1u);
            }
          }
          asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(uc_smem_u32(&sh.mbar_tile[mt])) : "memory");
        }
      }
      __syncwarp();
    }
  };
  bool alive = true;
  // Halo rows travel as bulk copies (TMA engine, shared::cta -> shared::cluster): after a conv input (r*h or h') is in this
  // CTA's window, the issuer warp sends the first / last own row of every (part, channel quad) -- Wp contiguous float4 each --
  // into the halo row of the strip above / below, completing on THAT CTA's mbarrier; no thread issues remote stores and no
  // cluster barrier is needed for the halos.  Tiles that read a halo row (outputs of the first / last row of the strip) wait
  // for the incoming bytes; the tiles between them only need this CTA's own rows and are issued first.
  // Hazards: a neighbour overwrites my halo rows only after the stats barrier that follows my MMAs over them (#B / #E); my
  // own rows are overwritten only after the next stats barrier, which the neighbours reach after consuming my copy.
  const bool nb_up = nrows > 0 && y0 > 0, nb_dn = nrows > 0 && y0 + nrows < L.h;
  const unsigned halo_row_bytes = (unsigned)Wp * 16u;
  const unsigned halo_in_bytes = (unsigned)((nb_up ? 1 : 0) + (nb_dn ? 1 : 0)) * 2u * NQ * halo_row_bytes;
  unsigned ph_halo = 0;
  const int t_first = (Wp - 1) / 128 + 1, t_last = nrows >= 3 ? ((nrows - 1) * Wp) / 128 : 0;     // interior tiles [t_first, t_last)
  auto issue_conv_split = [&](const float4* wts, int NB, int N2, int col0) {
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();                                                      // this CTA's rows are in the window (every writer fenced the async proxy)
    if (warp == kTcIssuer && halo_in_bytes > 0) {
      const unsigned bar = uc_smem_u32(&sh.mbar_halo);
      if (lane == 0) asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(halo_in_bytes) : "memory");
      // lane = dir * 2*NQ + part * NQ + quad
      const int dir = lane / (2 * NQ), pq = lane - dir * (2 * NQ);
      if (dir < 2 && ((dir == 0 && nb_up) || (dir == 1 && nb_dn))) {
        const int src_row = dir == 0 ? 1 : nrows, dst_row = dir == 0 ? R + 1 : 0;
        const unsigned nb = dir == 0 ? rank - KG : rank + KG;
        const unsigned src = win_s + (unsigned)((pq * PWa + src_row * Wp) * 16);
        const unsigned dst = tc_mapa(win_s + (unsigned)((pq * PWa + dst_row * Wp) * 16), nb);
        asm volatile("cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(dst), "r"(src), "r"(halo_row_bytes), "r"(tc_mapa(bar, nb)) : "memory");
      }
      __syncwarp();
    }
    if (t_first < t_last) issue_tiles(wts, NB, N2, col0, t_first, t_last);
    if (warp == kTcIssuer && halo_in_bytes > 0) { alive = uc_wait(&sh.mbar_halo, ph_halo) && alive; ph_halo ^= 1u; }
    if (t_first < t_last) { issue_tiles(wts, NB, N2, col0, 0, t_first); issue_tiles(wts, NB, N2, col0, t_last, MTa); }
    else issue_tiles(wts, NB, N2, col0, 0, MTa);
  };

  unsigned ph_mma = 0, ph_pre = 0;
  auto lane_base = [&](int j, int NB, int col0) {                         // TMEM address of this thread's row of tile j's accumulators
    return tmem + ((unsigned)(32 * (warp & 3)) << 16) + (unsigned)(col0 + ((warp >> 2) + 4 * j) * NB);
  };
  // accumulator read-back of one position: v[part*CK + c] = this CTA's own columns (raw + lo halves added); on the K-split
  // levels the columns owned by the other K-groups are pushed into their receive buffers.  Warp-uniform call.
  auto read_pos = [&](int j, int N, int NB, int col0, int nparts, float (&v)[2 * CK]) {
    alive = uc_wait(&sh.mbar_tile[(warp >> 2) + 4 * j], ph_mma) && alive;
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const unsigned tbase = lane_base(j, NB, col0);
#pragma unroll
    for (int part = 0; part < 2; ++part) {                                 // gates: r columns, u columns
      if (part >= nparts) break;
#pragma unroll
      for (int owner = 0; owner < KG; ++owner) {
#pragma unroll
        for (int c8 = 0; c8 < CK / 8; ++c8) {
          const int col = part * CH + owner * CK + c8 * 8;
          float a[8], b[8];
          tc_tmem_ld8(tbase + (unsigned)col, a);
          tc_tmem_ld8(tbase + (unsigned)(N + col), b);
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
          for (int i = 0; i < 8; ++i) a[i] += b[i];
          if (KG == 1 || owner == kg) {
#pragma unroll
            for (int i = 0; i < 8; ++i) v[part * CK + c8 * 8 + i] = a[i];
          } else if (BULK) {
            const int f4 = part * (CK / 4) + c8 * 2;                       // staged locally (the whole [f4][NPP] block travels)
            if (q_[j] < NPP) {
              send[f4 * NPP + q_[j]] = make_float4(a[0], a[1], a[2], a[3]);
              send[(f4 + 1) * NPP + q_[j]] = make_float4(a[4], a[5], a[6], a[7]);
            }
          } else if (fl_[j] & 1) {
            const int slot = kg < owner ? kg : kg - 1;
            const int f4 = part * (CK / 4) + c8 * 2;
            const unsigned la = recv_s + (unsigned)(((slot * (2 * CK / 4) + f4) * NPP + q_[j]) * 16);
            const unsigned ra = tc_mapa(la, (unsigned)(strip * KG + owner));
            tc_st_remote(ra, make_float4(a[0], a[1], a[2], a[3]));
            tc_st_remote(ra + (unsigned)NPP * 16u, make_float4(a[4], a[5], a[6], a[7]));
          }
        }
      }
    }
  };
  // own partial + the partials received from the other K-groups (fixed order) + the x-half
  auto add_recv_pre = [&](int j, int nparts, float (&v)[2 * CK]) {
#pragma unroll
    for (int part = 0; part < 2; ++part) {
      if (part >= nparts) break;
#pragma unroll
      for (int f = 0; f < CK / 4; ++f) {
        float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
        if (KG > 1) {
          for (int slot = 0; slot < KG - 1; ++slot) {
            const float4 r4 = recv[(slot * (2 * CK / 4) + part * (CK / 4) + f) * NPP + q_[j]];
            t.x += r4.x; t.y += r4.y; t.z += r4.z; t.w += r4.w;
          }
        }
        const float* pp = pre + (size_t)(part * CK + 4 * f) * Rw + sp_[j];
        v[part * CK + 4 * f + 0] += t.x + pp[0]; v[part * CK + 4 * f + 1] += t.y + pp[Rw];
        v[part * CK + 4 * f + 2] += t.z + pp[2 * Rw]; v[part * CK + 4 * f + 3] += t.w + pp[3 * Rw];
      }
    }
  };
  // TMEM as a register-file extension (STASH levels): CK values of this thread's row at column `col` of tile j's gate region
  auto stash_addr = [&](int j, int which) {                               // which: 0 = update-gate pre-activation, 1 = h
    return STASH == 1 ? lane_base(j, NBG, 0) + (unsigned)((1 + which) * CK)
                      : lane_base(j, 2 * CK, scol0) + (unsigned)(which * CK);
  };
  auto stash_st = [&](int j, int which, const float* v) {
    const unsigned ta = stash_addr(j, which);
#pragma unroll
    for (int c8 = 0; c8 < CK / 8; ++c8)
      asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
                   :: "r"(ta + (unsigned)(8 * c8)), "r"(__float_as_uint(v[8 * c8])), "r"(__float_as_uint(v[8 * c8 + 1])),
                      "r"(__float_as_uint(v[8 * c8 + 2])), "r"(__float_as_uint(v[8 * c8 + 3])), "r"(__float_as_uint(v[8 * c8 + 4])),
                      "r"(__float_as_uint(v[8 * c8 + 5])), "r"(__float_as_uint(v[8 * c8 + 6])), "r"(__float_as_uint(v[8 * c8 + 7])) : "memory");
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  };
  auto stash_ld = [&](int j, int which, float* v) {
    const unsigned ta = stash_addr(j, which);
#pragma unroll
    for (int c8 = 0; c8 < CK / 8; ++c8) {
      float t[8];
      tc_tmem_ld8(ta + (unsigned)(8 * c8), t);
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
      for (int i = 0; i < 8; ++i) v[8 * c8 + i] = t[i];
    }
  };
  // K-split exchange.  The wait behind per-thread st.shared::cluster pushes is the drain of those stores (~4 k cycles for 19 KB at
  // level 2, whatever the barrier: a K-group-only mbarrier handshake was measured no faster, profiles/r02_red_overlap_notes.md).
  // K-pair levels therefore stage their partner's columns locally and send them as ONE cp.async.bulk shared::cta ->
  // shared::cluster (TMA engine) that completes on the partner's mbarrier; the staging buffer is free again once the partner has
  // passed the next cluster barrier (it waited for these bytes before arriving there).
  unsigned ph_x = 0;
  auto exchange = [&](int nparts) {
    if (!BULK) { tc_cluster_sync(); return; }
    const unsigned bytes = (unsigned)(nparts * (CK / 4) * NPP) * 16u;
    asm volatile("fence.proxy.async;" ::: "memory");                     // staged by the generic proxy, read by the TMA engine
    __syncthreads();
    if (tid == 0) {
      const unsigned bar = uc_smem_u32(&sh.mbar_x), peer = (unsigned)(strip * KG + (kg ^ 1));
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");   // what I will receive
      asm volatile("cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                   ::"r"(tc_mapa(recv_s, peer)), "r"(uc_smem_u32(send)), "r"(bytes), "r"(tc_mapa(bar, peer)) : "memory");
    }
    alive = uc_wait(&sh.mbar_x, ph_x) && alive; ph_x ^= 1u;
  };
  // block-wide sums of four quantities, pushed into stat_in[which][my rank] of every CTA of the cluster (the cluster barrier
  // that follows publishes them): no remote load sits behind the barrier
  auto publish_stats = [&](int which, float (&v)[4]) {
    // a warp holds <= 96 positions x CK channels: fp32 is exact enough for their sum (the reference accumulates GroupNorm
    // statistics in fp32 throughout); across warps and CTAs the sums continue in fp64
#pragma unroll
    for (int k = 0; k < 4; ++k)
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) v[k] += __shfl_xor_sync(0xffffffffu, v[k], off);
    if (lane == 0)
#pragma unroll
      for (int k = 0; k < 4; ++k) sh.red[k][warp] = (double)v[k];
    __syncthreads();
    if (warp == 0) {
      double t[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) t[k] = lane < kTcEpiWarps ? sh.red[k][lane] : 0.0;
#pragma unroll
      for (int k = 0; k < 4; ++k)
#pragma unroll
        for (int off = 8; off > 0; off >>= 1) t[k] += __shfl_xor_sync(0xffffffffu, t[k], off);
      if (lane < kTcCluster) {                                             // lane l -> rank l
        const unsigned ra = tc_mapa(uc_smem_u32(&sh.stat_in[which][rank][0]), (unsigned)lane);
        asm volatile("st.shared::cluster.v2.f64 [%0], {%1, %2};" :: "r"(ra), "d"(t[0]), "d"(t[1]) : "memory");
        asm volatile("st.shared::cluster.v2.f64 [%0], {%1, %2};" :: "r"(ra + 16u), "d"(t[2]), "d"(t[3]) : "memory");
      }
    }
  };
  // after the cluster barrier: warp 0 adds the 16 ranks' sums (butterfly in rank order: every CTA gets the same bits) and
  // lanes < CK turn them into scale / shift; norm k uses sums (2k, 2k+1) and coef slot cslot + k
  auto gather_coef = [&](int which, int nnorm, int cslot, const float* w0, const float* b0, const float* w1, const float* b1) {
    if (warp == 0) {
      double t[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) t[k] = sh.stat_in[which][lane & 15][k];
#pragma unroll
      for (int k = 0; k < 4; ++k)
#pragma unroll
        for (int off = 8; off > 0; off >>= 1) t[k] += __shfl_xor_sync(0xffffffffu, t[k], off);
      for (int k = 0; k < nnorm; ++k) {
        if (lane < CK) {
          const double mean = t[2 * k] * L.inv_n;
          const float var = (float)fmax(t[2 * k + 1] * L.inv_n - mean * mean, 0.0);
          const float rstd = rsqrtf(var + 1e-5f);
          const float* gw = k == 0 ? w0 : w1; const float* gb = k == 0 ? b0 : b1;
          const float ca = __ldg(gw + kg * CK + lane) * rstd;
          sh.coef[cslot + k][lane][0] = ca; sh.coef[cslot + k][lane][1] = __ldg(gb + kg * CK + lane) - (float)mean * ca;
        }
      }
    }
    __syncthreads();
  };
  // write CK channels of one position into the window (raw + lo); the halo rows leave as bulk copies (issue_conv_split)
  auto store_window = [&](int j, const float (&v)[CK]) {
    const int wi = q_[j] + Wp + 1;
#pragma unroll
    for (int f = 0; f < NQ; ++f) {
      const float4 raw = make_float4(v[4 * f], v[4 * f + 1], v[4 * f + 2], v[4 * f + 3]);
      const float4 lo = make_float4(tc_lo(raw.x), tc_lo(raw.y), tc_lo(raw.z), tc_lo(raw.w));
      win[f * PWa + wi] = raw;
      win[(NQ + f) * PWa + wi] = lo;
    }
  };

  if (tid == 0) for (int i = 0; i < 20; ++i) sh.t_acc[i] = 0;
  auto mark = [&](int slot) {
    if (dbg != nullptr && tid == 0) { const long long t = clock64(); if (slot >= 0) sh.t_acc[slot] += t - sh.t_prev; sh.t_prev = t; }
  };

  // Per-position state that lives across the cluster barriers: keep = reset-gate pre-activation, later the output-conv
  // pre-activation, later h'.  The update-gate pre-activation and h of the own channels are parked in TMEM on the STASH
  // levels (columns [CK, 2CK) and [2CK, 3CK) of the position's gate accumulators, dead after the read-back).
  float keep[MAXP][CK];
  float ukeep[MAXP][STASH ? 1 : CK], hkeep[MAXP][STASH ? 1 : CK];
  issue_tiles(wg, NBG, NG, 0, 0, MTa);                                     // gates of plane 0
  for (int d = d_begin; d < D; ++d) {
    mark(-1);
    // ================= gates: G = GX[d] + conv(h; Wg)  (MMAs already in flight) =================
    float st[4] = {0.f, 0.f, 0.f, 0.f};
    {
      float v[MAXP][2 * CK];
      if (KG > 1) {
#pragma unroll
        for (int j = 0; j < MAXP; ++j)
          if ((warp >> 2) + 4 * j < MTa && warp < kTcEpiWarps) read_pos(j, NG, NBG, 0, 2, v[j]);
        mark(12);
        exchange(2);                                                       // #A: every K-group's partial sums have landed
        mark(13);
      }
      alive = uc_wait(&sh.mbar_pre, ph_pre) && alive; ph_pre ^= 1u;
#pragma unroll
      for (int j = 0; j < MAXP; ++j) {
        if ((warp >> 2) + 4 * j >= MTa || warp >= kTcEpiWarps) continue;  // warp-uniform
        if (KG == 1) read_pos(j, NG, NBG, 0, 2, v[j]);
        if (fl_[j] & 1) {
          add_recv_pre(j, 2, v[j]);
          float s0 = 0.f, q0 = 0.f, s1 = 0.f, q1 = 0.f;
#pragma unroll
          for (int c = 0; c < CK; ++c) {
            s0 += v[j][c]; q0 = fmaf(v[j][c], v[j][c], q0);
            s1 += v[j][CK + c]; q1 = fmaf(v[j][CK + c], v[j][CK + c], q1);
          }
          st[0] += s0; st[1] += q0; st[2] += s1; st[3] += q1;
        }
#pragma unroll
        for (int c = 0; c < CK; ++c) keep[j][c] = v[j][c];
        if (STASH) stash_st(j, 0, &v[j][CK]);
        else {
#pragma unroll
          for (int c = 0; c < CK; ++c) ukeep[j][STASH ? 0 : c] = v[j][CK + c];
        }
      }
    }
    ph_mma ^= 1u;
    mark(0);
    publish_stats(0, st);
    mark(1);
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    tc_cluster_sync();                                                     // #B: sums of every CTA are published; all gate MMAs are complete
    mark(2);
    issue_pre(1, d);                                                       // every thread is past its reads of the gate x-halves
    gather_coef(0, 2, 0, L.rn_w, L.rn_b, L.un_w, L.un_b);
    mark(10);
    // r*h -> window (own rows + halos); h of the own channels is kept for the state update
#pragma unroll
    for (int j = 0; j < MAXP; ++j) {
      if ((warp >> 2) + 4 * j >= MTa || warp >= kTcEpiWarps) continue;    // warp-uniform
      float hv[CK], rh[CK];
      const int wi = q_[j] + Wp + 1;
#pragma unroll
      for (int f = 0; f < NQ; ++f) {
        // lanes past the strip's positions would read the halo row the strip below is writing right now: they read nothing
        const float4 h4 = (fl_[j] & 1) ? win[f * PWa + wi] : make_float4(0.f, 0.f, 0.f, 0.f);
        hv[4 * f] = h4.x; hv[4 * f + 1] = h4.y; hv[4 * f + 2] = h4.z; hv[4 * f + 3] = h4.w;
      }
#pragma unroll
      for (int c = 0; c < CK; ++c) {
        const float2 cf = *reinterpret_cast<const float2*>(&sh.coef[0][c][0]);
        rh[c] = tc_sigmoid(fmaf(keep[j][c], cf.x, cf.y)) * hv[c];
      }
      if (fl_[j] & 1) store_window(j, rh);
      if (STASH) stash_st(j, 1, hv);
      else {
#pragma unroll
        for (int c = 0; c < CK; ++c) hkeep[j][STASH ? 0 : c] = hv[c];
      }
    }
    mark(11);
    asm volatile("fence.proxy.async;" ::: "memory");
    // (#C: the halos of r*h leave and arrive as bulk copies inside issue_conv_split: no cluster barrier)
    mark(3);
    // ================= output: O = OX[d] + conv(r*h; Wo) =================
    issue_conv_split(wo, NBO, N2O, ocol0);
    mark(4);
    // u = sigmoid(GN_u(G_u)) while the tensor core works on the output conv
#pragma unroll
    for (int j = 0; j < MAXP; ++j) {
      if ((warp >> 2) + 4 * j >= MTa || warp >= kTcEpiWarps) continue;
      float uv[CK];
      if (STASH) stash_ld(j, 0, uv);
#pragma unroll
      for (int c = 0; c < CK; ++c) {
        const float2 cu = *reinterpret_cast<const float2*>(&sh.coef[1][c][0]);
        if (STASH) uv[c] = tc_sigmoid(fmaf(uv[c], cu.x, cu.y));
        else ukeep[j][STASH ? 0 : c] = tc_sigmoid(fmaf(ukeep[j][STASH ? 0 : c], cu.x, cu.y));
      }
      if (STASH) stash_st(j, 0, uv);
    }
    st[0] = st[1] = st[2] = st[3] = 0.f;
    {
      float v[MAXP][2 * CK];
      if (KG > 1) {
#pragma unroll
        for (int j = 0; j < MAXP; ++j)
          if ((warp >> 2) + 4 * j < MTa && warp < kTcEpiWarps) read_pos(j, NO, NBO, ocol0, 1, v[j]);
        mark(14);
        exchange(1);                                                       // #D
        mark(15);
      }
      alive = uc_wait(&sh.mbar_pre, ph_pre) && alive; ph_pre ^= 1u;
#pragma unroll
      for (int j = 0; j < MAXP; ++j) {
        if ((warp >> 2) + 4 * j >= MTa || warp >= kTcEpiWarps) continue;
        if (KG == 1) read_pos(j, NO, NBO, ocol0, 1, v[j]);
        if (fl_[j] & 1) {
          add_recv_pre(j, 1, v[j]);
          float s0 = 0.f, q0 = 0.f;
#pragma unroll
          for (int c = 0; c < CK; ++c) { s0 += v[j][c]; q0 = fmaf(v[j][c], v[j][c], q0); }
          st[0] += s0; st[1] += q0;
        }
#pragma unroll
        for (int c = 0; c < CK; ++c) keep[j][c] = v[j][c];
      }
    }
    ph_mma ^= 1u;
    mark(5);
    publish_stats(1, st);
    mark(6);
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    tc_cluster_sync();                                                     // #E
    mark(7);
    if (d + 1 < D) issue_pre(0, d + 1);
    gather_coef(1, 1, 2, L.on_w, L.on_b, L.on_w, L.on_b);
    mark(16);
    // h' = u*h + (1-u)*tanh(GN_o(O)) -> window and halos now, the state history after the barrier arrive (module.py:57)
#pragma unroll
    for (int j = 0; j < MAXP; ++j) {
      if ((warp >> 2) + 4 * j >= MTa || warp >= kTcEpiWarps) continue;
      float uv[CK], hv[CK];
      if (STASH) { stash_ld(j, 0, uv); stash_ld(j, 1, hv); }
      else {
#pragma unroll
        for (int c = 0; c < CK; ++c) { uv[c] = ukeep[j][STASH ? 0 : c]; hv[c] = hkeep[j][STASH ? 0 : c]; }
      }
#pragma unroll
      for (int c = 0; c < CK; ++c) {
        const float2 co = *reinterpret_cast<const float2*>(&sh.coef[2][c][0]);
        const float uu = uv[c];
        keep[j][c] = uu * hv[c] + (1.0f - uu) * tc_tanh(fmaf(keep[j][c], co.x, co.y));
      }
      if (fl_[j] & 1) store_window(j, keep[j]);
    }
    mark(17);
    asm volatile("fence.proxy.async;" ::: "memory");
    mark(18);
    // (#F: likewise for h[d+1])
    mark(19);
#pragma unroll
    for (int j = 0; j < MAXP; ++j) {                                       // global stores drain behind the barrier and the next MMAs
      if (!(fl_[j] & 1)) continue;
      float* sp = L.s + (long long)(kg * CK) * L.s_cs + (long long)(d + 1) * L.px + (long long)y0 * w + sp_[j];
#pragma unroll
      for (int c = 0; c < CK; ++c) sp[(long long)c * L.s_cs] = keep[j][c];
    }
    mark(8);
    if (d + 1 < D) issue_conv_split(wg, NBG, NG, 0);
    mark(9);
  }
  if (!alive && tid == 0) *reinterpret_cast<volatile int*>(err) = 1;
  if (dbg != nullptr && tid == 0 && (rank == 0))
#pragma unroll
    for (int i = 0; i < 20; ++i) dbg[(CH == 8 ? 0 : CH == 16 ? 1 : CH == 32 ? 2 : 3) * 20 + i] = sh.t_acc[i];
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  tc_cluster_sync();                                                       // no CTA leaves while a peer may still write into its shared memory
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(kTmemCols));
}

__global__ void __launch_bounds__(kTcThreads, 1)
red_tc_kernel(const __grid_constant__ TcArgs a) {
  extern __shared__ __align__(128) unsigned char tc_smem[];
  __shared__ TcShared sh;
  const int lv = a.level0 + blockIdx.x / kTcCluster;
#ifndef TC_ONLY
#define TC_ONLY -1
#endif
  if (lv == 0 && (TC_ONLY < 0 || TC_ONLY == 0)) tc_level_run<8, kTcKG[0], kTcMaxP[0], 1>(a.l[0], a.d_begin, a.d_end, a.err, a.dbg, tc_smem, sh);
  else if (lv == 1 && (TC_ONLY < 0 || TC_ONLY == 1)) tc_level_run<16, kTcKG[1], kTcMaxP[1], 2>(a.l[1], a.d_begin, a.d_end, a.err, a.dbg, tc_smem, sh);
  else if (lv == 2 && (TC_ONLY < 0 || TC_ONLY == 2)) tc_level_run<32, kTcKG[2], kTcMaxP[2], 2>(a.l[2], a.d_begin, a.d_end, a.err, a.dbg, tc_smem, sh);
  else if (TC_ONLY < 0 || TC_ONLY == 3) tc_level_run<64, kTcKG[3], kTcMaxP[3], 0>(a.l[3], a.d_begin, a.d_end, a.err, a.dbg, tc_smem, sh);
}

inline size_t tc_pack_bytes(int ch) { return (size_t)216 * ch * ch; }     // KG * 9 * (CK/8) * 2 * 6ch * 16

// Launches the recurrence on the tensor cores; *launched stays false when a level's shape does not fit (rows not a multiple
// of 4 pixels, more tiles per CTA than a thread can hold, shared memory), and the caller then uses the FFMA cluster kernel
// or the per-plane chain.  `wpack[l]` = tc_pack_bytes(ch_l) bytes of scratch per level.
inline int red_tc_launch(TcArgs& a, const float* const* gate_w_h, const float* const* out_w_h, const long long* w_co,
                         char* const* wpack, int* err_flag, long long* dbg, cudaStream_t st, bool* launched, bool pack = true,
                         int only_level = -1) {   // only_level >= 0: a one-cluster grid running that level alone
  *launched = false;
  static const bool verbose = getenv("SATMVS_RED_DEBUG") != nullptr;
  int dev = 0, optin = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
  size_t smem = 0;
  for (int l = 0; l < 4; ++l) {
    TcLevel& L = a.l[l];
    if (L.ch != (8 << l) || L.w % 4 || L.w < 4 || L.h < 1) return SATMVS_OK;
    const TcGeom g = tc_geom(L.ch, kTcKG[l], L.h, L.w);
    // TMEM columns: MT tiles x 4ch gate columns; level 0 keeps its 2ch output-conv columns next to them (the gate region
    // parks u and h meanwhile)
    // (level 0: 6ch; levels 1, 2: 6ch + 2*CK parked columns; level 3: 4ch, output accumulators reuse them)
    const int tmem_cols = g.MT * (l == 0 ? 6 * L.ch : l == 3 ? 4 * L.ch : 6 * L.ch + 2 * L.ch / kTcKG[l]);
    if (g.MT > 4 * kTcMaxP[l] || tmem_cols > 512 || (size_t)g.PWa * 16 >= (1u << 18)) {
      if (verbose) fprintf(stderr, "red_tc_launch: level %d needs %d tiles per CTA\n", l, g.MT);
      return SATMVS_OK;
    }
    L.R = g.R; L.MT = g.MT; L.PWa = g.PWa; L.NPP = g.NPP;
    smem = g.smem > smem ? g.smem : smem;
  }
  if (smem + sizeof(TcShared) + 1024 > (size_t)optin) {
    if (verbose) fprintf(stderr, "red_tc_launch: %zu bytes of shared memory needed\n", smem);
    return SATMVS_OK;
  }
  auto declined = [&](const char* why, cudaError_t e) {
    if (verbose) fprintf(stderr, "red_tc_launch: falling back: %s (%s)\n", why, cudaGetErrorString(e));
    cudaGetLastError();
    return SATMVS_OK;
  };
  cudaError_t e;
  if ((e = cudaFuncSetAttribute(red_tc_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1)) != cudaSuccess)
    return declined("non-portable cluster size", e);
  if ((e = cudaFuncSetAttribute(red_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)) != cudaSuccess)
    return declined("dynamic shared memory", e);
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((only_level < 0 ? 4 : 1) * kTcCluster); cfg.blockDim = dim3(kTcThreads); cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = kTcCluster; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  int nclusters = 0;
  a.level0 = only_level < 0 ? 0 : only_level;
  if ((e = cudaOccupancyMaxActiveClusters(&nclusters, red_tc_kernel, &cfg)) != cudaSuccess || nclusters < 4)
    return declined("fewer than 4 co-resident clusters", e);
  for (int l = 0; l < 4; ++l) {
    if (pack && (only_level < 0 || only_level == l)) {
      TcPack p{gate_w_h[l], out_w_h[l], w_co[l], a.l[l].ch, kTcKG[l], reinterpret_cast<float4*>(wpack[l])};
      const int total = (int)(tc_pack_bytes(a.l[l].ch) / 16);
      tc_pack_kernel<<<ceil_div(total, 256), 256, 0, st>>>(p);
    }
    a.l[l].wpack = reinterpret_cast<const float4*>(wpack[l]);
  }
  a.err = err_flag; a.dbg = dbg;
  if ((e = cudaLaunchKernelEx(&cfg, red_tc_kernel, a)) != cudaSuccess) return declined("launch", e);
  *launched = true;
  return SATMVS_OK;
}

}  // namespace satmvs
