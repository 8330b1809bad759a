// umma_probe.cu — stand-alone check of the tcgen05 building blocks used by the split-precision conv:
// SWIZZLE_NONE K-major shared-memory descriptors over a [channel quad][position] float4 tile (the
// sweep's re-pack layout), shifted start addresses (= conv taps), kind::tf32 MMAs accumulated in
// TMEM, tcgen05.commit -> mbarrier, tcgen05.ld epilogue.  Prints max |err| against a CPU GEMM.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o umma_probe umma_probe.cu
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <vector>
#include <cuda_runtime.h>

typedef unsigned long long u64;
constexpr int M = 128, N = 32, K = 32;      // K = 32 channels = 8 quads = 4 MMAs of K 8
constexpr int PLANE = 200;                   // positions per quad plane in shared memory
constexpr int SHIFT = 37;                    // tap shift: A row m is position SHIFT + m

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

__device__ __forceinline__ u64 make_desc(unsigned saddr, unsigned lbo_bytes, unsigned sbo_bytes) {
  u64 d = 0;
  d |= (u64)((saddr >> 4) & 0x3fff);
  d |= (u64)((lbo_bytes >> 4) & 0x3fff) << 16;
  d |= (u64)((sbo_bytes >> 4) & 0x3fff) << 32;
  d |= (u64)1 << 46;                         // version 1 (sm_100)
  return d;                                  // layout_type 0 = SWIZZLE_NONE, base_offset 0
}

__global__ void probe(const float* A, const float* B, float* C, int* status, int mode) {
  extern __shared__ __align__(128) unsigned char sm[];
  float4* a_t = reinterpret_cast<float4*>(sm);                       // [K/4][PLANE]
  float4* b_t = a_t + (K / 4) * PLANE;                               // [K/4][N]
  __shared__ unsigned tmem_base;
  __shared__ __align__(8) u64 bar;
  const int tid = threadIdx.x, warp = tid >> 5;

  // A[m][k] row-major in global -> a_t[k/4][SHIFT + m]; everything else zero
  for (int i = tid; i < (K / 4) * PLANE; i += blockDim.x) a_t[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  __syncthreads();
  for (int i = tid; i < M * (K / 4); i += blockDim.x) {
    const int m = i % M, q = i / M;
    a_t[q * PLANE + SHIFT + m] = make_float4(A[m * K + 4 * q], A[m * K + 4 * q + 1], A[m * K + 4 * q + 2], A[m * K + 4 * q + 3]);
  }
  for (int i = tid; i < N * (K / 4); i += blockDim.x) {
    const int n = i % N, q = i / N;
    b_t[q * N + n] = make_float4(B[n * K + 4 * q], B[n * K + 4 * q + 1], B[n * K + 4 * q + 2], B[n * K + 4 * q + 3]);
  }
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base)), "r"(32));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");       // generic smem writes -> async proxy (tensor core)
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const unsigned tm = tmem_base;

  if (tid == 0) {
    // instruction descriptor: D f32, A/B tf32, both K-major, N >> 3 at bit 17, M >> 4 at bit 24
    const unsigned idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((unsigned)(N >> 3) << 17) | ((unsigned)(M >> 4) << 24);
    for (int ks = 0; ks < K / 8; ++ks) {
      const u64 da = make_desc(smem_u32(a_t + (2 * ks) * PLANE + SHIFT), PLANE * 16, 128);
      const u64 db = make_desc(smem_u32(b_t + (2 * ks) * N), N * 16, 128);
      const unsigned acc = ks > 0;
      asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                   "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                   ::"r"(tm), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
  }
  // bounded wait on the mbarrier (phase 0)
  unsigned ok = 0;
  for (int it = 0; it < 2000000 && !ok; ++it)
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(smem_u32(&bar)), "r"(0) : "memory");
  if (!ok) { if (tid == 0) *status = 1; }
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  if (ok) {
    // warp w reads TMEM lanes 32w .. 32w+31 (rows), 8 columns per load
    for (int c0 = 0; c0 < N; c0 += 8) {
      unsigned r[8];
      const unsigned taddr = tm + ((unsigned)(warp * 32) << 16) + (unsigned)c0;
      asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                   : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "r"(taddr));
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      for (int j = 0; j < 8; ++j) C[tid * N + c0 + j] = __uint_as_float(r[j]);
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tm), "r"(32));
  (void)mode;
}

static float tf32_trunc(float x) { unsigned u; memcpy(&u, &x, 4); u &= 0xffffe000u; memcpy(&x, &u, 4); return x; }
static float tf32_rn(float x) { unsigned u; memcpy(&u, &x, 4); u += 0xfffu + ((u >> 13) & 1u); u &= 0xffffe000u; memcpy(&x, &u, 4); return x; }

int main() {
  std::vector<float> A(M * K), B(N * K), C(M * N, -1.f);
  srand(1);
  for (auto& v : A) v = (float)rand() / RAND_MAX * 2.f - 1.f;
  for (auto& v : B) v = (float)rand() / RAND_MAX * 2.f - 1.f;
  float *dA, *dB, *dC; int* dS;
  cudaMalloc(&dA, A.size() * 4); cudaMalloc(&dB, B.size() * 4); cudaMalloc(&dC, C.size() * 4); cudaMalloc(&dS, 4);
  cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(dC, C.data(), C.size() * 4, cudaMemcpyHostToDevice);
  cudaMemset(dS, 0, 4);
  const size_t smem = ((K / 4) * PLANE + (K / 4) * N) * 16;
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  probe<<<1, 128, smem>>>(dA, dB, dC, dS, 0);
  cudaError_t e = cudaDeviceSynchronize();
  int st = 0;
  cudaMemcpy(&st, dS, 4, cudaMemcpyDeviceToHost);
  cudaMemcpy(C.data(), dC, C.size() * 4, cudaMemcpyDeviceToHost);
  printf("cuda: %s, status %d\n", cudaGetErrorString(e), st);
  double e_full = 0, e_tr = 0, e_rn = 0;
  for (int m = 0; m < M; ++m)
    for (int n = 0; n < N; ++n) {
      double s0 = 0, s1 = 0, s2 = 0;
      for (int k = 0; k < K; ++k) {
        s0 += (double)A[m * K + k] * B[n * K + k];
        s1 += (double)tf32_trunc(A[m * K + k]) * tf32_trunc(B[n * K + k]);
        s2 += (double)tf32_rn(A[m * K + k]) * tf32_rn(B[n * K + k]);
      }
      e_full = fmax(e_full, fabs(C[m * N + n] - s0)); e_tr = fmax(e_tr, fabs(C[m * N + n] - s1)); e_rn = fmax(e_rn, fabs(C[m * N + n] - s2));
    }
  printf("max |err| vs fp32 inputs %.3e, vs tf32-truncated inputs %.3e, vs tf32-rounded inputs %.3e\n", e_full, e_tr, e_rn);
  printf("C[0][0..3] = %f %f %f %f ; C[127][31] = %f\n", C[0], C[1], C[2], C[3], C[127 * N + 31]);
  return 0;
}
