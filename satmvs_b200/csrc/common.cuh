// common.cuh — error plumbing shared by every translation unit of libsatmvs_b200.so
#pragma once
#include <cuda_runtime.h>
#include <cstdarg>
#include <cstdint>
#include <cstdio>

#include "../../include/satmvs_b200.h"

namespace satmvs {

// thread-local last-error text (satmvs_last_error), defined in abi.cu
void set_error(const char* fmt, ...);

inline int fail_invalid(const char* what) {
  set_error("invalid argument: %s", what);
  return SATMVS_EINVAL;
}

inline int check_launch(const char* kernel) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("%s: %s", kernel, cudaGetErrorString(e));
    return SATMVS_ECUDA;
  }
  return SATMVS_OK;
}

#define SATMVS_REQUIRE(cond) \
  do { if (!(cond)) return ::satmvs::fail_invalid(#cond); } while (0)

// Device-side failures that cannot be returned by the launch that detects them (a tensor-core completion that never arrives,
// a flag that a co-resident CTA never sets) are written to a pinned, device-mapped int per (host thread, device): the kernel
// finishes without trapping the context, and the NEXT library call of the thread -- or satmvs_async_error() after a
// synchronisation -- reports it.  Codes: 1 tcgen05 completion timeout, 2 producer timeout, 3 cluster flag timeout.
int* async_error_devptr();    // device pointer the kernels write to (abi.cu)
int async_error_poll();       // host: returns and clears the code (0 = none)

#define SATMVS_CHECK_ASYNC() \
  do { if (int ae_ = ::satmvs::async_error_poll()) { ::satmvs::set_error("an earlier launch of this thread reported device-side error %d " \
       "(1 tensor-core completion timeout, 2 producer timeout, 3 cluster flag timeout): its results are invalid", ae_); return SATMVS_ECUDA; } } while (0)

constexpr int kNumSMs = 148;  // B200

inline int ceil_div(int64_t a, int64_t b) { return (int)((a + b - 1) / b); }

}  // namespace satmvs
