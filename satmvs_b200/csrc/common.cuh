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

constexpr int kNumSMs = 148;  // B200

inline int ceil_div(int64_t a, int64_t b) { return (int)((a + b - 1) / b); }

}  // namespace satmvs
