// abi.cu — version and thread-local error text of libsatmvs_b200.so
#include "common.cuh"
#include "prof.cuh"
#include <algorithm>
#include <cstring>
#include <utility>
#include <vector>
#include <cstdlib>

namespace satmvs {
static thread_local char g_error[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_error, sizeof(g_error), fmt, ap);
  va_end(ap);
}
struct AsyncErr { int* host = nullptr; int* dev = nullptr; int devid = -1; };
static AsyncErr& async_err() {
  static thread_local AsyncErr slot[16];
  int dev = 0;
  cudaGetDevice(&dev);
  AsyncErr& a = slot[dev & 15];
  if (a.devid != dev) {
    a.devid = dev;
    if (cudaHostAlloc(reinterpret_cast<void**>(&a.host), sizeof(int), cudaHostAllocMapped) == cudaSuccess) {
      *a.host = 0;
      if (cudaHostGetDevicePointer(reinterpret_cast<void**>(&a.dev), a.host, 0) != cudaSuccess) a.dev = nullptr;
    }
    if (a.dev == nullptr) cudaGetLastError();
  }
  return a;
}
int* async_error_devptr() { return async_err().dev; }
int async_error_poll() {
  AsyncErr& a = async_err();
  if (a.host == nullptr) return 0;
  const int v = *reinterpret_cast<volatile int*>(a.host);
  if (v) *reinterpret_cast<volatile int*>(a.host) = 0;
  return v;
}
ProfState& prof_state() {
  static ProfState s;
  return s;
}
}  // namespace satmvs

extern "C" {
int satmvs_profile_begin(void) {
  satmvs::ProfState& s = satmvs::prof_state();
  std::lock_guard<std::mutex> lk(s.mu);
  for (auto& v : s.ev) { for (cudaEvent_t e : v) if (e) cudaEventDestroy(e); v.clear(); }
  ++s.generation;
  s.on = true;
  return SATMVS_OK;
}

int satmvs_profile_end_ex(float* ms_by_class, int* launches_by_class, float* busy_ms_by_class) {
  satmvs::ProfState& s = satmvs::prof_state();
  s.on = false;
  cudaDeviceSynchronize();
  std::lock_guard<std::mutex> lk(s.mu);
  ++s.generation;
  for (int p = 0; p < satmvs::kProfCount; ++p)       // a scope still open on another thread: drop its unpaired start
    for (size_t i = 0; i + 1 < s.ev[p].size(); i += 2)
      if (s.ev[p][i + 1] == nullptr) { cudaEventDestroy(s.ev[p][i]); s.ev[p].erase(s.ev[p].begin() + i, s.ev[p].begin() + i + 2); i -= 2; }
  if (getenv("SATMVS_PROF_DUMP")) {   // timeline of every instrumented launch, relative to the first recorded event
    cudaEvent_t ref = nullptr;
    float best = 0.0f;
    for (int p = 0; p < satmvs::kProfCount; ++p)
      for (cudaEvent_t e : s.ev[p]) {
        if (!ref) { ref = e; continue; }
        float ms = 0.0f;
        if (cudaEventElapsedTime(&ms, ref, e) == cudaSuccess && ms < best) { best = ms; }
      }
    for (int p = 0; p < satmvs::kProfCount && ref; ++p)
      for (size_t i = 0; i + 1 < s.ev[p].size(); i += 2) {
        float a = 0.0f, b = 0.0f;
        cudaEventElapsedTime(&a, ref, s.ev[p][i]);
        cudaEventElapsedTime(&b, ref, s.ev[p][i + 1]);
        fprintf(stderr, "prof class %d launch %zu: %.1f .. %.1f us\n", p, i / 2, (a - best) * 1e3f, (b - best) * 1e3f);
      }
  }
  cudaEvent_t ref0 = nullptr;
  for (int p = 0; p < satmvs::kProfCount && !ref0; ++p)
    if (!s.ev[p].empty()) ref0 = s.ev[p][0];
  for (int p = 0; p < satmvs::kProfCount; ++p) {
    float tot = 0.0f;
    std::vector<std::pair<float, float>> iv;
    for (size_t i = 0; i + 1 < s.ev[p].size(); i += 2) {
      float ms = 0.0f;
      cudaEventElapsedTime(&ms, s.ev[p][i], s.ev[p][i + 1]);
      tot += ms;
      if (busy_ms_by_class) {
        float a = 0.0f, b = 0.0f;
        cudaEventElapsedTime(&a, ref0, s.ev[p][i]);
        cudaEventElapsedTime(&b, ref0, s.ev[p][i + 1]);
        iv.emplace_back(a, b);
      }
    }
    if (busy_ms_by_class) {       // time during which at least one launch of the class was running (launches on concurrent streams overlap)
      std::sort(iv.begin(), iv.end());
      float busy = 0.0f, lo = 0.0f, hi = 0.0f;
      bool open = false;
      for (const auto& x : iv) {
        if (!open) { lo = x.first; hi = x.second; open = true; }
        else if (x.first <= hi) { if (x.second > hi) hi = x.second; }
        else { busy += hi - lo; lo = x.first; hi = x.second; }
      }
      if (open) busy += hi - lo;
      busy_ms_by_class[p] = busy;
    }
    if (ms_by_class) ms_by_class[p] = tot;
    if (launches_by_class) launches_by_class[p] = (int)(s.ev[p].size() / 2);
  }
  for (int p = 0; p < satmvs::kProfCount; ++p) {     // after every class has been measured against the common reference event
    for (cudaEvent_t e : s.ev[p]) cudaEventDestroy(e);
    s.ev[p].clear();
  }
  return SATMVS_OK;
}

int satmvs_profile_end(float* ms_by_class, int* launches_by_class) {
  return satmvs_profile_end_ex(ms_by_class, launches_by_class, nullptr);
}

int satmvs_async_error(void) { return satmvs::async_error_poll(); }
int satmvs_abi_version(void) { return SATMVS_ABI_VERSION; }
const char* satmvs_last_error(void) { return satmvs::g_error; }
}
